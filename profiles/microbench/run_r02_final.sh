#!/bin/bash
# Final evidence of round 2 on ONE B200 (run under gpurun from the repo root), after the staged-epilogue change: GPU tests,
# smoke(), the default bench line + the reference arm, C2 / C4 lines, the ncu launch list and the CUPTI graph trace of the
# C3 step.  Outputs under gpurun_out/r02m_*; every profiled command first runs without the profiler.
O=gpurun_out
NCU="ncu --clock-control none"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/r02m_pytest.log; tail -2 $O/r02m_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > $O/r02m_bench.json 2> $O/r02m_bench.err; echo bench rc=$?
timeout 300 python bench.py --impl reference > $O/r02m_bench_ref.json 2> $O/r02m_bench_ref.err; echo ref rc=$?
timeout 300 python bench.py --workload c2 --no-extras > $O/r02m_bench_c2.json 2>> $O/r02m_bench.err
timeout 300 python bench.py --workload c4 --no-extras > $O/r02m_bench_c4.json 2>> $O/r02m_bench.err
timeout 200 python profiles/microbench/trace_step.py c3 > $O/r02m_c3_graph_trace.txt 2>&1
timeout 200 python profiles/microbench/trace_step.py c2 > $O/r02m_c2_graph_trace.txt 2>&1
$NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file $O/r02m_c3_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --tf32-peak 625 > $O/r02m_ncu_c3.log 2>&1
python - <<'PY'
import json
for f in ("r02m_bench", "r02m_bench_c2", "r02m_bench_c4", "r02m_bench_ref"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d.get("ms_per_step"), d.get("value"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
