#!/bin/bash
# Round-2 evidence on ONE B200 (run under gpurun from the repo root): GPU tests, smoke(), the default bench line + the
# reference arm, ncu launch list of the C3 step, `ncu --set full` of the contraction kernels (C3 conv layer, C4 layer), of
# the KL kernels and of the one-sweep prune kernels.  Outputs under gpurun_out/r02f_*; every profiled command first runs
# without the profiler.
O=gpurun_out
NCU="ncu --clock-control none"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/r02f_pytest.log; tail -2 $O/r02f_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > $O/r02f_bench.json 2> $O/r02f_bench.err; echo bench rc=$?
timeout 300 python bench.py --impl reference > $O/r02f_bench_ref.json 2> $O/r02f_bench_ref.err; echo ref rc=$?
timeout 300 python bench.py --workload c2 --no-extras > $O/r02f_bench_c2.json 2>> $O/r02f_bench.err
timeout 300 python bench.py --workload c4 --no-extras > $O/r02f_bench_c4.json 2>> $O/r02f_bench.err
python profiles/microbench/prof_conv.py c3 shared > $O/r02f_conv.txt 2>&1
python profiles/microbench/prof_conv.py c3 per >> $O/r02f_conv.txt 2>&1
python profiles/microbench/prof_conv.py c2 shared >> $O/r02f_conv.txt 2>&1
python profiles/microbench/prof_c4.py 32 > $O/r02f_c4.txt 2>&1
python profiles/microbench/prof_prune_into.py 16 > $O/r02f_prune.txt 2>&1
python profiles/microbench/prof_prune_into.py 64 --trace >> $O/r02f_prune.txt 2>&1
cat $O/r02f_conv.txt $O/r02f_c4.txt; head -4 $O/r02f_prune.txt
$NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file $O/r02f_c3_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --tf32-peak 625 > $O/r02f_ncu_c3.log 2>&1
# prof_conv.py: 9 warm-up launches (3 x fwd, dgrad, wgrad), then 5 x fwd (9-13), 5 x dgrad (14-18), 5 x wgrad (19-23)
$NCU --set full --import-source on -k regex:"contract_|wgrad_tma" -s 13 -c 7 -o $O/r02f_conv_c3 python profiles/microbench/prof_conv.py c3 shared > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"contract_|wgrad_tma" -s 6 -c 3 -o $O/r02f_c4 python profiles/microbench/prof_c4.py 32 > /dev/null 2>&1
$NCU --set full -k regex:"kl_kernel" -s 2 -c 2 -o $O/r02f_kl python profiles/microbench/prof_klprune.py 16 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"prune_sweep_into|prune_sample_kernel|prune_resolve|prune_finish|prune_bracket" -s 10 -c 5 -o $O/r02f_prune python profiles/microbench/prof_prune_into.py 16 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -k regex:prune_ -c 40 --csv --log-file $O/r02f_prune_launches.csv python profiles/microbench/prof_prune_into.py 16 > /dev/null 2>&1
ls -la $O/r02f*.ncu-rep
