"""Markdown table of one step out of an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: the launches
between the last two `adam_kl_kernel` launches of the list (one eager C3 / C2 step), aggregated by kernel.
`python profiles/microbench/launch_list_table.py profiles/r02_c3_launches_ncu.csv [min_share_percent]`."""
import csv
import re
import sys


LIB = re.compile(r"contract_pair|contract_tma|contract_kernel|wgrad_tma|wgrad_kernel|nchw_to_nhwc|adam_kl|kl_kernel|ce_fwd|ce_bwd|"
                 r"stddev_kernel|weight_layout|weight_unlayout|pack_kernel|prune_|materialize|bias_grad|im2col|col2im|peer_")


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("bnn::contract::", "").replace("bnn::", "")
    name = name.replace("contract::", "")
    name = re.sub(r"^at::native::", "", name)
    return name.split("(")[0][:80]


def main():
    path = sys.argv[1]
    min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    head, data = rows[0], rows[1:]
    ki, vi, ui = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
    launches = []
    for r in data:
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        us = v / 1e3 if unit.startswith("n") else (v * 1e3 if unit.startswith("m") else v)
        launches.append((r[ki], us))
    ends = [i for i, (n, _) in enumerate(launches) if "adam_kl_kernel" in n]
    # a step ends with the optimizer's launches (two adam_kl launches per step at C3: groups of tensors)
    step_ends = [e for j, e in enumerate(ends) if j + 1 == len(ends) or ends[j + 1] != e + 1]
    lo, hi = step_ends[-2] + 1, step_ends[-1] + 1
    step = launches[lo:hi]
    total = sum(us for _, us in step)
    agg = {}
    for n, us in step:
        k = short(n)
        a = agg.setdefault(k, [0.0, 0, n])
        a[0] += us
        a[1] += 1
    lib = sum(a[0] for k, a in agg.items() if LIB.search(k))
    print(f"one eager step (launches {lo}..{hi - 1} of the list): {len(step)} launches, {total:.1f} us of kernel time "
          f"(cold caches, serialised)")
    print(f"library kernels: {lib:.1f} us = {100 * lib / total:.1f} % of the step; the rest is the deterministic torch / cuDNN "
          f"trunk, the full-covariance head and fills\n")
    print("| us | share | launches | kernel |\n|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if 100 * a[0] / total >= min_share:
            tag = " (library)" if LIB.search(k) else ""
            print(f"| {a[0]:.1f} | {100 * a[0] / total:.1f} % | {a[1]} | `{k}`{tag} |")


if __name__ == "__main__":
    main()
