#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --eager`: per-kernel time share of
ONE fine-tune step (the launches between two consecutive optimizer-clock kernels).  Usage:
    python scripts/summarize_launches.py gpurun_out/launches.csv [step_index] > profiles/<name>.md"""
import collections
import csv
import re
import sys


def short(n: str) -> str:
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("jl::", "")
    return n[:64]


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    # a step starts with the optimizer-clock kernel (jl_adamw_advance, first launch of the step body)
    idx = [i for i, r in enumerate(rows) if "adamw_advance" in r["Kernel Name"]]
    a, b = idx[which], idx[which + 1]
    step = rows[a: b]
    tot = sum(float(r["Metric Value"]) for r in step)
    agg = collections.OrderedDict()
    for r in step:
        d = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        d[0] += 1
        d[1] += float(r["Metric Value"])
    print(f"# per-kernel device time of one fine-tune step (ncu gpu__time_duration.sum, cold-cache, serialised)\n")
    print(f"source: `{path}`, launches {a + 1}..{b} ({len(step)} launches, {tot / 1e3:.1f} us in total)\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {c} | {t / 1e3:.1f} | {100 * t / tot:.1f} % |")


if __name__ == "__main__":
    main()
