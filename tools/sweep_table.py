#!/usr/bin/env python
"""Markdown table of two `bench.py --workload sweep` lines, cell by cell: python tools/sweep_table.py before.json after.json"""
import json
import sys


def cells(path):
    line = json.loads(open(path).read().strip().splitlines()[-1])
    return {(c["bins"], c["window"]): c for c in line["cells"]}, line


def main():
    before, _ = cells(sys.argv[1])
    after, line = cells(sys.argv[2])
    print("| bins | window | ms per ≈100 k pairs, before → after | speed-up | HBM fraction before → after | FP64 fraction after |")
    print("|---|---|---|---|---|---|")
    for key in sorted(after, key=lambda k: (k[0], ["r1", "a3", "a5", "a7", "a9", "a11"].index(k[1]))):
        a, b = after[key], before.get(key)
        win = "-r 1" if key[1] == "r1" else key[1]
        if b is None:
            print("| %d | %s | %.2f | — | **%.3f** | %.3f |" % (key[0], win, a["kernel_us"] / 1e3, a["roofline_frac"], a["fp64_frac"]))
            continue
        print("| %d | %s | %.2f → %.2f | %.2f | %.3f → **%.3f** | %.3f |" % (
            key[0], win, b["kernel_us"] / 1e3, a["kernel_us"] / 1e3, b["kernel_us"] / a["kernel_us"],
            b["roofline_frac"], a["roofline_frac"], a["fp64_frac"]))
    print("\naggregate: %.4g pairs/s, roofline %.3f" % (line["value"], line["roofline"]["frac"]))


if __name__ == "__main__":
    main()
