#!/usr/bin/env python
"""Per source line: executed warp instructions and stall samples of one kernel, from
`ncu -i x.ncu-rep --page source --csv --print-source cuda,sass` (needs -lineinfo and --import-source on).

    python tools/ncu_lines_summary.py /tmp/x.csv [top]
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    lines = {}
    cur_file = None
    head = None
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            head = None
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            head = r
            i_inst = head.index("Instructions Executed")
            i_samp = head.index("# Samples")
            i_noinst = head.index("stall_no_inst")
            i_long = head.index("stall_long_sb")
            continue
        if head is None or not r[0].strip().isdigit() or r[2].strip() not in ("", "-"):
            continue          # SASS rows carry an address
        try:
            inst, samp = int(r[i_inst] or 0), int(r[i_samp] or 0)
        except ValueError:
            continue
        if inst == 0 and samp == 0:
            continue
        key = (cur_file, int(r[0]))
        old = lines.get(key, (0, 0, 0, 0, ""))
        lines[key] = (old[0] + inst, old[1] + samp, old[2] + int(r[i_noinst] or 0), old[3] + int(r[i_long] or 0),
                      r[1].strip()[:100])
    tot_i = sum(v[0] for v in lines.values())
    tot_s = sum(v[1] for v in lines.values())
    by_file = defaultdict(lambda: [0, 0])
    for (f, _), v in lines.items():
        by_file[f][0] += v[0]
        by_file[f][1] += v[1]
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    for f, (i, s) in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
        print("  %-24s %5.1f%% inst %5.1f%% samples" % (f, 100.0 * i / max(tot_i, 1), 100.0 * s / max(tot_s, 1)))
    print("\n%-26s %6s %6s %7s %7s  source" % ("file:line", "inst%", "samp%", "no_inst", "long_sb"))
    for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-26s %5.1f%% %5.1f%% %7d %7d  %s" % ("%s:%d" % (f, ln), 100.0 * v[0] / max(tot_i, 1),
                                                    100.0 * v[1] / max(tot_s, 1), v[2], v[3], v[4]))


if __name__ == "__main__":
    main()
