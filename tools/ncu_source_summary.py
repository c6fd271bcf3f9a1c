#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output (SASS view) of one kernel: instruction mix by opcode, share of
executed warp instructions and of stall samples, and the hottest instructions.  Run in the build container:

    ncu -i gpurun_out/x.ncu-rep --page source --csv > /tmp/x.csv && python tools/ncu_source_summary.py /tmp/x.csv
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path)))
    print(rows[0][1])
    head = rows[1]
    ix = {h: i for i, h in enumerate(head)}
    i_src, i_inst, i_samp = ix["Source"], ix["Instructions Executed"], ix["# Samples"]
    stalls = [h for h in head if h.startswith("stall_") and "Not Issued" not in h]
    by_op = defaultdict(lambda: [0, 0])
    tot_inst = tot_samp = 0
    per = []
    stall_tot = defaultdict(int)
    for r in rows[2:]:
        if len(r) <= i_samp:
            continue
        try:
            inst, samp = int(r[i_inst] or 0), int(r[i_samp] or 0)
        except ValueError:
            continue
        src = r[i_src].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDG", "LDS", "STS", "STG", "SHFL", "MUFU", "LDL", "STL")) and "." in op else "")
        by_op[op][0] += inst
        by_op[op][1] += samp
        tot_inst += inst
        tot_samp += samp
        st = {s: int(r[ix[s]] or 0) for s in stalls}
        for s, v in st.items():
            stall_tot[s] += v
        per.append((samp, inst, r[ix["Address"]], src, st))
    print("warp instructions executed: %d, stall samples: %d" % (tot_inst, tot_samp))
    print("\n%-12s %14s %7s %9s %7s" % ("opcode", "warp inst", "share", "samples", "share"))
    for op, (inst, samp) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-12s %14d %6.1f%% %9d %6.1f%%" % (op, inst, 100.0 * inst / max(tot_inst, 1), samp, 100.0 * samp / max(tot_samp, 1)))
    print("\nstall reasons (all samples):")
    for s, v in sorted(stall_tot.items(), key=lambda kv: -kv[1])[:10]:
        print("  %-24s %8d %5.1f%%" % (s, v, 100.0 * v / max(tot_samp, 1)))
    # how often each instruction ran: set-up / finish code runs once per item, the row loop once per warp-row --
    # classes of equal execution count (within 10 %) split the kernel into its phases
    print("\ninstructions by execution count (phases of the kernel):")
    classes = []
    for samp, inst, addr, src, st in sorted(per, key=lambda t: -t[1]):
        if inst == 0:
            continue
        if classes and inst >= 0.9 * classes[-1][0]:
            classes[-1][1] += 1
            classes[-1][2] += inst
            classes[-1][3] += samp
        else:
            classes.append([inst, 1, inst, samp])
    print("  %14s %8s %14s %7s %7s" % ("runs (each)", "SASS", "warp inst", "share", "samples"))
    for runs, count, inst, samp in sorted(classes, key=lambda c: -c[2])[:12]:
        print("  %14d %8d %14d %6.1f%% %6.1f%%" % (runs, count, inst, 100.0 * inst / max(tot_inst, 1),
                                                    100.0 * samp / max(tot_samp, 1)))
    print("\nhottest instructions by samples:")
    for samp, inst, addr, src, st in sorted(per, key=lambda t: -t[0])[:top]:
        main_st = max(st.items(), key=lambda kv: kv[1]) if st else ("", 0)
        print("  %6d %5.2f%% %10d  %s  [%s %d]" % (samp, 100.0 * samp / max(tot_samp, 1), inst, src[:90], main_st[0], main_st[1]))


if __name__ == "__main__":
    main()
