#!/usr/bin/env python
"""Turn `.ncu-rep` captures (gpurun_out/) into the two tracked summaries under profiles/:

    python tools/ncu_summaries.py gpurun_out/r02_hg38_10kb_r1_final.ncu-rep [...]

  <name>_ncu_full_summary.csv   the launch / throughput / pipe / stall / local-memory metrics of `--page raw`
  <name>_source_summary.txt     instruction mix, stall reasons and hottest instructions of `--page source`
                                (tools/ncu_source_summary.py)
Runs in the build container (no GPU needed)."""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("Kernel Name", "gpu__time_duration.sum", "launch__", "dram__bytes", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__inst_executed", "sm__inst_executed_pipe", "sm__pipe_fp64", "smsp__average_warps_issue_stalled",
        "smsp__issue_active", "smsp__warps_eligible", "sm__warps_active", "l1tex__t_bytes_pipe_lsu_mem_local",
        "smsp__thread_inst_executed_per_inst", "sm__throughput", "gpu__dram_throughput", "l1tex__data_bank_conflicts",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__sass_thread_inst_executed_op_d")


def main():
    for rep in sys.argv[1:]:
        name = os.path.basename(rep)[:-len(".ncu-rep")]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        head, units, vals = rows[0], rows[1], rows[2]
        with open(os.path.join(ROOT, "profiles", name + "_ncu_full_summary.csv"), "w") as fh:
            fh.write("metric,value,unit\n")
            for a, b, c in zip(head, units, vals):
                if a.startswith(KEEP):
                    fh.write('"%s","%s","%s"\n' % (a, c, b))
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        tmp = "/tmp/" + name + "_src.csv"
        open(tmp, "w").write(src)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_source_summary.py"), tmp, "30"],
                             capture_output=True, text=True).stdout
        open(os.path.join(ROOT, "profiles", name + "_source_summary.txt"), "w").write(out)
        print(name, "->", len(out.splitlines()), "lines")


if __name__ == "__main__":
    main()
