"""bench.py's contract, the parts that need no GPU: the `config` object is a function of the command line only
(both arms print the same one), the CPU reference arm's generators load without the package (so that arm maps no
CUDA library), the roofline arithmetic, and the per-pass DRAM traffic read from ncu's csv."""
import json
import subprocess
import sys

import numpy as np
import pytest

import bench
from tests.conftest import ROOT


@pytest.mark.parametrize("extra", [[], ["--window", "a7"], ["--workload", "chr2_25kb"],
                                   ["--workload", "dlbcl_5kb_bg1000"], ["--workload", "sweep"],
                                   ["--workload", "chr2_25kb", "--bin-size", "5000", "--window-bp", "3000000"]])
def test_config_is_a_function_of_the_command_line(extra):
    ours = bench.workload_config(bench.parse_args(extra + ["--gpus", "8", "--steps", "7"]))
    ref = bench.workload_config(bench.parse_args(extra + ["--impl", "reference"]))
    assert ours == ref and json.loads(json.dumps(ours)) == ours
    assert "workload" in ours and "model" not in ours and "pairs" in ours and "bins_per_pair" in ours


def test_default_workload_is_the_north_star_configuration():
    w = bench.workload_spec(bench.parse_args([]))
    assert (w["name"], w["cfg"], w["bin"], w["window_bp"], w["pairs"], w["bins_per_pair"], w["scaling"]) == \
        ("hg38_10kb", 2, 10000, 2000000, 101354, 201, "strong")
    assert bench.parse_args([]).gpus == 1 and bench.parse_args([]).warmup >= 3


def test_reference_arm_generators_load_without_the_package():
    code = ("import sys, bench; S = bench.synth_module(standalone=True); "
            "w = bench.workload_spec(bench.parse_args(['--impl', 'reference', '--workload', 'chr2_25kb'])); "
            "from oracle import chess_oracle as O; "
            "g, r, q, p = bench.make_workload(w, region_cls=O.Region, truncate_bins=400); "
            "bad = [m for m in sys.modules if m == 'chess_b200' or m.startswith('chess_b200.')]; "
            "maps = open('/proc/self/maps').read(); "
            "print(len(p), bad, 'libchess_b200' in maps, 'libcudart' in maps)")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-400:]
    n_pairs, bad, lib, cudart = out.stdout.strip().rsplit(" ", 3)
    assert int(n_pairs.split()[0]) > 0 and "[]" in bad and lib == "False" and cudart == "False"


def test_roofline_arithmetic():
    r = bench.roofline(alg_bytes=65.5e9, kernel_ms=12.5, n_pixels=101354 * 201.0 * 201, window="r1", compute_sn=True,
                       what="x", traffic=5)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["traffic"] == 5
    assert np.isclose(r["achieved"], 65.5e9 / 12.5e-3 / 1e9) and np.isclose(r["frac"], r["achieved"] / r["peak"])
    sec = r["secondary"]
    assert sec["bound"] == "fp64" and sec["flops_per_pixel"] == bench.flops_per_pixel("r1", True) == 24
    assert np.isclose(sec["achieved"], 101354 * 201.0 * 201 * 24 / 12.5e-3 / 1e12)
    assert bench.flops_per_pixel("a7", True) == 59 and bench.flops_per_pixel("a7", False) == 43


def _ncu_csv(launches):
    head = ["ID", "Process ID", "Process Name", "Host Name", "Kernel Name", "Context", "Stream", "Block Size",
            "Grid Size", "Device", "CC", "Section Name", "Metric Name", "Metric Unit", "Metric Value"]
    rows = [head]
    for k, (name, rd, wr) in enumerate(launches):
        for metric, (unit, v) in (("dram__bytes_read.sum", rd), ("dram__bytes_write.sum", wr)):
            rows.append([str(k), "1", "python", "h", name + "(chess_item_args)", "1", "7", "(128, 1, 1)", "(296, 1, 1)",
                         "0", "10.0", "Command line profiler metrics", metric, unit, v])
    return "\n".join(",".join('"%s"' % c for c in r) for r in rows)


def test_traffic_is_summed_over_the_launches_of_a_pass():
    big, small = ("Gbyte", "3.50"), ("Gbyte", "1.90")
    wr = ("Mbyte", "2.00")
    name = "void <unnamed>::compare_r1_tri_kernel<7, 1, 0, 2>"
    retry = "void <unnamed>::compare_win_other<1>"
    cold = [(name, ("Gbyte", "4.00"), wr), (retry, ("Kbyte", "1.00"), wr), (name, ("Gbyte", "2.50"), wr)]
    warm = [(name, big, wr), (retry, ("Kbyte", "1.00"), wr), (name, small, wr)]
    got, note = bench.parse_ncu_traffic(_ncu_csv(cold + warm + warm), probe_steps=3)
    assert got == int(3.5e9 + 1.9e9 + 2 * 2e6) and "one pass = 2 launch" in note
    # one launch per pass (a chromosome's worth of pairs)
    got, note = bench.parse_ncu_traffic(_ncu_csv([cold[0], warm[0], warm[0]]), probe_steps=3)
    assert got == int(3.5e9 + 2e6) and "one pass = 1 launch" in note
    # launches that do not divide into passes: the mean per launch, and the note says so
    got, note = bench.parse_ncu_traffic(_ncu_csv([cold[0], warm[0], warm[2], warm[0]]), probe_steps=3)
    assert got == int(np.mean([3.5e9 + 2e6, 1.9e9 + 2e6, 3.5e9 + 2e6])) and "unknown" in note
    assert bench.parse_ncu_traffic("", 1, "no device")[0] is None
