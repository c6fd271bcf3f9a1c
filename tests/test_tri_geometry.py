"""The lane geometry of the triangle and block kernels, checked on the CPU: `tri_mode`, `tri_lane` and `blk_lane`
(chess_b200/csrc/tri_rows.cuh) compile for the host too, and tests/tri_geometry_harness.cu walks every layout the
launchers and the compaction can produce -- strips of 3 .. 8 columns, compacted matrices of 1 .. 248 bins in one,
two or three row bands, 2 x 2 .. 7 x 7 blocks of 8 .. 29 strips up to 1024 bins -- verifying that the lanes tile the symmetric matrix exactly
(every cell and its mirror image weigh 2 together, the diagonal 1), that the strips of a band are contiguous and
that the lanes beside a band are idle (lane 31 always), which is what the neighbour exchange relies on."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_lanes_tile_the_triangle_exactly(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "tri_geometry")
    build = subprocess.run([nvcc, "-O1", "-std=c++17", "-o", exe, os.path.join(HERE, "tri_geometry_harness.cu")],
                           capture_output=True, text=True, timeout=600)
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:]
    last = run.stdout.strip().splitlines()[-1]
    assert last.endswith(" 0 failures") and int(last.split()[1]) > 10000, last
