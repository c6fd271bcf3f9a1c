"""Build chess_b200/csrc/libchess_b200.so with nvcc for sm_100a (in-tree).

    python -m chess_b200.build [--force]

nvcc cross-compiles without a GPU.  The shared library links only the CUDA
runtime (static) -- no torch, no Python headers: it is the C-ABI boundary
declared in include/chess_b200.h.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libchess_b200.so")
SOURCES = ["api.cu", "oe.cu", "compare_generic.cu", "compare_fast.cu", "background.cu", "ingest.cu", "extract.cu",
           "items_h0.cu", "items_tri.cu", "items_blk.cu", "items_blkw.cu", "items_tri7.cu", "items_triw.cu", "items_h1.cu", "items_h2.cu", "items_h3.cu"]
HEADERS = ["common.cuh", "compare_kernels.cuh", "compare_tri.cuh", "tri_rows.cuh", "compare_blk.cuh", "compare_blk7.cuh", "compare_blkw.cuh", "compare_tri7.cuh", "compare_triw.cuh", "parse_number.cuh", "pow5_table.inc", os.path.join("..", "..", "include", "chess_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]
LINK_FLAGS = ["--shared", "-cudart", "static"]


def _stamp():
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS + LINK_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp_file = LIB + ".stamp"
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file):
        if open(stamp_file).read().strip() == stamp:
            return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    obj_dir = os.path.join(CSRC, "build")
    os.makedirs(obj_dir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, proc

    # one nvcc per translation unit, in parallel (the kernel instantiations dominate the build time)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    log = "".join(r[2].stdout + r[2].stderr for r in results)
    if any(r[2].returncode != 0 for r in results):
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libchess_b200.so")
    proc = subprocess.run([nvcc] + NVCC_FLAGS + LINK_FLAGS + [r[1] for r in results] + ["-o", LIB],
                          capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(log + proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed linking libchess_b200.so")
    if verbose:
        sys.stderr.write(log + proc.stdout + proc.stderr)
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
