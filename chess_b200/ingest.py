"""Sparse-matrix text -> COO arrays on the GPU.

Replaces the per-line Python loop of ``edges_from_sparse_matrix``
(chess/helpers.py:309-339) for index-addressed files ("source<TAB>sink<TAB>weight").
The file bytes go to HBM once; line splitting, ``int()`` / ``float()`` (bit-identical
to Python, see csrc/parse_number.cuh) and the stable compaction run there.  Lines the
device cannot decide are re-parsed here with Python itself, so the result never
differs from the reference's.
"""
from __future__ import absolute_import

import ctypes as C
import gzip
import logging

import numpy as np

logger = logging.getLogger(__name__)

# files whose share of host-parsed lines exceeds this are left to the host parser altogether
MAX_FALLBACK_FRACTION = 0.01


def read_bytes(file_name):
    with open(file_name, "rb") as fh:
        magic = fh.read(2)
    if magic == b"\x1f\x8b":
        with gzip.open(file_name, "rb") as fh:
            return np.frombuffer(bytearray(fh.read()), dtype=np.uint8)
    return np.fromfile(file_name, dtype=np.uint8)


_STAGE = {}


def read_to_device(file_name, device):
    """The bytes of a (plain) file as a uint8 tensor on ``device``: read in 32 MB pieces into two pinned staging
    buffers, each piece copied to the GPU while the next one is being read."""
    import os
    import torch
    dev = torch.device("cuda", int(device))
    size = os.path.getsize(file_name)
    text = torch.empty(size, dtype=torch.uint8, device=dev)
    chunk = 32 << 20
    key = int(device)
    if key not in _STAGE:
        _STAGE[key] = ([torch.empty(chunk, dtype=torch.uint8).pin_memory() for _ in range(2)],
                       [torch.cuda.Event() for _ in range(2)])
    bufs, events = _STAGE[key]
    views = [b.numpy() for b in bufs]
    with torch.cuda.device(dev), open(file_name, "rb", buffering=0) as fh:
        off, k = 0, 0
        while off < size:
            slot = k & 1
            if k >= 2:
                events[slot].synchronize()        # the copy that last used this buffer is done
            got = fh.readinto(memoryview(views[slot])[:min(chunk, size - off)])
            if not got:
                break
            text[off:off + got].copy_(bufs[slot][:got], non_blocking=True)
            events[slot].record()
            off += got
            k += 1
        torch.cuda.current_stream(dev).synchronize()
    return text[:off] if off != size else text


def _host_line(raw):
    """One line exactly as the reference parses it (helpers.py:313-335) -> (kind, src, sink, weight)."""
    line = raw.decode("utf-8")
    if line.startswith("#"):
        return 1, 0, 0, 0.0
    line = line.rstrip()
    if line == "":
        return 1, 0, 0, 0.0          # blank lines are skipped (as the array parser always did)
    fields = line.split("\t")
    source, sink, weight = int(fields[0]), int(fields[1]), float(fields[2])
    if not np.isfinite(weight):
        return 2, 0, 0, 0.0
    if source > sink:
        source, sink = sink, source
    if source < 0 or sink >= 2 ** 31:
        raise ValueError("region index out of range: {!r}".format(line))
    return 0, source, sink, weight


def parse_sparse_text(data, device=None, return_device=False):
    """(src, sink, weight, n_data_lines, n_nonfinite) from the bytes of a sparse-matrix file.

    ``data``: file name, bytes or uint8 array.  Returns None when the text needs the host
    parser (lone carriage returns, or too many lines the device cannot decide).
    src / sink are int64 and weight float64 numpy arrays in file order (canonical
    src <= sink, non-finite weights dropped); with ``return_device`` the int32 / float64
    device tensors are returned instead.
    """
    import torch
    from . import _native
    from ._native import lib, check
    from .engine import get_context, visible_devices

    device = visible_devices()[0] if device is None else int(device)
    text = None
    if isinstance(data, str):
        with open(data, "rb") as fh:
            gz = fh.read(2) == b"\x1f\x8b"
        if gz:
            data = read_bytes(data)
        else:
            text = read_to_device(data, device)       # straight to the GPU; lines come back only if needed
            data = None
    elif isinstance(data, (bytes, bytearray)):
        data = np.frombuffer(bytearray(data), dtype=np.uint8)
    ctx = get_context(device)
    dev = torch.device("cuda", device)
    n_bytes = int(data.size) if data is not None else int(text.numel())
    empty = (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float64), 0, 0)
    if n_bytes == 0:
        return empty
    ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    with torch.cuda.device(dev):
        sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if text is None:
            text = torch.from_numpy(np.ascontiguousarray(data)).to(dev)
        n_lines, n_cr = C.c_int64(0), C.c_int64(0)
        check(lib.chess_sparse_text_lines(ctx.handle, ptr(text), n_bytes, C.byref(n_lines), C.byref(n_cr), sp),
              ctx.handle, "chess_sparse_text_lines")
        n_lines = n_lines.value
        if n_cr.value:
            logger.debug("sparse text has carriage returns without newline: host parser")
            return None
        if n_lines == 0:
            return empty
        src = torch.empty(n_lines, dtype=torch.int32, device=dev)
        sink = torch.empty(n_lines, dtype=torch.int32, device=dev)
        weight = torch.empty(n_lines, dtype=torch.float64, device=dev)
        kind = torch.empty(n_lines, dtype=torch.uint8, device=dev)
        line_start = torch.empty(n_lines + 1, dtype=torch.int64, device=dev)
        fb_cap = max(1024, int(n_lines * MAX_FALLBACK_FRACTION) + 1)
        fb_lines = torch.empty(fb_cap, dtype=torch.int64, device=dev)
        n_fb = C.c_int64(0)
        check(lib.chess_sparse_text_parse(ctx.handle, ptr(text), n_bytes, n_lines, ptr(src), ptr(sink), ptr(weight),
                                          ptr(kind), ptr(line_start), ptr(fb_lines), fb_cap, C.byref(n_fb), sp),
              ctx.handle, "chess_sparse_text_parse")
        n_fb = n_fb.value
        if n_fb > fb_cap or n_fb > max(1024, n_lines * MAX_FALLBACK_FRACTION):
            logger.debug("%d of %d lines need the host: host parser", n_fb, n_lines)
            return None
        if n_fb:
            # lines outside the device grammar: Python decides (and raises what the reference raises)
            lines = torch.sort(fb_lines[:n_fb]).values
            begin = line_start[lines].cpu().numpy()
            end = (line_start[lines + 1] - 1).cpu().numpy()
            if data is not None:
                patch = [_host_line(data[b:e].tobytes()) for b, e in zip(begin, end)]
            else:
                patch = [_host_line(text[int(b):int(e)].cpu().numpy().tobytes()) for b, e in zip(begin, end)]
            kind[lines] = torch.tensor([p[0] for p in patch], dtype=torch.uint8, device=dev)
            src[lines] = torch.tensor([p[1] for p in patch], dtype=torch.int32, device=dev)
            sink[lines] = torch.tensor([p[2] for p in patch], dtype=torch.int32, device=dev)
            weight[lines] = torch.tensor([p[3] for p in patch], dtype=torch.float64, device=dev)
        counts = (C.c_int64 * 3)()
        check(lib.chess_sparse_text_compact(ctx.handle, n_lines, ptr(src), ptr(sink), ptr(weight), ptr(kind),
                                            None, None, None, 0, counts, sp), ctx.handle,
              "chess_sparse_text_compact")
        kept = int(counts[0])
        o_src = torch.empty(kept, dtype=torch.int32, device=dev)
        o_sink = torch.empty(kept, dtype=torch.int32, device=dev)
        o_w = torch.empty(kept, dtype=torch.float64, device=dev)
        if kept:
            check(lib.chess_sparse_text_compact(ctx.handle, n_lines, ptr(src), ptr(sink), ptr(weight), ptr(kind),
                                                ptr(o_src), ptr(o_sink), ptr(o_w), kept, counts, sp), ctx.handle,
                  "chess_sparse_text_compact")
        n_data, n_bad = int(counts[1]), int(counts[2])
        if return_device:
            return o_src, o_sink, o_w, n_data, n_bad
        return (o_src.cpu().numpy().astype(np.int64), o_sink.cpu().numpy().astype(np.int64), o_w.cpu().numpy(),
                n_data, n_bad)
