"""First stage of `chess extract` on the GPU (chess/get_structures.py:80-121).

`log_ratio_stage` is the loop head of `extract_structures`: gather both submatrices with default
weight 0, OR = log(query / reference) with NaN / inf -> 0, the 0.5-std thresholds.  The denoising,
Otsu, morphology and labelling stages that follow are scikit-image / scipy code operating on the
small per-pair images this stage returns; they are not part of this package.
"""
from __future__ import absolute_import

import ctypes as C

import numpy as np


def log_ratio_stage(reference_edges, reference_regions, query_edges, query_regions, pairs, device=None):
    """For every (ID, reference_region, query_region): (positive, negative, std, mean_positive,
    mean_negative), or None where the reference would fail (region not found, unequal sizes).

    `positive` / `negative` are the n x n arrays of get_structures.py:112-113; the means are what the
    reference passes to denoise_bilateral as sigma_color (get_structures.py:117, 124).
    """
    import torch
    from . import _native
    from ._native import lib, check
    from .engine import (as_device_contacts, build_pair_table, get_context, index_bins, region_spans,
                         visible_devices)

    pairs = list(pairs)
    if not pairs:
        return []
    device = visible_devices()[0] if device is None else int(device)
    ctx = get_context(device)
    dev = torch.device("cuda", device)
    ref = as_device_contacts(reference_edges, index_bins(reference_regions))
    qry = as_device_contacts(query_edges, index_bins(query_regions))
    rf, rn = region_spans(reference_regions, [p[1] for p in pairs])
    qf, qn = region_spans(query_regions, [p[2] for p in pairs])
    ref.check_span(rf, rn, "reference region")
    qry.check_span(qf, qn, "query region")
    max_n = int(max(rn.max(), qn.max(), 1))
    ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    out = []
    with torch.cuda.device(dev):
        sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rw, rband, _ = ref.band_on(device, max_n, 0.0)       # default_weight=0. (get_structures.py:99-108)
        qw, qband, _ = qry.band_on(device, max_n, 0.0)
        table = build_pair_table(rf, rn, rw, qf, qn, qw, np.zeros(len(pairs), dtype=np.int32))
        chunk = max(1, (256 << 20) // (16 * max_n * max_n))  # ~256 MB of tiles per launch
        for lo in range(0, len(pairs), chunk):
            hi = min(len(pairs), lo + chunk)
            d_tab = torch.from_numpy(table[lo:hi].view(np.uint8)).to(dev)
            pos = torch.empty((hi - lo, max_n * max_n), dtype=torch.float64, device=dev)
            neg = torch.empty_like(pos)
            stats = torch.empty((hi - lo, 4), dtype=torch.float64, device=dev)
            status = torch.empty(hi - lo, dtype=torch.int32, device=dev)
            check(lib.chess_log_ratio_batch(ctx.handle, ptr(rband), ptr(qband), ptr(d_tab), hi - lo, max_n,
                                            ptr(pos), ptr(neg), ptr(stats), ptr(status), sp),
                  ctx.handle, "chess_log_ratio_batch")
            h_pos, h_neg = pos.cpu().numpy(), neg.cpu().numpy()
            h_stats, h_status = stats.cpu().numpy(), status.cpu().numpy()
            for k in range(hi - lo):
                if h_status[k] != 0:
                    out.append(None)
                    continue
                n = int(h_stats[k, 3])
                out.append((h_pos[k, :n * n].reshape(n, n).copy(), h_neg[k, :n * n].reshape(n, n).copy(),
                            float(h_stats[k, 0]), float(h_stats[k, 1]), float(h_stats[k, 2])))
    return out
