"""Pairwise comparison of Hi-C submatrices: SSIM + signal-to-noise.

Mirror of chess/sim.py (reference 0.3.8): same public names, signatures,
return types and NaN / exception behaviour.  The arithmetic (gather, masking,
SSIM, SN) runs in the CUDA kernels behind ``chess_compare_batch``; the
functions below translate the reference's call shapes into batches.
"""
import logging
import uuid

import numpy as np
import torch

from . import _native
from ._native import lib, check
from .engine import (CompareEngine, compare_dense_pairs, get_context, visible_devices, _ptr,
                     _stream_ptr)
from .helpers import sub_matrix_from_edges_dict  # noqa: F401  (re-exported like the reference)

logger = logging.getLogger(__name__)


def SN(M1, M2, size=7):
    """Average local signal-to-noise ratio of M1 - M2; chess/sim.py:15-30.

    mean over all pixels of |mean(window)| / var(window) with a 7x7 window
    that is NaN-padded at the border; NaN ratios (0/0) are skipped.
    """
    if size != 7:
        raise NotImplementedError("chess_b200 computes SN with the reference's fixed 7x7 window")
    M1 = np.ascontiguousarray(M1, dtype=np.float64)
    M2 = np.ascontiguousarray(M2, dtype=np.float64)
    if M1.shape != M2.shape or M1.ndim != 2 or M1.shape[0] != M1.shape[1]:
        raise ValueError("SN expects two square matrices of the same shape")
    device = visible_devices()[0]
    ctx = get_context(device)
    dev = torch.device("cuda", device)
    with torch.cuda.device(dev):
        d1 = torch.from_numpy(M1).to(dev)
        d2 = torch.from_numpy(M2).to(dev)
        out = torch.empty(1, dtype=torch.float64, device=dev)
        check(lib.chess_sn_dense(ctx.handle, _ptr(d1), _ptr(d2), M1.shape[0], _ptr(out), _stream_ptr(dev)),
              ctx.handle, "chess_sn_dense")
        return float(out.cpu()[0])


def find_masked_rows(m, masking_value=0):
    """Indices of columns whose sum equals n * masking_value; chess/sim.py:74-87."""
    return np.where(np.sum(m, 0) == m.shape[0] * masking_value)[0]


def remove_rows(m, target_rows):
    """Delete the given rows and columns; chess/sim.py:90-102."""
    return np.delete(np.delete(m, target_rows, 0), target_rows, 1)


def _result(pair_ix, ssim, sn, status, compute_sn):
    if status != _native.PAIR_OK:
        return pair_ix, np.nan, np.nan
    return pair_ix, np.float64(ssim), (float(sn) if compute_sn else None)


def compare_matrix_pair(reference, reference_region, query, query_region, pair_ix=None,
                        min_bins=20, keep_unmappable_bins=False, absolute_window_size=None,
                        relative_window_size=1., mappability_cutoff=0.1, default_value=1,
                        compute_sn=True):
    """(pair_ix, ssim, sn) for two dense matrices; chess/sim.py:306-380."""
    reference = np.asarray(reference, dtype=np.float64)
    query = np.asarray(query, dtype=np.float64)
    if query.shape[0] < min_bins or reference.shape[0] < min_bins:
        return pair_ix, np.nan, np.nan
    if mappability_cutoff <= 0 and not keep_unmappable_bins:
        raise TypeError("union1d() missing 2 required positional arguments: 'ar1' and 'ar2'")
    flip = reference_region.strand != query_region.strand
    ssim, sn, status = compare_dense_pairs(
        [reference], [query], [flip], min_bins=min_bins, keep_unmappable_bins=keep_unmappable_bins,
        absolute_window_size=absolute_window_size, relative_window_size=relative_window_size,
        mappability_cutoff=mappability_cutoff, default_value=default_value, compute_sn=compute_sn)
    if status[0] == _native.PAIR_WINDOW_EXCEEDS:
        raise ValueError("win_size exceeds image extent.")
    return _result(pair_ix, ssim[0], sn[0], status[0], compute_sn)


def _engine(reference_edges, reference_regions, query_edges, query_regions):
    return CompareEngine(reference_edges, reference_regions, query_edges, query_regions)


def compare_pair(pair, reference_edges, reference_regions, query_edges, query_regions,
                 min_bins=20, keep_unmappable_bins=False, absolute_window_size=None,
                 relative_window_size=1., mappability_cutoff=0.1, default_value=1,
                 compute_sn=True):
    """Gather and compare one (ID, reference_region, query_region); chess/sim.py:225-303."""
    ids, ssim, sn, status = _engine(reference_edges, reference_regions, query_edges, query_regions).compare(
        [pair], compute_sn=compute_sn, min_bins=min_bins, keep_unmappable_bins=keep_unmappable_bins,
        absolute_window_size=absolute_window_size, relative_window_size=relative_window_size,
        mappability_cutoff=mappability_cutoff, default_value=default_value)
    return _result(ids[0], ssim[0], sn[0], status[0], compute_sn)


def compare_structures(reference_edges, reference_regions, query_edges, query_regions, pairs,
                       worker_ID, min_bins=20, keep_unmappable_bins=False,
                       absolute_window_size=None, relative_window_size=1.,
                       mappability_cutoff=0.1, default_value=1):
    """{pair_id: (ssim, sn)} without background model; chess/sim.py:383-459."""
    ids, ssim, sn, status = _engine(reference_edges, reference_regions, query_edges, query_regions).compare(
        pairs, compute_sn=True, min_bins=min_bins, keep_unmappable_bins=keep_unmappable_bins,
        absolute_window_size=absolute_window_size, relative_window_size=relative_window_size,
        mappability_cutoff=mappability_cutoff, default_value=default_value)
    results = {}
    for k, pid in enumerate(ids):
        if status[k] != _native.PAIR_OK or np.isnan(ssim[k]):
            results[pid] = (np.nan, np.nan)
        else:
            results[pid] = (np.float64(ssim[k]), float(sn[k]))
    logger.debug('[WORKER #{}]: Done!'.format(worker_ID))
    return results


def chunk_comparison_worker(input_queue, output_queue, reference_edges, reference_regions,
                            query_edges, query_regions, min_bins=20, keep_unmappable_bins=False,
                            absolute_window_size=None, relative_window_size=1.,
                            mappability_cutoff=0.1, default_value=1):
    """Queue-driven worker with the reference's protocol; chess/sim.py:147-222.

    Each item is ``[chunk_of_pairs, compute_sn]`` (``None`` stops the worker);
    the reply is a list of ``[pair_ix, ssim, sn]``, or the exception object if
    anything other than a per-pair soft failure happened.  One chunk is one
    batched kernel launch.
    """
    worker_id = uuid.uuid4()
    try:
        engine = _engine(reference_edges, reference_regions, query_edges, query_regions)
        while True:
            data = input_queue.get(block=True)
            if data is None:
                break
            chunk, compute_sn = data
            ids, ssim, sn, status = engine.compare(
                chunk, compute_sn=compute_sn, min_bins=min_bins,
                keep_unmappable_bins=keep_unmappable_bins, absolute_window_size=absolute_window_size,
                relative_window_size=relative_window_size, mappability_cutoff=mappability_cutoff,
                default_value=default_value)
            output_queue.put([list(_result(ids[k], ssim[k], sn[k], status[k], compute_sn))
                              for k in range(len(ids))])
    except Exception as e:
        logger.error("Worker {} encountered a problem: {}".format(worker_id, e))
        output_queue.put(e)
    logger.debug("Worker {} exited gracefully".format(worker_id))


def comparison_worker(input_queue, output_queue, reference_edges, reference_regions, query_edges,
                      query_regions, min_bins=20, keep_unmappable_bins=False,
                      absolute_window_size=None, relative_window_size=1., mappability_cutoff=0.1,
                      default_value=1):
    """Per-pair variant of the worker; chess/sim.py:105-144."""
    worker_id = uuid.uuid4()
    try:
        engine = _engine(reference_edges, reference_regions, query_edges, query_regions)
        while True:
            data = input_queue.get(block=True)
            if data is None:
                break
            pair, compute_sn = data
            ids, ssim, sn, status = engine.compare(
                [pair], compute_sn=compute_sn, min_bins=min_bins,
                keep_unmappable_bins=keep_unmappable_bins, absolute_window_size=absolute_window_size,
                relative_window_size=relative_window_size, mappability_cutoff=mappability_cutoff,
                default_value=default_value)
            output_queue.put(_result(ids[0], ssim[0], sn[0], status[0], compute_sn))
    except Exception as e:
        logger.error("Worker {} encountered a problem: {}".format(worker_id, e))
        output_queue.put(e)
    logger.info("Worker {} exited gracefully".format(worker_id))
