"""Sharding of region pairs over GPUs and the host-side gather of results.

The comparison path shards naturally: every pair is independent
(chess/sim.py:306-380), so there is no data-path collective.  The reference
splits the pair list into ceil(P/p) chunks, one per worker
(chess/helpers.py:530-534, bin/chess:212-214) and collects (pair_ix, ssim, sn)
through a queue; here contiguous ranges go to ranks / devices and the results
(24 bytes per pair) are concatenated in pair order on rank 0.

Two ways to use several GPUs:
  * one process, several devices: ``CompareEngine(devices=[...])`` uses
    :func:`shard_bounds` and copies results back itself;
  * one process per GPU under ``torchrun``: each rank calls
    :func:`my_shard`, compares its slice, then :func:`gather_to_rank0`.
Works with the ``gloo`` backend on CPU tensors (tests) and ``nccl``/``gloo`` on
a GPU box; the payload is tiny, so the backend is irrelevant for speed.
"""
import numpy as np


def shard_bounds(total, world):
    """world+1 monotone offsets: shard k is [b[k], b[k+1]); sizes differ by at most 1."""
    world = max(int(world), 1)
    base, extra = divmod(int(total), world)
    sizes = [base + (1 if k < extra else 0) for k in range(world)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def my_shard(total, rank, world):
    b = shard_bounds(total, world)
    return int(b[rank]), int(b[rank + 1])


def gather_to_rank0(local_arrays, total, rank, world, dist=None, group=None, device=None):
    """Concatenate per-rank result arrays in rank (= pair) order on rank 0.

    ``local_arrays`` is a tuple of 1-D numpy arrays covering this rank's shard
    (``shard_bounds(total, world)``).  All arrays travel in ONE collective: each
    rank packs its arrays back to back into a byte buffer padded to the longest
    shard (20 bytes per pair for ssim, SN and status), rank 0 unpacks.  ``group``
    selects the process group (a gloo group for host buffers when the default
    group is NCCL); with ``device`` the buffer is staged through that device
    instead (NCCL groups).  Returns the full arrays on rank 0, ``None`` elsewhere.
    """
    if world == 1 or dist is None:
        return tuple(np.asarray(a) for a in local_arrays)
    import torch
    bounds = shard_bounds(total, world)
    longest = int(np.max(np.diff(bounds)))
    arrays = [np.ascontiguousarray(a) for a in local_arrays]
    sizes = [a.dtype.itemsize for a in arrays]
    buf = np.zeros(longest * sum(sizes), dtype=np.uint8)
    off = 0
    for a, size in zip(arrays, sizes):
        buf[off:off + a.nbytes] = a.view(np.uint8)
        off += longest * size
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    bucket = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, gather_list=bucket, dst=0, group=group)
    if rank != 0:
        return None
    out = [np.empty(total, dtype=a.dtype) for a in arrays]
    for k in range(world):
        part = bucket[k].cpu().numpy() if device is not None else bucket[k].numpy()
        n = int(bounds[k + 1] - bounds[k])
        off = 0
        for full, size in zip(out, sizes):
            full[bounds[k]:bounds[k + 1]] = part[off:off + n * size].view(full.dtype)
            off += longest * size
    return tuple(out)


_GATHER_BUFFERS = {}


def _gather_buffers(device, world, nbytes, rank):
    """Send buffer, rank 0's receive block and its page-locked landing area, kept between calls (a step of the
    sharded run is ~2 ms: allocating and zeroing these each time was a tenth of it)."""
    import torch
    key = (str(device), world, nbytes, rank)
    got = _GATHER_BUFFERS.get(key)
    if got is None:
        if len(_GATHER_BUFFERS) > 8:
            _GATHER_BUFFERS.clear()
        buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        block = torch.empty((world, nbytes), dtype=torch.uint8, device=device) if rank == 0 else None
        host = None
        if rank == 0:
            host = torch.empty((world, nbytes), dtype=torch.uint8)
            if device.type == "cuda":
                host = host.pin_memory()
        got = _GATHER_BUFFERS[key] = (buf, block, host)
    return got


def gather_tensors_to_rank0(local_tensors, total, rank, world, dist=None, group=None):
    """The same gather for results that are still on the GPU: the per-rank tensors (this rank's shard, e.g. ssim,
    SN, status on its device) travel as one padded byte buffer in one collective over the default group (NCCL:
    device to device over NVLink), and rank 0 alone copies the gathered block to the host -- one D2H into
    page-locked memory instead of one per rank plus a host-side collective.  Works unchanged on CPU tensors over
    gloo.  Returns numpy arrays in pair order on rank 0, ``None`` elsewhere."""
    import torch
    if world == 1 or dist is None:
        return tuple(t.cpu().numpy() for t in local_tensors)
    bounds = shard_bounds(total, world)
    longest = int(np.max(np.diff(bounds)))
    sizes = [t.element_size() for t in local_tensors]
    ref = local_tensors[0]
    buf, block, host = _gather_buffers(ref.device, world, longest * sum(sizes), rank)
    off = 0
    for t, size in zip(local_tensors, sizes):
        n = t.numel() * size
        if n:                      # a rank may hold no pair at all
            buf[off:off + n] = t.contiguous().view(torch.uint8).reshape(-1)
        off += longest * size
    dist.gather(buf, gather_list=list(block.unbind(0)) if rank == 0 else None, dst=0, group=group)
    if rank != 0:
        return None
    host.copy_(block, non_blocking=True)              # one copy for everything
    if ref.device.type == "cuda":
        torch.cuda.current_stream(ref.device).synchronize()
    view = host.numpy()
    dtypes = [torch.empty(0, dtype=t.dtype).numpy().dtype for t in local_tensors]
    out = [np.empty(total, dtype=dt) for dt in dtypes]
    for k in range(world):
        n = int(bounds[k + 1] - bounds[k])
        off = 0
        for full, size in zip(out, sizes):
            full[bounds[k]:bounds[k + 1]] = view[k, off:off + n * size].view(full.dtype)
            off += longest * size
    return tuple(out)


def run_zscores(ssim):
    """z_ssim over the whole run from the gathered ssim values (bin/chess:228-233): nanmean / nanstd
    (ddof 0) over the non-NaN entries, NaN kept.  Rank 0 calls this after `gather_to_rank0`."""
    ssim = np.asarray(ssim, dtype=np.float64)
    valid = ssim[~np.isnan(ssim)]
    if valid.size == 0:
        return np.full(len(ssim), np.nan)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (ssim - valid.mean()) / valid.std()         # a NaN ssim stays NaN


def max_over_ranks(value, world, dist=None, device=None):
    """Max of a Python float over ranks (timings are always max-over-ranks)."""
    if world == 1 or dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
