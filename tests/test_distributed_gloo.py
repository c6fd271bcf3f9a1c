"""World-size-2 gloo tests (CPU): pair sharding and the host-side result gather."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from chess_b200.distributed import shard_bounds, my_shard


def test_shard_bounds_cover_everything_in_order():
    for total in (0, 1, 7, 2402, 100000):
        for world in (1, 2, 3, 4, 8):
            b = shard_bounds(total, world)
            assert b[0] == 0 and b[-1] == total and len(b) == world + 1
            sizes = np.diff(b)
            assert sizes.min() >= 0 and sizes.max() - sizes.min() <= 1
            assert [my_shard(total, r, world) for r in range(world)] == list(zip(b[:-1], b[1:]))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_path):
    import torch.distributed as dist
    from chess_b200.distributed import gather_to_rank0, max_over_ranks, my_shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = my_shard(total, rank, world)
    idx = np.arange(lo, hi)
    ssim = np.sin(idx * 0.37)          # stands in for this rank's kernel output
    ssim[idx % 17 == 0] = np.nan       # soft failures stay NaN through the gather
    sn = idx * 0.5
    status = (idx % 17 == 0).astype(np.int32) * 3
    got = gather_to_rank0((ssim, sn, status), total, rank, world, dist)
    # the same through an explicit (second) gloo group, as bench.py does beside an NCCL default group
    again = gather_to_rank0((ssim, sn, status), total, rank, world, dist, group=dist.new_group(backend="gloo"))
    # ... and the tensor form (what the ranks use when the results are still on their GPUs)
    import torch
    from chess_b200.distributed import gather_tensors_to_rank0
    third = gather_tensors_to_rank0(tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in (ssim, sn, status)),
                                    total, rank, world, dist)
    slowest = max_over_ranks(1.0 + rank, world, dist)
    if rank == 0:
        assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, again))
        assert all(np.array_equal(a, b, equal_nan=True) and a.dtype == b.dtype for a, b in zip(got, third))
        from chess_b200.distributed import run_zscores
        np.savez(out_path, ssim=got[0], sn=got[1], status=got[2], slowest=slowest, z=run_zscores(got[0]))
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_odd_total_three_ranks(tmp_path):
    """Shards of unequal length (padding of the packed buffer) and a rank with an empty shard."""
    for total, world in ((7, 3), (2, 3)):
        out = str(tmp_path / ("g%d.npz" % total))
        mp.spawn(_worker, args=(world, _free_port(), total, out), nprocs=world, join=True)
        got = np.load(out)
        idx = np.arange(total)
        want = np.sin(idx * 0.37)
        want[idx % 17 == 0] = np.nan
        assert np.array_equal(got["ssim"], want, equal_nan=True) and np.array_equal(got["sn"], idx * 0.5)


def test_two_rank_gather_matches_single_process(tmp_path):
    total, world = 2402, 2
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), total, out), nprocs=world, join=True)
    got = np.load(out)
    idx = np.arange(total)
    want = np.sin(idx * 0.37)
    want[idx % 17 == 0] = np.nan
    assert np.array_equal(got["ssim"], want, equal_nan=True)     # identical bytes, identical order
    assert np.array_equal(got["sn"], idx * 0.5)
    assert np.array_equal(got["status"], (idx % 17 == 0).astype(np.int32) * 3)
    assert float(got["slowest"]) == 2.0
    ok = ~np.isnan(want)                                          # bin/chess:228-233 on the gathered values
    zw = np.full(total, np.nan)
    zw[ok] = (want[ok] - np.nanmean(want)) / np.nanstd(want)
    assert np.allclose(got["z"], zw, rtol=1e-12, atol=0, equal_nan=True)
