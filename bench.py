#!/usr/bin/env python
"""Benchmark of the `chess sim` comparison hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over the whole pair list of the workload:
for every region pair gather both submatrices from the resident O/E band,
mask / compact, SSIM, SN, then the run-level z-score.  Workload = BASELINE.json
configs[1]: human chr2 at 25 kb (9 688 bins), 2 Mb windows sliding by 100 kb
(2 402 pairs, 81 bins each with the reference's BED lookup), default
`chess sim` parameters (-r 1, SN on); synthetic, seeded.

Numbers on the JSON line
  value      pairs/s, whole job, samples and pair table already resident in HBM,
             CUDA-event time of K steps (L2 flushed before every step).
  e2e        same work through the host-buffer C-ABI call
             (chess_compare_band_batch_host): pair table from host memory, results
             back to host memory and z-scores, every step.
  roofline   compare kernel only: algorithmic bytes (2*n*n*8 + 24 per pair,
             SURVEY 8d) / its CUDA-event duration, against the measured HBM
             copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle (reference algorithm restated, oracle/) with a
             multiprocessing pool on this box's cores, on a bounded sample.

  synthetic_one_vs_many  (N = 1 only) BASELINE configs[0]: one dense reference against truth + 1000 decoys
             (utils/run_synthetic_test.py), SSIM only, per window of the script; device-resident and e2e.

  genome_wide  (N = 1 only) the same measurements for BASELINE configs[2] -- hg38 at 10 kb, ~100k pairs
             of 2 Mb (201 bins), the workload the north-star target is quoted on -- 5 timed steps.

Under torchrun (N > 1) every rank owns one GPU and an independent chr2-sized
sample pair (weak scaling, no data-path collective); rank 0 prints the line
with the max-over-ranks time.
"""
import argparse
import ctypes as C
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHR2_BP = 242193529
BIN = 25000
WINDOW_BP = 2000000
STEP_BP = 100000
METRIC = "region pairs/sec (SSIM+SN+z)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--window", default="r1", choices=["r1", "a3", "a5", "a7", "a9", "a11"],
                    help="r1 = default `chess sim` (-r 1); aN = -a N")
    ap.add_argument("--workload", default="chr2_25kb", choices=["chr2_25kb", "hg38_10kb"],
                    help="chr2_25kb = BASELINE configs[1] (the contract's bench line); hg38_10kb = configs[2], "
                         "genome-wide 10 kb, ~100k pairs of 2 Mb, pairs sharded over the ranks (strong scaling)")
    ap.add_argument("--bin-size", type=int, default=BIN, help="chr2 workload only: bin size in bp")
    ap.add_argument("--window-bp", type=int, default=WINDOW_BP, help="chr2 workload only: region size in bp")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-genome-wide", action="store_true",
                    help="skip the extra configs[2] (hg38 @ 10 kb, ~100k pairs) measurement at N=1")
    ap.add_argument("--no-ingest", action="store_true",
                    help="skip the extra sparse-text ingest measurement (8f-1) at N=1")
    ap.add_argument("--no-synthetic", action="store_true",
                    help="skip the extra configs[0] (dense one-vs-many synthetic comparison) measurement at N=1")
    ap.add_argument("--no-background", action="store_true",
                    help="skip the extra --background-query measurement (a-14) at N=1")
    ap.add_argument("--kernel", type=int, default=0,
                    help="0 auto, 1 force the generic kernel, 2 the CTA-per-pair tuned kernel")
    return ap.parse_args()


def window_params(tag):
    return dict(absolute_window_size=int(tag[1:])) if tag.startswith("a") else dict()


def make_workload(rank, workload="chr2_25kb", world=1, bin_size=BIN, window_bp=WINDOW_BP):
    """Synthetic sample pair and the sliding pair list.

    chr2_25kb: every rank gets its own chr2-sized samples (seeded by rank) and the full pair list.
    hg38_10kb: all ranks build the same genome; rank r keeps a contiguous shard of the pair list."""
    from chess_b200 import synth
    from chess_b200.distributed import my_shard
    if workload == "hg38_10kb":
        genome = synth.Genome(synth.chromosome_bins(synth.HG38_CHROM_BP, 10000), 10000, synth.HG38_CHROM_BP)
        ref = synth.synthetic_sample(genome, 215, seed=100, structure_seed=7)
        qry = synth.synthetic_sample(genome, 215, seed=101, structure_seed=7, perturb=0.1)
        pairs = synth.sliding_pairs(genome, WINDOW_BP, 30000)
        lo, hi = my_shard(len(pairs), rank, world)
        return genome, ref, qry, pairs[lo:hi]
    genome = synth.Genome(synth.chromosome_bins([("chr2", CHR2_BP)], bin_size), bin_size, [("chr2", CHR2_BP)])
    band = max(100, window_bp // bin_size + 20)
    ref = synth.synthetic_sample(genome, band, seed=100 + 2 * rank, structure_seed=7 + rank)
    qry = synth.synthetic_sample(genome, band, seed=101 + 2 * rank, structure_seed=7 + rank, perturb=0.1)
    pairs = synth.sliding_pairs(genome, window_bp, STEP_BP)
    return genome, ref, qry, pairs


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(window):
    """DRAM bytes per launch of the compare kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(window)
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------- CPU arms
_G = {}


def _cpu_chunk(idx):
    from oracle import chess_oracle as O
    chunk = [_G["pairs"][i] for i in idx]
    return O.compare_chunk(chunk, _G["re"], _G["rt"], _G["qe"], _G["qt"], True, **_G["kw"])


def cpu_setup(genome, ref, qry, pairs, window, oe=None):
    """Oracle-side state: O/E edges as dict of dicts (the reference's own layout)."""
    from oracle import chess_oracle as O
    regions = [O.Region(r.start, r.end, r.chromosome, ix=r.ix) for r in genome.regions()]
    trees = O.region_trees(regions)
    if oe is None:
        oe = []
        for s in (ref, qry):
            triples = list(zip(s[0].tolist(), s[1].tolist(), s[2].tolist()))
            oe.append(O.observed_expected(regions, triples))
    _G.update(re=O.edges_dict(oe[0]), qe=O.edges_dict(oe[1]), rt=trees, qt=trees, kw=window_params(window),
              pairs=[(pid, O.Region(r.start, r.end, r.chromosome, strand=r.strand),
                      O.Region(q.start, q.end, q.chromosome, strand=q.strand)) for pid, r, q in pairs])


def cpu_subset(genome, ref, qry, pairs, max_bins=12000):
    """The CPU oracle keeps its contacts in a dict of dicts (like the reference); for genome-wide
    workloads only the first `max_bins` bins and the pairs inside them are handed to it."""
    if genome.n_bins <= max_bins:
        return genome, ref, qry, pairs
    from chess_b200 import synth
    limit_bp = max_bins * genome.bin_size
    first = genome.names[0]
    sub = synth.Genome([(first, max_bins)], genome.bin_size, [(first, limit_bp)])
    keep = lambda s: tuple(a[(s[1] < max_bins)] for a in s)  # noqa: E731  (src <= sink)
    # strictly inside: a region ending on the last bin boundary would also touch the next bin
    # (the BED lookup of chess/helpers.py:212-237), which the subset does not have
    inside = limit_bp - genome.bin_size
    pairs = [p for p in pairs if p[1].chromosome == first and p[1].end <= inside and p[2].end <= inside]
    return sub, keep(ref), keep(qry), pairs


def cpu_run(sample_idx, pool, cores):
    """One pass over `sample_idx` with the reference's parallel structure:
    ceil(P / p)-sized chunks, one per worker (bin/chess:196-214)."""
    size = int(math.ceil(len(sample_idx) / cores))
    chunks = [sample_idx[i:i + size] for i in range(0, len(sample_idx), size)]
    t0 = time.perf_counter()
    out = pool.map(_cpu_chunk, chunks)
    dt = time.perf_counter() - t0
    return dt, [row for part in out for row in part]


def cpu_sample(n_pairs, target):
    target = max(1, min(n_pairs, target))
    return [int(v) for v in np.linspace(0, n_pairs - 1, target).round().astype(int)]


def run_reference(args, rank, world):
    """--impl reference: the CPU oracle port, all host cores, bounded sample per step."""
    if rank != 0:
        return
    import multiprocessing as mp
    genome, ref, qry, pairs = make_workload(0, args.workload, 1, args.bin_size, args.window_bp)
    genome, ref, qry, pairs = cpu_subset(genome, ref, qry, pairs)
    cpu_setup(genome, ref, qry, pairs, args.window)
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        probe = cpu_sample(len(pairs), cores)
        dt, _ = cpu_run(probe, pool, cores)
        per_pair_core = dt  # one pair per core
        budget = 150.0 / max(1, args.steps + args.warmup)
        target = int(max(cores, min(len(pairs), cores * budget / max(per_pair_core, 1e-3))))
        idx = cpu_sample(len(pairs), target)
        for _ in range(args.warmup):
            cpu_run(idx, pool, cores)
        times = [cpu_run(idx, pool, cores)[0] for _ in range(args.steps)]
    total = float(np.sum(times))
    value = len(idx) * args.steps / total
    n_bins = len(genome.regions())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {k: v for k, v in workload_config(args, len(pairs), n_bins, sample=len(idx)).items() if k != "l2"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": "%d of %d pairs per step, evenly spaced; oracle/chess_oracle.py (reference "
                                   "algorithm restated: dict-of-dict gather, restated skimage SSIM, reference SN) "
                                   "under a %d-process fork pool with ceil(P/p) chunks" % (len(idx), len(pairs), cores)},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_pairs, n_bins, sample=None):
    if args.workload == "chr2_25kb" and (args.bin_size != BIN or args.window_bp != WINDOW_BP):
        what = "human chr2 @ %g kb (%d bins), %g Mb windows step 100 kb" % (args.bin_size / 1e3, n_bins,
                                                                              args.window_bp / 1e6)
    else:
        what = ("configs[1]: human chr2 @ 25 kb (%d bins), 2 Mb windows step 100 kb" if args.workload == "chr2_25kb"
                else "configs[2]: human hg38 @ 10 kb genome-wide (%d bins), 2 Mb windows step 30 kb") % n_bins
    cfg = {"workload": "%s, chess sim %s, SN on" % (what, "-r 1 (default)" if args.window == "r1" else "-a " + args.window[1:]),
           "pairs_per_gpu": n_pairs,
           "bins_per_pair": args.window_bp // args.bin_size + 1 if args.workload == "chr2_25kb" else 201,
           "window": args.window,
           "l2": "flushed before every step (512 MiB write)",
           "parallelism": "pairs sharded, one process per GPU, no collective"}
    if sample is not None:
        cfg["pairs_per_step_cpu_sample"] = sample
    return cfg


# --------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world, local_rank):
    import copy
    import torch
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    line = measure(args, rank, world, local_rank, dist)
    if rank == 0 and world == 1 and args.workload == "chr2_25kb" and not args.no_genome_wide:
        # BASELINE configs[2] (the north-star target is quoted on it) rides along on the same line
        try:
            extra_args = copy.copy(args)
            extra_args.workload, extra_args.steps, extra_args.warmup, extra_args.no_cpu_baseline = \
                "hg38_10kb", 5, 3, True
            extra = measure(extra_args, rank, world, local_rank, None)
            line["genome_wide"] = {k: extra[k] for k in ("value", "unit", "ms_per_step", "roofline", "e2e", "steps",
                                                         "warmup", "gpu_launches")}
            line["genome_wide"]["workload"] = extra["config"]["workload"]
            line["genome_wide"]["pairs"] = extra["config"]["pairs_per_gpu"]
        except Exception as exc:  # never lose the contract line over the extra measurement
            line["genome_wide"] = {"error": repr(exc)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def measure(args, rank, world, local_rank, dist):
    """One workload on this rank's GPU; returns the JSON line on rank 0, None elsewhere."""
    import torch
    import chess_b200 as cb
    from chess_b200 import _native
    from chess_b200._native import lib, check
    from chess_b200.engine import get_context, region_spans, build_pair_table

    dev = torch.device("cuda", local_rank)
    genome, ref, qry, pairs = make_workload(rank, args.workload, world, args.bin_size, args.window_bp)
    regions = genome.regions()
    trees = cb.region_interval_trees(regions)
    params = window_params(args.window)
    p = _native.make_params(compute_sn=True, kernel=args.kernel,
                            flags=_native.FLAG_SYMMETRIC | _native.FLAG_EQUAL_SIZES, **params)
    ctx = get_context(local_rank)
    sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

    # ---- cold end-to-end once: host COO -> H2D -> O/E -> band -> compare -> D2H (also the setup)
    rf, rn = region_spans(trees, [q[1] for q in pairs])
    qf, qn = region_spans(trees, [q[2] for q in pairs])
    flip = np.array([1 if a.strand != b.strand else 0 for _, a, b in pairs], dtype=np.int32)
    max_n = int(max(rn.max(), qn.max()))
    n_pairs = len(pairs)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    re_ = cb.observed_expected_arrays(regions, ref, device=local_rank)
    qe = cb.observed_expected_arrays(regions, qry, device=local_rank)
    rw, rband, rsample = re_.sample_on(local_rank, max_n)
    qw, qband, qsample = qe.sample_on(local_rank, max_n)
    table = build_pair_table(rf, rn, rw, qf, qn, qw, flip)
    h_ssim, h_sn = np.empty(n_pairs), np.empty(n_pairs)
    h_status = np.empty(n_pairs, dtype=np.int32)

    h_z = np.empty(n_pairs)

    def host_step():
        # one C-ABI call with host buffers: pair table in, ssim / SN / status / z_ssim out
        check(lib.chess_sim_band_batch_host(ctx.handle, C.byref(rsample), C.byref(qsample),
                                            table.ctypes.data_as(C.c_void_p), n_pairs, max_n, C.byref(p),
                                            h_ssim.ctypes.data_as(C.c_void_p), h_sn.ctypes.data_as(C.c_void_p),
                                            h_status.ctypes.data_as(C.c_void_p), h_z.ctypes.data_as(C.c_void_p), sp),
              ctx.handle, "chess_sim_band_batch_host")
        return h_z

    host_step()
    cold_s = time.perf_counter() - t0
    coo_bytes = int(sum(len(s[2]) * 16 for s in (ref, qry)))

    # ---- resident buffers for the device-timed steps
    d_table = torch.from_numpy(table.view(np.uint8)).to(dev)
    d_ssim = torch.empty(n_pairs, dtype=torch.float64, device=dev)
    d_sn = torch.empty(n_pairs, dtype=torch.float64, device=dev)
    d_z = torch.empty(n_pairs, dtype=torch.float64, device=dev)
    d_status = torch.empty(n_pairs, dtype=torch.int32, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def device_step(e0=None, e1=None, e2=None):
        if e0 is not None:
            e0.record()
        check(lib.chess_compare_band_batch(ctx.handle, C.byref(rsample), C.byref(qsample), ptr(d_table), n_pairs,
                                           max_n, C.byref(p), ptr(d_ssim), ptr(d_sn), ptr(d_status), sp),
              ctx.handle, "chess_compare_band_batch")
        if e1 is not None:
            e1.record()
        check(lib.chess_zscore(ctx.handle, ptr(d_ssim), n_pairs, ptr(d_z), sp), ctx.handle, "chess_zscore")
        if e2 is not None:
            e2.record()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        device_step()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = ctx.launches()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        device_step(*ev[k])
    barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.launches() - launches0
    step_ms = [e[0].elapsed_time(e[2]) for e in ev]
    kern_ms = [e[0].elapsed_time(e[1]) for e in ev]
    total_ms = float(np.sum(step_ms))

    # ---- end to end through the host-buffer C-ABI call
    for _ in range(3):
        host_step()
    barrier()
    e2e_t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
        host_step()
    barrier()
    e2e_s = time.perf_counter() - e2e_t0
    # the flush is not part of the step: time it alone and subtract
    torch.cuda.synchronize(dev)
    f0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
    e2e_s -= time.perf_counter() - f0
    sampler.stop_flag = True
    sampler.join(timeout=2)

    ok = int((h_status == 0).sum())
    z_dev = d_z.cpu().numpy()
    host_step()
    valid = h_ssim[~np.isnan(h_ssim)]
    z_numpy = (h_ssim - valid.mean()) / valid.std()            # bin/chess:228-233
    z_dev_ok = bool(np.allclose(z_dev, z_numpy, rtol=1e-9, atol=1e-12, equal_nan=True) and
                    np.allclose(h_z, z_numpy, rtol=1e-9, atol=1e-12, equal_nan=True))

    # ---- max over ranks
    if dist is not None:
        t = torch.tensor([total_ms, e2e_s, float(np.mean(kern_ms))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = float(t[0]), float(t[1])
        cnt = torch.tensor([n_pairs], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt)
        all_pairs = int(cnt[0])
    else:
        all_pairs = n_pairs
    if rank != 0:
        return None

    value = all_pairs * args.steps / (total_ms * 1e-3)
    peak, peak_src = measured_peak()
    alg_bytes = float(np.sum(2.0 * np.maximum(rn, qn).astype(np.float64) ** 2 * 8 + 24))
    kavg_ms = float(np.mean(kern_ms))
    achieved = alg_bytes / (kavg_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak" if args.workload == "chr2_25kb" else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": dict(workload_config(args, n_pairs, len(regions)),
                       pairs_ok=ok, wall_s_timed_region=wall, zscore_device_matches_host=z_dev_ok,
                       e2e_cold={"seconds": cold_s, "pairs_per_s": n_pairs / cold_s,
                                 "h2d_bytes": coo_bytes + table.nbytes,
                                 "what": "host COO -> H2D -> expected+O/E -> band -> compare -> D2H, first call "
                                         "(includes CUDA context/module load)"}),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": recorded_traffic(args.window) if (args.workload == "chr2_25kb" and args.bin_size == BIN
                                                                  and args.window_bp == WINDOW_BP) else None,
                     "kernel": "compare (gather+mask+SSIM+SN), %.1f us per launch, %d pairs" % (kavg_ms * 1e3, n_pairs),
                     "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src},
        "e2e": {"value": all_pairs * args.steps / e2e_s, "unit": "pairs/s",
                "h2d_bytes_per_step": int(table.nbytes), "d2h_bytes_per_step": int(n_pairs * 28),
                "call": "chess_sim_band_batch_host (host pair table in; ssim, SN, status, z_ssim out to host arrays; up to 8192 "
                        "pairs the kernels read the table from / write the results to the pinned staging buffer over "
                        "PCIe, larger batches use one DMA each way)"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    if world == 1 and args.workload == "chr2_25kb" and not args.no_background:
        try:
            line["background"] = measure_background(args, local_rank, genome, regions, trees, re_, qe, pairs,
                                                    h_ssim, params)
        except Exception as exc:  # never lose the contract line over the extra measurement
            line["background"] = {"error": repr(exc)}
    if world == 1 and args.workload == "chr2_25kb" and not args.no_ingest:
        try:
            line["ingest"] = measure_ingest(local_rank, ref)
        except Exception as exc:
            line["ingest"] = {"error": repr(exc)}
    if world == 1 and args.workload == "chr2_25kb" and not args.no_synthetic:
        try:
            line["synthetic_one_vs_many"] = measure_synthetic(local_rank)
        except Exception as exc:
            line["synthetic_one_vs_many"] = {"error": repr(exc)}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(genome, ref, qry, pairs, args, re_, qe, h_ssim, h_sn)
    re_.drop_device_copies()
    qe.drop_device_copies()
    return line


def measure_background(args, local_rank, genome, regions, trees, re_, qe, pairs, pair_ssim, params, n_sample=64):
    """`chess sim --background-query` (bin/chess:259-332) for a sample of the pairs: every
    query bin start x both strands as background regions, compute_sn off, reduced to z_bg / p_bg on the
    device.  Throughput in background comparisons per second: device-timed (arrays resident) and through
    CompareEngine.background (host arrays in, z_bg / p_bg out)."""
    import torch
    import chess_b200 as cb
    from chess_b200.engine import region_spans
    from chess_b200.cli import _background_regions_from_query
    dev = torch.device("cuda", local_rank)
    step = max(1, len(pairs) // n_sample)
    pick = pairs[::step][:n_sample]
    idx = list(range(0, len(pairs), step))[:n_sample]
    t0 = time.perf_counter()
    regs = _background_regions_from_query(regions, pick[0][2], True)
    bf, bn = region_spans(trees, regs)
    bstrand = np.array([0 if r.strand == '+' else 1 for r in regs], dtype=np.int32)
    host_prep_s = time.perf_counter() - t0
    rf, rn = region_spans(trees, [p[1] for p in pick])
    zeros = np.zeros(len(pick), dtype=np.int32)
    s = np.asarray(pair_ssim)[idx]
    eng = cb.CompareEngine(re_, trees, qe, trees, devices=[local_rank])
    n_cmp = len(pick) * len(bf)
    for _ in range(2):
        z, pv = eng.background(rf, rn, zeros, s, bf, bn, bstrand, own=(rf, rn, zeros), **params)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    z, pv = eng.background(rf, rn, zeros, s, bf, bn, bstrand, own=(rf, rn, zeros), **params)
    e2e_s = time.perf_counter() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx = cb.engine.get_context(local_rank)
    from chess_b200 import _native
    from chess_b200._native import lib, check
    p = _native.make_params(compute_sn=False, flags=_native.FLAG_SYMMETRIC | _native.FLAG_EQUAL_SIZES, **params)
    max_n = int(max(rn.max(), bn.max()))
    rw, rband, rsample = re_.sample_on(local_rank, max_n)
    qw, qband, qsample = qe.sample_on(local_rank, max_n)
    up = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in
          (rf.astype(np.int64), rn.astype(np.int32), zeros, s, bf.astype(np.int64), bn.astype(np.int32), bstrand)]
    d_z = torch.empty(len(pick), dtype=torch.float64, device=dev)
    d_p = torch.empty(len(pick), dtype=torch.float64, device=dev)
    ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def device_call():
        check(lib.chess_background_batch(ctx.handle, C.byref(rsample), C.byref(qsample), ptr(up[0]), ptr(up[1]),
                                         ptr(up[2]), ptr(up[0]), ptr(up[1]), ptr(up[2]), ptr(up[3]), len(pick),
                                         ptr(up[4]), ptr(up[5]), ptr(up[6]), len(bf), max_n, C.byref(p), ptr(d_z),
                                         ptr(d_p), None, sp), ctx.handle, "chess_background_batch")

    for _ in range(2):
        device_call()
    torch.cuda.synchronize(dev)
    launches0 = ctx.launches()
    e0.record()
    device_call()
    e1.record()
    torch.cuda.synchronize(dev)
    dev_s = e0.elapsed_time(e1) * 1e-3
    same = bool(np.allclose(d_z.cpu().numpy(), z, rtol=1e-9, atol=1e-12, equal_nan=True) and
                np.array_equal(d_p.cpu().numpy(), pv, equal_nan=True))
    return {"what": "chess sim --background-query --limit-background: %d pairs x %d background regions "
                    "(every query bin start, both strands), compute_sn off, z_bg / p_bg reduced on the device"
                    % (len(pick), len(bf)),
            "comparisons": n_cmp, "value": n_cmp / dev_s, "unit": "background comparisons/s",
            "seconds_device": dev_s, "gpu_launches": int(ctx.launches() - launches0),
            "e2e": {"value": n_cmp / e2e_s, "unit": "background comparisons/s", "seconds": e2e_s,
                    "call": "CompareEngine.background (host spans in, z_bg / p_bg out)",
                    "host_region_list_s": host_prep_s},
            "whole_run_estimate_s": len(pairs) * len(bf) / (n_cmp / dev_s),
            "finite_z": int(np.isfinite(z).sum()), "device_call_matches_engine": same}


def measure_synthetic(local_rank, n=100, pool=1000):
    """BASELINE configs[0] (utils/run_synthetic_test.py:97-121): one dense reference against its truth and
    `pool` decoys, SSIM only, data_range 2.0, per window of the script.  `device`: the pool is resident, one
    chess_compare_batch launch per window, CUDA events.  `e2e`: chess_b200.synthetic_test.OneVsMany from
    host matrices (one upload of the pool, then every window of the script with results back on the host)."""
    import torch
    from chess_b200 import synth, _native
    from chess_b200 import synthetic_test as S
    from chess_b200.engine import get_context, PAIR_DTYPE, _ptr, _stream_ptr, check
    from oracle import chess_oracle as O
    rng = np.random.default_rng(0)
    base = synth.dense_decay(n)
    feat = synth.dense_features(n, rng)
    ref = synth.dense_oe(synth.dense_noisy(base + feat, rng))
    queries = [synth.dense_oe(synth.dense_noisy(base + (feat if k == 0 else synth.dense_features(n, rng)), rng))
               for k in range(pool + 1)]
    P = len(queries)
    dev = torch.device("cuda", local_rank)
    ctx = get_context(local_rank)
    with torch.cuda.device(dev):
        d_ref = torch.from_numpy(ref.ravel().copy()).to(dev)
        d_pool = torch.from_numpy(np.concatenate([q.ravel() for q in queries])).to(dev)
        tab = np.zeros(P, dtype=PAIR_DTYPE)
        for k in range(P):
            tab[k] = (0, k * n * n, n, n, n, n, 0, 0)
        d_tab = torch.from_numpy(tab.view(np.uint8)).to(dev)
        ssim = torch.empty(P, dtype=torch.float64, device=dev)
        sn = torch.empty(P, dtype=torch.float64, device=dev)
        status = torch.empty(P, dtype=torch.int32, device=dev)
        stream = _stream_ptr(dev)
        peak, _ = measured_peak()
        alg = P * (2.0 * n * n * 8 + 24)
        rows, launches = [], 0
        for rel in S.distinct_windows(n):
            win = S.window_size(n, rel)
            prm = _native.make_params(absolute_window_size=win, min_bins=0, keep_unmappable_bins=True,
                                      mappability_cutoff=0.0, compute_sn=False, data_range=2.0)
            best = None
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                check(_native.lib.chess_compare_batch(ctx.handle, _ptr(d_ref), _ptr(d_pool), _ptr(d_tab), P, n,
                                                      C.byref(prm), _ptr(ssim), _ptr(sn), _ptr(status), stream),
                      ctx.handle, "chess_compare_batch")
                e1.record()
                torch.cuda.synchronize()
                launches += 1
                dt = e0.elapsed_time(e1) * 1e-3
                best = dt if rep and (best is None or dt < best) else best
            rows.append({"window": win, "pairs_per_s": P / best, "roofline_frac": alg / best / (peak * 1e9)})
        rel = 0.1
        rels = S.distinct_windows(n)
        t0 = time.perf_counter()
        batch = S.OneVsMany(ref, queries[0], queries[1:], device=local_rank)
        for r_ in rels:
            out = batch.run(r_)
            if r_ == rel:
                truth, control, pval, z = out
        e2e_s = time.perf_counter() - t0
        k = 8
        t0 = time.perf_counter()
        want = np.array([O.synthetic_compare_matrices(ref, q, rel) for q in queries[1:1 + k]])
        cpu_s = (time.perf_counter() - t0) / k
        dev_rel = float(np.max(np.abs(control[:k] - want) / np.abs(want)))
    return {"what": "configs[0]: one dense O/E reference vs truth + %d decoys, %d bins, SSIM only, data_range 2.0 "
                    "(utils/run_synthetic_test.py:97-121), generic kernel" % (pool, n),
            "pairs_per_launch": P, "device": rows, "gpu_launches": launches,
            "e2e": {"windows": len(rels), "pairs_per_s": P * len(rels) / e2e_s, "seconds": e2e_s,
                    "h2d_bytes": int((P + 1) * n * n * 8), "d2h_bytes": int(P * 8 * len(rels)),
                    "call": "chess_b200.synthetic_test.OneVsMany (host matrices uploaded once; per window of the "
                            "script: truth, controls, p, z back on the host)",
                    "p_window_9": pval, "z_window_9": z},
            "cpu_oracle_1_core_pairs_per_s": 1.0 / cpu_s, "max_rel_dev_gpu_vs_cpu": dev_rel}


def measure_ingest(local_rank, sample):
    """Sparse-matrix text of the reference sample ("src<TAB>sink<TAB>repr(weight)" lines) -> COO:
    chess_sparse_text_* on the device (host bytes in, host arrays out) beside the host parser
    (pandas C reader, round-trip floats) and the reference's own Python loop on a sub-sample."""
    import torch
    from chess_b200.ingest import parse_sparse_text
    src, sink, w = sample[0], sample[1], sample[2]
    text = "".join("%d\t%d\t%r\n" % t for t in zip(src.tolist(), sink.tolist(), w.tolist())).encode()
    data = np.frombuffer(bytearray(text), dtype=np.uint8)
    n_lines = len(w)
    parse_sparse_text(data, device=local_rank)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, b, v, n_data, n_bad = parse_sparse_text(data, device=local_rank)
    dev_s = time.perf_counter() - t0
    exact = bool(v.tobytes() == np.asarray(w, dtype=np.float64).tobytes() and
                 np.array_equal(a, np.minimum(src, sink)) and np.array_equal(b, np.maximum(src, sink)))
    import io
    import pandas as pd
    t0 = time.perf_counter()
    frame = pd.read_csv(io.BytesIO(text), sep="\t", comment="#", header=None, usecols=[0, 1, 2],
                        dtype={2: np.float64}, engine="c", float_precision="round_trip")
    host_s = time.perf_counter() - t0
    sub = text[:text.index(b"\n", len(text) // 20) + 1].decode()
    t0 = time.perf_counter()
    k = 0
    for ln in sub.splitlines():
        f = ln.rstrip().split("\t")
        _ = (int(f[0]), int(f[1]), float(f[2]))
        k += 1
    loop_s = time.perf_counter() - t0
    return {"what": "sparse-matrix text of one chr2 sample -> canonical COO (int, int, float per line)",
            "lines": n_lines, "bytes": len(text), "bit_identical_to_source": exact,
            "device": {"lines_per_s": n_lines / dev_s, "GB_per_s": len(text) / dev_s / 1e9, "seconds": dev_s,
                       "call": "chess_b200.ingest.parse_sparse_text (H2D of the text + parse + D2H of the arrays)"},
            "host_pandas": {"lines_per_s": len(frame) / host_s, "seconds": host_s},
            "reference_python_loop": {"lines_per_s": k / loop_s, "sample_lines": k}}


def cpu_baseline(genome, ref, qry, pairs, args, re_, qe, gpu_ssim, gpu_sn):
    """Oracle on the host cores, bounded sample (~10-30 s), same inputs as the GPU run."""
    import multiprocessing as mp
    id_to_k = {p[0]: k for k, p in enumerate(pairs)}
    limit = 12000
    if genome.n_bins > limit:
        genome, ref, qry, pairs = cpu_subset(genome, ref, qry, pairs, limit)
    # the very O/E values the GPU used (restricted to the bins the CPU sample covers)
    oe = []
    for smp in (re_, qe):
        a, b, w = smp.triples()
        sel = b < genome.n_bins
        oe.append(list(zip(a[sel].tolist(), b[sel].tolist(), w[sel].tolist())))
    cpu_setup(genome, ref, qry, pairs, args.window, oe=oe)
    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(cores) as pool:
        probe = cpu_sample(len(pairs), cores)
        dt, _ = cpu_run(probe, pool, cores)
        target = int(max(cores, min(len(pairs), cores * 15.0 / max(dt, 1e-3))))
        idx = cpu_sample(len(pairs), target)
        dt, rows = cpu_run(idx, pool, cores)
    worst = 0.0
    for i, (pid, s, sn) in zip(idx, rows):
        k = id_to_k[pid]
        for got, want in ((gpu_ssim[k], s), (gpu_sn[k], sn)):
            if not (np.isnan(want) and np.isnan(got)):
                worst = max(worst, abs(got - want) / max(abs(want), 1e-300))
    return {"value": len(idx) / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": "%d of %d pairs, evenly spaced, one pass; oracle/chess_oracle.py under a %d-process fork "
                      "pool with ceil(P/p) chunks (reference algorithm restated, not skimage itself)"
                      % (len(idx), len(pairs), cores),
            "seconds": dt, "max_rel_dev_gpu_vs_cpu": worst}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
