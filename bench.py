#!/usr/bin/env python
"""Benchmark of the `chess sim` comparison hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload hg38_10kb|chr2_25kb|dlbcl_5kb_bg1000|sweep] [--window r1|a3|a5|a7|a9|a11]

A *step* is one pass of the hot path over the whole pair list of the workload: for every region pair gather
both submatrices from the resident O/E band, mask / compact, SSIM, SN, then the run-level z-score.

Workloads (BASELINE.json configs; all synthetic, seeded)
  hg38_10kb         configs[2], the DEFAULT and the configuration the north-star target is quoted on: human
                    hg38 at 10 kb (308 839 bins), 2 Mb windows sliding by 30 kb = 101 354 pairs of 201 bins,
                    default `chess sim` parameters (-r 1, SN on).  STRONG scaling: under torchrun the one pair
                    list is cut into contiguous ranges, one per rank (bin/chess:196-214 cuts it into ceil(P/p)
                    chunks the same way), results are gathered on rank 0 (20 B per pair) and z-scored there.
  chr2_25kb         configs[1]: chr2 at 25 kb, 2 Mb windows step 100 kb, 2 402 pairs of 81 bins.  Weak scaling
                    (every rank its own chr2-sized samples).  Rides along on the default line as `chr2_25kb`.
  dlbcl_5kb_bg1000  configs[3]: chr2 at 5 kb, 1 Mb windows (201 bins; --window-bp 3000000 for 601 bins) step
                    100 kb, every pair also against 1000 background regions (--background-regions style, SN
                    off) reduced to z_bg / p_bg on the device.  Strong scaling over the pairs.
  sweep             configs[4]: -r 1 and -a 3 ... 11 x submatrix sizes 41 ... 401 bins, ~100 k pairs per cell
                    (hg38 at 25 kb, step 30 kb); one line with a `cells` table.  Strong scaling.

Numbers on the JSON line
  value         pairs/s, whole job over all ranks, samples and pair table already resident in HBM; CUDA-event
                time of K steps, L2 flushed before every step, max over ranks.  At N > 1 a step includes the
                all-gather of the ssim values (NCCL, 8 B per pair) that the run-level z-score needs.
  e2e           the same work through the host-buffer C-ABI calls: pair table from host memory, results back
                in host memory and z-scores, every step, host arrays page-locked.  N = 1: one
                `chess_sim_band_batch_host` call.  N > 1: every rank copies its range of the pair table to its
                GPU and compares it, the ssim values are all-gathered and z-scored on the device, and
                `distributed.gather_tensors_to_rank0` brings ssim, SN, status and z (28 B per pair) to rank 0's
                host memory in one NCCL gather and one copy -- all inside the timed region.
  roofline      compare kernel only: algorithmic bytes (2*n*n*8 + 24 per pair) / its CUDA-event
                duration, against the measured HBM copy bandwidth (MEASURED_PEAKS.json); `traffic` = DRAM bytes
                of one pass from an ncu pass this very run makes on a child process (N = 1; null if ncu is
                not usable); `secondary` = the FP64 bound: algorithmic flops / the same duration
                against the measured FP64 FMA peak (profiles/r01_microbench.txt).
  run           run-specific facts (pairs that gave a result, sharding identity, cold start, ...); `config`
                holds only what defines the workload and is identical in both arms.
  cpu_baseline  the CPU oracle (reference algorithm restated, oracle/) under a fork pool on this box's cores,
                on a bounded sample, with the deviation of the GPU results from it.

`--impl reference` runs that CPU oracle port alone (no GPU, nothing of chess_b200 imported).
"""
import argparse
import ctypes as C
import importlib.util
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHR2_BP = 242193529
METRIC = "region pairs/sec (SSIM+SN+z)"
FP64_PEAK_TFLOPS = 33.9          # measured DFMA rate on this pool's B200 (profiles/r01_microbench.txt)
WINDOWS = ["r1", "a3", "a5", "a7", "a9", "a11"]
CPU_BINS = 12000                 # the CPU arms work on this many bins (dict-of-dict store, like the reference)
SWEEP_WINDOW_BP = [1000000, 2000000, 3000000, 5000000, 10000000]   # 41, 81, 121, 201, 401 bins at 25 kb

WORKLOADS = {
    # name: (BASELINE config, chromosomes, bin size, region size, step, scaling)
    "hg38_10kb": dict(cfg=2, genome="hg38", bin=10000, window_bp=2000000, step_bp=30000, scaling="strong"),
    "chr2_25kb": dict(cfg=1, genome="chr2", bin=25000, window_bp=2000000, step_bp=100000, scaling="weak"),
    "dlbcl_5kb_bg1000": dict(cfg=3, genome="chr2", bin=5000, window_bp=1000000, step_bp=100000, scaling="strong",
                             background=1000),
    "sweep": dict(cfg=4, genome="hg38", bin=25000, window_bp=None, step_bp=30000, scaling="strong"),
}


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--window", default="r1", choices=WINDOWS, help="r1 = default `chess sim` (-r 1); aN = -a N")
    ap.add_argument("--workload", default="hg38_10kb", choices=sorted(WORKLOADS))
    ap.add_argument("--bin-size", type=int, default=None, help="override the workload's bin size in bp")
    ap.add_argument("--window-bp", type=int, default=None, help="override the workload's region size in bp")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="N = 1 default line only: skip chr2_25kb, background, ingest, synthetic_one_vs_many")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu child pass that measures DRAM traffic")
    ap.add_argument("--kernel", type=int, default=0,
                    help="0 auto, 1 force the generic kernel, 2 the CTA-per-pair tuned kernel, 3 full-square items")
    ap.add_argument("--sweep-windows", default=",".join(WINDOWS),
                    help="sweep workload: comma-separated subset of the windows (development runs)")
    ap.add_argument("--probe-kernel", action="store_true",
                    help="internal: set the workload up, launch the compare kernel three times, exit (run under ncu)")
    return ap.parse_args(argv)


def window_params(tag):
    return dict(absolute_window_size=int(tag[1:])) if tag.startswith("a") else dict()


_SYNTH = None


def synth_module(standalone=False):
    """chess_b200/synth.py.  The CPU reference arm loads it by file path (`standalone`): the generators need nothing
    of the package, and that arm must not import chess_b200 (its __init__ maps the CUDA library)."""
    global _SYNTH
    if _SYNTH is None:
        if standalone:
            spec = importlib.util.spec_from_file_location("chess_b200_synth_standalone",
                                                          os.path.join(ROOT, "chess_b200", "synth.py"))
            _SYNTH = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(_SYNTH)
        else:
            from chess_b200 import synth as _SYNTH
    return _SYNTH


def workload_spec(args, window_bp=None):
    """Everything that defines the workload, from the arguments alone (no data is generated here)."""
    S = synth_module()
    w = dict(WORKLOADS[args.workload])
    w["name"] = args.workload
    if args.bin_size:
        w["bin"] = args.bin_size
    if args.window_bp:
        w["window_bp"] = args.window_bp
    if window_bp:
        w["window_bp"] = window_bp
    w["chroms_bp"] = S.HG38_CHROM_BP if w["genome"] == "hg38" else [("chr2", CHR2_BP)]
    w["n_bins"] = int(sum(n for _, n in S.chromosome_bins(w["chroms_bp"], w["bin"])))
    if w["window_bp"]:
        w["pairs"] = S.count_sliding_pairs(w["chroms_bp"], w["window_bp"], w["step_bp"])
        w["bins_per_pair"] = w["window_bp"] // w["bin"] + 1        # the BED lookup touches the next bin too
    return w


def workload_config(args):
    """The `config` object of the JSON line.  A function of the command line only, so the GPU arm and the
    CPU reference arm print the same object; run-specific values live under `run`."""
    w = workload_spec(args)
    win = "-r 1 (default)" if args.window == "r1" else "-a " + args.window[1:]
    genome = ("human hg38 genome-wide" if w["genome"] == "hg38" else "human chr2") + " @ %g kb (%d bins)" % (
        w["bin"] / 1e3, w["n_bins"])
    cfg = {"baseline_config": "configs[%d]" % w["cfg"], "window": args.window,
           "l2": "GPU arm: flushed before every step (512 MiB write)",
           "parallelism": ("one pair list cut into contiguous ranges, one process per GPU, host-side gather + z on "
                           "rank 0 (strong scaling)" if w["scaling"] == "strong" else
                           "every rank its own chr2-sized samples and pair list (weak scaling)") +
                          "; no data-path collective except the 8 B/pair ssim all-gather the run-level z needs"}
    if args.workload == "sweep":
        cfg["workload"] = ("configs[4]: %s, SSIM window sweep -r 1 and -a 3/5/7/9/11 x regions of %s bins, step %g kb, "
                           "SN on" % (genome, "/".join(str(b // w["bin"] + 1) for b in SWEEP_WINDOW_BP), w["step_bp"] / 1e3))
        cfg["pairs"] = [workload_spec(args, b)["pairs"] for b in SWEEP_WINDOW_BP]
        cfg["bins_per_pair"] = [b // w["bin"] + 1 for b in SWEEP_WINDOW_BP]
        cfg["window"] = WINDOWS
        return cfg
    cfg["workload"] = "configs[%d]: %s, %g Mb windows step %g kb, chess sim %s, SN on" % (
        w["cfg"], genome, w["window_bp"] / 1e6, w["step_bp"] / 1e3, win)
    cfg["pairs"] = w["pairs"]
    cfg["bins_per_pair"] = w["bins_per_pair"]
    if w.get("background"):
        cfg["workload"] += ", every pair against %d background regions (--background-regions, SN off) -> z_bg, p_bg" \
                           % w["background"]
        cfg["background_regions"] = w["background"]
    return cfg


def make_workload(w, rank=0, region_cls=None, truncate_bins=None):
    """Synthetic sample pair and the sliding pair list of a workload spec.  Strong-scaling workloads are the
    same on every rank (the caller keeps its range of the pair list); weak-scaling ones are seeded by rank.
    `truncate_bins` (CPU reference arm): only that many bins of the first chromosome are generated -- the oracle
    port works on a bounded sample anyway."""
    S = synth_module()
    chroms_bp = w["chroms_bp"]
    if truncate_bins and chroms_bp[0][1] > truncate_bins * w["bin"]:
        chroms_bp = [(chroms_bp[0][0], truncate_bins * w["bin"])]
    genome = S.Genome(S.chromosome_bins(chroms_bp, w["bin"]), w["bin"], chroms_bp)
    band = max(100, (w.get("band_bp") or w["window_bp"]) // w["bin"] + 15)
    shift = rank if w["scaling"] == "weak" else 0
    ref = S.synthetic_sample(genome, band, seed=100 + 2 * shift, structure_seed=7 + shift)
    qry = S.synthetic_sample(genome, band, seed=101 + 2 * shift, structure_seed=7 + shift, perturb=0.1)
    pairs = S.sliding_pairs(genome, w["window_bp"], w["step_bp"], region_cls=region_cls) if w["window_bp"] else None
    return genome, ref, qry, pairs


def background_bed(w, genome):
    """The seeded --background-regions BED of configs[3]: `background` random regions of the pair size, written
    the way `chess background` writes them (1-based start, end = start + window; bin/chess:505-517), so that
    the BED lookup gives them as many bins as the pair regions have."""
    rng = np.random.default_rng(1)
    n = w["window_bp"] // w["bin"]
    starts = np.sort(rng.integers(0, genome.n_bins - n - 2, w["background"]))
    return [("chr2", int(s) * w["bin"] + 1, int(s) * w["bin"] + 1 + w["window_bp"]) for s in starts]


def flops_per_pixel(window, compute_sn=True):
    """Algorithmic FP64 operations per pixel of the n x n pair: windowed SSIM = products
    (3), five vertical and five horizontal running sums (20), pointwise formula (19) + 1 divide; -r 1 needs the
    five moments only (3 mul + 5 add); SN = difference, square, two running sums each way, mean / variance and
    the ratio (15) + 1 divide.  A divide counts as one operation."""
    ssim = 8 if window == "r1" else 43
    return ssim + (16 if compute_sn else 0)


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

    def stop(self):
        self.stop_flag = True
        self.join(timeout=2)

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


def roofline(alg_bytes, kernel_ms, n_pixels, window, compute_sn, what, traffic=None):
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    tflops = n_pixels * flops_per_pixel(window, compute_sn) / (kernel_ms * 1e-3) / 1e12
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": what, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
            "secondary": {"bound": "fp64", "achieved": tflops, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                          "frac": tflops / FP64_PEAK_TFLOPS,
                          "flops_per_pixel": flops_per_pixel(window, compute_sn),
                          "peak_source": "measured DFMA rate, profiles/r01_microbench.txt"}}


def ncu_traffic(args, timeout_s=240):
    """DRAM bytes (read + write) of ONE PASS of the compare kernel over this workload: a child process sets the
    same workload up and makes PROBE_STEPS passes under `ncu --metrics dram__bytes_*`; the passes after the first
    are averaged (parse_ncu_traffic).  Returns (bytes, note); bytes is None when ncu cannot run here."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
           "--kernel-name", "regex:^compare_(r1|win)", "--launch-count", "48", "--csv",
           sys.executable, os.path.abspath(__file__), "--probe-kernel", "--workload", args.workload,
           "--window", args.window, "--kernel", str(args.kernel)]
    if args.bin_size:
        cmd += ["--bin-size", str(args.bin_size)]
    if args.window_bp:
        cmd += ["--window-bp", str(args.window_bp)]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
    except Exception as exc:
        return None, "ncu pass failed: %r" % (exc,)
    return parse_ncu_traffic(out.stdout, out.returncode, out.stderr)


PROBE_STEPS = 3          # device-timed passes the --probe-kernel child makes ...
PROBE_PASSES = PROBE_STEPS + 1   # ... after the one cold host-buffer call every run starts with


def parse_ncu_traffic(text, rc=0, err="", probe_steps=PROBE_PASSES):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE PASS of the compare kernel from ncu's csv of the probe
    child (`probe_steps` passes in all).  The kernel that moves the most bytes is the dominant one; a pass
    launches it once per chunk of <= 65 536 pairs, so its launches are summed per pass (like `achieved`, which divides the
    algorithmic bytes of the whole pass by the time of the whole pass); the first pass (cold module, cold band) is
    left out."""
    import csv
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rows = [r for r in csv.reader(text.splitlines()) if len(r) > 10]
    if not rows:
        return None, "ncu printed no metric rows (rc %d): %s" % (rc, (err or text)[-200:].replace("\n", " "))
    head = rows[0]
    try:
        i_id, i_name, i_metric, i_unit, i_val = (head.index("ID"), head.index("Kernel Name"), head.index("Metric Name"),
                                                 head.index("Metric Unit"), head.index("Metric Value"))
    except ValueError:
        return None, "unexpected ncu csv header"
    per_launch, names = {}, {}
    for r in rows[1:]:
        if r[i_metric] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            per_launch[r[i_id]] = per_launch.get(r[i_id], 0.0) + float(r[i_val].replace(",", "")) * unit_scale.get(r[i_unit], 1.0)
            names[r[i_id]] = r[i_name].split("(")[0]
    if not per_launch:
        return None, "no compare kernel captured"
    by_name = {}
    for k in sorted(per_launch, key=int):
        by_name.setdefault(names[k], []).append(per_launch[k])
    name = max(by_name, key=lambda n: np.sum(by_name[n]))
    vals = by_name[name]
    if probe_steps > 1 and len(vals) % probe_steps == 0:
        per_pass = len(vals) // probe_steps
        warm = vals[per_pass:]
        return int(np.sum(warm) / (probe_steps - 1)), (
            "ncu dram__bytes_read.sum + dram__bytes_write.sum of one pass = %d launch(es) of %s, mean of %d passes in a "
            "child process of this run (L2 flushed before each)" % (per_pass, name, probe_steps - 1))
    vals = vals[1:] or vals
    return int(np.mean(vals)), ("ncu dram__bytes_read.sum + dram__bytes_write.sum, mean of %d launches of %s in a child "
                                "process of this run (L2 flushed before each; launches per pass unknown)"
                                % (len(vals), name))


# --------------------------------------------------------------------------- CPU arms (oracle port)
_G = {}


def _cpu_chunk(idx):
    from oracle import chess_oracle as O
    chunk = [_G["pairs"][i] for i in idx]
    return O.compare_chunk(chunk, _G["re"], _G["rt"], _G["qe"], _G["qt"], _G.get("sn", True), **_G["kw"])


def _cpu_bg_chunk(idx):
    """Background comparisons (reference region of a pair x background region, SN off; bin/chess:300-307)."""
    from oracle import chess_oracle as O
    out = []
    for i, j in idx:
        pid, rr, _ = _G["pairs"][i]
        out.append(O.compare_pair((pid, rr, _G["bg"][j]), _G["re"], _G["rt"], _G["qe"], _G["qt"], compute_sn=False,
                                  **_G["kw"])[1])
    return out


def oracle_pairs(pairs):
    from oracle import chess_oracle as O
    return [(pid, O.Region(r.start, r.end, r.chromosome, strand=r.strand),
             O.Region(q.start, q.end, q.chromosome, strand=q.strand)) for pid, r, q in pairs]


def cpu_setup(genome, ref, qry, pairs, window, oe=None):
    """Oracle-side state: O/E edges as dict of dicts (the reference's own layout)."""
    from oracle import chess_oracle as O
    regions = genome.regions(O.Region)
    trees = O.region_trees(regions)
    if oe is None:
        oe = []
        for s in (ref, qry):
            triples = list(zip(s[0].tolist(), s[1].tolist(), s[2].tolist()))
            oe.append(O.observed_expected(regions, triples))
    _G.update(re=O.edges_dict(oe[0]), qe=O.edges_dict(oe[1]), rt=trees, qt=trees, kw=window_params(window),
              pairs=oracle_pairs(pairs), sn=True)


def cpu_subset(genome, ref, qry, pairs, max_bins=12000):
    """The CPU oracle keeps its contacts in a dict of dicts (like the reference); for genome-wide
    workloads only the first `max_bins` bins and the pairs inside them are handed to it."""
    if genome.n_bins <= max_bins:
        return genome, ref, qry, pairs
    S = synth_module()
    limit_bp = max_bins * genome.bin_size
    first = genome.names[0]
    sub = S.Genome([(first, max_bins)], genome.bin_size, [(first, limit_bp)])
    keep = lambda s: tuple(a[(s[1] < max_bins)] for a in s)  # noqa: E731  (src <= sink)
    # strictly inside: a region ending on the last bin boundary would also touch the next bin
    # (the BED lookup of chess/helpers.py:212-237), which the subset does not have
    inside = limit_bp - genome.bin_size
    pairs = [p for p in pairs if p[1].chromosome == first and p[1].end <= inside and p[2].end <= inside]
    return sub, keep(ref), keep(qry), pairs


def cpu_run(sample_idx, pool, cores, fn=_cpu_chunk):
    """One pass over `sample_idx` with the reference's parallel structure:
    ceil(P / p)-sized chunks, one per worker (bin/chess:196-214)."""
    size = int(math.ceil(len(sample_idx) / cores))
    chunks = [sample_idx[i:i + size] for i in range(0, len(sample_idx), size)]
    t0 = time.perf_counter()
    out = pool.map(fn, chunks)
    dt = time.perf_counter() - t0
    return dt, [row for part in out for row in part]


def cpu_sample(n_pairs, target):
    target = max(1, min(n_pairs, target))
    return [int(v) for v in np.linspace(0, n_pairs - 1, target).round().astype(int)]


PORT_NOTE = ("oracle/chess_oracle.py (reference algorithm restated: dict-of-dict gather, restated skimage SSIM, "
             "reference SN; not skimage itself) under a %d-process fork pool with ceil(P/p) chunks")


def reference_pairs_rate(w, window, budget_s, steps, warmup, pool, cores):
    """pairs/s of the oracle port on a bounded, evenly spaced sample of one pair-list workload."""
    from oracle import chess_oracle as O
    genome, ref, qry, pairs = make_workload(w, 0, region_cls=O.Region, truncate_bins=CPU_BINS)
    cpu_setup(genome, ref, qry, pairs, window)
    return _timed_sample(len(pairs), budget_s, steps, warmup, cores, _cpu_chunk), len(pairs)


def _timed_sample(n_items, budget_s, steps, warmup, cores, fn, make_idx=cpu_sample):
    import multiprocessing as mp
    with mp.get_context("fork").Pool(cores) as pool:        # forked AFTER _G is set: workers inherit the dicts
        probe = make_idx(n_items, cores)
        dt, _ = cpu_run(probe, pool, cores, fn)             # one item per core
        per_step = budget_s / max(1, steps + warmup)
        target = int(max(cores, min(n_items, cores * per_step / max(dt, 1e-3))))
        idx = make_idx(n_items, target)
        for _ in range(warmup):
            cpu_run(idx, pool, cores, fn)
        times = [cpu_run(idx, pool, cores, fn)[0] for _ in range(steps)]
    return len(idx), float(np.sum(times))


def run_reference(args, rank, world):
    """--impl reference: the CPU oracle port, all host cores, bounded sample per step.  Imports nothing of
    chess_b200 (the generators are loaded by file path), so no CUDA library is mapped by this process."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    w = workload_spec(args)
    note = PORT_NOTE % cores
    if args.workload == "sweep":
        # every cell of the sweep on a few pairs; the aggregate is total pairs / total (extrapolated) time
        cells, total_pairs, total_s, spent = [], 0, 0.0, 0.0
        per_cell = 150.0 / (len(SWEEP_WINDOW_BP) * len(WINDOWS))
        for b in SWEEP_WINDOW_BP:
            from oracle import chess_oracle as O
            wc = workload_spec(args, b)
            genome, ref, qry, pairs = make_workload(dict(wc, band_bp=b), 0, region_cls=O.Region, truncate_bins=4000)
            for win in WINDOWS:
                cpu_setup(genome, ref, qry, pairs, win)
                n, secs = _timed_sample(len(pairs), per_cell, 1, 0, cores, _cpu_chunk)
                rate = n / secs
                cells.append({"bins": wc["bins_per_pair"], "window": win, "pairs_per_s": rate, "sample_pairs": n})
                total_pairs += wc["pairs"]
                total_s += wc["pairs"] / rate
                spent += secs
        value, ms = total_pairs / total_s, 1e3 * total_s
        sample = "%d cells, evenly spaced pairs of each, one pass per cell (%.0f s of CPU work in all); value = " \
                 "all cells' pairs / the time the port would need for them; %s" % (len(cells), spent, note)
        extra = {"cells": cells}
    elif w.get("background"):
        from oracle import chess_oracle as O
        genome, ref, qry, pairs = make_workload(w, 0, region_cls=O.Region, truncate_bins=CPU_BINS)
        bed = background_bed(w, genome)
        cpu_setup(genome, ref, qry, pairs, args.window)
        _G["bg"] = [O.Region(s, e, c, strand=None) for c, s, e in bed]
        n_bg = len(_G["bg"])

        def grid(n_items, target):
            target = max(1, min(n_items, target))
            flat = np.linspace(0, n_items - 1, target).round().astype(np.int64)
            return [(int(v // n_bg), int(v % n_bg)) for v in flat]

        budget = 150.0
        n_c, secs_c = _timed_sample(len(pairs) * n_bg, 0.8 * budget, args.steps, args.warmup, cores, _cpu_bg_chunk, grid)
        n_p, secs_p = _timed_sample(len(pairs), 0.2 * budget, args.steps, args.warmup, cores, _cpu_chunk)
        per_pair = (secs_p / (n_p * args.steps)) + w["background"] * (secs_c / (n_c * args.steps))
        value, ms = 1.0 / per_pair, 1e3 * (secs_c + secs_p) / args.steps
        sample = "per step %d of the pair x background-region comparisons (SN off) and %d of the pairs (SN on), evenly " \
                 "spaced; value = 1 / (time per pair + %d x time per background comparison); %s" % (
                     n_c, n_p, w["background"], note)
        extra = {}
    else:
        (n, secs), n_all = reference_pairs_rate(w, args.window, 150.0, args.steps, args.warmup, None, cores)
        value, ms = n * args.steps / secs, 1e3 * secs / args.steps
        sample = "%d pairs per step, evenly spaced over the %d pairs of a %d-bin stretch generated like the workload; %s" % (
            n, n_all, min(w["n_bins"], CPU_BINS), note)
        extra = {}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line.update(extra)
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
class Rank(object):
    """Process-group plumbing of one rank: NCCL for device buffers and timing, a gloo group for the
    host-side result gather (the north-star's "final host-side gather")."""

    def __init__(self, rank, world, local_rank):
        import torch
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.dev = torch.device("cuda", local_rank)
        torch.cuda.set_device(self.dev)
        self.dist = self.gloo = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
            self.gloo = dist.new_group(backend="gloo")

    def barrier(self):
        import torch
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max(self, *values):
        import torch
        if self.dist is None:
            return [float(v) for v in values]
        t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    def sum(self, value):
        import torch
        if self.dist is None:
            return int(value)
        t = torch.tensor([int(value)], dtype=torch.int64, device=self.dev)
        self.dist.all_reduce(t)
        return int(t[0])

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


class Resident(object):
    """One workload resident on this rank's GPU: both samples (O/E, band, index arrays) and the spans of the
    whole pair list; `table(lo, hi)` gives the chess_pair records of a contiguous range."""

    def __init__(self, w, R, max_n=None, shard=None):
        """``shard`` = (lo, hi): this rank compares pairs lo .. hi-1 only and holds the bands of the bin range they
        read (per-GPU placement); None = the whole bands."""
        import torch
        import chess_b200 as cb
        from chess_b200.engine import region_spans
        self.w, self.R = w, R
        self.genome, self.ref, self.qry, self.pairs = make_workload(w, R.rank)
        self.regions = self.genome.regions()
        self.trees = cb.region_interval_trees(self.regions)
        self.rf, self.rn = region_spans(self.trees, [q[1] for q in self.pairs])
        self.qf, self.qn = region_spans(self.trees, [q[2] for q in self.pairs])
        self.flip = np.array([1 if a.strand != b.strand else 0 for _, a, b in self.pairs], dtype=np.int32)
        # the band width (and with it the strip layout of the kernels) follows max_n: fix it over the WHOLE pair
        # list before any sharding, so that a pair gives the same bits on every rank count
        self.max_n = int(max_n or max(self.rn.max(), self.qn.max()))
        torch.cuda.synchronize(R.dev)
        t0 = time.perf_counter()
        self.re_ = cb.observed_expected_arrays(self.regions, self.ref, device=R.local_rank)
        self.qe = cb.observed_expected_arrays(self.regions, self.qry, device=R.local_rank)
        self.place(shard)
        torch.cuda.synchronize(R.dev)
        self.setup_s = time.perf_counter() - t0
        self.coo_bytes = int(sum(len(s[2]) * 16 for s in (self.ref, self.qry)))

    def place(self, shard=None):
        """Bands (and index arrays) on this rank's GPU: of the bins the pairs shard[0] .. shard[1]-1 read, or whole."""
        def rng(first, n):
            if shard is None:
                return None
            f, c = first[shard[0]:shard[1]], n[shard[0]:shard[1]]
            live = c > 0
            return (int(f[live].min()), int((f[live] + c[live]).max())) if live.any() else (0, 1)
        dev = self.R.local_rank
        self.rw, self.rband, self.rsample, self.r0 = self.re_.sample_on(dev, self.max_n, None, rng(self.rf, self.rn),
                                                                        with_origin=True)
        self.qw, self.qband, self.qsample, self.q0 = self.qe.sample_on(dev, self.max_n, None, rng(self.qf, self.qn),
                                                                        with_origin=True)
        self.band_bytes = int(self.rband.numel() * 8 + self.qband.numel() * 8)

    def table(self, lo=0, hi=None):
        from chess_b200.engine import build_pair_table
        hi = len(self.pairs) if hi is None else hi
        return build_pair_table(self.rf[lo:hi] - self.r0, self.rn[lo:hi], self.rw, self.qf[lo:hi] - self.q0,
                                self.qn[lo:hi], self.qw, self.flip[lo:hi])

    def alg_bytes(self, lo=0, hi=None):
        n = np.maximum(self.rn[lo:hi], self.qn[lo:hi]).astype(np.float64)
        return float(np.sum(2.0 * n * n * 8 + 24)), float(np.sum(n * n))

    def drop(self):
        self.re_.drop_device_copies()
        self.qe.drop_device_copies()


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def measure_pairs(args, R, w, res=None, steps=None, warmup=None, want_cpu=True, want_traffic=True):
    """Pair-list workloads (hg38_10kb strong, chr2_25kb weak): returns the JSON line on rank 0."""
    import torch
    from chess_b200 import _native
    from chess_b200._native import lib, check
    from chess_b200.engine import get_context
    from chess_b200.distributed import gather_tensors_to_rank0, my_shard

    steps = steps or args.steps
    warmup = max(warmup or args.warmup, 3)
    dev, rank, world = R.dev, R.rank, R.world
    strong = w["scaling"] == "strong" and world > 1
    own = res is None
    if res is None:
        res = Resident(w, R, shard=my_shard(workload_spec(args)["pairs"], rank, world) if strong else None)
    total = len(res.pairs)
    lo, hi = my_shard(total, rank, world) if strong else (0, total)
    n_local = hi - lo
    longest = int(math.ceil(total / world)) if strong else total
    params = window_params(args.window)
    p = _native.make_params(compute_sn=True, kernel=args.kernel,
                            flags=_native.FLAG_SYMMETRIC | _native.FLAG_EQUAL_SIZES, **params)
    ctx = get_context(R.local_rank)
    sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    table = res.table(lo, hi)
    # the caller's host buffers are page-locked (the contract's "pinned host memory"): the library then copies
    # straight from / into them instead of staging through its own pinned area
    def pinned(n, dtype):
        return torch.empty(n, dtype=dtype).pin_memory().numpy()
    pinned_table = torch.from_numpy(table.view(np.uint8)).pin_memory()
    h_table = pinned_table.numpy()
    h_ssim, h_sn, h_z = pinned(n_local, torch.float64), pinned(n_local, torch.float64), pinned(n_local, torch.float64)
    h_status = pinned(n_local, torch.int32)

    # ---- the two host-buffer C-ABI calls
    def host_sim():        # N = 1: pair table in; ssim, SN, status and z_ssim out
        check(lib.chess_sim_band_batch_host(ctx.handle, C.byref(res.rsample), C.byref(res.qsample), _np_ptr(h_table),
                                            n_local, res.max_n, C.byref(p), _np_ptr(h_ssim), _np_ptr(h_sn),
                                            _np_ptr(h_status), _np_ptr(h_z), sp), ctx.handle, "chess_sim_band_batch_host")
        return h_ssim, h_sn, h_status, h_z

    e_tab = torch.empty(table.nbytes, dtype=torch.uint8, device=dev) if strong else None
    e_pad = torch.full((longest,), float("nan"), dtype=torch.float64, device=dev)    # the pad entry stays NaN
    e_ssim = e_pad[:n_local]
    e_sn = torch.empty(n_local, dtype=torch.float64, device=dev)
    e_status = torch.empty(n_local, dtype=torch.int32, device=dev)
    e_all = torch.empty(longest * world, dtype=torch.float64, device=dev) if strong else None
    e_zall = torch.empty(longest * world, dtype=torch.float64, device=dev) if strong else None

    def host_sharded():    # N > 1: this rank's range; z_ssim on the device; everything gathered to rank 0's host
        e_tab.copy_(pinned_table, non_blocking=True)                      # the host pair table of this range
        check(lib.chess_compare_band_batch(ctx.handle, C.byref(res.rsample), C.byref(res.qsample), _ptr(e_tab),
                                           n_local, res.max_n, C.byref(p), _ptr(e_ssim), _ptr(e_sn), _ptr(e_status),
                                           sp), ctx.handle, "chess_compare_band_batch")
        # the run-level z (bin/chess:216-233) needs every rank's ssim: 8 B per pair all-gathered over NVLink,
        # K6 on every rank, and each rank's slice of z travels with its ssim, SN and status
        R.dist.all_gather_into_tensor(e_all, e_pad)
        check(lib.chess_zscore(ctx.handle, _ptr(e_all), e_all.numel(), _ptr(e_zall), sp), ctx.handle, "chess_zscore")
        e_z = e_zall[rank * longest:rank * longest + n_local]
        return gather_tensors_to_rank0((e_ssim, e_sn, e_status, e_z), total, rank, world, R.dist)

    host_step = host_sharded if strong else host_sim
    t0 = time.perf_counter()
    first = host_step()
    cold_s = res.setup_s + (time.perf_counter() - t0) if own else None

    # ---- resident buffers for the device-timed steps
    d_table = torch.from_numpy(table.view(np.uint8)).to(dev)
    d_ssim = torch.full((longest,), float("nan"), dtype=torch.float64, device=dev)   # the pad entry stays NaN
    d_sn = torch.empty(longest, dtype=torch.float64, device=dev)
    d_status = torch.empty(longest, dtype=torch.int32, device=dev)
    d_all = torch.empty(longest * world, dtype=torch.float64, device=dev) if strong else d_ssim
    d_z = torch.empty(longest * (world if strong else 1), dtype=torch.float64, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def device_step(e0=None, e1=None, e2=None):
        if e0 is not None:
            e0.record()
        check(lib.chess_compare_band_batch(ctx.handle, C.byref(res.rsample), C.byref(res.qsample), _ptr(d_table),
                                           n_local, res.max_n, C.byref(p), _ptr(d_ssim), _ptr(d_sn), _ptr(d_status),
                                           sp), ctx.handle, "chess_compare_band_batch")
        if e1 is not None:
            e1.record()
        if strong:   # the run-level z needs every rank's ssim values: 8 B per pair over NVLink, then K6 everywhere
            R.dist.all_gather_into_tensor(d_all, d_ssim)
        check(lib.chess_zscore(ctx.handle, _ptr(d_all), d_all.numel(), _ptr(d_z), sp), ctx.handle, "chess_zscore")
        if e2 is not None:
            e2.record()

    if args.probe_kernel:
        for _ in range(PROBE_STEPS):
            flush.zero_()
            device_step()
        torch.cuda.synchronize(dev)
        return None

    for _ in range(warmup):
        flush.zero_()
        device_step()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    sampler = ClockSampler(R.local_rank)
    R.barrier()
    sampler.start()
    launches0 = ctx.launches()
    wall0 = time.perf_counter()
    for k in range(steps):
        flush.zero_()
        device_step(*ev[k])
    R.barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.launches() - launches0
    total_ms = float(np.sum([e[0].elapsed_time(e[2]) for e in ev]))
    kern_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))

    # ---- end to end through the host-buffer C-ABI calls
    for _ in range(3):
        host_step()
    R.barrier()
    e2e_t0 = time.perf_counter()
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
        result = host_step()
    R.barrier()
    e2e_s = time.perf_counter() - e2e_t0
    torch.cuda.synchronize(dev)          # the flush is not part of the step: time it alone and subtract
    f0 = time.perf_counter()
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
    e2e_s -= time.perf_counter() - f0
    sampler.stop()

    total_ms, e2e_s, kern_ms = R.max(total_ms, e2e_s, kern_ms)
    all_pairs = total if (strong or world == 1) else R.sum(total)
    z_dev = d_z.cpu().numpy()
    if rank != 0:
        if own:
            res.drop()
        return None

    # ---- checks on rank 0: z-scores, and (strong) the gathered arrays against rank 0's own run of the WHOLE list
    g_ssim, g_sn, g_status, g_z = result
    valid = g_ssim[~np.isnan(g_ssim)]
    z_numpy = (g_ssim - valid.mean()) / valid.std()            # bin/chess:228-233
    if strong:
        bounds = [my_shard(total, r, world) for r in range(world)]
        z_dev = np.concatenate([z_dev[r * longest:r * longest + (b - a)] for r, (a, b) in enumerate(bounds)])
    z_ok = bool(np.allclose(z_dev, z_numpy, rtol=1e-9, atol=1e-12, equal_nan=True) and
                np.allclose(g_z, z_numpy, rtol=1e-9, atol=1e-12, equal_nan=True))
    run = {"pairs_ok": int((g_status == 0).sum()), "pairs_this_rank": n_local, "max_n": res.max_n,
           "band_bytes_this_rank": res.band_bytes,
           "wall_s_timed_region": wall, "zscore_device_matches_host": z_ok}
    if strong:
        res.place(None)            # rank 0 alone, after the timed regions: the whole bands for the reference run
        full_tab = res.table()
        f_ssim, f_sn = np.empty(total), np.empty(total)
        f_status = np.empty(total, dtype=np.int32)
        check(lib.chess_compare_band_batch_host(ctx.handle, C.byref(res.rsample), C.byref(res.qsample),
                                                _np_ptr(full_tab), total, res.max_n, C.byref(p), _np_ptr(f_ssim),
                                                _np_ptr(f_sn), _np_ptr(f_status), sp), ctx.handle, "full list")
        run["sharding_identical"] = bool(f_ssim.tobytes() == g_ssim.tobytes() and f_sn.tobytes() == g_sn.tobytes()
                                         and np.array_equal(f_status, g_status))
        run["sharding_identical_what"] = ("ssim, SN, status gathered from %d ranks == rank 0's own run of all %d "
                                          "pairs, byte for byte" % (world, total))
    if cold_s is not None:
        run["e2e_cold"] = {"seconds": cold_s, "pairs_per_s": n_local / cold_s, "h2d_bytes": res.coo_bytes + table.nbytes,
                           "what": "host COO -> H2D -> expected + O/E -> band + index arrays -> compare -> D2H, first "
                                   "call on this rank (includes CUDA module load)"}

    alg_bytes, pixels = res.alg_bytes(lo, hi)
    traffic, traffic_note = (None, "not measured (N > 1 or --no-traffic)")
    if want_traffic and world == 1 and not args.no_traffic:
        traffic, traffic_note = ncu_traffic(args)
    line = {
        "metric": METRIC, "value": all_pairs * steps / (total_ms * 1e-3), "unit": "pairs/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps, "higher_is_better": True,
        "scaling": w["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "run": run,
        "roofline": dict(roofline(alg_bytes, kern_ms, pixels, args.window, True,
                                  "compare (gather+mask+SSIM+SN), %.1f us per launch, %d pairs on this rank"
                                  % (kern_ms * 1e3, n_local), traffic), traffic_source=traffic_note),
        "e2e": {"value": all_pairs * steps / e2e_s, "unit": "pairs/s", "ms_per_step": 1e3 * e2e_s / steps,
                "h2d_bytes_per_step": int(table.nbytes), "d2h_bytes_per_step": int(n_local * 28),
                "call": ("every rank: its range of the host pair table -> GPU, chess_compare_band_batch; "
                         "all-gather of ssim (8 B per pair) + chess_zscore on every rank; "
                         "distributed.gather_tensors_to_rank0 (ssim, SN, status, z_ssim of all ranks in one NCCL "
                         "gather, 28 B per pair, then ONE copy into rank 0's page-locked host memory) -- all timed"
                         if strong else
                         "chess_sim_band_batch_host (host pair table in; ssim, SN, status, z_ssim out to host arrays; "
                         "up to 8192 pairs the kernels read / write the pinned staging buffer directly, larger "
                         "batches use one DMA each way)")},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    if world == 1 and want_cpu and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(res, args.window, g_ssim, g_sn)
    line["_resident"] = res           # for the extras that reuse the resident samples; popped before printing
    return line


def cpu_baseline(res, window, gpu_ssim, gpu_sn):
    """Oracle on the host cores, bounded sample (~10-30 s), same inputs as the GPU run; reports the largest
    relative deviation of the GPU's SSIM / SN from it."""
    import multiprocessing as mp
    genome, ref, qry, pairs = res.genome, res.ref, res.qry, res.pairs
    id_to_k = {p[0]: k for k, p in enumerate(pairs)}
    genome, ref, qry, pairs = cpu_subset(genome, ref, qry, pairs, 12000)
    oe = []      # the very O/E values the GPU used (restricted to the bins the CPU sample covers)
    for smp in (res.re_, res.qe):
        a, b, w = smp.triples()
        sel = b < genome.n_bins
        oe.append(list(zip(a[sel].tolist(), b[sel].tolist(), w[sel].tolist())))
    cpu_setup(genome, ref, qry, pairs, window, oe=oe)
    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(cores) as pool:
        probe = cpu_sample(len(pairs), cores)
        dt, _ = cpu_run(probe, pool, cores)
        target = int(max(cores, min(len(pairs), cores * 15.0 / max(dt, 1e-3))))
        idx = cpu_sample(len(pairs), target)
        dt, rows = cpu_run(idx, pool, cores)
    worst, compared = 0.0, 0
    for i, (pid, s, sn) in zip(idx, rows):
        k = id_to_k[pid]
        for got, want in ((gpu_ssim[k], s), (gpu_sn[k], sn)):
            if not (np.isnan(want) and np.isnan(got)):
                worst = max(worst, abs(got - want) / max(abs(want), 1e-300))
                compared += 1
    return {"value": len(idx) / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": "%d of the %d pairs inside the first %d bins, evenly spaced, one pass; %s"
                      % (len(idx), len(pairs), genome.n_bins, PORT_NOTE % cores),
            "seconds": dt, "max_rel_dev_gpu_vs_cpu": worst, "values_compared": compared}


def measure_dlbcl(args, R, w):
    """configs[3]: every pair (SN on, z_ssim) plus `background` background comparisons per pair (SN off) reduced to
    z_bg / p_bg on the device (bin/chess:259-332).  Pairs are cut into contiguous ranges over the ranks."""
    import torch
    from chess_b200 import _native
    from chess_b200._native import lib, check
    from chess_b200.engine import get_context, region_spans, CompareEngine
    from chess_b200.helpers import GenomicRegion
    from chess_b200.distributed import gather_to_rank0, my_shard
    dev, rank, world = R.dev, R.rank, R.world
    steps, warmup = args.steps, max(args.warmup, 3)
    res = Resident(w, R)
    total = len(res.pairs)
    lo, hi = my_shard(total, rank, world)
    n_local = hi - lo
    params = window_params(args.window)
    bed = background_bed(w, res.genome)
    bg = [GenomicRegion(chromosome=c, start=s, end=e, strand=None) for c, s, e in bed]   # load_regions keeps 0-based starts
    bf, bn = region_spans(res.trees, bg)
    bstrand = np.full(len(bg), 2, dtype=np.int32)                                        # None differs from '+' and '-'
    max_n = int(max(res.max_n, bn.max()))
    ctx = get_context(R.local_rank)
    sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p_sn = _native.make_params(compute_sn=True, kernel=args.kernel,
                               flags=_native.FLAG_SYMMETRIC | _native.FLAG_EQUAL_SIZES, **params)
    sizes = np.unique(np.concatenate([res.rn, bn]))
    p_bg = _native.make_params(compute_sn=False, kernel=args.kernel, flags=_native.FLAG_SYMMETRIC | (
        _native.FLAG_EQUAL_SIZES if len(sizes) == 1 else _native.FLAG_MOSTLY_EQUAL_SIZES), **params)
    table = res.table(lo, hi)
    d_table = torch.from_numpy(table.view(np.uint8)).to(dev)
    d_ssim = torch.empty(n_local, dtype=torch.float64, device=dev)
    d_sn = torch.empty(n_local, dtype=torch.float64, device=dev)
    d_status = torch.empty(n_local, dtype=torch.int32, device=dev)
    zeros = np.zeros(n_local, dtype=np.int32)
    up = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in
          (res.rf[lo:hi].astype(np.int64), res.rn[lo:hi].astype(np.int32), zeros, bf.astype(np.int64),
           bn.astype(np.int32), bstrand, res.qf[lo:hi].astype(np.int64), res.qn[lo:hi].astype(np.int32))]
    d_zbg = torch.empty(n_local, dtype=torch.float64, device=dev)
    d_pbg = torch.empty(n_local, dtype=torch.float64, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def device_step(e0=None, e1=None, e2=None):
        check(lib.chess_compare_band_batch(ctx.handle, C.byref(res.rsample), C.byref(res.qsample), _ptr(d_table),
                                           n_local, max_n, C.byref(p_sn), _ptr(d_ssim), _ptr(d_sn), _ptr(d_status), sp),
              ctx.handle, "chess_compare_band_batch")
        if e0 is not None:
            e0.record()
        check(lib.chess_background_batch(ctx.handle, C.byref(res.rsample), C.byref(res.qsample), _ptr(up[0]), _ptr(up[1]),
                                         _ptr(up[2]), _ptr(up[6]), _ptr(up[7]), _ptr(up[2]), _ptr(d_ssim), n_local,
                                         _ptr(up[3]), _ptr(up[4]), _ptr(up[5]), len(bf), max_n, C.byref(p_bg),
                                         _ptr(d_zbg), _ptr(d_pbg), None, sp), ctx.handle, "chess_background_batch")
        if e1 is not None:
            e1.record()

    eng = CompareEngine(res.re_, res.trees, res.qe, res.trees, devices=[R.local_rank])

    def host_step():
        # the public API: spans in host arrays -> ssim, SN, status, then z_bg / p_bg; gathered on rank 0
        ssim, sn, status = eng.compare_spans(res.rf[lo:hi], res.rn[lo:hi], res.qf[lo:hi], res.qn[lo:hi],
                                             res.flip[lo:hi], compute_sn=True, kernel=args.kernel, **params)
        z, pv = eng.background(res.rf[lo:hi], res.rn[lo:hi], zeros, ssim, bf, bn, bstrand,
                               own=(res.qf[lo:hi], res.qn[lo:hi], zeros), kernel=args.kernel, **params)
        got = gather_to_rank0((ssim, sn, status, z, pv), total, rank, world, R.dist, group=R.gloo)
        return got

    for _ in range(warmup):
        flush.zero_()
        device_step()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    sampler = ClockSampler(R.local_rank)
    R.barrier()
    sampler.start()
    launches0 = ctx.launches()
    for k in range(steps):
        flush.zero_()
        ev[k][2].record()
        device_step(ev[k][0], ev[k][1])
    R.barrier()
    launches = ctx.launches() - launches0
    total_ms = float(np.sum([e[2].elapsed_time(e[1]) for e in ev]))
    bg_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    host_step()
    R.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        got = host_step()
    R.barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop()
    total_ms, e2e_s, bg_ms = R.max(total_ms, e2e_s, bg_ms)
    z_dev, p_dev = d_zbg.cpu().numpy(), d_pbg.cpu().numpy()
    if rank != 0:
        return None
    n = float(max_n)
    n_cmp = n_local * len(bf)
    line = {
        "metric": METRIC, "value": total * steps / (total_ms * 1e-3), "unit": "pairs/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "run": {"comparisons_per_step": total * (len(bf) + 1),
                "background_comparisons_per_s": total * len(bf) * steps / (total_ms * 1e-3),
                "finite_z_bg": int(np.isfinite(got[3]).sum()), "pairs_ok": int((got[2] == 0).sum()),
                "device_call_matches_engine": bool(
                    np.allclose(z_dev, got[3][lo:hi], rtol=1e-9, atol=1e-12, equal_nan=True) and
                    np.array_equal(p_dev, got[4][lo:hi], equal_nan=True))},
        "roofline": roofline(n_cmp * (2.0 * n * n * 8 + 24), bg_ms, n_cmp * n * n, args.window, False,
                             "background batch (descriptor table + compare SN off + z_bg / p_bg reduction), %.2f ms "
                             "for %d comparisons on this rank" % (bg_ms, n_cmp)),
        "e2e": {"value": total * steps / e2e_s, "unit": "pairs/s", "ms_per_step": 1e3 * e2e_s / steps,
                "h2d_bytes_per_step": int(table.nbytes + n_local * 24 + len(bf) * 16),
                "d2h_bytes_per_step": int(n_local * 36),
                "call": "CompareEngine.compare_spans + CompareEngine.background (host spans in; ssim, SN, status, z_bg, "
                        "p_bg out), distributed.gather_to_rank0"},
        "gpu_launches": int(launches), "clocks": sampler.summary(),
    }
    return line


def measure_sweep(args, R, w):
    """configs[4]: one cell per (region size, window); the samples are generated once, the band is rebuilt per
    region size (its width follows the largest region of a run)."""
    import copy
    cells, launches, tot_pairs, tot_ms, tot_e2e, tot_bytes, tot_kern, tot_flops = [], 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0
    base = dict(workload_spec(args, max(SWEEP_WINDOW_BP)))
    res = Resident(base, R)
    from chess_b200.engine import region_spans
    S = synth_module()
    steps = min(args.steps, 5)
    clocks = None
    for b in SWEEP_WINDOW_BP:
        wc = workload_spec(args, b)
        res.w = wc
        res.pairs = S.sliding_pairs(res.genome, b, wc["step_bp"])
        res.rf, res.rn = region_spans(res.trees, [q[1] for q in res.pairs])
        res.qf, res.qn, res.flip = res.rf, res.rn, np.zeros(len(res.pairs), dtype=np.int32)
        res.max_n = int(res.rn.max())
        res.re_._bands.clear(), res.re_._index.clear(), res.qe._bands.clear(), res.qe._index.clear()
        from chess_b200.distributed import my_shard as _shard
        res.place(_shard(len(res.pairs), R.rank, R.world) if R.world > 1 else None)
        for win in [x for x in WINDOWS if x in args.sweep_windows.split(",")]:
            a = copy.copy(args)
            a.window = win
            line = measure_pairs(a, R, wc, res=res, steps=steps, warmup=3, want_cpu=False, want_traffic=False)
            if line is None:
                continue
            line.pop("_resident", None)
            clocks = line["clocks"]
            kern_us = float(line["roofline"]["kernel"].split(" us")[0].split(", ")[-1])
            cells.append({"bins": wc["bins_per_pair"], "window": win, "pairs": wc["pairs"],
                          "pairs_per_s": line["value"], "ms_per_step": line["ms_per_step"],
                          "kernel_us": kern_us, "roofline_frac": line["roofline"]["frac"],
                          "fp64_frac": line["roofline"]["secondary"]["frac"],
                          "e2e_pairs_per_s": line["e2e"]["value"],
                          "sharding_identical": line["run"].get("sharding_identical")})
            launches += line["gpu_launches"]
            tot_pairs += wc["pairs"]
            tot_ms += line["ms_per_step"]
            tot_e2e += wc["pairs"] / line["e2e"]["value"]
            tot_bytes += line["roofline"]["algorithmic_bytes_per_launch"]
            tot_kern += kern_us * 1e-3
            tot_flops += line["roofline"]["secondary"]["achieved"] * kern_us * 1e-3
    if R.rank != 0:
        return None
    peak, peak_src = measured_peak()
    achieved = tot_bytes / (tot_kern * 1e-3) / 1e9
    return {"metric": METRIC, "value": tot_pairs / (tot_ms * 1e-3), "unit": "pairs/s", "n_gpus": R.world,
            "steps": steps, "warmup": 3, "ms_per_step": tot_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
            "run": {"what": "a step = one pass over every cell; value = all cells' pairs / the sum of the cells' step times"},
            "cells": cells,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "compare kernels of all cells (rank 0), %.2f ms per pass" % tot_kern,
                         "algorithmic_bytes_per_launch": tot_bytes, "peak_source": peak_src,
                         "secondary": {"bound": "fp64", "achieved": tot_flops / tot_kern, "peak": FP64_PEAK_TFLOPS,
                                       "unit": "TFLOP/s", "frac": tot_flops / tot_kern / FP64_PEAK_TFLOPS}},
            "e2e": {"value": tot_pairs / tot_e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(40 * tot_pairs / R.world),
                    "d2h_bytes_per_step": int(28 * tot_pairs / R.world), "call": "per cell as for hg38_10kb"},
            "gpu_launches": int(launches), "clocks": clocks}


def measure_background(args, R, res, pair_ssim, n_sample=64):
    """`chess sim --background-query` (bin/chess:259-332) for a sample of the pairs: every
    query bin start x both strands as background regions, compute_sn off, reduced to z_bg / p_bg on the
    device.  Throughput in background comparisons per second: device-timed (arrays resident) and through
    CompareEngine.background (host arrays in, z_bg / p_bg out)."""
    import torch
    import chess_b200 as cb
    from chess_b200 import _native
    from chess_b200._native import lib, check
    from chess_b200.engine import region_spans, get_context
    from chess_b200.cli import _background_regions_from_query
    dev, local_rank = R.dev, R.local_rank
    pairs, params = res.pairs, window_params(args.window)
    step = max(1, len(pairs) // n_sample)
    pick = pairs[::step][:n_sample]
    idx = list(range(0, len(pairs), step))[:n_sample]
    t0 = time.perf_counter()
    regs = _background_regions_from_query(res.regions, pick[0][2], True)
    bf, bn = region_spans(res.trees, regs)
    bstrand = np.array([0 if r.strand == '+' else 1 for r in regs], dtype=np.int32)
    host_prep_s = time.perf_counter() - t0
    rf, rn = region_spans(res.trees, [p[1] for p in pick])
    zeros = np.zeros(len(pick), dtype=np.int32)
    s = np.asarray(pair_ssim)[idx]
    eng = cb.CompareEngine(res.re_, res.trees, res.qe, res.trees, devices=[local_rank])
    n_cmp = len(pick) * len(bf)
    for _ in range(2):
        z, pv = eng.background(rf, rn, zeros, s, bf, bn, bstrand, own=(rf, rn, zeros), **params)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    z, pv = eng.background(rf, rn, zeros, s, bf, bn, bstrand, own=(rf, rn, zeros), **params)
    e2e_s = time.perf_counter() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx = get_context(local_rank)
    p = _native.make_params(compute_sn=False, flags=_native.FLAG_SYMMETRIC | _native.FLAG_EQUAL_SIZES, **params)
    max_n = int(max(rn.max(), bn.max()))
    rw, rband, rsample = res.re_.sample_on(local_rank, max_n)
    qw, qband, qsample = res.qe.sample_on(local_rank, max_n)
    up = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in
          (rf.astype(np.int64), rn.astype(np.int32), zeros, s, bf.astype(np.int64), bn.astype(np.int32), bstrand)]
    d_z = torch.empty(len(pick), dtype=torch.float64, device=dev)
    d_p = torch.empty(len(pick), dtype=torch.float64, device=dev)
    sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def device_call():
        check(lib.chess_background_batch(ctx.handle, C.byref(rsample), C.byref(qsample), _ptr(up[0]), _ptr(up[1]),
                                         _ptr(up[2]), _ptr(up[0]), _ptr(up[1]), _ptr(up[2]), _ptr(up[3]), len(pick),
                                         _ptr(up[4]), _ptr(up[5]), _ptr(up[6]), len(bf), max_n, C.byref(p), _ptr(d_z),
                                         _ptr(d_p), None, sp), ctx.handle, "chess_background_batch")

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
    return {"what": "chess sim --background-query --limit-background on the chr2_25kb samples: %d pairs x %d "
                    "background regions (every query bin start, both strands), compute_sn off, z_bg / p_bg reduced "
                    "on the device" % (len(pick), len(bf)),
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
    from chess_b200.engine import get_context, PAIR_DTYPE, _stream_ptr, check
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
                                      mappability_cutoff=0.0, compute_sn=False, data_range=2.0,
                                      flags=_native.FLAG_SYMMETRIC | _native.FLAG_EQUAL_SIZES)
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
                    "(utils/run_synthetic_test.py:97-121); windows 7 and 9 on the upper-triangle kernels (symmetric "
                    "matrices), larger windows on the generic kernel" % (pool, n),
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


def run_ours(args, rank, world, local_rank):
    import copy
    R = Rank(rank, world, local_rank)
    w = workload_spec(args)
    if args.workload == "sweep":
        line = measure_sweep(args, R, w)
    elif w.get("background"):
        line = measure_dlbcl(args, R, w)
    else:
        line = measure_pairs(args, R, w)
        res = line.pop("_resident", None) if line else None
        if args.probe_kernel:
            R.close()
            return
        if res is not None:
            res.drop()
        if rank == 0 and world == 1 and args.workload == "hg38_10kb" and not args.no_extras:
            # the other BASELINE configs ride along on the default line (never lose the line over an extra)
            try:
                a = copy.copy(args)
                a.workload, a.bin_size, a.window_bp = "chr2_25kb", None, None
                extra = measure_pairs(a, R, workload_spec(a), steps=max(args.steps, 50), want_cpu=False)
                res2 = extra.pop("_resident")
                line["chr2_25kb"] = {k: extra[k] for k in ("value", "unit", "ms_per_step", "roofline", "e2e", "steps",
                                                           "warmup", "gpu_launches", "run")}
                line["chr2_25kb"]["workload"] = extra["config"]["workload"]
                for name, fn in (("background", lambda: measure_background(a, R, res2, _last_ssim(res2, a, R))),
                                 ("ingest", lambda: measure_ingest(local_rank, res2.ref)),
                                 ("synthetic_one_vs_many", lambda: measure_synthetic(local_rank))):
                    try:
                        line[name] = fn()
                    except Exception as exc:
                        line[name] = {"error": repr(exc)}
                res2.drop()
            except Exception as exc:
                line["chr2_25kb"] = {"error": repr(exc)}
    if rank == 0 and line is not None:
        line.pop("_resident", None)
        print(json.dumps(line), flush=True)
    R.close()


def _last_ssim(res, args, R):
    """ssim of every pair of a resident workload (one host-buffer call)."""
    import torch
    from chess_b200 import _native
    from chess_b200._native import lib, check
    from chess_b200.engine import get_context
    n = len(res.pairs)
    table = res.table()
    ssim, sn, st = np.empty(n), np.empty(n), np.empty(n, dtype=np.int32)
    p = _native.make_params(compute_sn=True, flags=_native.FLAG_SYMMETRIC | _native.FLAG_EQUAL_SIZES,
                            **window_params(args.window))
    ctx = get_context(R.local_rank)
    sp = C.c_void_p(torch.cuda.current_stream(R.dev).cuda_stream)
    check(lib.chess_compare_band_batch_host(ctx.handle, C.byref(res.rsample), C.byref(res.qsample), _np_ptr(table), n,
                                            res.max_n, C.byref(p), _np_ptr(ssim), _np_ptr(sn), _np_ptr(st), sp),
          ctx.handle, "chess_compare_band_batch_host")
    return ssim


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        synth_module(standalone=True)
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
