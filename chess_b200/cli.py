"""The `chess` command: global flags, dispatch, and the `sim` sub-command.

Interface-compatible with bin/chess of the reference (0.3.8) for `sim`: same
arguments, same log lines, same output table (bin/chess:80-359).  The body is
re-plumbed: instead of a multiprocessing pool fed through Manager queues, the
two samples are made resident on the GPUs once and all pairs -- and all
background comparisons -- go through the batched engine.
"""
import logging
import os
import sys

import numpy as np

from . import commands

logger = logging.getLogger('')

# stage -> seconds of the last `run_sim` (development aid, tools/cli_e2e.py prints it)
LAST_STAGES = {}
# the command-line entry points set this: a run may then hide GPUs it does not need from the CUDA driver
# (_limit_visible_gpus); programmatic calls of Chess(argv) leave the process environment alone
LIMIT_GPUS = False


class _StageClock(object):
    """Wall-clock split of run_sim; every stage ends with a device synchronisation so that asynchronous GPU work is
    charged to the stage that launched it."""

    def __init__(self):
        import time
        self._time = time.perf_counter
        self._t = self._time()
        LAST_STAGES.clear()

    def mark(self, name, sync=True):
        if sync:
            try:
                import torch
                from .engine import _contexts
                if torch.cuda.is_available():
                    for d in list(_contexts):        # only GPUs this run has touched (a sync would create a context)
                        torch.cuda.synchronize(d)
            except Exception:
                pass
        now = self._time()
        LAST_STAGES[name] = LAST_STAGES.get(name, 0.0) + (now - self._t)
        logger.debug("stage %s: %.3f s", name, now - self._t)
        self._t = now



class Chess(object):

    def __init__(self, argv=None):
        argv = sys.argv if argv is None else argv
        parser = commands.chess_parser()
        flag_increments = {'-l': 2, '--log-file': 2}
        option_ix = 1
        while option_ix < len(argv) and argv[option_ix].startswith('-'):
            option_ix += flag_increments.get(argv[option_ix], 1)
        args = parser.parse_args(argv[1:option_ix + 1])

        if args.verbosity == 1:
            log_level = logging.WARN
        elif args.verbosity > 2:
            log_level = logging.DEBUG
        else:
            log_level = logging.INFO
        logger.setLevel(log_level)
        formatter = logging.Formatter("%(asctime)s %(levelname)s %(message)s")
        if args.log_file is None:
            handler = logging.StreamHandler()
        else:
            handler = logging.FileHandler(os.path.expanduser(args.log_file), mode='a')
        handler.setFormatter(formatter)
        handler.setLevel(log_level)
        logger.addHandler(handler)

        if args.command is None or not hasattr(self, args.command) or args.command.startswith('_'):
            print('Unrecognized command')
            parser.print_help()
            sys.exit(1)

        logger.info("Running '{}'".format(" ".join(argv)))
        import chess_b200
        logger.info("CHESS version: {}".format(chess_b200.__version__))
        getattr(self, args.command)([argv[0]] + argv[option_ix:])
        logger.info("Finished '{}'".format(" ".join(argv)))

    def sim(self, argv):
        args = commands.sim_parser().parse_args(argv[2:])
        run_sim(args)

    def oe(self, argv):
        """bin/chess:360-380: O/E of a sparse-format matrix, values written with six decimals."""
        args = commands.oe_parser().parse_args(argv[2:])
        import gzip
        from .oe import observed_expected
        from .helpers import is_gzipped
        _, edges = observed_expected(args.regions, args.input_matrix)
        text = "".join("{}\t{}\t{:.6e}\n".format(a, b, w) for a, b, w in edges)
        if is_gzipped(args.output_matrix):
            with gzip.open(args.output_matrix, 'w') as o:
                o.write(text.encode('utf-8'))
        else:
            with open(args.output_matrix, 'w') as o:
                o.write(text)

    @staticmethod
    def _checked_output(path):
        output_dir = os.path.dirname(path)
        if output_dir != '' and not os.path.exists(output_dir):
            logger.error("Specified output directory does not exist!")
            raise RuntimeError("Specified output path not found.")
        return path

    @staticmethod
    def _ucsc_sizes(genome):
        """pybedtools' chromosome-size table if that package is there (bin/chess:414-432)."""
        try:
            import pybedtools as pbt  # third-party, optional
        except ImportError:
            raise ValueError("pybedtools is not installed")
        sizes = pbt.chromsizes(genome)
        if len(sizes) == 0:
            raise ValueError("Genome not recognised in pybedtools.chromsizes: {}".format(genome))
        return {k: v[1] for k, v in sizes.items()}

    def pairs(self, argv):
        """bin/chess:382-458."""
        args = commands.pairs_parser().parse_args(argv[2:])
        from . import emit
        self._checked_output(args.output)
        if args.file_input:
            sizes = emit.load_chromosome_sizes(args.genome)
        else:
            try:
                sizes = self._ucsc_sizes(args.genome)
            except (OSError, ValueError):
                logger.info('No entry found with pybedtools. Trying to read from file.')
                sizes = emit.load_chromosome_sizes(args.genome)
        emit.write_pairs(args.output, sizes, args.window, args.step, args.chromosome)

    def background(self, argv):
        """bin/chess:460-518."""
        args = commands.background_parser().parse_args(argv[2:])
        from . import emit
        from .helpers import read_chromosome_sizes, GenomicRegion
        self._checked_output(args.output)
        if args.strand not in ['-', '+', None]:
            raise ValueError("--strand must be either - or +")
        try:
            sizes = self._ucsc_sizes(args.genome_or_region)
            regions = [(c, 1, end) for c, end in sizes.items()]
        except ValueError:
            logger.info('No entry found with pybedtools. Trying to read from file.')
            if os.path.exists(args.genome_or_region):
                sizes = read_chromosome_sizes(args.genome_or_region)
                regions = [(c, 1, end) for c, end in sizes.items()]
            else:
                region = GenomicRegion.from_string(args.genome_or_region)
                regions = [(region.chromosome, region.start, region.end)]
        emit.write_background(args.output, regions, args.window, args.step, args.strand)


def _zscores(ssim):
    """z_ssim over the run; bin/chess:216-233 (nanmean / nanstd, ddof 0)."""
    import warnings
    ssim = np.asarray(ssim, dtype=np.float64)
    valid = ssim[~np.isnan(ssim)]
    with warnings.catch_warnings(), np.errstate(invalid="ignore", divide="ignore"):
        warnings.simplefilter("ignore")
        mean, sd = np.nanmean(valid), np.nanstd(valid)
        z = np.full(len(ssim), np.nan)
        fin = np.isfinite(ssim)
        z[fin] = (ssim[fin] - mean) / sd
    return z


def _background_regions_from_query(query_regions, query_region, limit_background):
    """Every query bin start x {+,-} with the size of `query_region`; bin/chess:272-298."""
    from .helpers import GenomicRegion
    from collections import defaultdict
    ends = defaultdict(int)
    for r in query_regions:
        ends[r.chromosome] = max(ends[r.chromosome], r.end)
    size = query_region.end - query_region.start
    out = []
    for region in query_regions:
        if limit_background and region.chromosome != query_region.chromosome:
            continue
        if region.start + size <= ends[region.chromosome]:
            out.append(GenomicRegion(chromosome=region.chromosome, start=region.start,
                                     end=region.start + size, strand='+'))
            out.append(GenomicRegion(chromosome=region.chromosome, start=region.start,
                                     end=region.start + size, strand='-'))
    return out


def _float_column(values):
    """Every value as '{}'.format(np.float64(v)) writes it (numpy prints float64 like Python prints float: shortest
    round-trip digits, exponent form below 1e-4 and from 1e16; tests/test_host_logic.py compares the two)."""
    return list(map(repr, np.asarray(values, dtype=np.float64).tolist()))


def _background_spans_from_query(query_regions, query_trees, query_region, limit_background):
    """bin/chess:272-298 as columns: every query bin start x {+, -} with the size of `query_region`, kept where it
    ends inside its chromosome, optionally on the pair's chromosome only -> (first_bin, n_bins, strand code)
    of every background region, in the reference's order.  No region objects are made."""
    from .engine import region_spans
    from .helpers import RegionTable
    if not isinstance(query_regions, RegionTable):
        regs = _background_regions_from_query(query_regions, query_region, limit_background)
        bf, bn = region_spans(query_trees, regs)
        return bf, bn, np.array([0 if r.strand == '+' else 1 for r in regs], dtype=np.int32)
    t = query_regions
    size = query_region.end - query_region.start
    ends = np.zeros(len(t.chrom_names), dtype=np.int64)
    np.maximum.at(ends, t.chrom_codes, t.ends)
    keep = t.starts + size <= ends[t.chrom_codes]
    if limit_background:
        if query_region.chromosome not in t.chrom_names:
            keep[:] = False
        else:
            keep &= t.chrom_codes == t.chrom_names.index(query_region.chromosome)
    rows = np.flatnonzero(keep)
    first = np.full(len(rows), -1, dtype=np.int64)
    count = np.zeros(len(rows), dtype=np.int64)
    codes = t.chrom_codes[rows]
    for c in np.unique(codes):
        sel = np.flatnonzero(codes == c)
        starts = t.starts[rows[sel]]
        lo, hi = query_trees[t.chrom_names[int(c)]].span(starts - 1, starts + size)   # region [start, start + size]
        ok = lo >= 0
        first[sel[ok]] = lo[ok]
        count[sel[ok]] = hi[ok] - lo[ok] + 1
    return (np.repeat(first, 2), np.repeat(count, 2).astype(np.int64),
            np.tile(np.array([0, 1], dtype=np.int32), len(rows)))


def _background_set_from_query(query_regions, query_trees, query_region, limit_background):
    """The same set, enumerated and looked up on the GPU (engine.QueryBackgroundSet) when the regions are a regular
    RegionTable; otherwise the host columns above.  Returns what CompareEngine.background takes as
    (bg_first, bg_n, bg_strand)."""
    from .engine import QueryBackgroundSet
    from .helpers import RegionTable
    if isinstance(query_regions, RegionTable):
        bg = QueryBackgroundSet(query_regions, query_region.end - query_region.start,
                                query_region.chromosome if limit_background else None)
        if bg.ok:
            return bg, None, None
    return _background_spans_from_query(query_regions, query_trees, query_region, limit_background)


def _count_lines(path):
    with open(os.path.expanduser(path), "rb") as fh:
        return sum(chunk.count(b"\n") for chunk in iter(lambda: fh.read(1 << 22), b""))


def _estimate_work(args):
    """Comparisons x bins^2 of a run from the first lines and the line counts of its input files -- host only,
    before CUDA starts.  None when the files do not say (matrix formats that carry their own regions)."""
    try:
        if not args.reference_regions or not args.query_regions:
            return None
        with open(os.path.expanduser(args.pairs)) as fh:
            first = fh.readline().split()
        span = int(first[2]) - int(first[1])
        with open(os.path.expanduser(args.reference_regions)) as fh:
            fields = fh.readline().split()
        bins = span // max(int(fields[2]) - int(fields[1]), 1) + 1
        n_pairs = _count_lines(args.pairs)
        per_pair = 1.0
        if args.background_query:
            n_query_bins = _count_lines(args.query_regions)
            per_pair += 2.0 * (n_query_bins / 24.0 if args.limit_background else n_query_bins)
        elif args.background_regions is not None:
            per_pair += _count_lines(args.background_regions)
        return float(n_pairs) * per_pair * bins * bins
    except (OSError, ValueError, IndexError):
        return None


def _limit_visible_gpus(args):
    """On a multi-GPU box the CUDA driver initialises every visible GPU when the first one is used (6.7 s of a
    9.9 s first run on eight B200s, whatever the number of contexts created afterwards).  A run whose work is
    worth fewer GPUs (engine.plan_devices) hides the others from the driver BEFORE CUDA starts in this process.
    A CUDA_VISIBLE_DEVICES list is narrowed to its first entries; nothing happens when the GPUs were named with
    CHESS_B200_DEVICES or CUDA is up already."""
    if not LIMIT_GPUS or os.environ.get("CHESS_B200_DEVICES"):
        return
    import torch
    if torch.cuda.is_initialized():
        return
    work = _estimate_work(args)
    if work is None:
        return
    chosen = os.environ.get("CUDA_VISIBLE_DEVICES")
    if chosen is not None:                       # the user's set is kept as the superset: its first n entries
        ids = [v for v in chosen.split(",") if v.strip() != ""]
    else:
        try:
            import pynvml
            pynvml.nvmlInit()
            ids = [str(d) for d in range(pynvml.nvmlDeviceGetCount())]
        except Exception:
            return
    from .engine import plan_devices
    n = len(plan_devices(list(range(len(ids))), work))
    if n < len(ids):
        os.environ["CUDA_VISIBLE_DEVICES"] = ",".join(ids[:n])
        logger.info("Using {} of {} GPUs for this run (CHESS_B200_DEVICES chooses them explicitly)".format(n, len(ids)))


def run_sim(args):
    """Body of `chess sim` (bin/chess:80-359) on the batched engine."""
    from .engine import CompareEngine, region_spans, visible_devices
    from .helpers import load_pairs_table, load_regions, load_contacts, load_oe_contacts, RegionTable
    _limit_visible_gpus(args)

    output_file = args.out
    output_dir = os.path.dirname(output_file)
    if output_dir == '':
        output_file = os.path.join('.', output_file)
    elif not os.path.exists(output_dir):
        logger.error("Specified output directory does not exist!")
        raise RuntimeError("Specified output path not found.")
    logger.debug('Parameters:')
    logger.debug(args)
    clock = _StageClock()

    loader = load_oe_contacts if args.oe_input else load_contacts
    all_devices = visible_devices()
    devices = all_devices[:2]          # the two samples are loaded side by side; more GPUs join when the work
                                       # is worth their context (engine.plan_devices, below)

    def start_contexts(devs):
        # CUDA contexts cost about a second each: start them side by side
        from concurrent.futures import ThreadPoolExecutor
        from .engine import get_context
        import torch
        with ThreadPoolExecutor(max(len(devs), 1)) as pool:
            list(pool.map(lambda d: (torch.cuda.synchronize(d), get_context(d)), devs))

    if len(devices) > 1:
        start_contexts(devices)
        clock.mark("create GPU contexts", sync=False)
    query_job = None
    if len(devices) > 1:
        # two samples, two GPUs: the query is read, parsed and normalised on the second one meanwhile
        from concurrent.futures import ThreadPoolExecutor
        query_job = ThreadPoolExecutor(1).submit(loader, args.query_contacts, args.query_regions, devices[1])
    logger.info('Loading reference contact data')
    try:
        reference_edges, reference_trees, reference_regions = loader(
            args.reference_contacts, args.reference_regions, devices[0])
    except ValueError as error:
        logger.error(error)
        logger.error("Reference contact data could not be loaded. Please specify a valid input "
                     "file. Files in sparse format can only be loaded if --reference-regions is "
                     "specified.")
        sys.exit()
    clock.mark("load reference contacts (+O/E)")
    logger.info('Loading query contact data')
    try:
        if query_job is not None:
            query_edges, query_trees, query_regions = query_job.result()
        else:
            query_edges, query_trees, query_regions = loader(args.query_contacts, args.query_regions, devices[0])
    except ValueError as error:
        logger.error(error)
        logger.error("Query contact data could not be loaded. Please specify a valid input file. "
                     "Files in sparse format can only be loaded if --query-regions is specified.")
        sys.exit()

    clock.mark("load query contacts (+O/E)")
    def bin_size(regions):
        return regions.max_bin_size() if isinstance(regions, RegionTable) else max(r.end - r.start + 1 for r in regions)

    reference_bin_size = bin_size(reference_regions)
    query_bin_size = bin_size(query_regions)

    logger.info('Loading region pairs')
    pairs, dropped = load_pairs_table(
        args.pairs, set(reference_trees.keys()) | set(query_trees.keys()))
    if dropped > 0:
        logger.warning("{} region pairs have been dropped, because they involve chromosomes that "
                       "are not present in the provided contact data.".format(dropped))
    if len(pairs) == 0:
        logger.error("No valid region pairs found; aborting.")
        sys.exit()

    ref_bp, qry_bp = pairs.max_spans_bp()
    max_ref_span = int(ref_bp / reference_bin_size)
    max_qry_span = int(qry_bp / query_bin_size)
    for span, what, bin_size in ((max_ref_span, "reference", reference_bin_size),
                                 (max_qry_span, "query", query_bin_size)):
        if span < 20:  # bin/chess:169-187
            logger.error("All regions need to span at least 20 bins. The provided {} regions span "
                         "at most {} bins. Please try again with larger regions or a smaller bin "
                         "size. The bin size of the input data has been detected to be {} bp."
                         .format(what, span, bin_size))
            sys.exit()

    clock.mark("load pairs, bin-size guard")
    # GPUs for the comparisons: as many as the work is worth (comparisons x bins^2 against a second per context)
    from .engine import plan_devices
    bins_ref = max(ref_bp // max(reference_bin_size - 1, 1), 1) + 1
    work = float(len(pairs)) * bins_ref * bins_ref
    if args.background_query:
        n_bins_query = len(query_regions)
        per_pair = 2.0 * (n_bins_query / max(len(query_trees), 1) if args.limit_background else n_bins_query)
        work += float(len(pairs)) * per_pair * bins_ref * bins_ref
    elif args.background_regions is not None:
        with open(os.path.expanduser(args.background_regions)) as fh:
            work += float(len(pairs)) * sum(1 for _ in fh) * bins_ref * bins_ref
    planned = plan_devices(all_devices, work)
    if len(planned) > len(devices):
        start_contexts(planned[len(devices):])
        clock.mark("create GPU contexts", sync=False)
    devices = planned
    logger.info("Launching comparison engine on GPU(s) {}".format(devices))
    engine = CompareEngine(reference_edges, reference_trees, query_edges, query_trees, devices=devices)
    params = dict(min_bins=20, keep_unmappable_bins=args.keep_unmappable_bins,
                  absolute_window_size=args.absolute_windowsize,
                  relative_window_size=args.relative_windowsize,
                  mappability_cutoff=args.mappability_cutoff)

    logger.info("Submitting pairs for comparison")
    ids, ssim, sn, status = engine.compare(pairs, compute_sn=True, **params)
    clock.mark("compare (spans, bands, kernels, gather)")
    z_ssim = _zscores(ssim)
    clock.mark("z-scores")

    background = args.background_regions is not None or args.background_query
    z_bg = p_bg = None
    if background:
        logger.info("Running background calculations")
        if args.background_regions is not None:
            background_regions, _, _ = load_regions(os.path.expanduser(args.background_regions), ignore_ix=True)
        else:
            background_regions = None
            logger.info("Generating background regions from query")
        z_bg = np.full(len(pairs), np.nan)
        p_bg = np.full(len(pairs), np.nan)
        rf, rn, qf, qn = engine.last_spans
        # strands as integer codes: a comparison is flipped iff the codes differ (sim.py:332-333)
        codes = {'+': 0, '-': 1, None: 2}

        def strand_code(strand):
            if strand not in codes:
                codes[strand] = len(codes)
            return codes[strand]

        ref_strand = np.array([strand_code(v) for v in pairs.ref[3]], dtype=np.int32)
        qry_strand = np.array([strand_code(v) for v in pairs.qry[3]], dtype=np.int32)
        # pairs that share a background set (the BED, or query size [+ chromosome]) go to the device
        # together: one chess_background_batch per set, nothing but z_bg / p_bg comes back
        if background_regions is not None:
            groups = {"bed": np.arange(len(pairs))}
        else:
            import pandas as pd
            size = pairs.qry[2] - pairs.qry[1]
            keys = pd.DataFrame({"c": pairs.qry[0] if args.limit_background else np.zeros(len(pairs), dtype=object),
                                 "s": size})
            groups = {key: np.asarray(ix) for key, ix in keys.groupby(["c", "s"], sort=False).indices.items()}
        for key, ks in groups.items():
            if background_regions is not None:
                bf, bn = region_spans(query_trees, background_regions)
                bstrand = np.array([strand_code(r.strand) for r in background_regions], dtype=np.int32)
            else:
                bf, bn, bstrand = _background_set_from_query(query_regions, query_trees, pairs[int(ks[0])][2],
                                                             args.limit_background)
            z, pv = engine.background(rf[ks], rn[ks], ref_strand[ks], np.asarray(ssim)[ks], bf, bn, bstrand,
                                      own=(qf[ks], qn[ks], qry_strand[ks]), **params)
            z_bg[ks] = z
            p_bg[ks] = pv

    if background:
        clock.mark("background")
    # rows in pairs-file order; numbers as '{}'.format(np.float64) prints them (shortest round-trip repr, 'nan');
    # soft failures carry (nan, nan, nan), successful pairs their SN (bin/chess:236-256, 334-353)
    ssim = np.asarray(ssim, dtype=np.float64)
    failed = np.isnan(ssim)
    nan_counter = int(failed.sum())
    cols = [np.where(failed, np.nan, np.asarray(sn, dtype=np.float64)), ssim, np.asarray(z_ssim, dtype=np.float64)]
    if background:
        cols += [np.asarray(z_bg, dtype=np.float64), np.asarray(p_bg, dtype=np.float64)]
    text = [[str(i) for i in ids]] + [_float_column(c) for c in cols]
    with open(output_file, 'w') as o:
        o.write("ID\tSN\tssim\tz_ssim\tz_bg\tp_bg\n" if background else "ID\tSN\tssim\tz_ssim\n")
        if len(ids):
            o.write("\n".join(map("\t".join, zip(*text))))
            o.write("\n")
    clock.mark("write table")
    if nan_counter > 0:
        logger.info("Could not compute similarity for {}/{} region pairs. This can be due to faulty "
                    "coordinates, too small region sizes or too many unmappable bins".format(
                        nan_counter, len(pairs)))


def main(argv=None):
    global LIMIT_GPUS
    LIMIT_GPUS = True
    Chess(argv)
