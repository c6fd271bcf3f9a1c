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
    valid = [s for s in ssim if not np.isnan(s)]
    with warnings.catch_warnings(), np.errstate(invalid="ignore", divide="ignore"):
        warnings.simplefilter("ignore")
        mean, sd = np.nanmean(valid), np.nanstd(valid)
        z = np.full(len(ssim), np.nan)
        fin = np.isfinite(ssim)
        z[fin] = (np.asarray(ssim)[fin] - mean) / sd
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


def run_sim(args):
    """Body of `chess sim` (bin/chess:80-359) on the batched engine."""
    from .engine import CompareEngine, region_spans, visible_devices
    from .helpers import load_pairs_for_chroms, load_regions, load_contacts, load_oe_contacts

    output_file = args.out
    output_dir = os.path.dirname(output_file)
    if output_dir == '':
        output_file = os.path.join('.', output_file)
    elif not os.path.exists(output_dir):
        logger.error("Specified output directory does not exist!")
        raise RuntimeError("Specified output path not found.")
    logger.debug('Parameters:')
    logger.debug(args)

    loader = load_oe_contacts if args.oe_input else load_contacts
    logger.info('Loading reference contact data')
    try:
        reference_edges, reference_trees, reference_regions = loader(
            args.reference_contacts, args.reference_regions)
    except ValueError as error:
        logger.error(error)
        logger.error("Reference contact data could not be loaded. Please specify a valid input "
                     "file. Files in sparse format can only be loaded if --reference-regions is "
                     "specified.")
        sys.exit()
    logger.info('Loading query contact data')
    try:
        query_edges, query_trees, query_regions = loader(args.query_contacts, args.query_regions)
    except ValueError as error:
        logger.error(error)
        logger.error("Query contact data could not be loaded. Please specify a valid input file. "
                     "Files in sparse format can only be loaded if --query-regions is specified.")
        sys.exit()

    reference_bin_size = max(r.end - r.start + 1 for r in reference_regions)
    query_bin_size = max(r.end - r.start + 1 for r in query_regions)

    logger.info('Loading region pairs')
    pairs, dropped = load_pairs_for_chroms(
        args.pairs, set(reference_trees.keys()) | set(query_trees.keys()))
    if dropped > 0:
        logger.warning("{} region pairs have been dropped, because they involve chromosomes that "
                       "are not present in the provided contact data.".format(dropped))
    if len(pairs) == 0:
        logger.error("No valid region pairs found; aborting.")
        sys.exit()

    max_ref_span = int(max(p[1].end - p[1].start + 1 for p in pairs) / reference_bin_size)
    max_qry_span = int(max(p[2].end - p[2].start + 1 for p in pairs) / query_bin_size)
    for span, what, bin_size in ((max_ref_span, "reference", reference_bin_size),
                                 (max_qry_span, "query", query_bin_size)):
        if span < 20:  # bin/chess:169-187
            logger.error("All regions need to span at least 20 bins. The provided {} regions span "
                         "at most {} bins. Please try again with larger regions or a smaller bin "
                         "size. The bin size of the input data has been detected to be {} bp."
                         .format(what, span, bin_size))
            sys.exit()

    devices = visible_devices()
    logger.info("Launching comparison engine on GPU(s) {}".format(devices))
    engine = CompareEngine(reference_edges, reference_trees, query_edges, query_trees, devices=devices)
    params = dict(min_bins=20, keep_unmappable_bins=args.keep_unmappable_bins,
                  absolute_window_size=args.absolute_windowsize,
                  relative_window_size=args.relative_windowsize,
                  mappability_cutoff=args.mappability_cutoff)

    logger.info("Submitting pairs for comparison")
    ids, ssim, sn, status = engine.compare(pairs, compute_sn=True, **params)
    z_ssim = _zscores(ssim)

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
        rf, rn = region_spans(reference_trees, [p[1] for p in pairs])
        # strands as integer codes: a comparison is flipped iff the codes differ (sim.py:332-333)
        codes = {'+': 0, '-': 1, None: 2}

        def strand_code(strand):
            if strand not in codes:
                codes[strand] = len(codes)
            return codes[strand]

        ref_strand = np.array([strand_code(p[1].strand) for p in pairs], dtype=np.int32)
        qry_strand = np.array([strand_code(p[2].strand) for p in pairs], dtype=np.int32)
        qf, qn = region_spans(query_trees, [p[2] for p in pairs])
        # pairs that share a background set (the BED, or query size [+ chromosome]) go to the device
        # together: one chess_background_batch per set, nothing but z_bg / p_bg comes back
        groups = {}
        for k, pair in enumerate(pairs):
            if background_regions is not None:
                key = "bed"
            else:
                q = pair[2]
                key = (q.chromosome if args.limit_background else None, q.end - q.start)
            groups.setdefault(key, []).append(k)
        for key, ks in groups.items():
            if background_regions is not None:
                regs = background_regions
            else:
                regs = _background_regions_from_query(query_regions, pairs[ks[0]][2], args.limit_background)
            bf, bn = region_spans(query_trees, regs)
            bstrand = np.array([strand_code(r.strand) for r in regs], dtype=np.int32)
            ks = np.asarray(ks)
            z, pv = engine.background(rf[ks], rn[ks], ref_strand[ks], np.asarray(ssim)[ks], bf, bn, bstrand,
                                      own=(qf[ks], qn[ks], qry_strand[ks]), **params)
            z_bg[ks] = z
            p_bg[ks] = pv

    nan_counter = 0
    with open(output_file, 'w') as o:
        o.write("ID\tSN\tssim\tz_ssim\tz_bg\tp_bg\n" if background else "ID\tSN\tssim\tz_ssim\n")
        for k, pair_id in enumerate(ids):
            s = np.float64(ssim[k])
            # soft failures carry (nan, nan, nan); successful pairs a Python float SN (sim.py:30)
            row_sn = np.nan if np.isnan(s) else float(sn[k])
            if np.isnan(s):
                nan_counter += 1
            if background:
                o.write("{}\t{}\t{}\t{}\t{}\t{}\n".format(pair_id, row_sn, s, np.float64(z_ssim[k]),
                                                        np.float64(z_bg[k]), np.float64(p_bg[k])))
            else:
                o.write("{}\t{}\t{}\t{}\n".format(pair_id, row_sn, s, np.float64(z_ssim[k])))
    if nan_counter > 0:
        logger.info("Could not compute similarity for {}/{} region pairs. This can be due to faulty "
                    "coordinates, too small region sizes or too many unmappable bins".format(
                        nan_counter, len(pairs)))


def main(argv=None):
    Chess(argv)
