"""Argument parsers of the `chess` command line.

The `sim` sub-command keeps the reference's contract exactly -- positionals,
flag names, dests, types and defaults of chess/commands.py:48-185 -- because
scripts are written against it.  Engine knobs (which GPUs, kernel choice) are
deliberately NOT flags: they come from environment variables
(CHESS_B200_DEVICES) so the command line stays interchangeable.
"""
import argparse
import sys
import textwrap


class MyParser(argparse.ArgumentParser):
    """Print the full help on a usage error (chess/commands.py:428-432)."""

    def error(self, message):
        sys.stderr.write('error: %s\n' % message)
        self.print_help()
        sys.exit(2)


def chess_parser():
    """Global flags and the sub-command name; chess/commands.py:6-45."""
    usage = '''\
        chess <command> [options]

        Commands:
            sim             Calculate structural similarity (B200 engine)
            oe              Observed/expected transform of a sparse-format matrix
            pairs           Sliding-window region pairs for a genome scan (BEDPE)
            background      Sliding-window background regions (BED)

        Run chess <command> -h for help on a specific command.
        '''
    parser = argparse.ArgumentParser(
        description='CHESS: Compare Hi-C Experiments using Structural Similarity',
        usage=textwrap.dedent(usage))
    from .version import __version__
    parser.add_argument('--version', action='version', version='CHESS {} (chess_b200)'.format(__version__))
    parser.add_argument('--verbose', '-v', dest='verbosity', action='count', default=0,
                        help='Verbosity: -v errors and warnings, -vv adds info (the default), '
                             '-vvv adds debug messages.')
    parser.add_argument('-l', '--log-file', dest='log_file', help='Append the log to this file.')
    parser.add_argument('command', nargs='?', help='Subcommand to run')
    return parser


def sim_parser():
    """`chess sim`; flags, dests, types and defaults as chess/commands.py:48-185."""
    parser = MyParser(
        description='Compare pairs of regions between two balanced chromatin contact matrices '
                    'with the structural similarity index (on observed/expected values) and '
                    'report a signal-to-noise score and z-scores. With --background-query or '
                    '--background-regions every reference region is also compared against a '
                    'background set, giving z_bg and p_bg. Pass --oe-input if the matrices '
                    'already hold observed/expected values.',
        formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('reference_contacts', type=str,
                        help='Balanced reference contact matrix. Sparse text format '
                             '(<row bin> <column bin> <value> per line) together with '
                             '--reference-regions; formats that need FAN-C (.hic, .cool, .mcool) '
                             'only if fanc is installed.')
    parser.add_argument('--reference-regions', dest="reference_regions", type=str,
                        help='BED file (no header), one line per bin of the reference matrix.')
    parser.add_argument('query_contacts', type=str,
                        help='Balanced query contact matrix, same formats as the reference.')
    parser.add_argument('--query-regions', dest="query_regions", type=str,
                        help='BED file (no header), one line per bin of the query matrix.')
    parser.add_argument('pairs', type=str,
                        help='Region pairs in 10-column BEDPE: chrom1/start1/end1 = reference, '
                             'chrom2/start2/end2 = query, then ID, score (ignored), strand1, strand2.')
    parser.add_argument('out', type=str, help='Path to outfile.')
    parser.add_argument('--background-regions', dest='background_regions', type=str,
                        help='BED file of background regions; enables z_bg and p_bg.')
    parser.add_argument('--background-query', dest='background_query', default=False,
                        action='store_true',
                        help='Use every same-sized region of the query genome as background.')
    parser.add_argument('-p', dest='threads', type=int, default=1,
                        help='Kept for compatibility (CPU worker count in the reference); the '
                             'GPUs to use are taken from CHESS_B200_DEVICES.')
    parser.add_argument('--keep-unmappable-bins', action='store_true', default=False,
                        help='Disable deletion of unmappable bins.')
    parser.add_argument('--mappability-cutoff', type=float, default=0.1,
                        help='Largest tolerated fraction of unmappable bins in a matrix; matrices '
                             'above it are not compared, below it the bins are deleted.')
    parser.add_argument('-r', '--relative-windowsize', type=float, default=1,
                        help='SSIM window as a fraction of the matrix size.')
    parser.add_argument('-a', '--absolute-windowsize', type=int, default=None,
                        help='SSIM window in bins. Overwrites -r.')
    parser.add_argument('--oe-input', action='store_true', default=False,
                        help='Input contacts are already observed/expected transformed.')
    parser.add_argument('--limit-background', action='store_true', default=False,
                        help='Restrict --background-query to the chromosome of the paired query region.')
    return parser


def oe_parser():
    """`chess oe`; positionals as chess/commands.py:188-212."""
    parser = MyParser(description='Observed/expected transform of a balanced matrix in sparse format '
                                  '(expected values per chromosome and distance, computed on the GPU).',
                      formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('input_matrix', type=str, help='Balanced matrix, sparse text format.')
    parser.add_argument('regions', type=str, help='BED file (no header), one line per matrix bin.')
    parser.add_argument('output_matrix', type=str, help='Output matrix, same format (gzipped if it ends in .gz).')
    return parser


def pairs_parser():
    """`chess pairs`; positionals and flags as chess/commands.py:215-262."""
    parser = MyParser(description='Write every position of a sliding window over a genome as a region pair '
                                  '(reference = query) in the 10-column BEDPE format `chess sim` reads.',
                      formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('genome', type=str,
                        help='Tab- or space-separated chromosome sizes file (<name> <size>). A UCSC genome name '
                             'is looked up with pybedtools only if that package is installed.')
    parser.add_argument('window', type=int, help='Window size in base pairs')
    parser.add_argument('step', type=int, help='Step size in base pairs')
    parser.add_argument('output', type=str, help='Path to output file')
    parser.add_argument('--file-input', action='store_true', default=False,
                        help='Treat `genome` as a file without trying pybedtools first.')
    parser.add_argument('--chromosome', type=str, help='Write pairs for this chromosome only.')
    return parser


def background_parser():
    """`chess background`; positionals and flags as chess/commands.py:265-301."""
    parser = MyParser(description='Write sliding-window regions for the background model of `chess sim` (BED).',
                      formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('genome_or_region', type=str,
                        help='Chromosome sizes file (<name><TAB><size>), or a region <chromosome>:<start>-<end>; '
                             'a UCSC genome name only if pybedtools is installed.')
    parser.add_argument('window', type=int, help='Window size in base pairs')
    parser.add_argument('step', type=int, help='Step size in base pairs')
    parser.add_argument('output', type=str, help='Path to output file')
    parser.add_argument('--strand', type=str, help='[+/-] Regions on this strand only. Default: both strands')
    return parser
