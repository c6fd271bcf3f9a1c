"""Sliding-window generators behind `chess pairs` and `chess background`
(bin/chess:382-518), as arrays.

The reference writes one text line per window in a Python loop and `chess sim`
parses the file again; these functions produce the same windows as numpy
arrays, so a genome scan can go from chromosome sizes to the comparison engine
without the text round trip (`window_pair_tuples` gives the (ID, region,
region) tuples every `chess_b200.sim` entry point accepts).  The writers below
emit byte-identical files to the reference's.
"""
from __future__ import absolute_import

import os

import numpy as np


def load_chromosome_sizes(path):
    """Whitespace-separated <chromosome> <size> lines, blank lines skipped; bin/chess:396-412."""
    sizes = {}
    with open(os.path.expanduser(path), "r") as fh:
        for line in fh:
            if line == "\n":
                continue
            fields = line.strip().split()
            sizes[fields[0]] = int(fields[1])
    return sizes


def window_pairs(chromosome_sizes, window, step, chromosome=None):
    """(chromosome names, chromosome index, start, end, pair id) of `chess pairs`.

    Every position `start = 0, step, 2*step, ...` with `start <= size - window`, chromosomes in
    the order of `chromosome_sizes`, ids running over the whole genome (bin/chess:444-458).
    """
    names = list(chromosome_sizes.keys())
    if chromosome:
        if chromosome not in chromosome_sizes:
            raise ValueError('Specified chromosome has no chromosome size entry.')
        names = [chromosome]
    if step <= 0:
        raise ValueError("step must be positive")
    chrom_ix, starts = [], []
    for k, name in enumerate(names):
        last = chromosome_sizes[name] - window
        if last < 0:
            continue
        s = np.arange(0, last + 1, step, dtype=np.int64)
        starts.append(s)
        chrom_ix.append(np.full(len(s), k, dtype=np.int64))
    if starts:
        starts = np.concatenate(starts)
        chrom_ix = np.concatenate(chrom_ix)
    else:
        starts = np.zeros(0, np.int64)
        chrom_ix = np.zeros(0, np.int64)
    return names, chrom_ix, starts, starts + window, np.arange(len(starts), dtype=np.int64)


def window_pair_tuples(chromosome_sizes, window, step, chromosome=None):
    """The pairs `chess sim` would load from the file `chess pairs` writes (helpers.py:483-527:
    BEDPE start + 1, end kept, both strands '+', reference region = query region)."""
    from .helpers import GenomicRegion
    names, chrom_ix, starts, ends, ids = window_pairs(chromosome_sizes, window, step, chromosome)
    out = []
    for c, s, e, k in zip(chrom_ix.tolist(), starts.tolist(), ends.tolist(), ids.tolist()):
        out.append((str(k), GenomicRegion(chromosome=names[c], start=s + 1, end=e, strand='+'),
                    GenomicRegion(chromosome=names[c], start=s + 1, end=e, strand='+')))
    return out


def write_pairs(path, chromosome_sizes, window, step, chromosome=None):
    names, chrom_ix, starts, ends, ids = window_pairs(chromosome_sizes, window, step, chromosome)
    with open(path, "w") as fh:
        fh.write("".join("%s\t%d\t%d\t%s\t%d\t%d\t%d\t.\t+\t+\n" % (names[c], s, e, names[c], s, e, k)
                         for c, s, e, k in zip(chrom_ix.tolist(), starts.tolist(), ends.tolist(), ids.tolist())))
    return len(ids)


def background_windows(input_regions, window, step, strand=None):
    """Rows (chromosome, start, end, index, strand) of `chess background`; bin/chess:503-518.

    `input_regions`: (chromosome, start, end) triples; windows start at `range(start, end - window,
    step)`; the index restarts for every input region and counts both strands."""
    if strand not in ('-', '+', None):
        raise ValueError("--strand must be either - or +")
    strands = ['+', '-'] if strand is None else [strand]
    rows = []
    for chromosome, start, end in input_regions:
        ix = 0
        for s in range(start, end - window, step):
            for st in strands:
                rows.append((chromosome, s, s + window, ix, st))
                ix += 1
    return rows


def write_background(path, input_regions, window, step, strand=None):
    rows = background_windows(input_regions, window, step, strand)
    with open(path, "w") as fh:
        fh.write("".join("%s\t%d\t%d\t%d\t.\t%s\n" % r for r in rows))
    return len(rows)
