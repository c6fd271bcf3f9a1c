"""Seeded synthetic Hi-C samples, region files and pair lists.

Stands in for the missing example data of the reference (its .hic inputs are
absent blobs) and follows the recipe of utils/make_synthetic_references.py:
power-law distance decay d^-0.85 (:93-109) with nested TAD blocks (:66-90),
here sampled as Poisson counts with per-bin biases, so that two samples built
from the same `structure_seed` share structure and differ by noise.  Entries
with a zero count are absent from the sparse output (their fraction rises with
distance), and runs of bins are unmappable (no entries at all).

Used by tests/ and bench.py; everything is vectorised numpy on the host.
The module imports nothing of the package at import time, so bench.py's CPU
reference arm can load it by file path without mapping the CUDA library;
functions that make region objects take the class to use (`region_cls`).
"""
import os

import numpy as np

HG38_CHROM_BP = [
    ("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555),
    ("chr5", 181538259), ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636),
    ("chr9", 138394717), ("chr10", 133797422), ("chr11", 135086622), ("chr12", 133275309),
    ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189), ("chr16", 90338345),
    ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
    ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415),
]


def chromosome_bins(chroms_bp, bin_size):
    """[(name, n_bins)] for chromosome lengths in bp."""
    return [(name, int(-(-length // bin_size))) for name, length in chroms_bp]


class Genome(object):
    """Bin layout: chromosome names, first bin of each, total bins."""

    def __init__(self, chrom_bins, bin_size, chroms_bp=None):
        self.names = [c for c, _ in chrom_bins]
        self.sizes = np.array([n for _, n in chrom_bins], dtype=np.int64)
        self.starts = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)
        self.n_bins = int(self.starts[-1])
        self.bin_size = int(bin_size)
        self.lengths_bp = dict(chroms_bp) if chroms_bp else {c: int(n) * bin_size for c, n in chrom_bins}

    def regions(self, region_cls=None):
        if region_cls is None:
            from .helpers import GenomicRegion as region_cls
        GenomicRegion = region_cls
        out = []
        for c, name in enumerate(self.names):
            for k in range(int(self.sizes[c])):
                end = min((k + 1) * self.bin_size, self.lengths_bp[name])
                out.append(GenomicRegion(chromosome=name, start=k * self.bin_size, end=end,
                                         ix=int(self.starts[c]) + k))
        return out

    def write_bed(self, path):
        with open(path, "w") as fh:
            for c, name in enumerate(self.names):
                for k in range(int(self.sizes[c])):
                    end = min((k + 1) * self.bin_size, self.lengths_bp[name])
                    fh.write("%s\t%d\t%d\n" % (name, k * self.bin_size, end))


def _domains(n, rng, mean_size):
    """Domain id per bin for one TAD level."""
    ids = np.empty(n, dtype=np.int64)
    pos, k = 0, 0
    while pos < n:
        size = max(3, int(rng.gamma(4.0, mean_size / 4.0)))
        ids[pos:pos + size] = k
        pos += size
        k += 1
    return ids


def synthetic_sample(genome, band, seed, structure_seed=0, depth=60.0, unmappable_frac=0.03,
                     perturb=0.0):
    """Canonical COO (src, sink, weight) of one balanced-like sample.

    band            largest bin distance generated
    seed            noise seed (differs between the two samples of a comparison)
    structure_seed  TAD layout / unmappable runs (shared between samples)
    perturb         fraction of top-level domains whose intensity differs in this
                    sample (gives the comparison something to find)
    """
    srng = np.random.default_rng(structure_seed)
    nrng = np.random.default_rng(seed)
    src_all, sink_all, w_all = [], [], []
    for c in range(len(genome.names)):
        n = int(genome.sizes[c])
        off = int(genome.starts[c])
        b = min(band, n - 1)
        lvl1, lvl2 = _domains(n, srng, 60), _domains(n, srng, 15)
        dead = np.zeros(n, dtype=bool)
        n_runs = max(1, int(unmappable_frac * n / 4))
        for s in srng.integers(0, max(n - 8, 1), n_runs):
            dead[s:s + int(srng.integers(1, 8))] = True
        bias = srng.lognormal(0.0, 0.15, n)
        boost = np.ones(int(lvl1.max()) + 1)
        if perturb > 0:
            pick = nrng.random(len(boost)) < perturb
            boost[pick] = nrng.uniform(1.5, 2.5, int(pick.sum()))
        i = np.repeat(np.arange(n, dtype=np.int64), b + 1)
        d = np.tile(np.arange(b + 1, dtype=np.int64), n)
        j = i + d
        keep = j < n
        i, d, j = i[keep], d[keep], j[keep]
        lam = depth * np.power(d + 1.0, -0.85)
        same1 = lvl1[i] == lvl1[j]
        lam = lam * np.where(same1, 1.75 * boost[lvl1[i]], 1.0) * np.where(lvl2[i] == lvl2[j], 1.5, 1.0)
        lam *= bias[i] * bias[j]
        counts = nrng.poisson(lam)
        ok = (counts > 0) & ~dead[i] & ~dead[j]
        i, j, counts = i[ok], j[ok], counts[ok]
        w = counts / (bias[i] * bias[j])
        src_all.append(i + off)
        sink_all.append(j + off)
        w_all.append(w)
    return (np.concatenate(src_all), np.concatenate(sink_all), np.concatenate(w_all).astype(np.float64))


def write_sparse(path, src, sink, w):
    import pandas as pd
    frame = pd.DataFrame({"a": src, "b": sink, "w": w})
    frame.to_csv(path, sep="\t", header=False, index=False, float_format="%.17g")


def count_sliding_pairs(chroms_bp, window_bp, step_bp):
    """Number of pairs `sliding_pairs` makes for these chromosome lengths."""
    return int(sum(len(range(0, length - window_bp + 1, step_bp)) for _, length in chroms_bp))


def sliding_pairs(genome, window_bp, step_bp, chroms=None, strand_query="+", region_cls=None):
    """[(ID, reference_region, query_region)] sliding windows, same locus both sides
    (the layout `chess pairs` emits, bin/chess:444-458)."""
    if region_cls is None:
        from .helpers import GenomicRegion as region_cls
    GenomicRegion = region_cls
    out, pid = [], 0
    for name in genome.names:
        if chroms is not None and name not in chroms:
            continue
        length = genome.lengths_bp[name]
        for start in range(0, length - window_bp + 1, step_bp):
            out.append((str(pid),
                        GenomicRegion(chromosome=name, start=start + 1, end=start + window_bp, strand="+"),
                        GenomicRegion(chromosome=name, start=start + 1, end=start + window_bp,
                                      strand=strand_query)))
            pid += 1
    return out


def write_pairs(path, pairs):
    with open(path, "w") as fh:
        for pid, r, q in pairs:
            fh.write("%s\t%d\t%d\t%s\t%d\t%d\t%s\t.\t%s\t%s\n" % (
                r.chromosome, r.start - 1, r.end, q.chromosome, q.start - 1, q.end, pid, r.strand, q.strand))


def write_dataset(directory, genome, ref, qry, pairs):
    os.makedirs(directory, exist_ok=True)
    genome.write_bed(os.path.join(directory, "regions.bed"))
    write_sparse(os.path.join(directory, "reference.sparse"), *ref)
    write_sparse(os.path.join(directory, "query.sparse"), *qry)
    write_pairs(os.path.join(directory, "pairs.bedpe"), pairs)
    return {k: os.path.join(directory, v) for k, v in (
        ("regions", "regions.bed"), ("reference", "reference.sparse"),
        ("query", "query.sparse"), ("pairs", "pairs.bedpe"))}


# ---------------------------------------------------------------- dense matrices of BASELINE configs[0]
def dense_decay(n):
    """Distance-decay background of utils/make_synthetic_references.py:93-109:
    linspace(1e-4, 1, 1000)^-0.85 along the diagonals (first n values), clipped at 1."""
    decay = np.power(np.linspace(0.0001, 1, 1000), -0.85)
    d = np.abs(np.arange(n)[:, None] - np.arange(n)[None, :])
    return np.maximum(decay[np.minimum(d, 999)], 1.0)


def dense_features(n, rng, levels=(75.0, 100.0, 150.0), loop_prob=0.5):
    """Nested TAD blocks (+75 / +100 / +150 per level) with corner loops, after
    make_synthetic_references.py:66-90; returns the symmetric feature matrix to add
    to `dense_decay`.  Level k cuts every level k-1 domain into smaller ones."""
    m = np.zeros((n, n))
    borders = [0, n]
    for lvl, inc in enumerate(levels):
        nxt = [0]
        for a, b in zip(borders[:-1], borders[1:]):
            k = int(rng.integers(2, 5)) if b - a >= 12 else 1
            cuts = np.sort(rng.choice(np.arange(a + 3, max(b - 3, a + 4)), size=k - 1, replace=False)) if k > 1 else []
            nxt.extend(int(c) for c in cuts)
            nxt.append(b)
        borders = sorted(set(nxt))
        for a, b in zip(borders[:-1], borders[1:]):
            m[a:b, a:b] += inc
            if rng.random() < loop_prob and b - a >= 6:
                rad = max(int((b - a) * 0.05), 1)
                m[max(a - rad, 0):a + rad, b - rad:min(b + rad, n)] += inc * float(rng.uniform(0.5, 1.5))
    return np.triu(m) + np.triu(m, 1).T


def dense_noisy(m, rng, noise=0.2):
    """Multiplicative lognormal noise, symmetric, strictly positive (so that log2(O/E) stays finite)."""
    e = rng.lognormal(0.0, noise, m.shape)
    e = np.triu(e) + np.triu(e, 1).T
    return m * e


def dense_oe(m):
    """Matrix divided by its per-diagonal mean (utils/run_synthetic_test.py:335-346)."""
    n = m.shape[0]
    d = np.abs(np.arange(n)[:, None] - np.arange(n)[None, :])
    sums = np.bincount(d.ravel(), weights=m.ravel(), minlength=n)
    cnt = np.bincount(d.ravel(), minlength=n)
    return m / (sums / cnt)[d]
