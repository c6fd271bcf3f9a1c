"""Diagnostic: do a pair's bits depend on how the pair list is cut into launches?  (chr2 @ 25 kb, per regime)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import chess_b200 as cb

genome, ref, qry, pairs = bench.make_workload(bench.workload_spec(bench.parse_args(["--workload", "chr2_25kb"])))
regions = genome.regions()
trees = cb.region_interval_trees(regions)
re_ = cb.observed_expected_arrays(regions, ref)
qe = cb.observed_expected_arrays(regions, qry)
eng = cb.CompareEngine(re_, trees, qe, trees)
for kw in (dict(), dict(absolute_window_size=7), dict(compute_sn=False), dict(absolute_window_size=5)):
    ids, s, sn, st = eng.compare(pairs, **kw)
    ids, s2, sn2, st2 = eng.compare(pairs, **kw)
    print(kw, "repeat identical:", s.tobytes() == s2.tobytes(), sn.tobytes() == sn2.tobytes())
    for size in (517, 97, 1200):
        cut = [eng.compare(pairs[lo:lo + size], **kw) for lo in range(0, len(pairs), size)]
        cs = np.concatenate([c[1] for c in cut]); csn = np.concatenate([c[2] for c in cut])
        ds = np.nonzero(~((cs == s) | (np.isnan(cs) & np.isnan(s))))[0]
        dn = np.nonzero(~((csn == sn) | (np.isnan(csn) & np.isnan(sn))))[0]
        print("  cut", size, "ssim differs:", len(ds), ds[:8], "max rel", (np.abs(cs[ds]-s[ds])/np.abs(s[ds])).max() if len(ds) else 0,
              "| sn differs:", len(dn), dn[:8], "max rel", (np.abs(csn[dn]-sn[dn])/np.abs(sn[dn])).max() if len(dn) else 0)
    g = eng.compare(pairs, kernel=1, **kw)
    ok = ~np.isnan(g[1])
    print("  vs generic: ssim", np.max(np.abs(s[ok]-g[1][ok])/np.abs(g[1][ok])), "sn", np.nanmax(np.abs(sn[ok]-g[2][ok])/np.abs(g[2][ok])) if kw.get("compute_sn", True) else None)
