#!/usr/bin/env python
"""Golden vectors for the head of `chess extract`, produced by the REFERENCE's own
`extract_structures` (chess/get_structures.py:80-121).  Build container only.

    python tests/golden/make_golden_extract.py

`get_structures.py` is loaded from /root/reference with the shims of make_golden.py plus stand-ins
for the scikit-image names it imports.  The stand-in for `restoration.denoise_bilateral` is the
probe: the reference calls it with `positive` (then `negative`) and `sigma_color=np.mean(...)`,
i.e. exactly what the first stage produces; the probe records both calls and stops the function
after the second (everything after it is scikit-image / scipy post-processing, out of scope).
Everything executed up to that point -- the dict-of-dict gather with default weight 0, the log
ratio, the NaN / inf clean-up, np.std, the two thresholded maps -- is the reference's code.

Output: tests/golden/extract_golden.npz (float64 arrays, exact) for pairs of the toy data set
read through the reference's own loaders.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402

TOY = os.path.join(HERE, "toy")
P = lambda name: os.path.join(TOY, name)  # noqa: E731


class _Stop(Exception):
    pass


def main():
    mg.install_shims()
    calls = []

    def denoise_bilateral(image, sigma_color=None, **kwargs):
        calls.append((np.array(image, dtype=np.float64, copy=True), float(sigma_color)))
        if len(calls) % 2 == 0:
            raise _Stop()
        return image

    sk = sys.modules["skimage"]
    rest = types.ModuleType("skimage.restoration")
    rest.denoise_bilateral = denoise_bilateral
    morph = types.ModuleType("skimage.morphology")
    morph.square = morph.closing = None
    filt = types.ModuleType("skimage.filters")
    meas = types.ModuleType("skimage.measure")
    meas.label = meas.regionprops = None
    sk.restoration, sk.morphology, sk.filters, sk.measure = rest, morph, filt, meas
    for name, mod in (("restoration", rest), ("morphology", morph), ("filters", filt), ("measure", meas)):
        sys.modules["skimage." + name] = mod

    ref = mg.load_reference()
    H = ref["helpers"]
    spec = importlib.util.spec_from_file_location("chess.get_structures",
                                                  os.path.join(mg.REF, "chess", "get_structures.py"))
    GS = importlib.util.module_from_spec(spec)
    sys.modules["chess.get_structures"] = GS
    spec.loader.exec_module(GS)

    # the toy samples as O/E edges, through the reference's own loader (--oe-input style files)
    redges, rtrees, _ = H.load_oe_contacts(P("ref.oe.sparse"), P("ref.bed"))
    qedges, qtrees, _ = H.load_oe_contacts(P("qry.oe.sparse"), P("qry.bed"))
    pairs, _ = H.load_pairs_for_chroms(P("pairs.bedpe"), set(rtrees.keys()) | set(qtrees.keys()))
    out = {}
    kept = []
    import tempfile
    tmp = tempfile.mkdtemp()
    for pair in pairs:
        pid, rr, qr = pair
        del calls[:]
        try:
            GS.extract_structures(redges, rtrees, qedges, qtrees, [pair], tmp, 3, 3, 5, 8)
        except _Stop:
            pass
        except Exception as exc:          # unequal sizes (np.divide broadcast), region not found ...
            out["fail_" + pid] = np.array([type(exc).__name__])
            continue
        (pos, mean_pos), (neg, mean_neg) = calls
        # std is not an argument of the probe; it follows from the thresholds only implicitly, so recompute
        # it from the reference's own gather (same statements as get_structures.py:109-115, on ITS matrices)
        reference, _ = H.sub_matrix_from_edges_dict(redges, rtrees, rr, default_weight=0.)
        query, _ = H.sub_matrix_from_edges_dict(qedges, qtrees, qr, default_weight=0.)
        with np.errstate(all="ignore"):
            orm = np.log(np.divide(query, reference))
        orm[np.isnan(orm)] = 0.
        orm[np.isinf(orm)] = 0.
        out["pos_" + pid], out["neg_" + pid] = pos, neg
        out["stats_" + pid] = np.array([np.std(orm), mean_pos, mean_neg])
        kept.append(pid)
    out["ids"] = np.array(kept)
    out["all_ids"] = np.array([p[0] for p in pairs])
    path = os.path.join(HERE, "extract_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(kept), "pairs with arrays,",
          sum(1 for k in out if k.startswith("fail_")), "failing pairs")


if __name__ == "__main__":
    main()
