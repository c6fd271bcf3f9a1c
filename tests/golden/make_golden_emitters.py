#!/usr/bin/env python
"""Golden files for `chess pairs`, `chess background` and `chess oe`, produced by the REFERENCE's own
bin/chess (build container only; same shims as make_golden.py plus a `pybedtools` stand-in whose
`chromsizes` knows no genome, which sends the reference down its chromosome-sizes-file branch,
bin/chess:423-432, 487-497).

    python tests/golden/make_golden_emitters.py
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402

TOY = os.path.join(HERE, "toy")
P = lambda name: os.path.join(TOY, name)  # noqa: E731


def main():
    mg.install_shims()
    pbt = types.ModuleType("pybedtools")

    def chromsizes(genome):
        raise OSError("pybedtools shim: no genome table for {}".format(genome))

    pbt.chromsizes = chromsizes
    sys.modules["pybedtools"] = pbt
    mg.load_reference()
    Chess = mg.load_cli_class()
    with open(P("chrom.sizes"), "w") as fh:
        fh.write("chrA\t60000\nchrB\t25500\n\nchrC\t9000\n")

    def run(method, argv):
        obj = object.__new__(Chess)
        getattr(obj, method)(["chess", method] + argv)

    run("pairs", [P("chrom.sizes"), "20000", "7000", P("emit_pairs.bedpe"), "--file-input"])
    run("pairs", [P("chrom.sizes"), "9000", "4000", P("emit_pairs_chrB.bedpe"), "--chromosome", "chrB"])
    # the reference catches only ValueError around pybedtools here (bin/chess:487-497): give it one
    pbt.chromsizes = lambda genome: (_ for _ in ()).throw(ValueError("pybedtools shim: unknown genome"))
    run("background", [P("chrom.sizes"), "20000", "9000", P("emit_background.bed")])
    run("background", [P("chrom.sizes"), "8000", "5000", P("emit_background_minus.bed"), "--strand", "-"])
    run("background", ["chrA:1001-40000", "10000", "6000", P("emit_background_region.bed"), "--strand", "+"])
    run("oe", [P("ref.sparse"), P("ref.bed"), P("emit_oe.sparse")])
    for name in sorted(os.listdir(TOY)):
        if name.startswith("emit_") or name == "chrom.sizes":
            print(name, os.path.getsize(P(name)), "bytes")


if __name__ == "__main__":
    main()
