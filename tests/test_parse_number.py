"""Decimal text -> double / int of the device parser (chess_b200/csrc/parse_number.cuh), host build.

The same header compiles for host and device; here g++ builds it into a small harness and every
accepted conversion is compared bit by bit with Python's float() / int() -- the functions the
reference applies to each field (chess/helpers.py:316-326).  A conversion may give up (kind 3: the
host parses that line), it must never differ.
"""
import math
import os
import random
import shutil
import struct
import subprocess
from decimal import Decimal, getcontext

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))

FIXED = ["0", "-0", "+0.0", "1", "1.5", "0.1", "1e22", "1e23", "9007199254740993", "9007199254740992",
         "9007199254740991", "1.7976931348623157e308", "1.7976931348623159e308", "1e309", "-1e400",
         "2.2250738585072014e-308", "2.2250738585072011e-308", "5e-324", "4.9406564584124654e-324",
         "0.000000000000000000000000001", "123456789012345678", "1234567890123456789", "12345678901234567890",
         "12345678901234567891", "1.", ".5", ".", "1e", "1e+", "inf", "-Infinity", "NaN", "nan", "+INF", "1_0",
         " 1", "1 ", "0x1p3", "8.5e-5", "0.3", "2.5", "1e-22", "1e-23", "123456789.123456789", "0.1e1",
         "17976931348623157e292", "1" + "0" * 30, "0." + "0" * 30 + "1", "9223372036854775807",
         "9223372036854775808", "18446744073709551615", "18014398509481985e-1", "45035996273704965e-1",
         "1.00000000000000011102230246251565404236316680908203125", "6.0221409e+23", "3.141592653589793",
         "2.718281828459045e0", "1e-400", "0e500", "0.0e-500", "00012.500", "1E5", "1e05", "--1", "1e5.0", "", "e5",
         "9.999999999999999e22", "8.41e21", "7.3177701707893310e+15", "1.0000000000000002", "0.9999999999999999",
         "4.35", "0.000001", "1e-5", "123456.789e3", "1.2345678901234567e-05", "72057594037927933e0"]


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("parse") / "parse_harness")
    subprocess.run([gxx, "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "parse_number_harness.cpp")], check=True)
    return exe


def _run(exe, cases, mode="f"):
    out = subprocess.run([exe, mode], input="\n".join(cases) + "\n", capture_output=True, text=True, check=True)
    rows = out.stdout.split("\n")[:len(cases)]
    assert len(rows) == len(cases)
    return [r.split() for r in rows]


def _random_cases(n, seed):
    rng = random.Random(seed)
    getcontext().prec = 60
    out = []
    for _ in range(n):
        k = rng.random()
        if k < 0.25:
            out.append(repr(rng.lognormvariate(0, 3)))
        elif k < 0.5:
            x = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(63)))[0]
            if math.isfinite(x):
                out.append(repr(x))
        elif k < 0.7:
            nd = rng.randint(1, 19)
            out.append("%de%d" % (rng.randint(1, 10 ** nd - 1), rng.randint(-340, 310)))
        elif k < 0.85:
            m = str(rng.randint(1, 10 ** rng.randint(1, 19) - 1))
            p = rng.randint(0, len(m))
            out.append(m[:p] + "." + m[p:])
        else:  # close to the midpoint between two adjacent doubles
            b = rng.getrandbits(62) | (1 << 61)
            x = struct.unpack("<d", struct.pack("<Q", b))[0]
            y = struct.unpack("<d", struct.pack("<Q", b + 1))[0]
            mid = (Decimal(x) + Decimal(y)) / 2
            out.append(("%." + str(rng.randint(15, 18)) + "e") % mid)
    return out


def test_float_conversion_is_bit_identical_or_deferred(harness):
    cases = FIXED + _random_cases(300000, 7)
    cases = [c for c in cases if "\n" not in c]
    rows = _run(harness, cases)
    accepted = deferred = 0
    for c, (rc, bits) in zip(cases, rows):
        rc = int(rc)
        try:
            ref = float(c)
        except ValueError:
            ref = None
        if rc == 3:
            deferred += 1
            continue
        assert ref is not None, "accepted a spelling float() rejects: %r" % c
        if rc == 2:
            assert not math.isfinite(ref), c
            continue
        assert math.isfinite(ref), c
        assert struct.unpack("<Q", struct.pack("<d", ref))[0] == int(bits, 16), (c, ref.hex(), bits)
        accepted += 1
    assert accepted > 0.97 * len(cases), (accepted, deferred)
    # everyday spellings must not be deferred
    quick = dict(zip(cases, rows))
    for c in ("1", "0.1", "2.5", "1e22", "1e23", "8.5e-5", "1.2345678901234567e-05", "6.0221409e+23", "nan", "inf",
              "1.7976931348623157e308", "2.2250738585072014e-308", "9007199254740993", "1234567890123456789"):
        assert quick[c][0] != "3", c


def test_index_conversion(harness):
    cases = ["0", "7", "007", "+5", "123456789", "1234567890", "-1", "1.0", "", " 3", "3 ", "1_0", "12a", "2147483647"]
    rows = _run(harness, cases, "i")
    for c, (rc, v) in zip(cases, rows):
        if rc == "0":
            assert int(c) == int(v), c
        else:
            assert rc == "3", c
    ok = {c for c, (rc, _) in zip(cases, rows) if rc == "0"}
    assert ok == {"0", "7", "007", "+5", "123456789"}
