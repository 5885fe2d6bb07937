"""CPU model of the device-side "%10.8e" formatter (csrc/vtk.cu: fmt_e8) on the generated
128-bit power-of-ten table (csrc/pow10_table.h): the same integer steps in Python ints,
checked against C printf semantics ('%.8e' in Python is correctly rounded like glibc's).
Values the fast path cannot decide with certainty must come back as EXCEPTIONS (the
library then formats that chunk with snprintf on the host), never as wrong digits."""
import os
import random
import re
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M64 = (1 << 64) - 1


def load_table():
    text = open(os.path.join(ROOT, "multigrid_parallel_b200", "csrc", "pow10_table.h")).read()
    qmin = int(re.search(r"kPow10Min = (-?\d+)", text).group(1))
    pairs = re.findall(r"\{0x([0-9a-f]{16})ull, 0x([0-9a-f]{16})ull\}", text)
    exps = re.search(r"kPow10ExpHost\[\d+\] = \{(.*?)\};", text, re.S).group(1)
    exps = [int(x) for x in exps.replace("\n", " ").split(",") if x.strip()]
    tab = [((int(h, 16) << 64) | int(l, 16), t) for (h, l), t in zip(pairs, exps)]
    assert len(tab) == len(exps) == 633
    return qmin, tab


QMIN, TAB = load_table()


def test_table_entries_are_truncated_powers_of_ten():
    for q in (-300, -1, 0, 1, 27, 38, 39, 332):
        p, t = TAB[q - QMIN]
        assert (1 << 127) <= p < (1 << 128)
        if q >= 0:
            x = 10 ** q
            assert (p << t if t >= 0 else p >> -t) <= x < ((p + 1) << t if t >= 0 else x + 1)
        else:  # p * 2^t <= 10^q < (p+1) * 2^t  <=>  p * 10^-q <= 2^-t < (p+1) * 10^-q
            x = 10 ** (-q)
            assert p * x <= (1 << -t) < (p + 1) * x


def fmt_e8_model(bits):
    """returns the 14/15-char string, or None for an exception (host formats it)"""
    sign = bits >> 63
    ex = (bits >> 52) & 0x7FF
    frac = bits & ((1 << 52) - 1)
    if ex == 0x7FF:
        return None
    pre = "-" if sign else ""
    if ex == 0 and frac == 0:
        return pre + "0.00000000e+00"
    if ex == 0:  # subnormal: normalise
        m, e2 = frac, -1074
        sh = 53 - m.bit_length()
        m <<= sh
        e2 -= sh
    else:
        m, e2 = frac | (1 << 52), ex - 1075
    E = ((e2 + 52) * 1233) >> 12  # floor(log10(2^(e2+52))) or one less
    for _ in range(3):
        q = 8 - E
        if q < QMIN or q - QMIN >= len(TAB):
            return None
        p, t = TAB[q - QMIN]
        prod = m * p  # < 2^181
        s = -(e2 + t)
        if not (130 <= s <= 190):
            return None
        r2, r1, r0 = prod >> 128, (prod >> 64) & M64, prod & M64
        D = r2 >> (s - 128)
        if D >= 10 ** 9:
            E += 1
            continue
        if D < 10 ** 8:
            E -= 1
            continue
        f2 = r2 & ((1 << (s - 128)) - 1)
        h2 = 1 << (s - 129)
        exact = 0 <= q <= 38  # 10^q < 2^128: the table entry is 10^q itself
        if f2 > h2 or (f2 == h2 and (r1 | r0)):
            D += 1
        elif f2 == h2:
            if not exact:
                return None  # looks like a tie, but the table entry was truncated
            D += D & 1  # a true tie: to even, like printf in round-to-nearest
        elif f2 == h2 - 1 and r1 == M64 and not exact:
            return None  # within the truncation error of the table below one half
        if D == 10 ** 9:
            D, E = 10 ** 8, E + 1
        if not -99 <= E <= 99:
            return None
        ds = str(D)
        return f"{pre}{ds[0]}.{ds[1:]}e{'+' if E >= 0 else '-'}{abs(E):02d}"
    return None


def bits_of(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def check(x, allow_exception=False):
    got = fmt_e8_model(bits_of(x))
    if got is None:
        assert allow_exception, x
        return False
    assert got == "%.8e" % x, (x, got, "%.8e" % x)
    return True


def test_model_special_values():
    for x in (0.0, -0.0, 1.0, -1.0, 0.5, 1e-10, 9.999999995, 9.9999999949999, 123456789.0,
              999999999.4, 0.1, 1 / 3, 2 / 3, 1e22, 1e23, 5e-324 * 2 ** 60, 1.7976931348623157e308,
              2.2250738585072014e-308, 1e-99, 9.99999999e-99, 9.5e98):
        check(x, allow_exception=abs(x) > 1e99 or (x != 0 and abs(x) < 1e-98))
    # exact ties are exceptions (decided by the host's printf), never silently rounded
    # (1.000000005e9 and 1.000000015e9 through the truncated 10^-1, 1.000000025e7 through
    # the exact 10^1)
    for x in (1000000005.0, 1000000015.0):
        assert fmt_e8_model(bits_of(x)) is None, x
    # ... unless the power of ten is exact: then the tie is real and goes to even
    for x in (10000000.25, -10000000.25, 10000000.75, 0.01025390625, 103 / 1024, 0.5, 2.5e-7):
        assert fmt_e8_model(bits_of(x)) == "%.8e" % x, x
    assert fmt_e8_model(bits_of(float("inf"))) is None and fmt_e8_model(bits_of(float("nan"))) is None
    assert fmt_e8_model(bits_of(1e100)) is None and fmt_e8_model(bits_of(1e-100)) is None


def test_model_random_doubles():
    rng = random.Random(7)
    done = 0
    for _ in range(60000):
        b = rng.getrandbits(64)
        ex = (b >> 52) & 0x7FF
        if ex == 0x7FF:
            continue
        x = struct.unpack("<d", struct.pack("<Q", b))[0]
        done += check(x, allow_exception=not (1e-99 < abs(x) < 1e99))
    # the range a Poisson solution / its error lives in
    for _ in range(60000):
        x = rng.uniform(-1, 1) * 10.0 ** rng.randint(-16, 3)
        done += check(x)
    for i in range(0, 2049):
        done += check(i * (1.0 / 2048)) + check(i * (1.0 / 512)) + check(i / 1000.0)
    assert done > 70000


def test_model_near_decimal_boundaries():
    # doubles next to d.dddddddd5 boundaries and to powers of ten
    rng = random.Random(11)
    for _ in range(20000):
        D = rng.randint(10 ** 8, 10 ** 9 - 1)
        E = rng.randint(-20, 20)
        x = float(f"{D}5e{E - 9}")
        for y in (x, np.nextafter(x, np.inf), np.nextafter(x, -np.inf)):
            check(float(y), allow_exception=True)
    for E in range(-98, 99):
        x = float(f"1e{E}")
        for y in (x, np.nextafter(x, np.inf), np.nextafter(x, -np.inf)):
            check(float(y), allow_exception=True)


def test_oracle_vtk_writer_vs_compiled_reference(tmp_path, orc, ref):
    """oracle/mg_oracle.c: orc_write_vtk against the reference's own writeOutputData
    (postprocess.h:5-47, compiled into oracle/_ref), byte for byte"""
    if ref is None or not hasattr(ref.L, "ref_write_vtk"):
        import pytest
        pytest.skip("oracle/_ref not built")
    from oracle_lib import seeded
    for N, seed in ((5, 1), (9, 2), (17, 3)):
        v = seeded((N,) * 3, seed) * 10.0 ** np.random.default_rng(seed).integers(-12, 3, (N,) * 3)
        v[0, 0, :3] = [0.0, -0.0, 103 / 1024]
        h = 1.0 / (N - 1)
        a, b = tmp_path / f"orc{N}.vtk", tmp_path / f"ref{N}.vtk"
        orc.write_vtk(a, v, h)
        ref.write_vtk(b, v, h)
        assert a.read_bytes() == b.read_bytes()


def test_oracle_vtk_writer_lines(tmp_path, orc):
    """... and, for a box, against the format strings themselves"""
    from oracle_lib import seeded
    v = seeded((3, 4, 5), 8)
    orc.write_vtk(tmp_path / "box.vtk", v, 0.25)
    lines = (tmp_path / "box.vtk").read_text().split("\n")
    assert lines[4] == "DIMENSIONS 3 4 5" and lines[5] == "POINTS 60 float"
    assert lines[6 + 7] == "%10.8e %10.8e %10.8e" % (0.0, 0.25 * 1, 0.25 * 2)  # point 7 = (0, 1, 2)
    assert lines[6 + 60:6 + 64] == ["", "POINT_DATA 60", "SCALARS data float 1", "LOOKUP_TABLE default"]
    assert lines[6 + 64 + 59] == "%10.8e" % v.reshape(-1)[59] and lines[-1] == ""


def test_oracle_vtk_writer_golden(tmp_path, orc):
    """hash of the file the reference's own writer produced for a seeded 9^3 grid
    (tests/golden/operators.json, oracle/gen_golden.py)"""
    import hashlib
    import json
    from oracle_lib import seeded
    ops = json.load(open(os.path.join(ROOT, "tests", "golden", "operators.json")))
    g = seeded((9,) * 3, 21) * 10.0 ** np.random.default_rng(21).integers(-12, 3, (9,) * 3)
    orc.write_vtk(tmp_path / "g.vtk", g, 0.125)
    assert hashlib.sha256((tmp_path / "g.vtk").read_bytes()).hexdigest() == ops["vtk_9_sha256"]
