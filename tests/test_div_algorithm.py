"""CPU check of the exact-division algorithm the encode kernel uses (two FMA corrections of
v * RN(1/d)); see tests/native/div_rcp_check.c."""
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def test_known_reciprocal_division_is_ieee_exact():
    src = os.path.join(HERE, "native", "div_rcp_check.c")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "div_rcp_check")
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"])
        out = subprocess.run([exe, "25", "7"], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout[-500:]
        assert out.stdout.strip().endswith("0 mismatches")


def test_known_reciprocal_float64_division_is_ieee_exact():
    """decode_one's division (v * 2*alpha*n) / ((2^e - 1) * n) through RN(1/den) + two FMA corrections."""
    src = os.path.join(HERE, "native", "ddiv_rcp_check.c")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "ddiv_rcp_check")
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"])
        out = subprocess.run([exe, "23", "11"], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout[-500:]
        assert out.stdout.strip().endswith("0 mismatches")
