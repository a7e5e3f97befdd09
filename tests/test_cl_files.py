"""Text outputs in the reference's format (camb/modules.f90:1309-1417): Fortran E15.5 fields, header line, and a round
trip of the golden C_l through the file."""
import os
import re

import numpy as np

from cosmomc_galileon_b200.cl_files import fortran_e, write_lensed_cls, write_scalar_cls

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_fortran_e_descriptor():
    assert fortran_e(1.0) == "    0.10000E+01"
    assert fortran_e(-6.0221e23) == "   -0.60221E+24"
    assert fortran_e(0.0) == "    0.00000E+00"
    assert fortran_e(9.999996e-11) == "    0.10000E-09"  # mantissa rounds up to the next decade
    assert fortran_e(1234.5678) == "    0.12346E+04"
    assert fortran_e(1e-100) == "    0.10000E-99"
    assert fortran_e(1e-101) == "    0.10000-100"      # three-digit exponent drops the letter
    assert fortran_e(1000.0) == "    0.10000E+04" and fortran_e(0.001) == "    0.10000E-02"
    assert len(fortran_e(-3.3e-7)) == 15


def test_scalar_and_lensed_files_round_trip(tmp_path):
    z = np.load(os.path.join(GOLD, "config2_golden.npz"))
    Cl = z["Cl"]
    p = tmp_path / "test_scalCls.dat"
    write_scalar_cls(str(p), Cl, factor=7.42835025e12)
    lines = open(p).read().splitlines()
    # "#", A5, " ", then each LEN=12 tag right-justified in A15
    assert lines[0] == "#    L " + "".join("   " + t + " " * 10 for t in ("TT", "EE", "TE"))
    assert len(lines) == 1 + Cl.shape[1]
    assert re.fullmatch(r" {5}2( {3,4}-?0\.\d{5}E[+-]\d\d){3}", lines[1])
    back = np.loadtxt(str(p))
    assert back.shape == (Cl.shape[1], 4) and back[0, 0] == 2 and back[-1, 0] == Cl.shape[1] + 1
    np.testing.assert_allclose(back[:, 1:], 7.42835025e12 * Cl.T, rtol=6e-5)
    L = np.load(os.path.join(GOLD, "config3_lensing_golden.npz"))["Cl_lensed_7"]
    q = tmp_path / "test_lensedCls.dat"
    write_lensed_cls(str(q), L)
    back = np.loadtxt(str(q))
    assert back.shape == (L.shape[1], 5)
    np.testing.assert_allclose(back[:, 1:], L.T, rtol=6e-5)
    assert open(q).readline().split() == ["#", "L", "TT", "EE", "BB", "TE"]
