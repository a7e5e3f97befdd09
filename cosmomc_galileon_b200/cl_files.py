"""Text outputs of the scalar path in the reference's format: output_cl_files (camb/modules.f90:1309-1417, the
ScalFile and LensFile branches; headers as open_file_header :1277-1294) -- what camb/inidriver.F90:345-347 writes as
<root>_scalCls.dat and <root>_lensedCls.dat.  Host-side formatting only; the numbers come from GalCamb.Cls() /
GalCamb.Cl_lensed()."""
import math

C_NAME_TAGS = ["TT", "EE", "TE", "PP", "TP", "EP"]  # camb/modules.f90:1142
CT_NAME_TAGS = ["TT", "EE", "BB", "TE"]             # :1143
NAME_TAG_LEN = 12                                   # :1141


def fortran_e(x, w=15, d=5):
    """Fortran Ew.d edit descriptor (0.ddddd E+ee normalisation; three-digit exponents drop the letter)."""
    if x != x:
        body = "NaN"
    elif math.isinf(x):
        body = "Infinity" if x > 0 else "-Infinity"
    else:
        sign = "-" if (x < 0 or (x == 0 and math.copysign(1.0, x) < 0)) else ""
        a = abs(x)
        if a == 0.0:
            e, digits = 0, "0" * d
        else:
            e = int(math.floor(math.log10(a))) + 1
            m = round(a / 10.0 ** e * 10 ** d)
            if m >= 10 ** d:  # 0.99999|9 -> 1.00000
                m //= 10
                e += 1
            if m < 10 ** (d - 1):  # log10 rounding at exact powers of ten
                m *= 10
                e -= 1
            digits = "%0*d" % (d, m)
        es = "%+03d" % e
        body = sign + "0." + digits + ("E" + es if len(es) <= 3 else es)
    return body.rjust(w) if len(body) <= w else "*" * w


def _header(columns, nn=6):
    # write(unit,'("#",1A<nn-1>," ",*(A15))') 'L', Columns   with character(LEN=12) column tags
    return "#" + "L".rjust(nn - 1) + " " + "".join(c.ljust(NAME_TAG_LEN).rjust(15) for c in columns) + "\n"


def write_scalar_cls(path, Cl, factor=1.0, headers=True):
    """ScalFile branch: Cl[type][l-2] with 3 (TT,EE,TE) or 6 types; columns up to TP (last_C = min(C_PhiTemp, C_last))."""
    ntype, nl = len(Cl), len(Cl[0])
    last_c = min(5, ntype)
    lmax = nl + 1
    with open(path, "w") as f:
        if headers:
            f.write(_header(C_NAME_TAGS[:last_c]))
        for l in range(2, min(10000, lmax) + 1):
            f.write("%6d" % l + "".join(fortran_e(factor * Cl[t][l - 2]) for t in range(last_c)) + "\n")
        for l in range(10100, lmax + 1, 100):
            f.write(fortran_e(float(l)) + "".join(fortran_e(factor * Cl[t][l - 2]) for t in range(last_c)) + "\n")


def write_lensed_cls(path, Cl_lensed, factor=1.0, headers=True):
    """LensFile branch: Cl_lensed[TT,EE,BB,TE][l-2], l = 2..lmax_lensed."""
    nl = len(Cl_lensed[0])
    with open(path, "w") as f:
        if headers:
            f.write(_header(CT_NAME_TAGS))
        for l in range(2, nl + 2):
            f.write("%6d" % l + "".join(fortran_e(factor * Cl_lensed[t][l - 2]) for t in range(4)) + "\n")
