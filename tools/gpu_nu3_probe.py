"""Three-eigenstate Galileon model through one form of K1 (GALCAMB_OCTET_ITEMS selects it): per-column source errors
against the oracle.  usage: python tools/gpu_nu3_probe.py [lmax]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as T  # noqa: E402
from tables_from_oracle import tables_from_oracle  # noqa: E402
from cosmomc_galileon_b200 import GalCamb  # noqa: E402


def main():
    lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 800
    if len(sys.argv) > 2 and sys.argv[2] in ("nu1", "lcdm"):
        from oracle_py import Oracle
        P = T.GAL if sys.argv[2] == "nu1" else T.LCDM
        o = Oracle(lmax=lmax, **P)
        o.run()
        t = tables_from_oracle(o, P, lmax=lmax)
    else:
        o = T.oracle_three_eigenstates(T.GAL, lmax)
        o.run()
        t = tables_from_oracle(o, dict(T.GAL, **T.NU3), lmax=lmax)
    print("nu_tau_massive", t.d["nu_tau_massive"], "nonrel", t.d["nu_tau_nonrelativistic"], "tau0", t.d["tau0"], "last taus", t.d["tau"][-3:])
    g = GalCamb(0)
    tmpl = np.fromfile(os.path.join(T.GOLD, "highL_template.f64")).reshape(7999, 4).T.copy()
    g.set_template(tmpl)
    T.run_gpu(g, t)
    S, So = g.Src(), o.Src()
    num = np.max(np.abs(S - So), axis=0)
    den = np.max(np.abs(So), axis=0)
    rel = num / np.maximum(den, 1e-300)
    for s in range(S.shape[1]):
        k = int(np.argmax(rel[s]))
        print("source %d: worst column k index %d of %d rel %.3e (column max %.3e, source max %.3e); 2nd worst %.3e"
              % (s, k, S.shape[2], rel[s, k], den[s, k], den[s].max(), np.sort(rel[s])[-2]))
    d = np.abs(S[:, 0] - So[:, 0])
    for k in (0, 5, 60, 117):
        j = int(np.argmax(d[:, k]))
        print("k %d: worst time index %d of %d (tau %.2f): gpu %.12e oracle %.12e; neighbours rel %s" % (
            k, j, S.shape[0], t.d["tau"][j], S[j, 0, k], So[j, 0, k],
            ["%.1e" % (d[jj, k] / den[0, k]) for jj in range(max(0, j - 2), min(S.shape[0], j + 3))]))
    print("elemmax(floor 1e-6) %.3e elemmax(floor 1e-9) %.3e" % (T.elemmax(S, So), T.elemmax(S, So, 1e-9)))
    print("colmax %.3e cl %.3e n_derivs gpu %.0f oracle %.0f" % (T.colmax(S, So), float(np.max(np.abs(g.Cls()[1] / o.Cl() - 1))),
                                                               g.counters()["n_derivs"], o.get("n_derivs")))


if __name__ == "__main__":
    main()
