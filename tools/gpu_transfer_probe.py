"""Matter transfer functions (outtransf taps of K1 + the transfer-only wavenumbers + sigma_8) of N copies of a point
against the oracle.  The launcher's choice between the two forms of K1 follows GALCAMB_OCTET_ITEMS (tests force the
lane-per-item form with 0).  usage: python tools/gpu_transfer_probe.py [nmodels] [lcdm|galileon]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cosmomc_galileon_b200 import GalCamb, make_params, prestage  # noqa: E402

GAL = dict(ombh2=0.0221743, omch2=0.1177806, omnuh2=0.00064, H0=78.07227, use_galileon=1, c2=-8.438179, c3=-3.366555,
           c4=-1.03411, cG=0.001188243)
LCDM = dict(ombh2=0.0226, omch2=0.112, omnuh2=0.00064, H0=70.0)
Z = (2.0, 0.5, 0.0)
KMAX = 1.0
LMAX = 800


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    P = LCDM if (len(sys.argv) > 2 and sys.argv[2] == "lcdm") else GAL
    from oracle_py import Oracle
    o = Oracle(lmax=LMAX, **P)
    o.want_transfer(redshifts=Z, kmax=KMAX)
    o.run()
    To = o.transfer()
    p = make_params(Max_l=LMAX, Max_eta_k=2.0 * LMAX, WantTransfer=1, transfer_kmax=KMAX, transfer_num_redshifts=len(Z),
                    transfer_redshifts=Z, **P)
    tabs, st = prestage([p] * n)
    assert all(s == 0 for s in st), st
    g = GalCamb(0)
    g.set_template(np.fromfile(os.path.join(ROOT, "tests", "golden", "highL_template.f64")).reshape(7999, 4).T.copy())
    g.InitSpherBessels(tabs[0].ls, tabs[0].d["kmaxfile"], tabs[0].d["AccuracyBoost"])
    g.set_models(tabs)
    g.CalcScalarSources()
    g.SourceToTransfers()
    g.ClTransferToCl()
    worst_t = worst_s = worst_q = worst_src = 0.0
    where = (0, 0, 0)
    for slot in range(n):
        T = g.transfer(slot)
        assert T["TransferData"].shape == To["TransferData"].shape, (T["TransferData"].shape, To["TransferData"].shape)
        worst_q = max(worst_q, float(np.max(np.abs(T["q_trans"] / To["q_trans"] - 1))))
        a, b = T["TransferData"], To["TransferData"]
        # per row (transfer variable) and redshift: relative to the row's largest entry over k
        den = np.max(np.abs(b), axis=1, keepdims=True)
        den[den == 0] = 1.0
        err = np.abs(a - b) / den
        if float(np.max(err)) > worst_t:
            worst_t = float(np.max(err))
            where = np.unravel_index(int(np.argmax(err)), err.shape)
        worst_s = max(worst_s, float(np.max(np.abs(T["sigma8"] / To["sigma8"] - 1))))
    S, So = g.Src(), o.Src()
    worst_src = float(np.max(np.abs(S - So)) / np.max(np.abs(So)))
    cl = g.Cls()[1]
    if os.environ.get("GALCAMB_PROBE_VERBOSE"):
        T = g.transfer(0)
        np.set_printoptions(linewidth=200, precision=10)
        for iz in range(len(Z)):
            for ik in sorted({0, 60, o.geti("nk") - 1, To["q_trans"].size - 1, where[1]}):
                print("z", Z[iz], "k", To["q_trans"][ik], "\n  gpu   ", T["TransferData"][iz, ik], "\n  oracle", To["TransferData"][iz, ik],
                      "\n  rel   ", T["TransferData"][iz, ik] / To["TransferData"][iz, ik] - 1)
    print("worst transfer element: z index %d, k index %d (k = %.4g), row %d" % (where[0], where[1], To["q_trans"][where[1]], where[2] + 1))
    print("n=%d form=%s nq_trans=%d (source k %d) transfer err %.3e sigma8 err %.3e q_trans err %.3e Src err %.3e Cl err %.3e "
          "sigma8 %s n_derivs gpu %.0f oracle %.0f"
          % (n, os.environ.get("GALCAMB_OCTET_ITEMS", "auto"), To["q_trans"].size, o.geti("nk"), worst_t, worst_s, worst_q, worst_src,
             float(np.max(np.abs(cl / o.Cl() - 1))), np.array2string(To["sigma8"], precision=5), g.counters()["n_derivs"] / n,
             o.get("n_derivs")))


if __name__ == "__main__":
    main()
