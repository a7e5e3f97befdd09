"""Several GPUs from one process: (i) one spectrum (CosmoMC settings) over G = 1 .. ndev devices -- ms per call, K1 and
line-of-sight parts, agreement with G = 1; (ii) a batch of distinct points sharded over G devices -- spectra/s.
usage: python tools/gpu_multi_probe.py [points in the batch, default 256]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosmomc_galileon_b200 import GalCamb, make_params, prestage  # noqa: E402
from cosmomc_galileon_b200.proposal import CENTRE, valid_points  # noqa: E402


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    g = GalCamb(0)
    g.set_template(np.fromfile(os.path.join(ROOT, "tests", "golden", "highL_template.f64")).reshape(7999, 4).T.copy())
    ndev = g.device_count()
    pc = make_params(Max_l=2650, Max_eta_k=6625.0, DoLateRadTruncation=0, Num_Nu_massive=3, Num_Nu_massless=0.046, **CENTRE)
    ref = None
    for G in [n for n in (1, 2, 4, 8) if n <= ndev]:
        g.multi_single_cls(G, pc)  # warm-up: contexts, Bessel tables, NCCL communicators
        calls = []
        for _ in range(3):
            cl, ms = g.multi_single_cls(G, pc)
            calls.append(ms)
        best = min(calls, key=lambda m: m["call"])
        if ref is None:
            ref = cl
        print("single spectrum on %d GPU(s): %.1f ms per call (pre-stage %.1f, K1 %.1f, LOS slice %.2f); max rel. dev. from 1 GPU %.2e"
              % (G, best["call"], best["prestage"], best["evolve"], best["los"], float(np.max(np.abs(cl / ref - 1)))), flush=True)

    def to_params(P):
        return make_params(Max_l=2500, Max_eta_k=5000.0, **P)

    def accept(cand):
        _, st = prestage([to_params(P) for P in cand], copy=False)
        return [s == 0 for s in st]

    pts, _ = valid_points(nb, 20250101, accept)
    params = [to_params(P) for P in pts]
    ref = None
    for G in [n for n in (1, 2, 4, 8) if n <= ndev]:
        g.multi_params_to_cls(G, params[: 4 * G])
        t0 = time.time()
        out = g.multi_params_to_cls(G, params)
        dt = time.time() - t0
        if ref is None:
            ref = out
        print("%d distinct points over %d GPU(s), parameters -> C_l: %.2f s = %.1f spectra/s; max rel. dev. from 1 GPU %.2e"
              % (nb, G, dt, nb / dt, float(np.nanmax(np.abs(out / ref - 1)))), flush=True)


if __name__ == "__main__":
    main()
