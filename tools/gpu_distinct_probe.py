"""K1 on a batch of DISTINCT parameter points (seeded proposal draws, tables from the product pre-stage): stage timings
for a list of batch sizes.  usage: python tools/gpu_distinct_probe.py [batch sizes ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosmomc_galileon_b200 import GalCamb, make_params, prestage  # noqa: E402
from cosmomc_galileon_b200.proposal import valid_points  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [128]
    B = max(sizes)
    g = GalCamb(0)
    g.set_template(np.fromfile(os.path.join(ROOT, "tests", "golden", "highL_template.f64")).reshape(7999, 4).T.copy())

    def to_params(P):
        return make_params(Max_l=2500, Max_eta_k=5000.0, **P)

    def accept(cand):
        _, st = prestage([to_params(P) for P in cand], copy=False)
        return [s == 0 for s in st]

    pts, drawn = valid_points(B, 20250101, accept)
    params = [to_params(P) for P in pts]
    for n in sizes:
        prestage(params[:n], copy=False)
        g.prestaged_upload()
        for rep in range(2):
            t0 = time.time()
            g.CalcScalarSources()
            g.SourceToTransfers()
            g.ClTransferToCl()
            wall = time.time() - t0
        tm = g.timings()
        print("B=%d distinct: evolve %.1f ms los %.1f ms wall %.2f s -> %.1f spectra/s; n_derivs %.4g executed %.4g"
              % (n, tm["evolve"], tm["los"], wall, n / wall, g.counters()["n_derivs"], g.evolve_stats()["rhs_executed"]), flush=True)


if __name__ == "__main__":
    main()
