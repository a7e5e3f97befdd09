"""Galileon backgrounds of a batch on the device (galcamb_set_prestage_device_) against the host threads: wall clock of
the pre-stage with either, largest table difference.  usage: python tools/gpu_bg_probe.py [batch] [host threads]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosmomc_galileon_b200 import GalCamb, make_params, prestage  # noqa: E402
from cosmomc_galileon_b200.proposal import valid_points  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 8)
    g = GalCamb(0)

    def to_params(P):
        return make_params(Max_l=2500, Max_eta_k=5000.0, **P)

    def accept(cand):
        _, st = prestage([to_params(P) for P in cand], nthreads=threads, copy=False)
        return [s == 0 for s in st]

    pts, _ = valid_points(B, 20250101, accept)
    params = [to_params(P) for P in pts]
    t0 = time.time()
    th, sh = prestage(params, nthreads=threads, copy=True)
    t_host = time.time() - t0
    g.set_prestage_device(0)
    prestage(params[:2], nthreads=threads, copy=False)  # buffers, stream
    t0 = time.time()
    td, sd = prestage(params, nthreads=threads, copy=True)
    t_dev = time.time() - t0
    err = 0.0
    for a, b in zip(th, td):
        for key in ("gal_h", "gal_x", "vis", "dotmu"):
            x, y = np.asarray(a.d[key]), np.asarray(b.d[key])
            err = max(err, float(np.max(np.abs(x - y)) / np.max(np.abs(x))))
    print("B=%d, %d host threads: pre-stage %.3f s on the host, %.3f s with the backgrounds on the device; status equal %s; "
          "largest table difference %.2e" % (B, threads, t_host, t_dev, sh == sd, err))


if __name__ == "__main__":
    main()
