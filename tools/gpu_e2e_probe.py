"""Where the time of one end-to-end step (parameters -> C_l through galcamb_prestaged_to_cls_) goes: the bench's e2e
loop with the library's host-phase timers.  usage: python tools/gpu_e2e_probe.py [batch] [steps] [host threads] [dev: Galileon backgrounds on the device]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosmomc_galileon_b200 import GalCamb, make_params, prestage  # noqa: E402
from cosmomc_galileon_b200.proposal import valid_points  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 8)
    g = GalCamb(0)
    g.set_host_threads(threads)
    if len(sys.argv) > 4 and sys.argv[4] == "dev":
        g.set_prestage_device(0)
    g.set_template(np.fromfile(os.path.join(ROOT, "tests", "golden", "highL_template.f64")).reshape(7999, 4).T.copy())

    def to_params(P):
        return make_params(Max_l=2500, Max_eta_k=5000.0, **P)

    def accept(cand):
        _, st = prestage([to_params(P) for P in cand], nthreads=threads, copy=False)
        return [s == 0 for s in st]

    pts, _ = valid_points(B, 20250101, accept)
    params = [to_params(P) for P in pts]
    out = np.zeros((B, 3, 2499))
    g.params_to_cls(params, nthreads=threads, out=out)
    t0 = time.perf_counter()
    g.prestage_async(params, threads)
    for i in range(steps):
        ta = time.perf_counter()
        g.prestage_wait()
        tb = time.perf_counter()
        if i + 1 < steps:
            g.prestage_async(params, threads)
        g.prestaged_to_cls(out)
        tc = time.perf_counter()
        print("step %d: wait for pre-stage %.0f ms, prestaged_to_cls %.0f ms; host phases %s; device %s"
              % (i, 1e3 * (tb - ta), 1e3 * (tc - tb), {k: round(v) for k, v in g.host_timings().items()},
                 {k: round(v) for k, v in g.timings().items()}), flush=True)
    dt = time.perf_counter() - t0
    print("B=%d: %.1f spectra/s end to end (%d host threads)" % (B, B * steps / dt, threads))


if __name__ == "__main__":
    main()
