#!/usr/bin/env python
"""bench.py -- scalar C_l spectra/sec (Galileon, lmax=2500) on N B200s, beside the host-core CPU baseline.

Workload (config.workload): BASELINE.json configs[1] -- Galileon scalar TT/TE/EE C_l to lmax=2500 with the
camb/params.ini accuracy settings -- as a batch of B parameter points per GPU (throughput mode of configs[3]).
A "step" = one batch through the whole hot path: per-k ODE evolution + source sampling (K1), k-spline (K2),
line-of-sight Bessel integration (K3), C_l assembly + l-interpolation (K4).
  value : spectra/s with all model tables already resident in HBM when the timed region starts
  e2e   : spectra/s through the public C-ABI call galcamb_batch_cls_ with HOST tables (packing, H2D of every
          model's tables and D2H of every C_l inside the timed region)
Inputs are synthetic: B replicas of the config-2 table fixture (tests/golden/config2_tables.npz, produced by
the oracle's host pre-stage; the pre-stage itself is a "next" row).  Every replica has its own
copy of the tables in HBM and its own ODE state, so no work is shared between replicas.
--impl reference times the CPU restatement of the reference path (oracle/, kind "port": the Fortran/icc/GSL
reference cannot be built in this image) on the host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
METRIC = "scalar C_l spectra/sec (Galileon, lmax=2500)"
UNIT = "spectra/s"


def load_meta(t):
    return json.loads(bytes(t.d["meta_json"]).decode())


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in self.rows)
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][2]), "reasons": sorted(reasons),
                "samples": len(self.rows)}


_cpu_pool = {}


def cpu_reference_rate(nspectra, threads_per_model=None):
    """Host-core baseline: the oracle.  Two layouts are tried and the faster one is used (and named in `sample`):
    OpenMP inside a model like the reference (cmbmain.f90:201-206, :263-267) with several models side by side like
    its MPI x OpenMP job layout (cosmomc.sh:13-15), and one single-threaded model per core (best for independent
    parameter points).  Model objects persist between calls so the Bessel tables stay cached as in the reference
    (camb/bessels.f90:40-41); the first call warms them outside the timed region."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_py import Oracle, build  # test infrastructure: allowed here (cpu_baseline / --impl reference)
    build()
    meta = load_meta_fixture()
    P = meta["params"]
    cores = os.cpu_count() or 1

    def pool(tpm):
        key = ("models", tpm)
        if key not in _cpu_pool:
            lanes = max(1, cores // tpm)
            _cpu_pool[key] = [Oracle(nthreads=tpm, **P) for _ in range(lanes)]
            th = [threading.Thread(target=o.run) for o in _cpu_pool[key]]
            for t in th:
                t.start()
            for t in th:
                t.join()
        return _cpu_pool[key]

    def run(tpm, n):
        models = pool(tpm)
        lanes = len(models)
        counts = [n // lanes + (1 if i < n % lanes else 0) for i in range(lanes)]

        def work(o, c):
            for _ in range(c):
                o.run(reuse_bessels=True)

        th = [threading.Thread(target=work, args=(o, c)) for o, c in zip(models, counts) if c]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0, lanes

    if threads_per_model is None:
        if "best_tpm" not in _cpu_pool:
            cand = sorted({min(8, cores), 1})
            rates = {}
            for tpm in cand:
                n = max(2, cores // tpm)
                dt, _ = run(tpm, n)
                rates[tpm] = n / dt
            _cpu_pool["best_tpm"] = max(rates, key=rates.get)
            _cpu_pool["layout_rates"] = rates
        threads_per_model = _cpu_pool["best_tpm"]
    dt, lanes = run(threads_per_model, nspectra)
    return nspectra / dt, cores, lanes, threads_per_model, dt


def load_meta_fixture():
    z = np.load(os.path.join(GOLD, "config2_tables.npz"))
    return json.loads(bytes(z["meta_json"]).decode())


def run_reference(args, rank, world):
    if rank != 0:
        return
    per_step = max(4, os.cpu_count() or 8)
    for _ in range(max(args.warmup, 1)):
        cpu_reference_rate(max(2, (os.cpu_count() or 8) // 4))
    n, dt = 0, 0.0
    for _ in range(args.steps):
        r, cores, lanes, tpm, d = cpu_reference_rate(per_step)
        n += per_step
        dt += d
    v = n / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: Galileon scalar C_l lmax=2500 (params.ini accuracy); "
                                   "%d spectra per step on host cores" % per_step},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d spectra per step, %d models side by side x %d OpenMP threads" % (per_step, lanes, tpm)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=int(os.environ.get("GALCAMB_BENCH_BATCH", "1024")), help="parameter points per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from cosmomc_galileon_b200 import GalCamb, ModelTables
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    fixture = ModelTables.load(os.path.join(GOLD, "config2_tables.npz"))
    meta = load_meta(fixture)
    g = GalCamb(local)
    tmpl = np.fromfile(os.path.join(GOLD, "highL_template.f64")).reshape(7999, 4)[:, :3].T.copy()
    g.set_template(tmpl)
    g.InitSpherBessels(fixture.ls, fixture.d["kmaxfile"], 1.0)
    models = [fixture.copy() for _ in range(B)]
    for m in models:
        m.cstruct()
    lmax = fixture.d["lmax"]
    out = np.zeros((B, 3, lmax - 1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        g.L.galcamb_sync_()

    def exchange(cl_host):
        # the path's one exchange step (parameter-batch mode): all-gather of the final C_l
        if world > 1:
            from cosmomc_galileon_b200.sharding import gather_cls
            return gather_cls(cl_host, world * B, device=torch.device("cuda", local))

    def step_resident():
        g.CalcScalarSources()
        g.SourceToTransfers()
        g.ClTransferToCl()

    # ---------------- resident-input timing (value)
    g.set_models(models)
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    ev = {"evolve": 0.0, "spline_k": 0.0, "los": 0.0, "cls": 0.0}
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_resident()
        tm = g.timings()
        for k in ev:
            ev[k] += tm[k]
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.summary()
    counters = g.counters()
    # ---------------- end-to-end timing through the public call with host tables (e2e)
    for _ in range(min(args.warmup, 1)):
        g.batch_cls(models, out)
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        g.batch_cls(models, out)
        exchange(out)
    barrier()
    dt_e2e = time.perf_counter() - t1
    if world > 1:
        tt = torch.tensor([dt, dt_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_e2e = float(tt[0]), float(tt[1])
    # ---------------- the same batch with desynchronised lanes (N = 1 only; reported next to `value`, not instead of it)
    # Replicas march in lock step: all 32 lanes of a warp accept/reject the same steps.  A real MCMC batch holds
    # different cosmologies whose adaptive step sequences differ (tools/gpu_hetero_probe.py draws one with the oracle
    # as the host pre-stage).  The bench may not run the oracle on this arm, so it perturbs the one thing it can
    # without a pre-stage: each replica's source k grid is scaled by a seeded random factor within +-2 %, which
    # desynchronises the lanes the same way (measured within 4 % of the distinct-cosmology batch, profiles/README.md).
    desync = None
    if world == 1:
        rng = np.random.default_rng(20250101)
        jm = [fixture.copy() for _ in range(B)]
        for m in jm[1:]:
            m.d["k"] = m.d["k"] * (1 + 0.02 * rng.uniform(-1, 1))
        g.set_models(jm)
        step_resident()
        barrier()
        t2 = time.perf_counter()
        nd = max(1, min(2, args.steps))
        for _ in range(nd):
            step_resident()
        barrier()
        dt_d = time.perf_counter() - t2
        _, cl0 = g.Cls(0)
        desync = {"value": B * nd / dt_d, "unit": UNIT, "steps": nd, "ms_per_step": 1e3 * dt_d / nd,
                  "evolve_ms": g.timings()["evolve"],
                  "workload": "same batch, source k grids of replicas 1..B-1 scaled by seeded random factors within +-2 %"}
    # parity guard inside the bench: the last model's C_l against the committed golden
    gold = np.load(os.path.join(GOLD, "config2_golden.npz"))
    cl_err = float(np.max(np.abs(out[-1] / gold["Cl"] - 1)))
    if rank == 0:
        total = world * B * args.steps
        value = total / dt
        e2e = total / dt_e2e
        t_ev = ev["evolve"] / args.steps * 1e-3
        flop = meta["flop_ode"] * B
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        pk = np.zeros(3)
        g.L.galcamb_fp64_peak_(ctypes.c_void_p(pk.ctypes.data))  # FP64 FMA micro-benchmark on this GPU, now
        fp64_peak = float(pk[0]) if pk[0] > 0 else 37.2
        t_los = ev["los"] / args.steps * 1e-3
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: Galileon scalar C_l lmax=2500 (camb/params.ini accuracy), batch of "
                                   "%d parameter points per GPU (replicas of tests/golden/config2_tables.npz)" % B,
                       "batch_per_gpu": B, "l2_note": "inputs larger than L2: per-step working set = %.0f MB of model tables + %.0f MB of ODE "
                       "workspace vs 126 MB L2" % (B * fixture.nbytes() / 1e6, B * 197 * 10.3e-3), "parallelism": "parameter points sharded across ranks; NCCL all-gather of C_l"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(B * fixture.nbytes()),
                    "d2h_bytes_per_step": int(out.nbytes)},
            "gpu_launches": int(args.steps * 6),  # evolve, spline_k, zero_delta, los, cl, cl_interp per step
            "clocks": clocks,
            "roofline": {"kernel": "evolve_sources_kernel", "bound": "fp64", "achieved": flop / t_ev / 1e12, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": flop / t_ev / 1e12 / fp64_peak,
                         "peak_source": "measured: galcamb_fp64_peak_ DFMA micro-benchmark in this run (MEASURED_PEAKS.json has no FP64 figure; nominal 37.2)",
                         "traffic": 3.38e9 * B, "traffic_source": "DRAM bytes of the ncu capture profiles/r1b_evolve_b256_sections.txt "
                         "(441.09 GB/s x 1.96 s / 256 spectra = 3.38 GB per spectrum), scaled to this batch; bytes, per launch",
                         "share_of_step": ev["evolve"] / (ev["evolve"] + ev["spline_k"] + ev["los"] + ev["cls"]),
                         "flop_per_spectrum": meta["flop_ode"], "ms_per_launch": 1e3 * t_ev},
            "roofline_los": {"kernel": "los_kernel", "bound": "hbm", "achieved": meta["bytes_los_operand"] * B / t_los / 1e9,
                             "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s",
                             "frac": meta["bytes_los_operand"] * B / t_los / 1e9 / peaks.get("hbm_gbs", 6650.0),
                             "note": "operand bytes (SURVEY 8d); tables are L2 resident, compulsory bytes/spectrum = %d"
                                     % meta["bytes_los_compulsory"], "ms_per_launch": 1e3 * t_los},
            "stage_ms": {k: v / args.steps for k, v in ev.items()},
            "counters_last_step": counters,
            "parity": {"cl_max_rel_err_vs_golden": cl_err, "tolerance": 1e-5},
        }
        if desync is not None:
            line["desync"] = desync
        if world == 1 and not args.no_cpu:
            cpu_reference_rate(2)  # warm-up + layout choice
            r, cores, lanes, tpm, dtc = cpu_reference_rate(max(8, 2 * (os.cpu_count() or 8)))
            line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%.1f s of oracle runs, %d models side by side x %d OpenMP threads" % (dtc, lanes, tpm)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
