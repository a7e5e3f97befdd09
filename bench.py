#!/usr/bin/env python
"""bench.py -- scalar C_l spectra/sec (Galileon, lmax=2500) on N B200s, beside the host-core CPU baseline.

Workload (config.workload): BASELINE.json configs[1] -- Galileon scalar TT/TE/EE C_l to lmax=2500 with the
camb/params.ini accuracy settings -- as an MCMC proposal batch of B DISTINCT parameter points per GPU (the throughput
form of configs[3]): seeded Gaussian draws around the config-2 point, redrawn until the reference's own stability
checks accept them (cosmomc_galileon_b200/proposal.py).  Every table a point needs (Galileon background, RECFAST,
thermo tables, time / k grids) is built by the product's own host pre-stage (csrc/prestage.cu).
A "step" = one batch through the whole hot path: per-k ODE evolution + source sampling (K1), k-spline (K2),
line-of-sight Bessel integration (K3), C_l assembly + l-interpolation (K4).
  value : spectra/s with the tables of all points already resident in HBM when the timed region starts
  e2e   : spectra/s from PARAMETERS on the host to C_l on the host through the C ABI (galcamb_prestage_async_ /
          galcamb_prestage_wait_ / galcamb_prestaged_to_cls_): host pre-stage, table packing, H2D of every point's
          tables, the four device stages, D2H of every C_l and (N > 1) the NCCL gather of the C_l on rank 0, all inside
          the timed region; the pre-stage of step i+1 runs on the host threads while the device works on step i
Side legs at N = 1 (reported in the same JSON line, not instead of the headline): the same batch size as replicas of one
point, the tables -> C_l call galcamb_batch_cls_, configs[3] as worded (1024 points over 8 GPUs = 128 per GPU),
configs[2] (lensed, AccuracyBoost 2, lmax 5000) and configs[4] (one chain-step call with CosmoMC's settings).
--impl reference times the CPU restatement of the reference path (oracle/, kind "port": the Fortran/icc/GSL reference
cannot be built in this image; built -O3 -march=native as BASELINE.md section 2 prescribes) on the same kind of points.
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
SEED = 20250101
LMAX = 2500


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


def load_meta_fixture():
    z = np.load(os.path.join(GOLD, "config2_tables.npz"))
    return json.loads(bytes(z["meta_json"]).decode())


# ------------------------------------------------------------------------------------------------ CPU reference arm
_cpu = {}


def _oracle_class():
    """The oracle's Python front-end, switched once per process to its -O3 -march=native build."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py  # test infrastructure: allowed here (cpu_baseline / --impl reference)
    if "native" not in _cpu:
        _cpu["native"] = oracle_py.use_native_build()
        if not _cpu["native"]:
            oracle_py.build()
    return oracle_py.Oracle


def cpu_reference_rate(points, threads_per_model=None):
    """Host-core baseline: the oracle on the given parameter points (pre-stage included, like a CAMB_GetResults call).
    Two layouts are tried and the faster one is used (and named in `sample`): OpenMP inside a model like the reference
    (cmbmain.f90:201-206, :263-267) with several models side by side like its MPI x OpenMP job layout
    (cosmomc.sh:13-15), and one single-threaded model per core (best for independent parameter points).  Model objects
    persist between calls so the Bessel tables stay cached as in the reference (camb/bessels.f90:40-41); the first
    call warms them outside the timed region.  Points the stability checks reject cost next to nothing and are not
    counted.  Returns (spectra/s, cores, lanes, threads per model, seconds, spectra computed)."""
    Oracle = _oracle_class()
    cores = os.cpu_count() or 1
    centre = load_meta_fixture()["params"]

    def pool(tpm):
        key = ("models", tpm)
        if key not in _cpu:
            lanes = max(1, cores // tpm)
            _cpu[key] = [Oracle(nthreads=tpm, **centre) for _ in range(lanes)]
            th = [threading.Thread(target=o.run) for o in _cpu[key]]
            for t in th:
                t.start()
            for t in th:
                t.join()
        return _cpu[key]

    def run(tpm, pts):
        models = pool(tpm)
        lanes = len(models)
        done = [0] * lanes
        nxt = [0]
        lock = threading.Lock()

        def work(i, o):
            while True:
                with lock:
                    j = nxt[0]
                    nxt[0] += 1
                if j >= len(pts):
                    return
                try:
                    o.set_point(**pts[j])
                    o.run(reuse_bessels=True)
                    done[i] += 1
                except RuntimeError:
                    pass  # rejected point (unstable background ...): not a spectrum

        th = [threading.Thread(target=work, args=(i, o)) for i, o in enumerate(models)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0, lanes, sum(done)

    if threads_per_model is None:
        if "best_tpm" not in _cpu:
            cand = sorted({min(8, cores), 1})
            rates = {}
            for tpm in cand:
                n = max(2, cores // tpm)
                dt, _, nd = run(tpm, [centre] * n)
                rates[tpm] = nd / dt
            _cpu["best_tpm"] = max(rates, key=rates.get)
        threads_per_model = _cpu["best_tpm"]
    dt, lanes, nd = run(threads_per_model, points)
    return nd / dt, cores, lanes, threads_per_model, dt, nd


def oracle_valid_points(n, seed):
    """n points of the proposal distribution that the reference's stability checks accept, judged by the oracle's own
    background stage (the CPU arm may not touch the product)."""
    from concurrent.futures import ThreadPoolExecutor
    Oracle = _oracle_class()
    from cosmomc_galileon_b200.proposal import valid_points

    def ok(P):
        try:
            Oracle(template=False, **P).stage("background")
            return True
        except RuntimeError:
            return False

    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        pts, drawn = valid_points(n, seed, lambda cand: list(ex.map(ok, cand)))
    return pts, drawn


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 8
    per_step = max(4, cores)
    cpu_reference_rate([load_meta_fixture()["params"]] * 2)  # build, warm the Bessel caches, choose the layout
    pts, drawn = oracle_valid_points(per_step * (args.steps + max(args.warmup, 1)), SEED)
    for w in range(max(args.warmup, 1)):
        cpu_reference_rate(pts[w * per_step:(w + 1) * per_step][: max(2, cores // 4)])
    off = max(args.warmup, 1) * per_step
    n, dt = 0, 0.0
    for s in range(args.steps):
        r, cores, lanes, tpm, d, nd = cpu_reference_rate(pts[off + s * per_step: off + (s + 1) * per_step])
        n += nd
        dt += d
    v = n / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: Galileon scalar C_l lmax=2500 (camb/params.ini accuracy), distinct "
                                   "parameter points of the seeded proposal distribution, parameters -> C_l; %d spectra per "
                                   "step on host cores" % per_step},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d spectra per step, %d models side by side x %d OpenMP threads, oracle built %s"
                                       % (per_step, lanes, tpm, "-O3 -march=native" if _cpu.get("native") else "-O3 (portable)")},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=int(os.environ.get("GALCAMB_BENCH_BATCH", "1024")), help="parameter points per GPU per step")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads per rank (default: cores / ranks)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the N = 1 side legs (replicas, configs[2..4])")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from cosmomc_galileon_b200 import GalCamb, make_params, prestage
    from cosmomc_galileon_b200.proposal import CENTRE, valid_points
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    meta = load_meta_fixture()
    cores = os.cpu_count() or 8
    host_threads = args.host_threads if args.host_threads > 0 else max(1, cores // world)
    g = GalCamb(local)
    g.set_host_threads(host_threads)
    # Galileon backgrounds of a batch on the device (one lane per cosmology, next to the per-k kernel of the previous
    # batch) when the rank has few host cores: with 4 threads the pre-stage of 1024 points takes 6.2 s on the host alone
    # (longer than the device step) and 4.1 s this way; with 16 threads it makes no difference (2.4 s either way).
    device_background = host_threads < 12 and not os.environ.get("GALCAMB_BENCH_HOST_BACKGROUND")
    if device_background:
        g.set_prestage_device(local)
    tmpl = np.fromfile(os.path.join(GOLD, "highL_template.f64")).reshape(7999, 4).T.copy()
    g.set_template(tmpl)

    def to_params(P, **kw):
        q = dict(P)
        q.update(kw)
        return make_params(Max_l=q.pop("Max_l", LMAX), Max_eta_k=q.pop("Max_eta_k", 2.0 * LMAX), **q)

    # ---------------- the batch: B distinct points per rank that the stability checks accept (untimed set-up)
    def accept(cand):
        _, st = prestage([to_params(P) for P in cand], nthreads=host_threads, copy=False)
        return [s == 0 for s in st]

    t_setup = time.perf_counter()
    points, drawn = valid_points(B, SEED + 7919 * rank, accept)
    params = [to_params(P) for P in points]
    t_setup = time.perf_counter() - t_setup
    out = np.zeros((B, 3, LMAX - 1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        g.L.galcamb_sync_()

    def exchange(cl_host):
        # the path's one exchange step (parameter-batch mode): the final C_l go to the rank that feeds the
        # likelihoods, device to device over NVLink; no other rank copies anything back
        if world > 1:
            from cosmomc_galileon_b200.sharding import gather_cls_to_root
            return gather_cls_to_root(cl_host, world * B, device=torch.device("cuda", local))

    def step_resident():
        g.CalcScalarSources()
        g.SourceToTransfers()
        g.ClTransferToCl()

    # ---------------- resident-input timing (value)
    t_pre0 = time.perf_counter()
    tabs, st = prestage(params, nthreads=host_threads, copy=False)
    t_pre = time.perf_counter() - t_pre0
    assert all(s == 0 for s in st)
    table_bytes = int(sum(t.nbytes() for t in tabs))
    del tabs
    g.prestaged_upload()
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
    # ---------------- end to end from parameters (e2e): pre-stage of step i+1 on the host while the device runs step i
    # Steady state of the pipeline: the batch of the first timed step is pre-staged during the warm-up, and EVERY timed
    # step carries one complete pre-stage (of the following batch) on the host threads behind its device work.
    g.params_to_cls(params, nthreads=host_threads, out=out)  # warm-up of the whole call path
    exchange(out)  # communicator set-up and buffer registration of the gather belong to the warm-up
    g.prestage_async(params, host_threads)
    g.prestage_wait()
    barrier()
    t1 = time.perf_counter()
    e2e_ms = {"to_cls": 0.0, "exchange": 0.0, "wait_for_prestage": 0.0}
    for i in range(args.steps):
        ta = time.perf_counter()
        g.prestage_async(params, host_threads)
        g.prestaged_to_cls(out)
        tb = time.perf_counter()
        exchange(out)
        tc = time.perf_counter()
        g.prestage_wait()
        td = time.perf_counter()
        e2e_ms["to_cls"] += 1e3 * (tb - ta) / args.steps
        e2e_ms["exchange"] += 1e3 * (tc - tb) / args.steps
        e2e_ms["wait_for_prestage"] += 1e3 * (td - tc) / args.steps
    barrier()
    dt_e2e = time.perf_counter() - t1
    e2e_ms["library_phases_last_step"] = g.host_timings()
    e2e_ms["device_stages_last_step"] = g.timings()
    if world > 1:
        tt = torch.tensor([dt, dt_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_e2e = float(tt[0]), float(tt[1])
    # parity guard inside the bench: point 0 of every rank is the config-2 centre, whose C_l are a committed golden
    gold = np.load(os.path.join(GOLD, "config2_golden.npz"))
    cl_err = float(np.max(np.abs(out[0] / gold["Cl"] - 1)))
    n_nan = int(np.sum(np.isnan(out[:, 0, 0])))

    extras = {}
    if world == 1 and not args.no_extras:
        # (a) the same batch size as REPLICAS of one point: every lane of a warp takes identical accept/reject decisions
        reps = [to_params(CENTRE) for _ in range(B)]
        prestage(reps, nthreads=host_threads, copy=False)
        g.prestaged_upload()
        step_resident()
        barrier()
        t2 = time.perf_counter()
        nrep = max(1, min(2, args.steps))
        for _ in range(nrep):
            step_resident()
        barrier()
        d2 = time.perf_counter() - t2
        extras["replicas"] = {"value": B * nrep / d2, "unit": UNIT, "steps": nrep, "evolve_ms": g.timings()["evolve"],
                              "workload": "%d replicas of the config-2 point (tables resident)" % B}
        # (b) tables -> C_l through galcamb_batch_cls_ (the boundary a Fortran host that keeps its own pre-stage calls)
        tabs, _ = prestage(params[: min(B, 256)], nthreads=host_threads, copy=True)
        o2 = np.zeros((len(tabs), 3, LMAX - 1))
        g.batch_cls(tabs, o2)
        barrier()
        t3 = time.perf_counter()
        g.batch_cls(tabs, o2)
        barrier()
        d3 = time.perf_counter() - t3
        extras["tables_to_cls"] = {"value": len(tabs) / d3, "unit": UNIT, "batch": len(tabs),
                                   "h2d_bytes": int(sum(t.nbytes() for t in tabs)), "d2h_bytes": int(o2.nbytes),
                                   "workload": "galcamb_batch_cls_ on host tables of %d distinct points" % len(tabs)}
        del tabs
        # (c) BASELINE configs[3] as worded: 1024 points over 8 GPUs = 128 points per GPU and step, parameters -> C_l
        o3 = np.zeros((128, 3, LMAX - 1))
        g.params_to_cls(params[:128], nthreads=host_threads, out=o3)
        g.prestage_async(params[:128], host_threads)
        g.prestage_wait()
        barrier()
        t4 = time.perf_counter()
        for i in range(3):  # steady state, as in the headline loop
            g.prestage_async(params[:128], host_threads)
            g.prestaged_to_cls(o3)
            g.prestage_wait()
        barrier()
        d4 = time.perf_counter() - t4
        extras["config3_128_per_gpu"] = {"value": 3 * 128 / d4, "unit": UNIT, "ms_per_step": 1e3 * d4 / 3,
                                         "evolve_ms": g.timings()["evolve"],
                                         "workload": "128 distinct points per GPU and step (1024 over 8 GPUs), parameters -> C_l"}
        # (d) BASELINE configs[2]: lensed C_l, AccuracyBoost 2, lmax 5000 (every stage incl. the lensing post-step)
        n2 = 16
        p2 = [to_params(P, DoLensing=1, AccuracyBoost=2.0, Max_l=5000, Max_eta_k=10000.0)
              for P in points[:n2]]
        t2tabs, st2 = prestage(p2, nthreads=host_threads, copy=True)
        t2tabs = [t for t in t2tabs if t is not None]
        if t2tabs:
            g.batch_lensed_cls(t2tabs)
            barrier()
            t5 = time.perf_counter()
            g.batch_lensed_cls(t2tabs)
            barrier()
            d5 = time.perf_counter() - t5
            extras["config2_lensed"] = {"value": len(t2tabs) / d5, "unit": "lensed spectra/s", "batch": len(t2tabs),
                                        "stage_ms": {k: float(v) for k, v in g.timings().items() if k in ("evolve", "los", "cls", "lensing")},
                                        "workload": "Galileon lensed C_l, AccuracyBoost 2, lmax 5000, tables -> lensed C_l"}
            del t2tabs
        # (e) BASELINE configs[4]: one chain-step call with CosmoMC's fixed settings: Max_l 2650,
        # lensing on, DoLateRadTruncation F, three degenerate massive neutrinos; parameters -> lensed C_l, ms per call
        pc = to_params(CENTRE, DoLensing=1, DoLateRadTruncation=0, Max_l=2650, Max_eta_k=6625.0, Num_Nu_massive=3,
                       Num_Nu_massless=0.046)
        lat = []
        for _ in range(4):
            barrier()
            t6 = time.perf_counter()
            tb, _ = prestage([pc], nthreads=1, copy=True)
            g.batch_lensed_cls(tb)
            barrier()
            lat.append(1e3 * (time.perf_counter() - t6))
        extras["latency"] = {"ms_per_call": float(np.median(lat[1:])), "evolve_ms": float(g.timings()["evolve"]),
                             "los_ms": float(g.timings()["los"]), "lensing_ms": float(g.timings()["lensing"]),
                             "workload": "one CAMB_GetResults-equivalent call, CosmoMC settings, parameters -> lensed C_l"}

    if rank == 0:
        total = world * B * args.steps
        value = total / dt
        e2e = total / dt_e2e
        t_ev = ev["evolve"] / args.steps * 1e-3
        # algorithmic FLOPs of SURVEY 8(d): the oracle-counted figure of the config-2 point scaled by the number of
        # right-hand-side evaluations the reference algorithm takes on THIS batch (counted by the kernel)
        flop = meta["flop_ode"] * counters["n_derivs"] / meta["n_derivs"]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        pk = np.zeros(3)
        g.L.galcamb_fp64_peak_(ctypes.c_void_p(pk.ctypes.data))  # FP64 FMA micro-benchmark on this GPU, now
        fp64_peak = float(pk[0]) if pk[0] > 0 else 37.2
        t_los = ev["los"] / args.steps * 1e-3
        traffic, traffic_src = None, "no ncu capture on file"
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r2_k1_traffic.json")))
            traffic = tr["dram_bytes_per_spectrum"] * B  # measured at batch tr["batch"]; scaled to this batch per spectrum
            traffic_src = tr["source"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: Galileon scalar C_l lmax=2500 (camb/params.ini accuracy), batch of %d DISTINCT "
                                   "parameter points per GPU (seeded proposal draws around the config-2 point, %d drawn for %d "
                                   "accepted by the stability checks; tables from the product pre-stage)" % (B, drawn, B),
                       "batch_per_gpu": B,
                       "l2_note": "inputs larger than L2: per-step working set = %.0f MB of model tables + %.0f MB of ODE workspace vs 126 MB L2"
                                  % (table_bytes / 1e6, B * 197 * 10.3e-3),
                       "parallelism": "parameter points sharded across ranks; NCCL gather of C_l on rank 0",
                       "host_threads_per_rank": host_threads, "prestage_s_per_batch": t_pre,
                       "galileon_backgrounds_on": "device" if device_background else "host threads"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": table_bytes, "d2h_bytes_per_step": int(out.nbytes),
                    "what": "parameters -> C_l in the steady state of the pipeline: every timed step runs one complete host "
                            "pre-stage (of the following batch, on the host threads behind the device work), packing, H2D, "
                            "K1-K4, D2H" + (", NCCL gather to rank 0" if world > 1 else "")},
            "e2e_phases_ms": e2e_ms,
            "gpu_launches": int(args.steps * 6),  # evolve, spline_k, zero_delta, los, cl, cl_interp per step
            "clocks": clocks,
            "roofline": {"kernel": "evolve_batch_kernel", "bound": "fp64", "achieved": flop / t_ev / 1e12, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": flop / t_ev / 1e12 / fp64_peak,
                         "peak_source": "measured: galcamb_fp64_peak_ DFMA micro-benchmark in this run (MEASURED_PEAKS.json has no FP64 figure; nominal 37.2)",
                         "traffic": traffic, "traffic_source": traffic_src,
                         "share_of_step": ev["evolve"] / (ev["evolve"] + ev["spline_k"] + ev["los"] + ev["cls"]),
                         "flop_per_spectrum": flop / B, "ms_per_launch": 1e3 * t_ev},
            "roofline_los": {"kernel": "los_kernel", "bound": "hbm", "achieved": meta["bytes_los_operand"] * counters["n_iter"] / meta["n_iter"] / t_los / 1e9,
                             "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s",
                             "frac": meta["bytes_los_operand"] * counters["n_iter"] / meta["n_iter"] / t_los / 1e9 / peaks.get("hbm_gbs", 6650.0),
                             "note": "operand bytes (SURVEY 8d); tables are L2 resident, compulsory bytes/spectrum = %d"
                                     % meta["bytes_los_compulsory"], "ms_per_launch": 1e3 * t_los},
            "stage_ms": {k: v / args.steps for k, v in ev.items()},
            "counters_last_step": counters,
            "parity": {"cl_max_rel_err_vs_golden": cl_err, "tolerance": 1e-5, "nan_rows": n_nan},
        }
        line.update(extras)
        if world == 1 and not args.no_cpu:
            pts, _ = oracle_valid_points(max(8, 2 * cores) + 2, SEED)
            cpu_reference_rate(pts[:2])  # warm-up + layout choice
            r, ccores, lanes, tpm, dtc, nd = cpu_reference_rate(pts[2:])
            if "latency" in line:
                # the same chain-step call on the host: one model, OpenMP over k like the reference (cmbmain.f90:201-206,
                # :263-267, lensing.f90:288), all cores; parameters -> lensed C_l, Bessel tables cached as in the reference
                Oracle = _oracle_class()
                oc = Oracle(nthreads=cores, DoLensing=1, DoLateRadTruncation=0, lmax=2650, Num_Nu_massive=3, Num_Nu_massless=0.046,
                            **meta["params"])
                tc = []
                for _ in range(4):
                    t7 = time.perf_counter()
                    oc.run(reuse_bessels=True)
                    oc.stage("lensing")
                    tc.append(1e3 * (time.perf_counter() - t7))
                line["latency"]["cpu_ms_per_call"] = float(np.median(tc[1:]))
                line["latency"]["cpu_threads"] = cores
            line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": ccores, "kind": "port",
                                    "sample": "%.1f s of oracle runs (%d distinct points, parameters -> C_l), %d models side by side x %d "
                                              "OpenMP threads, oracle built %s" % (dtc, nd, lanes, tpm, "-O3 -march=native" if _cpu.get("native") else "-O3 (portable)")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
