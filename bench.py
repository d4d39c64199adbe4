#!/usr/bin/env python
"""bench.py -- likelihood evals/sec + wall-time-to-logZ on BASELINE.json's metric config.

A "step" is ONE complete nested-sampling run of the workload (BASELINE config[1]: 20-D Gaussian,
nlive=1000, num_repeats=40, nDerived=2, precision_criterion=1e-3) from live-point generation to
the final kill-off.  evals = the reference's nlike counter (calculate.f90:44).

  value     evals/s with everything resident on the device: pc_run(), no dumper, timed with the
            CUDA events the engine records around its persistent-kernel launches.
  e2e       the same metric through the reference-facing C ABI, polychord_c_interface(), with a
            dumper callback that receives HOST arrays (live/dead/logweights) at every update and
            at the end; host wall time, every host<->device copy inside the timed region.
  roofline  the dispatched run kernel (its name comes from the run): algorithmic bytes (DESIGN.md:
            8T+8D per slice step, 16T per chain, 8D^2 per generation) / CUDA-event kernel time vs
            MEASURED_PEAKS.json hbm_gbs; `traffic` from the committed ncu capture of the same workload.
  ensemble  the device filled with independent runs of the workload in ONE launch (dense chain phase),
            with its own roofline block -- the figure that fills the machine.
  configs   the other single-GPU BASELINE configurations (C1, C3, C4), one complete run each on the
            device, with a bounded CPU sample beside each.
  cpu_baseline  the CPU oracle (oracle/pc_oracle.cpp, reference schedule, 1 thread) on the box's host.

N > 1 (torchrun): ONE run with nlive*N live points sharded over the GPUs; during warm-up one seed is also
run on rank 0's GPU alone and must give the same run ("sharded_parity").

`--impl reference` times the CPU restatement of the reference's linear-mode algorithm (the Fortran
reference cannot be built in this image: no gfortran/MPI) on the same workload (nlive*N for --gpus N).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

WORKLOADS = {
    # name: (nDims, nDerived, nlive, num_repeats, like, prior box)
    "gaussian20_nlive1000_R40": dict(nDims=20, nDerived=2, nlive=1000, num_repeats=40, like="gaussian", box=None),
    "gaussian20_nlive500_R40": dict(nDims=20, nDerived=2, nlive=500, num_repeats=40, like="gaussian", box=None),
    "rastrigin10_nlive2000_R50": dict(nDims=10, nDerived=0, nlive=2000, num_repeats=50, like="rastrigin", box=5.12,
                                      clustering=True),
    "gaussian20_nlive8000_R40": dict(nDims=20, nDerived=2, nlive=8000, num_repeats=40, like="gaussian", box=None),
    # BASELINE config 4: random_gaussian.f90 (mu = 0.5, sigma_j = 0.1 * 0.01^((j-1)/(D-1)), Haar-random basis), R = 5 D
    "corr_gaussian50_nlive4000_R250": dict(nDims=50, nDerived=0, nlive=4000, num_repeats=250, like="corr_gaussian", box=None),
}
# analytic evidences
TRUE_LOGZ = {"gaussian20_nlive1000_R40": -1.15e-5, "gaussian20_nlive500_R40": -1.15e-5, "gaussian20_nlive8000_R40": -1.15e-5,
             "rastrigin10_nlive2000_R50": -23.2630, "corr_gaussian50_nlive4000_R250": 0.0}


def like_params(w):
    """Synthetic likelihood parameters of a workload (None: the built-in defaults)."""
    if w["like"] != "corr_gaussian":
        return None
    import numpy as np
    D = w["nDims"]
    rng = np.random.default_rng(0)
    Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    sig = float(np.float32(0.1)) * 0.01 ** (np.arange(D) / (D - 1))      # random_gaussian.f90:5: sigma is a single-precision literal
    invcov = (Q / sig ** 2) @ Q.T
    return np.concatenate([np.full(D, 0.5), invcov.flatten(order="F"), [2 * np.log(sig).sum()]])


def flops_per_eval(w):
    """Whole-evaluation flops of a workload's likelihood (a cos counted as one)."""
    D = w["nDims"]
    return {"gaussian": 11.0 * D, "rastrigin": 10.0 * D + D, "corr_gaussian": 2.0 * D * D + 7.0 * D}[w["like"]]


def committed_traffic(tag):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture of this workload."""
    for name in (f"r02_ncu_{tag}.json", f"r02_ncu_{tag}_single.json"):
        p = ROOT / "profiles" / name
        if p.exists():
            try:
                return float(json.loads(p.read_text())["dram_bytes_per_launch"]), f"profiles/{name}"
            except (KeyError, ValueError):
                pass
    return None, None
METRIC = "likelihood_evals_per_sec"
UNIT = "evals/s"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).

    Sampled in-process through NVML (the library nvidia-smi itself reads): a looping `nvidia-smi -lms`
    child holds driver locks for milliseconds at a time and was measured to stretch the host side of a
    17 ms run to 20-115 ms, i.e. it perturbs exactly the end-to-end number it is supposed to vouch for.
    Falls back to the nvidia-smi loop when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, uuid=None, period_s=0.02):
        self.index, self.uuid, self.period = index, uuid, period_s
        self.proc = None
        self.lines = []
        self.samples = []   # (sm_mhz, reasons bitmask)
        self.nv = None
        self.stop_flag = threading.Event()

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            if self.uuid:
                try:
                    u = self.uuid if str(self.uuid).startswith("GPU-") else "GPU-" + str(self.uuid)
                    h = nv.nvmlDeviceGetHandleByUUID(u.encode() if hasattr(u, "encode") else u)
                except Exception:
                    h = None
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.nv, self.h = nv, h
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((mhz, rs))
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nv is not None:
            self.stop_flag.set()
            self.t.join(timeout=2)
            nv = self.nv
            names = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40),
                     ("sw_thermal_slowdown", 0x20), ("hw_power_brake_slowdown", 0x80))
            reasons = set()
            for _, rs in self.samples:
                for name, bit in names:
                    if rs & bit:
                        reasons.add(name)
            sm = [m for m, _ in self.samples]
            try:
                nv.nvmlShutdown()
            except Exception:
                pass
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(reasons), "samples": len(sm), "source": "nvml, in-process, every 20 ms"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def _oracle_run(workload, seed, nlive=None, max_ndead=-1):
    """One oracle run in the reference schedule (one death + one birth per iteration)."""
    import oracle_lib as O
    w = WORKLOADS[workload]
    kw = dict(prior_lo=[-w["box"]] * w["nDims"], prior_hi=[w["box"]] * w["nDims"]) if w["box"] else {}
    lp = like_params(w)
    if lp is not None:
        kw["like_params"] = lp
    s = O.make_settings(w["nDims"], w["nDerived"], nlive=nlive or w["nlive"], num_repeats=w["num_repeats"], seed=seed,
                        batch_K=0, do_clustering=w.get("clustering", False), max_ndead=max_ndead)
    t0 = time.perf_counter()
    r, _ = O.run(s, like=w["like"], **kw)
    return r, time.perf_counter() - t0


def _reference_replica(job):
    """One oracle run in a worker process (run_reference's all_cores leg); returns its likelihood evaluations."""
    workload, seed, nlive, max_ndead = job
    return _oracle_run(workload, seed, nlive, max_ndead)[0].nlike


def cpu_sample(workload, nlive, budget_s=8.0):
    """A bounded CPU sample of a workload: complete runs when one fits the budget, else runs cut at max_ndead deaths
    (evals/s does not depend on where a run is cut).  Returns (evals/s, description, seconds per complete run or None)."""
    w = WORKLOADS[workload]
    # ~3e6 evals/s for the 20-D Gaussian; cost per evaluation grows with the likelihood's flops
    est_evals_per_s = 3.0e6 * 220.0 / max(flops_per_eval(w), 220.0)
    est_run_evals = nlive * 30.0 * w["num_repeats"] * 5.0
    cut = -1
    if est_run_evals / est_evals_per_s > budget_s:
        cut = max(2 * nlive // 10, int(budget_s * est_evals_per_s / (w["num_repeats"] * 5.0)))
    tot_e, tot_t, n = 0, 0.0, 0
    while tot_t < 0.6 * budget_s and n < 3:
        r, t = _oracle_run(workload, n, nlive, cut)
        tot_e += r.nlike; tot_t += t; n += 1
    what = (f"{n} complete run(s)" if cut < 0 else f"{n} run(s) cut at max_ndead={cut}") + \
           f" of the workload at nlive={nlive}, oracle reference schedule (1 death/iteration), 1 host thread of " \
           f"{os.cpu_count()}; {tot_t:.1f} s of CPU work"
    return tot_e / tot_t, what, (tot_t / n if cut < 0 else None)


def run_reference(args):
    """The reference's CPU path (restated: oracle, reference schedule batch_K=0) on the host cores.  `value` is one
    thread on one run (the linear mode is single-threaded); `all_cores` fills every core with an independent run, the
    only way that mode uses a box without MPI.  For --gpus N the workload is the sharded arm's: nlive * N."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    nlive = w["nlive"] * max(1, args.gpus)
    # each step is a bounded sample: a complete run when it takes a few seconds, else a run cut at max_ndead
    probe, t_probe = _oracle_run(args.workload, 999, nlive, max(200, nlive // 5))
    rate = probe.nlike / t_probe
    est_full = nlive * 30.0 * w["num_repeats"] * 5.0 / rate
    per_step_budget = max(1.0, 150.0 / max(1, args.steps + args.warmup))
    cut = -1 if est_full <= per_step_budget else max(nlive // 5, int(per_step_budget * rate / (w["num_repeats"] * 5.0)))
    for i in range(args.warmup):
        _oracle_run(args.workload, 1000 + i, nlive, max(200, nlive // 5))
    tot_e, tot_t, lz = 0, 0.0, []
    for i in range(args.steps):
        r, t = _oracle_run(args.workload, i, nlive, cut)
        tot_e += r.nlike; tot_t += t; lz.append(r.logZ)
    v = tot_e / tot_t
    ncores = os.cpu_count() or 1
    all_cores = None
    if ncores > 1 and not args.no_all_cores:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(ncores) as pool:
            t0 = time.perf_counter()
            res = pool.map(_reference_replica, [(args.workload, 2000 + i, nlive, cut) for i in range(ncores)])
            ta = time.perf_counter() - t0
        all_cores = {"value": sum(res) / ta, "unit": UNIT, "cores": ncores, "seconds": ta,
                     "sample": f"{ncores} independent runs of the workload, one per host core, started together"}
    sample = (f"{args.steps} complete runs (seeds 0..{args.steps - 1})" if cut < 0 else
              f"{args.steps} runs (seeds 0..{args.steps - 1}) cut at max_ndead={cut}") + \
             f" of the workload at nlive={nlive}, oracle/pc_oracle.cpp reference schedule, 1 host thread; the Fortran " \
             "reference cannot be built here (no gfortran/MPI) and its linear mode is single-threaded"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload + (f"_x{args.gpus}_sharded" if args.gpus > 1 else ""),
                   **{k: w[k] for k in ("nDims", "nDerived", "num_repeats")}, "nlive": nlive,
                   "precision_criterion": 1e-3, "schedule": "reference: one death + one birth per iteration"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_time_to_logZ_s": (tot_t / args.steps) if cut < 0 else None, "logZ_mean": (sum(lz) / len(lz)) if cut < 0 else None,
        "host_cores": os.cpu_count(), "all_cores": all_cores,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-all-cores", action="store_true", help="reference arm: skip the one-run-per-core ensemble leg")
    ap.add_argument("--workload", default="gaussian20_nlive1000_R40", choices=sorted(WORKLOADS))
    ap.add_argument("--batch-fraction", type=float, default=None)
    ap.add_argument("--warps-per-cta", type=int, default=None)
    ap.add_argument("--ensemble", type=int, default=74, help="replicas for the ensemble-throughput figure (0=skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the block of the other BASELINE configurations")
    ap.add_argument("--replicas", action="store_true",
                    help="N>1: N independent runs of the workload (one per GPU, no exchange) instead of ONE run "
                         "sharded over the GPUs with nlive scaled by N")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from polychordlite_b200 import _capi as capi

    rank, world, local = dist_env()
    if not torch.cuda.is_available() or capi.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    capi.set_option("device", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.batch_fraction is not None:
        capi.set_option("batch_fraction", args.batch_fraction)
    if args.warps_per_cta is not None:
        capi.set_option("warps_per_cta", args.warps_per_cta)

    w = WORKLOADS[args.workload]
    D, P, n, R = w["nDims"], w["nDerived"], w["nlive"], w["num_repeats"]
    clustering = bool(w.get("clustering", False))
    # N > 1 (weak scaling, BASELINE config 5's shape): ONE run with nlive*N live points sharded over the N GPUs --
    # every rank runs 1/N of each generation's chains, the new live points and the covariance statistics are
    # exchanged over NVLink inside the persistent kernel (polychordlite_b200/mgpu.py)
    sharded = world > 1 and not args.replicas
    if sharded:
        n = n * world
    box = dict(prior_lo=[-w["box"]] * D, prior_hi=[w["box"]] * D) if w["box"] else {}
    lp = like_params(w)
    if lp is not None:
        box["like_params"] = lp
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    peak, which = peaks()

    def settings(seed, nlive=None):
        return capi.make_settings(D, P, nlive=nlive or n, num_repeats=R, seed=seed, do_clustering=clustering)

    # ---- sharded parity (N > 1): the same seed alone on rank 0's GPU and sharded over all of them ---------------
    sharded_parity = None
    if sharded:
        from polychordlite_b200 import mgpu
        alone = None
        if rank == 0:   # (the same batch size as the sharded run: the automatic one follows the number of devices)
            capi.set_option("batch_K", capi.auto_batch_size(n, world))
            try:
                alone, _ = capi.run(settings(424242), like=w["like"], **box)
            finally:
                capi.set_option("batch_K", 0)
        dist.barrier()
        mgpu.attach(settings(0))
        both, _ = capi.run(settings(424242), like=w["like"], **box)
        if rank == 0:
            sharded_parity = bool(both.ndead == alone.ndead and both.ngenerations == alone.ngenerations and
                                  both.nupdates == alone.nupdates and abs(both.logZ - alone.logZ) < 1e-8)
            nl = torch.tensor([int(both.nlike)], dtype=torch.int64, device="cuda")
        else:
            nl = torch.tensor([int(both.nlike)], dtype=torch.int64, device="cuda")
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)
        if rank == 0:
            sharded_parity = sharded_parity and int(nl.item()) == int(alone.nlike)

    def step_device(seed):
        flush.zero_()
        torch.cuda.synchronize()
        if sharded:
            dist.barrier()
        info, _ = capi.run(settings(seed), like=w["like"], **box)
        return info

    # ---- the reference-facing call with host buffers -----------------------------------------
    L = capi.lib()
    like_fn = {"gaussian": L.pc_gaussian_loglikelihood, "rastrigin": L.pc_rastrigin_loglikelihood,
               "corr_gaussian": L.pc_corr_gaussian_loglikelihood}[w["like"]]
    if lp is not None:
        q = np.ascontiguousarray(lp, dtype=np.float64)
        L.pc_register_device_likelihood(C.cast(like_fn, capi.LL_CB), 2, q.ctypes.data_as(C.POINTER(C.c_double)), q.size)
    prior_fn = L.pc_uniform_prior if w["box"] else L.pc_unit_prior
    if w["box"]:
        pp = np.array(box["prior_lo"] + box["prior_hi"], dtype=np.float64)
        L.pc_register_device_prior(C.cast(prior_fn, capi.PRIOR_CB), 0, pp.ctypes.data_as(C.POINTER(C.c_double)), pp.size)
    sink = {"rows": 0, "logZ": None, "calls": 0}

    def _dumper(ndead, nlive, npars, live, dead, lw, logZ, logZerr):
        # wrap the host arrays without copying, like pypolychord's shim does (PyArray_SimpleNewFromData,
        # _pypolychord.cpp:88-94), and touch them
        if ndead:
            d = np.frombuffer((C.c_double * (ndead * npars)).from_address(C.addressof(dead.contents)),
                              dtype=np.float64).reshape(ndead, npars)
            ww = np.frombuffer((C.c_double * ndead).from_address(C.addressof(lw.contents)), dtype=np.float64)
            sink["last_logL"] = float(d[ndead - 1, npars - 1])
            sink["last_logw"] = float(ww[ndead - 1])
        sink["rows"] = ndead
        sink["logZ"] = logZ
        sink["calls"] += 1

    dcb = capi.DUMPER_CB(_dumper)
    grade_frac = (C.c_double * 1)(1.0)
    grade_dims = (C.c_int * 1)(D)
    comm = C.c_int(0)
    L.polychord_c_interface.restype = None
    L.polychord_c_interface.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_bool, C.c_int, C.c_double,
        C.c_double, C.c_int, C.c_double] + [C.c_bool] * 11 + [C.c_double, C.c_bool, C.c_int, C.c_int, C.c_char_p,
                                                                C.c_char_p, C.c_int, C.POINTER(C.c_double),
                                                                C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double),
                                                                C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]

    def step_e2e(seed):
        flush.zero_()
        torch.cuda.synchronize()
        if sharded:
            dist.barrier()
        t0 = time.perf_counter()
        # sharded run: every rank makes the call (like the ranks of an MPI run), rank 0 carries the dumper
        L.polychord_c_interface(C.cast(like_fn, C.c_void_p), C.cast(prior_fn, C.c_void_p),
                                C.cast(dcb, C.c_void_p) if (rank == 0 or not sharded) else None,
                                n, R, -1, -1, clustering, 0, 1e-3, -1e30, -1, 0.0,
                                False, False, False, False, False, False, False, False, False, False, False,
                                float(np.exp(-1)), True, D, P, b"chains", b"bench", 1, grade_frac, grade_dims, 0, None,
                                None, seed, C.byref(comm))
        t = time.perf_counter() - t0
        return capi.last_run_info(), t

    # ---- warm-up --------------------------------------------------------------------------------
    roff = 0 if sharded else 1  # a sharded run needs the same seed on every rank, replicas need different ones
    for i in range(args.warmup):
        step_device(10_000 + i + 100 * rank * roff)
        step_e2e(20_000 + i + 100 * rank * roff)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- timed region 1: device-resident ------------------------------------------------------
    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid=uuid)
    sampler.start()
    barrier()
    evals = launches = 0
    dev_ms, wall, algo_bytes, logZs, logZerrs, ndead = 0.0, 0.0, 0, [], [], 0
    t_region = time.perf_counter()
    for i in range(args.steps):
        t0 = time.perf_counter()
        info = step_device(i + 1000 * rank * roff)
        info_dev = info
        wall += time.perf_counter() - t0
        evals += info.nlike; dev_ms += info.device_ms + info.cluster_ms; launches += info.kernel_launches
        algo_bytes += info.algorithmic_bytes; logZs.append(info.logZ); logZerrs.append(info.logZerr); ndead += info.ndead
    barrier()
    t_region = time.perf_counter() - t_region
    # ---- timed region 2: end to end through polychord_c_interface ----------------------------
    e_evals, e_t, h2d, d2h, e_launch = 0, 0.0, 0, 0, 0
    barrier()
    for i in range(args.steps):
        info, t = step_e2e(i + 1000 * rank * roff)
        e_evals += info.nlike; e_t += t; h2d += info.h2d_bytes; d2h += info.d2h_bytes; e_launch += info.kernel_launches
    barrier()
    clocks = sampler.stop()

    fp64_peak = capi.measure_fp64_tflops() if rank == 0 else None

    def roof(info_list, ms, wl, tag):
        """HBM and FP64 roofline blocks of a set of runs that took `ms` of device time in all."""
        ab = sum(i.algorithmic_bytes for i in info_list)
        ev = sum(i.nlike for i in info_list)
        ach = ab / (ms * 1e-3) / 1e9
        tr, src = committed_traffic(tag)
        f64 = ev / (ms * 1e-3) * flops_per_eval(wl) / 1e12
        return ({"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": tr,
                 "traffic_source": src, "algorithmic_bytes_per_launch": ab, "peak_source": which,
                 "kernel": info_list[0].kernel},
                {"bound": "fp64_fma", "achieved": f64, "peak": fp64_peak, "unit": "TFLOP/s",
                 "frac": f64 / fp64_peak if fp64_peak else None, "flops_per_eval": flops_per_eval(wl),
                 "peak_source": "measured here (pc_measure_fp64_tflops: FMA microbenchmark)"})

    # ---- the device filled with independent runs (dense chain phase) -----------------------------
    ens = None
    if args.ensemble > 0 and world == 1 and w["like"] != "corr_gaussian":
        try:
            capi.run_ensemble(settings(0), list(range(5000, 5000 + args.ensemble)), like=w["like"], **box)
            flush.zero_(); torch.cuda.synchronize()
            infos = capi.run_ensemble(settings(0), list(range(args.ensemble)), like=w["like"], **box)
            ms = infos[0].device_ms
            lz = [i.logZ for i in infos]
            rh, rf = roof(infos, ms, w, args.workload + "_ensemble")
            ens = {"replicas": args.ensemble, "value": sum(i.nlike for i in infos) / (ms * 1e-3), "unit": UNIT, "device_ms": ms,
                   "ms_per_run_amortised": ms / args.ensemble,
                   "logZ_mean": float(np.mean(lz)), "logZ_sem": float(np.std(lz, ddof=1) / np.sqrt(len(lz))),
                   "logZerr_mean_reported": float(np.mean([i.logZerr for i in infos])),
                   "batch_K": int(infos[0].batch_K), "ctas_per_run": int(infos[0].ctas_per_run),
                   "warps_per_cta": int(infos[0].warps_per_cta), "roofline": rh, "arithmetic_roofline": rf}
        except RuntimeError as ex:  # e.g. too many replicas for one launch
            ens = {"error": str(ex)}

    # ---- the other single-GPU BASELINE configurations: one complete run each -------------------
    configs = None
    if world == 1 and not args.no_configs and args.workload == "gaussian20_nlive1000_R40":
        configs = []
        for name in ("gaussian20_nlive500_R40", "rastrigin10_nlive2000_R50", "corr_gaussian50_nlive4000_R250"):
            cw = WORKLOADS[name]
            ckw = dict(prior_lo=[-cw["box"]] * cw["nDims"], prior_hi=[cw["box"]] * cw["nDims"]) if cw["box"] else {}
            clp = like_params(cw)
            if clp is not None:
                ckw["like_params"] = clp
            cs = lambda seed: capi.make_settings(cw["nDims"], cw["nDerived"], nlive=cw["nlive"], num_repeats=cw["num_repeats"],
                                                 seed=seed, do_clustering=bool(cw.get("clustering", False)))
            try:
                capi.run(cs(77), like=cw["like"], **ckw)   # warm-up (allocations, code load)
                flush.zero_(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                ci, _ = capi.run(cs(1), like=cw["like"], **ckw)
                cwall = time.perf_counter() - t0
                ms = ci.device_ms + ci.cluster_ms
                rh, rf = roof([ci], ms, cw, name)
                entry = {"workload": name, "value": ci.nlike / (ms * 1e-3), "unit": UNIT, "ms_per_run": ms,
                         "wall_ms_per_run": 1e3 * cwall, "clustering_ms": ci.cluster_ms, "logZ": ci.logZ, "logZerr": ci.logZerr,
                         "logZ_true": TRUE_LOGZ[name], "ndead": int(ci.ndead), "evals": int(ci.nlike),
                         "ncluster_max": int(ci.ncluster_max), "batch_K": int(ci.batch_K), "roofline": rh,
                         "arithmetic_roofline": rf}
                if not args.no_cpu_baseline:
                    cv, cwhat, _ = cpu_sample(name, cw["nlive"], budget_s=6.0)
                    entry["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": 1, "kind": "port", "sample": cwhat}
                configs.append(entry)
            except RuntimeError as ex:
                configs.append({"workload": name, "error": str(ex)})

    # ---- reduce over ranks ------------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([dev_ms, e_t, wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_max, e_t_max, wall_max = t.tolist()
        c = torch.tensor([evals, e_evals, launches + e_launch, algo_bytes, h2d, d2h], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        evals_all, e_evals_all, launches_all, algo_all, h2d, d2h = c.tolist()
    else:
        dev_ms_max, e_t_max, wall_max = dev_ms, e_t, wall
        evals_all, e_evals_all, launches_all, algo_all = evals, e_evals, launches + e_launch, algo_bytes

    if rank == 0:
        # rank 0's kernel: a sharded run's counters are replicated, each rank moves 1/world of the algorithmic bytes
        achieved = (algo_bytes / (world if sharded else 1) / (dev_ms * 1e-3)) / 1e9
        traffic, traffic_src = committed_traffic(args.workload) if world == 1 else (None, None)
        value = evals_all / (dev_ms_max * 1e-3)
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cv, cwhat, crun = cpu_sample(args.workload, n, budget_s=8.0)
            cpu = {"value": cv, "unit": UNIT, "cores": 1, "kind": "port", "sample": cwhat, "wall_time_to_logZ_s": crun}
        fp64_ach = value / world * flops_per_eval(w) / 1e12
        K = int(info_dev.batch_K)
        # variance of log X per unit of compression of a generation that kills K of n, relative to one death at a time:
        # sum_{j<K} (n-j)^-2 / sum_{j<K} (n-j)^-1 * n
        var_ratio = float(sum(1.0 / (n - j) ** 2 for j in range(K)) / sum(1.0 / (n - j) for j in range(K)) * n)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload + (f"_x{world}_sharded" if sharded else ""), "nDims": D, "nDerived": P,
                       "nlive": n, "num_repeats": R, "do_clustering": clustering,
                       "precision_criterion": 1e-3, "batch_K": K, "ctas_per_run": int(info_dev.ctas_per_run),
                       "warps_per_cta": int(info_dev.warps_per_cta), "step": "one complete nested-sampling run",
                       "l2": "flushed (256 MiB memset) before every step",
                       "multi_gpu": ("n/a" if world == 1 else
                                     "independent replica runs per rank (no data-path collective)" if not sharded else
                                     f"ONE run, nlive={n} sharded over {world} GPUs: chains dealt k % world, last babies and "
                                     "covariance statistics exchanged over NVLink peer memory inside the persistent kernel")},
            "schedule": {"batch_K": K, "rule": "engine default: about nlive/2 for a run alone on the device -- a whole number of waves of chains ((SMs-1)*4 per device) where that stays within [0.5, 0.6] nlive, else nlive/2 -- and nlive/4 inside an ensemble",
                         "logX_variance_per_unit_compression_vs_one_death_at_a_time": var_ratio,
                         "note": "the reported logZerr carries this factor (the evidence recurrences are applied death by death)"},
            "wall_time_to_logZ_s": dev_ms / args.steps * 1e-3, "wall_ms_per_step_host": 1e3 * wall_max / args.steps,
            "logZ_mean": float(np.mean(logZs)), "logZ_sem": float(np.std(logZs, ddof=1) / np.sqrt(len(logZs))) if len(logZs) > 1 else None,
            "logZerr_mean_reported": float(np.mean(logZerrs)), "logZ_true": TRUE_LOGZ.get(args.workload),
            "ncluster_max": int(info_dev.ncluster_max), "ndead_per_step": ndead / args.steps, "evals_per_step": (evals_all if sharded else evals) / args.steps,
            "e2e": {"value": e_evals_all / e_t_max, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps,
                    "d2h_bytes_per_step": d2h / args.steps, "api": "polychord_c_interface + dumper (host arrays)",
                    "ms_per_step": 1e3 * e_t_max / args.steps, "dumper_calls": sink["calls"]},
            "gpu_launches": int(launches_all),
            "phase_ms_last_step": {k: round(v, 3) for k, v in info_dev.as_dict()["phase_ms"].items()},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": algo_bytes / args.steps / (world if sharded else 1),
                         "peak_source": which, "kernel": info_dev.kernel,
                         "note": "one run is latency-bound (its chains occupy a few hundred of the device's 9472 warp "
                                 "slots); 'ensemble' is the same kernel family with the device filled"},
            "arithmetic_roofline": {"bound": "fp64_fma", "achieved": fp64_ach, "peak": fp64_peak, "unit": "TFLOP/s per GPU",
                                    "frac": fp64_ach / fp64_peak if fp64_peak else None, "flops_per_eval": flops_per_eval(w),
                                    "peak_source": "measured here (pc_measure_fp64_tflops: FMA microbenchmark)"},
            "cpu_baseline": cpu, "clocks": clocks, "ensemble": ens, "configs": configs, "region_wall_s": t_region,
        }
        if sharded:
            line["sharded_parity"] = sharded_parity
        print(json.dumps(line), flush=True)
    if sharded:
        mgpu.detach()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
