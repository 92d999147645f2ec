#!/usr/bin/env python
"""Headline benchmark: BASELINE.json config[1] — successive model, 5 psites, 1 M synthetic parameter
sets per GPU, forward solve over the 14 experimental time points with the weighted-residual loss
and score_fit fused into the kernel epilogue.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of B parameter sets (per GPU; weak scaling).
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for how every field is obtained.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL, NS, B_PER_GPU = "succmod", 5, 1_000_000
T_GRID = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
METRIC = "ODE solves/sec (full horizon)"
UNIT = "solves/s"


def workload_config(n_gpus, b_per_gpu):
    return {"workload": f"{MODEL} ns={NS} (7 states, 14 params), {b_per_gpu} parameter sets/GPU ~U(0.05,3), "
                        f"14 output times 0..960 min, fused ssr+score_fit epilogue",
            "batch_per_gpu": b_per_gpu, "global_batch": b_per_gpu * n_gpus, "parallelism": f"shard{n_gpus}",
            "integrator": "ROS6L order 6(5) Rosenbrock, 7 solves/step, rtol=2e-5 atol=2e-9 (library defaults: error <= 0.15 of the "
                          "1e-6 parity bound vs the reference's tight solution)",
            "l2": "256 MiB written between timed steps (flush), excluded from the step timing"}


# ------------------------------------------------------------------ algorithmic work per step
def flops_per_step(model, ns, nsol=None):
    """FP64 operations of ONE integrator step attempt as the kernels perform them (FMA = 2,
    add/mul/div/max = 1) — DESIGN.md §Kernels derives each term.  nsol = solves per step: 7 for ROS6L (default of the
    thread-per-system kernels: dist/succ up to 8 sites), 6 for ROS5L / RODAS4 (dense kernel)."""
    if nsol is None:
        nsol = 7 if model in ("distmod", "succmod") and ns <= 8 else 6
    n = 2 + ns
    if model == "distmod":
        factor, rhs, solve = 10 * ns + 16, 4 * ns + 5, 4 * ns + 6
    elif model == "succmod":
        factor, rhs, solve = 13 * ns + 8, 4 * ns + 5, 5 * ns + 4
    else:
        n = 2 + (1 << ns) - 1
        nnz = 3 + ns + ((1 << ns) - 1) * (ns + 1)          # non-zeros of the transition-rate matrix
        factor, rhs, solve = (2 * n ** 3) // 3, 2 * nnz + n, 2 * n * n
    # v0 = h f (n) ; nsol solves ; y_new/err accumulation (4 nsol - 3) n ; error ratio 9 n ; finite check n ; controller 20
    return factor + rhs + n + nsol * solve + (4 * nsol - 3) * n + 9 * n + n + 20


def bytes_per_solve(model, ns):
    P = 4 + 2 * ns if model != "randmod" else 4 + ns + (1 << ns) - 1
    return 8 * P + 8 + 8 + 4 + 4 + 4       # params in; ssr, score, status, nsteps, nrej out


# ----------------------------------------------------------------------------- clock sampler
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the bench runs."""

    def __init__(self, cuda_index):
        self.samples = []
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = cuda_index
            if vis:
                parts = [p.strip() for p in vis.split(",") if p.strip()]
                if cuda_index < len(parts) and parts[cuda_index].isdigit():
                    idx = int(parts[cuda_index])
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.ok:
            self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.ok:
            self.thread.join(timeout=2)

    def summary(self, windows):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sel = [s for s in self.samples if any(a <= s[0] <= b for a, b in windows)] or self.samples
        mhz = sorted(s[1] for s in sel)
        bits = 0
        for s in sel:
            bits |= int(s[2])
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                 0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting",
                 0x10: "sync_boost", 0x100: "display_clock_setting"}
        reasons = [n for b, n in names.items() if bits & b]
        return {"sm_mhz": float(mhz[len(mhz) // 2]), "sm_max_mhz": float(self.max_mhz), "reasons": reasons,
                "samples": len(sel)}


# --------------------------------------------------------------------------- CPU baselines
def _cpu_chunk(args):
    """Worker: reference-equivalent solve + score for a chunk (oracle port; reference RHS is
    numba-jitted, so is the port's)."""
    params, y0, target = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import local_models as om
    import loss as ol
    out = np.empty(len(params))
    for i, p in enumerate(params):
        _, flat = om.solve_ode(MODEL, p, y0, NS, T_GRID)
        out[i] = ol.score_fit(p, target, flat)
    return out


def cpu_setup():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import local_models as om
    y0 = np.asarray(om.initial_condition(MODEL, NS))
    theta0 = np.random.default_rng(20).uniform(0.05, 3.0, 14)
    target = om.solve_ode(MODEL, theta0, y0, NS, T_GRID)[1]
    return y0, target


def cpu_baseline_single(n_sample):
    """Single-core oracle port on the first n_sample parameter sets of the workload."""
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ.setdefault(k, "1")
    y0, target = cpu_setup()
    params = np.random.default_rng(2).uniform(0.05, 3.0, (n_sample, 14))
    _cpu_chunk((params[:16], y0, target))            # JIT warm-up, excluded
    t0 = time.perf_counter()
    _cpu_chunk((params, y0, target))
    dt = time.perf_counter() - t0
    return {"value": n_sample / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {n_sample} of the 1M parameter sets (oracle/local_models.py: scipy LSODA + "
                      f"numba RHS + score_fit, one process), {dt:.1f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference
    itself is pure Python + SciPy/Numba and /root/reference does not exist on the GPU box) on all
    host cores, chunked over a process pool."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    per_core = 1500
    S = per_core * cores
    y0, target = cpu_setup()
    params = np.random.default_rng(2).uniform(0.05, 3.0, (S, 14))
    chunks = [(c, y0, target) for c in np.array_split(params, cores * 4)]
    warm = [(params[:8], y0, target)] * cores
    times = []
    with ProcessPoolExecutor(max_workers=cores) as ex:
        list(ex.map(_cpu_chunk, warm))                                   # spawn + JIT, excluded
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            list(ex.map(_cpu_chunk, chunks))
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = S / (ms * 1e-3)
    sample = f"{S} of the 1M parameter sets per step ({per_core}/core), ProcessPoolExecutor({cores}), chunked"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus, B_PER_GPU),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# -------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import phoskintime_b200 as pk
    from phoskintime_b200 import parallel
    from phoskintime_b200.models import succmod
    from phoskintime_b200.steady import initial_condition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    eng = pk.get_engine(local_rank)
    run = parallel.ShardedRun(engine=eng, backend="nccl" if world > 1 else None)

    B = args.batch
    n, P, L = pk.local_dims(MODEL, NS, len(T_GRID))
    y0 = np.asarray(initial_condition(NS, MODEL))
    # synthetic inputs: params ~ U(0.05,3) seed 2 (+rank); target = model output
    # at a hidden theta0 (seed 20); sigma = ones
    params_h = torch.empty((B, P), dtype=torch.float64).pin_memory()
    params_h.numpy()[:] = np.random.default_rng(2 + rank).uniform(0.05, 3.0, (B, P))
    theta0 = np.random.default_rng(20).uniform(0.05, 3.0, P)
    target = succmod.solve_ode(theta0, y0, NS, T_GRID)[1]

    params_d = params_h.to(dev)
    y0_d = torch.from_numpy(y0).to(dev)
    t_d = torch.from_numpy(T_GRID).to(dev)
    target_d = torch.from_numpy(target).to(dev)
    out_d = {k: torch.empty(B, dtype=torch.float64, device=dev) for k in ("ssr", "score")}
    out_d.update({k: torch.empty(B, dtype=torch.int32, device=dev) for k in ("status", "nsteps", "nrej")})
    gathered = torch.empty(world * B, dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    gather_chunks = args.gather_chunks if args.gather_chunks >= 0 else (2 if world >= 8 else 0)

    def step_device():
        # world > 1: the one collective of the path, the all-gather of the per-sample losses.  --gather-chunks K > 0
        # fuses it with the solve (K pieces, piece c's gather overlaps piece c+1: pk_local_solve_allgather); 0 = one
        # launch followed by one all-gather.
        if world > 1 and gather_chunks > 0:
            return eng.solve_local_batch(MODEL, params_d, y0_d, NS, t_d, want=("ssr", "score"), target=target_d,
                                         out=out_d, counters=True, gather=("score", gathered, gather_chunks))
        res = eng.solve_local_batch(MODEL, params_d, y0_d, NS, t_d, want=("ssr", "score"), target=target_d,
                                    out=out_d, counters=True)
        if world > 1:
            eng.allgather_f64(res["score"], gathered)
        return res

    sampler = ClockSampler(local_rank)
    sampler.start()
    fp64_peak = eng.measure_fp64_peak()

    for _ in range(args.warmup):
        step_device()
    run.barrier()
    torch.cuda.synchronize()
    step_ms, kern_ms, launches = [], [], 0
    w0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        eng.region_begin()
        res = step_device()
        step_ms.append(eng.region_end())
        nl, kms = eng.last_launch_info()
        # last_launch_info refers to the last pk_* call (the all-gather issues no kernel of ours)
        kern_ms.append(kms if world == 1 else None)
        launches += nl
    torch.cuda.synchronize()
    run.barrier()
    w1 = time.perf_counter()
    total_ms = run.max_over_ranks(float(np.sum(step_ms)))
    ms_per_step = total_ms / args.steps
    value = B * world / (ms_per_step * 1e-3)

    nsteps_total = int((res["nsteps"].long() + res["nrej"].long()).sum().item())
    n_bad = int((res["status"] != 0).sum().item())
    if world == 1:
        kms = float(np.mean([k for k in kern_ms if k is not None]))
    else:
        kms = float(np.mean(step_ms))
    fl = nsteps_total * flops_per_step(MODEL, NS)
    achieved_tf = fl / (kms * 1e-3) / 1e12
    hbm_gbs = B * bytes_per_solve(MODEL, NS) / (kms * 1e-3) / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    # ---- end to end through the reference-facing plugin call, host buffers, copies inside
    out_h = {k: torch.empty(B, dtype=torch.float64).pin_memory().numpy() for k in ("ssr", "score")}
    out_h["status"] = torch.empty(B, dtype=torch.int32).pin_memory().numpy()
    params_np = params_h.numpy()

    def step_e2e():
        return succmod.solve_ode_batch(params_np, y0, NS, T_GRID, want=("ssr", "score"), target=target,
                                       out=out_h, counters=False)
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    run.barrier()
    e0 = time.perf_counter()
    e2e_times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        r = step_e2e()
        _ = float(r["score"][0])
        e2e_times.append(time.perf_counter() - t0)
    run.barrier()
    e1 = time.perf_counter()
    e2e_s = run.max_over_ranks(float(np.mean(e2e_times)))
    e2e_value = B * world / e2e_s
    h2d = B * P * 8 + n * 8 + len(T_GRID) * 8 + L * 8
    d2h = B * 8 * 2 + B * 4

    sampler.stop()
    clocks = sampler.summary([(w0, w1), (e0, e1)])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_single(args.cpu_sample)
        # "rel err vs ref" of the metric (part of the CPU leg: the oracle is the checker): the first 512 parameter sets of
        # the workload, GPU trajectories at the library defaults vs the oracle's exact solution (matrix exponential) and
        # vs the stock reference path (LSODA defaults), in units of the parity bound 1e-6*|ref| + 1e-9
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import local_models as om
        nchk = 512
        pchk = params_h.numpy()[:nchk]
        got = eng.solve_local_batch(MODEL, pchk, y0, NS, T_GRID, want=("sol",))["sol"]
        worst_exact, worst_stock, rel_max = 0.0, 0.0, 0.0
        for b in range(nchk):
            ex = om.exact_linear(MODEL, pchk[b], y0, NS, T_GRID)
            worst_exact = max(worst_exact, float((np.abs(got[b] - ex) / (1e-6 * np.abs(ex) + 1e-9)).max()))
            rel_max = max(rel_max, float((np.abs(got[b] - ex) / np.maximum(np.abs(ex), 1e-12)).max()))
            if b < 64:
                st = om.solve_ode(MODEL, pchk[b], y0, NS, T_GRID)[0]
                worst_stock = max(worst_stock, float((np.abs(got[b] - st) / (1e-6 * np.abs(st) + 1e-7)).max()))
        cpu["parity"] = {"systems": nchk, "max_err_over_bound_vs_exact": worst_exact, "max_rel_err_vs_exact": rel_max,
                         "max_err_over_mixed_bound_vs_stock_reference": worst_stock,
                         "bounds": "1e-6*|ref|+1e-9 (exact), 1e-6*|ref|+1e-7 (stock LSODA path, 64 systems)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, B),
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this workload, one
                         # `ncu --set full` capture (profiles/r1_succ5_g_ros6l.txt: 116.4 MB + 129.3 MB); the
                         # algorithmic bytes are 140 MB (bytes_per_solve x B): the rest is write-back of the
                         # L2-resident trajectory slots
                         "traffic": 245.6e6 if (B == B_PER_GPU and world == 1) else None,
                         "peak_source": "pk_measure_fp64_peak (register-resident DFMA probe, this run); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "flops_per_step": flops_per_step(MODEL, NS),
                         "steps_per_solve": nsteps_total / B, "kernel_ms": kms,
                         "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                                 "bytes_per_solve": bytes_per_solve(MODEL, NS)}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "api": "phoskintime_b200.models.succmod.solve_ode_batch (host numpy, pinned)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "failed_systems": n_bad,
            "device": eng.device_name,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    run.barrier()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Native libraries (NCCL's "NCCL version ..." banner, NCCL_DEBUG output) write to file descriptor 1: move
    # everything except the JSON line to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="parameter sets per GPU")
    ap.add_argument("--cpu-sample", type=int, default=10000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather-chunks", type=int, default=-1,
                    help="N>1: fuse the all-gather with the solve in this many pieces (0 = one launch + one all-gather; "
                         "-1 = auto: 2 pieces at 8 GPUs, where the collective is long enough to be worth a second launch "
                         "tail — measured 3.50 -> 2.97 ms per step at 8 GPUs (pieces of 2/3 + 1/3), 2.87 -> 2.94 at 4, 2.78 -> 2.97 at 2 (4 pieces))")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
