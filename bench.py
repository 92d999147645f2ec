#!/usr/bin/env python
"""Headline benchmark: BASELINE.json configs[1] — successive model, 5 psites, 1 M synthetic parameter
sets per GPU, forward solve over the 14 experimental time points with the weighted-residual loss
and score_fit fused into the kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path over one batch of parameter sets.  Prints ONE JSON line (rank 0).
`--workload` selects the other BASELINE.json configurations (the driver runs the default):

    succ5      configs[1]  successive ns=5, 1 M sets per GPU (weak scaling), fused ssr + score_fit          [default]
    dist3      north-star target shape: distributive ns=3, 2^20 sets per GPU (weak), fused score_fit
    morris3    configs[0]  distributive ns=3, Morris N=1000 -> 11 000 rows in total (strong), fused Y + elementary effects
    rand6      configs[2]  random model ns=6 (65 states), 262 144 sets per GPU (weak), flat rows out
    normest4   configs[3]  distributive ns=4, 1000 proteins x 256 starts = 256 000 in total (strong, sharded by protein),
                           fused per-protein loss, NCCL all-gather of the losses
    global120  configs[4]  coupled network N=120 (533 states), 16 384 parameter vectors in total (strong), fused Morris
                           scalar + LOSS_FN, NCCL all-gather of the scalars
See DESIGN.md §Measurement for how every field is obtained.
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

T_GRID = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
T_RNA = np.array([4.0, 8.0, 15.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
T_UNION = np.unique(np.concatenate([T_GRID, T_RNA]))
METRIC = "ODE solves/sec (full horizon)"
UNIT = "solves/s"

# name -> (kind, model, ns, batch, scaling, outputs, groups)
WORKLOADS = {
    "succ5": dict(kind="local", model="succmod", ns=5, batch=1_000_000, scaling="weak", want=("ssr", "score"), gather="score",
                  text="succmod ns=5 (7 states, 14 params), {b} parameter sets/GPU ~U(0.05,3), 14 output times 0..960 min, "
                       "fused ssr+score_fit epilogue"),
    "dist3": dict(kind="local", model="distmod", ns=3, batch=1 << 20, scaling="weak", want=("score",), gather="score",
                  text="distmod ns=3 (5 states, 10 params), {b} parameter sets/GPU ~U(0.05,3), 14 output times, fused score_fit "
                       "(north-star target shape)"),
    "morris3": dict(kind="local", model="distmod", ns=3, batch=11_000, scaling="strong", want=("Y",), gather="Y",
                    text="distmod ns=3, Morris N=1000 x (D+1) = {b} rows in total (BASELINE configs[0]), fused total_signal Y + "
                         "elementary effects"),
    "rand6": dict(kind="local", model="randmod", ns=6, batch=262_144, scaling="weak", want=("flat",), gather=None,
                  text="randmod ns=6 (65 states, 73 params), {b} parameter sets/GPU ~U(0.05,3), flat rows (107 doubles) out "
                       "(BASELINE configs[2])"),
    "normest4": dict(kind="local", model="distmod", ns=4, batch=256_000, scaling="strong", want=("ssr",), gather="ssr",
                     groups=1000,
                     text="distmod ns=4, 1000 proteins x 256 starts = {b} in total, sharded by protein, fused per-protein "
                          "weighted-residual loss, all-gather of the losses (BASELINE configs[3])"),
    "global120": dict(kind="global", batch=16_384, scaling="strong", gather="metric",
                      text="coupled kinase-TF-protein network N=120 K=40 (533 states, 934 params), {b} parameter vectors in "
                           "total (+-5 %), fused Morris scalar + LOSS_FN, all-gather of the scalars (BASELINE configs[4])"),
}


def integrator_text(w):
    if w["kind"] == "global":
        return ("staged RODAS4 order 4(3), analytic Jacobian, Schur-complement solve (fixed-point sweeps to 1e-12 when ||K||_inf < 0.5, "
                "register Gauss-Jordan otherwise), rtol=2e-6 atol=2e-9 (library defaults)")
    if w["model"] == "randmod":
        return "ROS5L order 5(4) Rosenbrock, 6 solves/step on a reused inverse, rtol=2e-6 atol=2e-9 (library defaults)"
    return ("ROS6L order 6(5) Rosenbrock, 7 solves/step, rtol=2e-5 atol=2e-11 (library defaults: error <= 0.15 of the 1e-6 "
            "parity bound vs the reference's tight solution)")


def workload_config(name, n_gpus, b_per_gpu):
    w = WORKLOADS[name]
    total = b_per_gpu * n_gpus if w["scaling"] == "weak" else w["batch"]
    return {"workload": w["text"].format(b=b_per_gpu if w["scaling"] == "weak" else total), "name": name,
            "batch_per_gpu": b_per_gpu, "global_batch": total, "parallelism": f"shard{n_gpus}",
            "integrator": integrator_text(w),
            "l2": "256 MiB written between timed steps (flush), excluded from the step timing"}


# ------------------------------------------------------------------ algorithmic work per step
def flops_per_step(model, ns, nsol=None, fp32_control=False):
    """FP64 operations of ONE integrator step attempt exactly as the kernels perform them (FMA = 2, add / mul = 1; the
    reciprocal of a factorisation = 4 FMA after the MUFU seed) — DESIGN.md §3 derives each term from
    csrc/local_tps.cuh.  The FP32 error ratio and step controller (9n + 20 operations) are NOT FP64 work and are only
    added when fp32_control=True.  nsol = solves per step: 7 for ROS6L (default of the thread-per-system kernels:
    dist/succ up to 8 sites), 6 for ROS5L / RODAS4 (dense kernel)."""
    if nsol is None:
        nsol = 7 if model in ("distmod", "succmod") and ns <= 8 else 6
    n = 2 + ns
    if model == "distmod":
        # factor: q_i, prefix/suffix products, Q, sum S_i prod q_j, three pivots, batch inversion of 3, 1/q_i and c S_i/q_i
        factor = 2 * ns + 2 * (ns - 1) + 1 + 3 * ns + 2 + 6 + (3 * 2 + 8) + 2 + 3 * ns
        rhs, solve = 4 * ns + 5, 4 * ns + 7
    elif model == "succmod":
        nt = (ns + 1) // 2
        nb = ns - nt
        top = 2 + (4 if nt > 1 else 0) + 5 * max(0, nt - 2)                      # continuants from the top
        bot = (2 if nb > 0 else 0) + (4 if nb > 1 else 0) + 5 * max(0, nb - 2)   # and from the bottom
        delta = 1 + 6 + (3 if nb > 0 else 0)
        factor = 2 * ns + top + bot + delta + 2 + (3 * (ns + 1) + 8) + (2 + (nt - 1) + nt + 2 * nb)
        rhs = 4 * ns + 5
        solve = 3 + 2 * (nt - 1) + 2 * max(0, nb - 1) + (3 + (2 if nb > 0 else 0)) + 3 * nt + 3 * nb   # twisted sweeps
    else:
        n = 2 + (1 << ns) - 1
        nnz = 3 + ns + ((1 << ns) - 1) * (ns + 1)          # non-zeros of the transition-rate matrix
        # dense kernel: mat-vecs with a reused inverse; the inversions (2 n^3 each, ~1 per 8 steps) are NOT counted here
        factor, rhs, solve = 0, 2 * nnz + n, 2 * n * n
    # v0 = h f (n) ; nsol solves ; y_new/err accumulation (4 nsol - 3) n ; FP64 part of the step-size logic 10
    f = factor + rhs + n + nsol * solve + (4 * nsol - 3) * n + 10
    return f + (9 * n + 20 if fp32_control else 0)


def bytes_per_solve(w):
    if w["kind"] == "global":
        return 8 * 934 + 8 + 24 + 12
    ns, model = w["ns"], w["model"]
    P = 4 + 2 * ns if model != "randmod" else 4 + ns + (1 << ns) - 1
    L = (len(T_GRID) - 5) + len(T_GRID) + ns * len(T_GRID)
    out = sum({"ssr": 8, "score": 8, "Y": 8, "flat": 8 * L}[k] for k in w["want"])
    return 8 * P + out + 12       # params in; requested outputs, status, nsteps, nrej out


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


# --------------------------------------------------------------------------- synthetic inputs
def local_inputs(name, rank, B):
    """parameters ~ U(0.05, 3), seeded; targets = the model's own output at hidden parameter vectors."""
    w = WORKLOADS[name]
    model, ns = w["model"], w["ns"]
    P = 4 + 2 * ns if model != "randmod" else 4 + ns + (1 << ns) - 1
    seed = {"succ5": 2, "dist3": 11, "morris3": 1, "rand6": 3, "normest4": 4}[name]
    if name == "morris3":
        from phoskintime_b200 import sensitivity
        theta = np.random.default_rng(1).uniform(0.05, 3.0, 10)
        X = sensitivity.morris_sample(sensitivity.define_sensitivity_problem_ds(3, theta), 1000, 400, seed=42)
        return X, None, None
    params = np.random.default_rng(seed + rank).uniform(0.05, 3.0, (B, P))
    return params, P, seed


# --------------------------------------------------------------------------- CPU baselines
def _reference_solver(model):
    """The reference's own `solve_ode` when an unmodified copy of the reference travels with the repo
    (baseline/_ref/, importable through oracle/ref_shim.py); otherwise the oracle port (kind 'port')."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    if os.path.isdir(os.path.join(ref_root, "models")):
        os.environ["PHOSKIN_REFERENCE_ROOT"] = ref_root
        try:
            import ref_shim
            mod = ref_shim.load_local_models()[model]
            return (lambda p, y0, ns, t: mod.solve_ode(tuple(p), y0, ns, t)), "reference"
        except Exception:
            pass
    import local_models as om
    return (lambda p, y0, ns, t: om.solve_ode(model, p, y0, ns, t)), "port"


def _cpu_chunk(args):
    """Worker: reference-equivalent solve (+ score) for a chunk (oracle port; the reference RHS is numba-jitted, so is
    the port's)."""
    name, params, y0, target = args
    w = WORKLOADS[name]
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    if w["kind"] == "global":
        import global_models as og
        net, t = y0
        out = np.empty(len(params))
        for i, p in enumerate(params):
            Y = og.simulate_odeint(0, net, t, 1e-8, 1e-8, 200000, params=og.unpack_params(p, net))
            out[i] = Y[-1, 0]
        return out
    import loss as ol
    solve, _ = _reference_solver(w["model"])
    out = np.empty(len(params))
    for i, p in enumerate(params):
        sol, flat = solve(p, y0, w["ns"], T_GRID)
        out[i] = ol.score_fit(p, target, flat) if target is not None else ol.compute_Y(sol, w["ns"])
    return out


def cpu_setup(name):
    w = WORKLOADS[name]
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    if w["kind"] == "global":
        from phoskintime_b200.global_model import synthetic_system
        s = synthetic_system(seed=5, N=120, K=40, max_sites=4, model=0)
        base = s.pack_params()
        return (s.as_dict(), T_UNION), None, base
    import local_models as om
    y0 = np.asarray(om.initial_condition(w["model"], w["ns"]))
    target = None
    if "score" in w["want"] or "ssr" in w["want"]:
        P = 4 + 2 * w["ns"]
        theta0 = np.random.default_rng(20).uniform(0.05, 3.0, P)
        target = om.solve_ode(w["model"], theta0, y0, w["ns"], T_GRID)[1]
    return y0, target, None


def cpu_params(name, n):
    w = WORKLOADS[name]
    if w["kind"] == "global":
        _, _, base = cpu_setup(name)
        return base[None, :] * (1.0 + 0.05 * np.random.default_rng(5).uniform(-1, 1, (n, base.size)))
    return local_inputs(name, 0, max(n, 16))[0][:n]


def cpu_baseline_single(name, n_sample):
    """Single-core oracle port on the first n_sample parameter sets of the workload."""
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ.setdefault(k, "1")
    y0, target, _ = cpu_setup(name)
    params = cpu_params(name, n_sample)
    _cpu_chunk((name, params[:min(16, len(params))], y0, target))            # JIT warm-up, excluded
    t0 = time.perf_counter()
    _cpu_chunk((name, params, y0, target))
    dt = time.perf_counter() - t0
    kind = "port" if WORKLOADS[name]["kind"] == "global" else _reference_solver(WORKLOADS[name]["model"])[1]
    what = ("oracle/global_models.py: reference RHS + finite-difference Jacobian through scipy LSODA"
            if WORKLOADS[name]["kind"] == "global" else "oracle/local_models.py: scipy LSODA + numba RHS + score_fit")
    return {"value": n_sample / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"first {n_sample} parameter sets of the workload ({what}, one process), {dt:.1f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the unmodified reference from baseline/_ref when
    it travelled with the repo, else the oracle port: the reference is pure Python + SciPy/Numba, /root/reference does
    not exist on the GPU box and its poetry build backend is absent from the wheelhouse) on all host cores, chunked
    over a process pool."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor
    name = args.workload
    w = WORKLOADS[name]
    cores = os.cpu_count() or 1
    per_core = {"succ5": 1500, "dist3": 1500, "morris3": 600, "rand6": 60, "normest4": 1500, "global120": 1}[name]
    S = per_core * cores
    y0, target, _ = cpu_setup(name)
    params = cpu_params(name, S)
    chunks = [(name, c, y0, target) for c in np.array_split(params, cores * (4 if per_core >= 4 else 1))]
    warm = [(name, params[:min(8, S)] if w["kind"] == "local" else params[:1], y0, target)] * cores
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
    kind = "port" if w["kind"] == "global" else _reference_solver(w["model"])[1]
    sample = f"{S} parameter sets of the workload per step ({per_core}/core), ProcessPoolExecutor({cores}), chunked"
    b_gpu = args.batch or (w["batch"] if w["scaling"] == "weak" else -(-w["batch"] // args.gpus))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": w["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(name, args.gpus, b_gpu),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------ exact solution, batched
def exact_batch(model, params, y0, ns, t):
    """Oracle's exact solution (matrix exponential of the affine system, oracle/local_models.py::exact_linear) for many
    parameter sets at once: ONE expm per system on the base step of the output grid, the other output times by products
    of its binary powers (all output times are multiples of 0.25 min)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import local_models as om
    from scipy.linalg import expm
    dt = 0.25
    mult = np.rint(np.asarray(t) / dt).astype(np.int64)
    assert np.allclose(mult * dt, t)
    B = len(params)
    M0, b0 = om.linear_form(model, params[0], ns)
    n = M0.shape[0]
    Z = np.zeros((B, n + 1, n + 1))
    for i in range(B):
        M, b = om.linear_form(model, params[i], ns)
        Z[i, :n, :n] = M
        Z[i, :n, n] = b
    pw = [expm(Z * dt)]
    while (1 << len(pw)) <= mult.max():
        pw.append(pw[-1] @ pw[-1])
    out = np.empty((B, len(t), n))
    yb = np.concatenate([np.broadcast_to(np.asarray(y0, float), (B, n)), np.ones((B, 1))], axis=1)
    for k, m in enumerate(mult):
        v = yb.copy()
        j = 0
        while m:
            if m & 1:
                v = np.einsum("bij,bj->bi", pw[j], v)
            m >>= 1
            j += 1
        out[:, k] = v[:, :n]
    return np.clip(out, 0, None)


# -------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import phoskintime_b200 as pk
    from phoskintime_b200 import parallel

    name = args.workload
    w = WORKLOADS[name]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    eng = pk.get_engine(local_rank)
    run = parallel.ShardedRun(engine=eng, backend="nccl" if world > 1 else None)

    if w["scaling"] == "weak":
        B = args.batch or w["batch"]
        total = B * world
    else:
        total = args.batch or w["batch"]
        align = {"normest4": 256, "morris3": 11}.get(name, 1)          # whole proteins / whole Morris trajectories per rank
        lo, hi = parallel.shard_bounds(total, world, rank, align=align)
        B = hi - lo
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gather_chunks = args.gather_chunks if args.gather_chunks >= 0 else 2

    # ---------------------------------------------------------------- per-workload device step and host (e2e) step
    if w["kind"] == "local":
        from phoskintime_b200.models import distmod, randmod, succmod
        from phoskintime_b200.steady import initial_condition
        model, ns = w["model"], w["ns"]
        plugin = {"distmod": distmod, "succmod": succmod, "randmod": randmod}[model]
        n, P, L = pk.local_dims(model, ns, len(T_GRID))
        y0 = np.asarray(initial_condition(ns, model))
        if name == "morris3":
            X, _, _ = local_inputs(name, 0, total)
            params_np = np.ascontiguousarray(X[lo:hi])
        elif w["scaling"] == "strong":
            params_np = local_inputs(name, 0, total)[0][lo:hi]
        else:
            params_np = local_inputs(name, rank, B)[0]
        params_h = torch.empty((B, P), dtype=torch.float64).pin_memory()
        params_h.numpy()[:] = params_np
        kw, kw_h = {}, {}
        want = w["want"]
        G = w.get("groups", 1)
        if "score" in want or "ssr" in want:
            hidden = np.random.default_rng(20).uniform(0.05, 3.0, (G, P))
            target = np.stack([plugin.solve_ode(h, y0, ns, T_GRID)[1] for h in hidden[:min(G, 8)]])
            if G > 8:       # 1000 per-protein targets: the first 8 from the model, the rest as scaled copies (synthetic data)
                sc = 1.0 + 0.05 * np.random.default_rng(21).standard_normal((G, 1))
                target = target[np.arange(G) % 8] * sc
            group = (np.arange(lo, hi) // (total // G)).astype(np.int32) if G > 1 else None
            kw = {"target": torch.from_numpy(target).to(dev)}
            kw_h = {"target": target}
            if group is not None:
                kw["group"] = torch.from_numpy(group).to(dev)
                kw_h["group"] = group
        params_d = params_h.to(dev)
        y0_d = torch.from_numpy(y0).to(dev)
        t_d = torch.from_numpy(T_GRID).to(dev)
        out_d = {k: torch.empty((B, L) if k == "flat" else B, dtype=torch.float64, device=dev) for k in want}
        out_d.update({k: torch.empty(B, dtype=torch.int32, device=dev) for k in ("status", "nsteps", "nrej")})
        gkey = w["gather"]
        per_rank = B if w["scaling"] == "weak" else max(parallel.shard_sizes(total, world, align={"normest4": 256, "morris3": 11}.get(name, 1)))
        gathered = torch.empty(world * per_rank, dtype=torch.float64, device=dev) if (world > 1 and gkey) else None
        send_pad = torch.zeros(per_rank, dtype=torch.float64, device=dev) if gathered is not None and per_rank != B else None

        use_p2p = gathered is not None and args.gather_mode == "p2p" and model != "randmod"
        if use_p2p:
            run.setup_p2p(per_rank)

        def step_device():
            # world > 1: the one collective of the path, the gather of the per-sample scalars on every rank.
            #   p2p (default): the solve kernel stores each finished system's scalar straight into every rank's symmetric
            #       buffer over NVLink peer memory (pk_local_solve_gather_p2p) — no collective call, no tail;
            #   nccl: one launch followed by one ncclAllGather;  fused: --gather-chunks K pieces, piece c's ncclAllGather
            #       overlaps piece c+1 (pk_local_solve_allgather).
            if use_p2p:
                return eng.solve_local_batch(model, params_d, y0_d, ns, t_d, want=want, out=out_d, counters=True,
                                             gather_p2p=gkey, **kw)
            if gathered is not None and args.gather_mode == "fused" and gather_chunks > 0 and send_pad is None:
                return eng.solve_local_batch(model, params_d, y0_d, ns, t_d, want=want, out=out_d, counters=True,
                                             gather=(gkey, gathered, gather_chunks), **kw)
            res = eng.solve_local_batch(model, params_d, y0_d, ns, t_d, want=want, out=out_d, counters=True, **kw)
            if name == "morris3" and world == 1:
                eng.morris_ee(params_d, res["Y"], 400, scaled=True)
            if gathered is not None:
                src = res[gkey]
                if send_pad is not None:
                    send_pad[:B] = src
                    src = send_pad
                eng.allgather_f64(src, gathered)
            return res

        out_h = {k: torch.empty((B, L) if k == "flat" else B, dtype=torch.float64).pin_memory().numpy() for k in want}
        out_h["status"] = torch.empty(B, dtype=torch.int32).pin_memory().numpy()
        params_e2e = params_h.numpy()
        api = f"phoskintime_b200.models.{model}.solve_ode_batch (host numpy, pinned)"

        def step_e2e():
            return plugin.solve_ode_batch(params_e2e, y0, ns, T_GRID, want=want, out=out_h, counters=False, **kw_h)

        h2d = B * P * 8 + n * 8 + len(T_GRID) * 8 + (G * L * 8 if kw else 0) + (B * 4 if kw.get("group") is not None else 0)
        d2h = sum(B * 8 * (L if k == "flat" else 1) for k in want) + B * 4
        fl_step = flops_per_step(model, ns)
        first_key = want[0]
    else:
        from phoskintime_b200.global_model import metric_time_indices, simulate_batch, synthetic_loss_data, synthetic_system
        s = synthetic_system(seed=5, N=120, K=40, max_sites=4, model=0)
        base = s.pack_params()
        allp = base[None, :] * (1.0 + 0.05 * np.random.default_rng(5).uniform(-1, 1, (total, base.size)))
        params_h = torch.empty((B, base.size), dtype=torch.float64).pin_memory()
        params_h.numpy()[:] = allp[lo:hi]
        params_d = params_h.to(dev)
        ld = synthetic_loss_data(s, T_UNION, seed=12)
        mt = metric_time_indices(T_UNION, T_GRID, T_RNA, T_GRID)
        y0g = torch.from_numpy(s.y0()).to(dev)
        y0 = s.y0()
        per_rank = max(parallel.shard_sizes(total, world))
        gathered = torch.empty(world * per_rank, dtype=torch.float64, device=dev) if world > 1 else None
        send_pad = torch.zeros(per_rank, dtype=torch.float64, device=dev) if gathered is not None and per_rank != B else None

        def step_device():
            res = simulate_batch(s, params_d, T_UNION, ("metric", "loss"), y0=y0g, loss_data=ld, metric_times=mt, engine=eng)
            if gathered is not None:
                src = res["metric"]
                if send_pad is not None:
                    send_pad[:B] = src
                    src = send_pad
                eng.allgather_f64(src, gathered)
            return res

        params_e2e = params_h.numpy()
        api = "phoskintime_b200.global_model.simulate_batch (host numpy, pinned)"

        def step_e2e():
            return simulate_batch(s, params_e2e, T_UNION, ("metric", "loss"), y0=y0, loss_data=ld, metric_times=mt, engine=eng)

        h2d = B * base.size * 8
        d2h = B * (8 + 24 + 4)
        Q, nst, nnz_tf = 96, s.idx.state_dim, int(len(s.TF_data))
        # DESIGN.md §3.3.  Lower bound since round 2: most steps solve their six Schur systems by fixed-point sweeps (no
        # inversion, no dense apply; ~20 sweeps of 2 nnz(TF) flops per system) - only that work is counted here; the steps that
        # still invert the |Q| x |Q| block (||K||_inf >= 0.5) do 2 |Q|^3 + 12 |Q|^2 more.
        fl_step = 6 * (20 * 2 * nnz_tf + 2 * nnz_tf + 14 * nst) + 60 * nst
        first_key = "metric"

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
        run.barrier()              # (outside the timed region) all ranks enter the step together: what a rank then waits for
        eng.region_begin()         # inside the step is the collective, not the other ranks' L2 flush
        res = step_device()
        step_ms.append(eng.region_end())
        nl, kms = eng.last_launch_info()
        # last_launch_info refers to the last pk_* call (the all-gather issues no kernel of ours)
        kern_ms.append(kms if world == 1 and name != "morris3" else None)
        launches += nl
    torch.cuda.synchronize()
    run.barrier()
    w1 = time.perf_counter()
    total_ms = run.max_over_ranks(float(np.sum(step_ms)))
    ms_per_step = total_ms / args.steps
    value = total / (ms_per_step * 1e-3)

    nsteps_total = int((res["nsteps"].long() + res["nrej"].long()).sum().item())
    n_bad = int((res["status"] != 0).sum().item())
    kvals = [k for k in kern_ms if k is not None]
    kms = float(np.mean(kvals)) if kvals else float(np.mean(step_ms))
    fl = nsteps_total * fl_step
    achieved_tf = fl / (kms * 1e-3) / 1e12
    hbm_gbs = B * bytes_per_solve(w) / (kms * 1e-3) / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    # ---- end to end through the reference-facing plugin call, host buffers, copies inside
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    run.barrier()
    e0 = time.perf_counter()
    e2e_times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        r = step_e2e()
        _ = float(np.asarray(r[first_key]).reshape(-1)[0])
        e2e_times.append(time.perf_counter() - t0)
    run.barrier()
    e1 = time.perf_counter()
    e2e_s = run.max_over_ranks(float(np.mean(e2e_times)))
    e2e_value = total / e2e_s

    sampler.stop()
    clocks = sampler.summary([(w0, w1), (e0, e1)])

    cpu = None
    extra = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_single(name, args.cpu_sample or {"succ5": 10000, "dist3": 10000, "morris3": 5000, "rand6": 300,
                                                              "normest4": 10000, "global120": 4}[name])
        if name in ("succ5", "dist3"):
            cpu["parity"] = parity_block(eng, name, params_h.numpy(), params_d, y0, y0_d, t_d)
        if name == "succ5" and not args.no_extra:
            extra = extra_lines(eng, dev, flush)

    if rank == 0:
        roof = {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / fp64_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this workload, one `ncu --set full`
                # capture (profiles/r2_succ5_final.txt); the algorithmic bytes are bytes_per_solve x B
                "traffic": 124.4e6 if (name == "succ5" and B == 1_000_000 and world == 1) else None,
                "peak_source": "pk_measure_fp64_peak (register-resident DFMA probe, this run); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "flops_per_step": fl_step,
                "flops_note": "FP64 operations only (FMA = 2), as the kernel performs them; the FP32 error ratio and step "
                              "controller are not counted" + ("; dense kernel: mat-vecs only, the amortised inversions are "
                                                               "not counted" if name == "rand6" else "") +
                              ("; global kernel: the sweep-solved step is counted (a lower bound of the work done; the kernel is "
                               "latency bound, this fraction is not its figure of merit)" if w["kind"] == "global" else ""),
                "steps_per_solve": nsteps_total / B, "kernel_ms": kms,
                "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                        "bytes_per_solve": bytes_per_solve(w)}}
        if w["kind"] == "local" and w["model"] != "randmod":
            roof["flops_per_step_with_fp32_control"] = flops_per_step(w["model"], w["ns"], fp32_control=True)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": w["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(name, world, B if w["scaling"] == "weak" else per_rank),
            "roofline": roof,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "api": api},
            "gpu_launches": launches,
            "clocks": clocks,
            "failed_systems": n_bad,
            "device": eng.device_name,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if extra is not None:
            line["extra"] = extra
        emit(line)
    run.barrier()


def parity_block(eng, name, params_np, params_d, y0, y0_d, t_d):
    """"rel err vs ref" of the metric (part of the CPU leg: the oracle is the checker).
      * ALL systems of the workload: library defaults vs a tight run of a different integrator (RODAS4 at 1e-10/1e-14) on
        the GPU — max pure relative error (floor 1e-12) and max error in units of the parity bound 1e-6*|ref| + 1e-9;
      * the first 100 000 systems: library defaults AND the tight run vs the oracle's exact solution (matrix exponential);
      * the first 256 systems vs the stock reference path (LSODA at its defaults), and that path's own error vs exact."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import local_models as om
    w = WORKLOADS[name]
    model, ns = w["model"], w["ns"]
    got = eng.solve_local_batch(model, params_d, y0_d, ns, t_d, want=("sol",))["sol"]
    tight = eng.solve_local_batch(model, params_d, y0_d, ns, t_d, want=("sol",), method="rodas4", rtol=1e-10, atol=1e-14)["sol"]
    d = (got - tight).abs()
    all_rel = float((d / tight.abs().clamp_min(1e-12)).max())
    all_bound = float((d / (1e-6 * tight.abs() + 1e-9)).max())
    nchk = min(100_000, len(params_np))
    ex = exact_batch(model, params_np[:nchk], y0, ns, T_GRID)
    g, tt = got[:nchk].cpu().numpy(), tight[:nchk].cpu().numpy()
    worst_exact = float((np.abs(g - ex) / (1e-6 * np.abs(ex) + 1e-9)).max())
    rel_max = float((np.abs(g - ex) / np.maximum(np.abs(ex), 1e-12)).max())
    tight_vs_exact = float((np.abs(tt - ex) / (1e-6 * np.abs(ex) + 1e-9)).max())
    worst_stock, stock_rel, stock_bound = 0.0, 0.0, 0.0
    for b in range(256):
        st = om.solve_ode(model, params_np[b], y0, ns, T_GRID)[0]
        worst_stock = max(worst_stock, float((np.abs(g[b] - st) / (1e-6 * np.abs(st) + 1e-7)).max()))
        # the stock reference's OWN error against the exact solution (context for the figures above)
        stock_rel = max(stock_rel, float((np.abs(st - ex[b]) / np.maximum(np.abs(ex[b]), 1e-12)).max()))
        stock_bound = max(stock_bound, float((np.abs(st - ex[b]) / (1e-6 * np.abs(ex[b]) + 1e-9)).max()))
    return {"systems_vs_tight_run": int(got.shape[0]), "max_rel_err_vs_tight_run": all_rel,
            "max_err_over_bound_vs_tight_run": all_bound,
            "systems_vs_exact": nchk, "max_err_over_bound_vs_exact": worst_exact, "max_rel_err_vs_exact": rel_max,
            "tight_run_max_err_over_bound_vs_exact": tight_vs_exact,
            "max_err_over_mixed_bound_vs_stock_reference": worst_stock,
            "stock_reference_self_error_vs_exact": {"systems": 256, "max_rel_err": stock_rel, "max_err_over_bound": stock_bound},
            "bounds": "1e-6*|ref|+1e-9 (exact, tight run), 1e-6*|ref|+1e-7 (stock LSODA path, 256 systems); relative errors "
                      "with floor 1e-12"}


def extra_lines(eng, dev, flush):
    """Kernel-level rates of the other BASELINE shapes on this GPU (device-resident inputs, kernel time by CUDA events
    inside the library, best of 3 after one warm-up, L2 flushed) — context for the headline line, not the contract
    metric; `bench.py --workload NAME` gives each its own full line."""
    import torch
    import phoskintime_b200 as pk
    from phoskintime_b200.global_model import metric_time_indices, simulate_batch, synthetic_loss_data, synthetic_system
    from phoskintime_b200.steady import initial_condition
    out = {}
    t_d = torch.from_numpy(T_GRID).to(dev)

    def timed(fn, reps=3):
        best, res = 1e30, None
        for i in range(reps + 1):
            flush.fill_(1)
            torch.cuda.synchronize()
            res = fn()
            if i:
                best = min(best, eng.last_launch_info()[1])
        return best, res

    for key, model, ns, B, want in (("dist3_1M_fused_score", "distmod", 3, 1 << 20, ("score",)),
                                    ("normest4_256k_fused_loss", "distmod", 4, 256_000, ("ssr",)),
                                    ("rand6_65536_flat", "randmod", 6, 65_536, ("flat",))):
        n, P, L = pk.local_dims(model, ns, len(T_GRID))
        rng = np.random.default_rng(11)
        p = torch.from_numpy(rng.uniform(0.05, 3.0, (B, P))).to(dev)
        y0 = torch.tensor(initial_condition(ns, model), device=dev)
        kw = {}
        if "flat" not in want:
            G = 1000 if key.startswith("normest") else 1
            kw["target"] = torch.from_numpy(rng.uniform(0.1, 2.0, (G, L))).to(dev)
            if G > 1:
                kw["group"] = (torch.arange(B, device=dev, dtype=torch.int32) // (B // G)).to(torch.int32)
        ms, r = timed(lambda: eng.solve_local_batch(model, p, y0, ns, t_d, want=want, **kw))
        steps = float((r["nsteps"].double() + r["nrej"].double()).mean())
        out[key] = {"systems": B, "kernel_ms": ms, "solves_per_s": B / ms * 1e3, "steps_per_solve": steps,
                    "failed": int((r["status"] != 0).sum()),
                    "fp64_tflops": steps * B * flops_per_step(model, ns) / (ms * 1e-3) / 1e12}
    s = synthetic_system(seed=5, N=120, K=40, max_sites=4, model=0)
    B = 2368                                             # 16 systems per SM
    base = s.pack_params()
    P = torch.from_numpy(base[None, :] * (1.0 + 0.05 * np.random.default_rng(5).uniform(-1, 1, (B, base.size)))).to(dev)
    ld = synthetic_loss_data(s, T_UNION, seed=12)
    mt = metric_time_indices(T_UNION, T_GRID, T_RNA, T_GRID)
    y0g = torch.from_numpy(s.y0()).to(dev)
    ms, r = timed(lambda: simulate_batch(s, P, T_UNION, ("metric", "loss"), y0=y0g, loss_data=ld, metric_times=mt, engine=eng), reps=2)
    out["global120_2368_fused_metric_loss"] = {"systems": B, "kernel_ms": ms, "solves_per_s": B / ms * 1e3,
                                               "steps_per_solve": float((r["nsteps"].double() + r["nrej"].double()).mean()),
                                               "failed": int((r["status"] != 0).sum()), "state_dim": int(s.idx.state_dim)}
    return out


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
    ap.add_argument("--workload", default="succ5", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="parameter sets per GPU (weak workloads) or in total (strong workloads)")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the kernel-level rates of the other BASELINE shapes")
    ap.add_argument("--gather-mode", default="p2p", choices=["p2p", "nccl", "fused"],
                    help="N>1, thread-per-system workloads: how the per-sample scalars reach every rank (see step_device)")
    ap.add_argument("--gather-chunks", type=int, default=-1,
                    help="N>1: fuse the all-gather with the solve in this many pieces (0 = one launch + one all-gather; "
                         "-1 = auto: 2 pieces at 8 GPUs for the default workload)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
