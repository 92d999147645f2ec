"""All five BASELINE.json configs on one GPU (device-resident inputs, kernel time by CUDA events inside the
library) — evidence for DESIGN.md/profiles; bench.py stays the contract benchmark (config[1]).

    python tools/bench_configs.py [--quick] [--out profiles/r1_configs.json]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import phoskintime_b200 as pk
from phoskintime_b200 import sensitivity
from phoskintime_b200.global_model import metric_time_indices, simulate_batch, synthetic_loss_data, synthetic_system
from phoskintime_b200.steady import initial_condition

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--out", default=None)
args = ap.parse_args()
eng = pk.get_engine(0)
dev = torch.device("cuda", 0)
T14 = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
T15 = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 15.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
rows = []


def timed(fn, reps=3):
    best, res = 1e30, None
    for _ in range(reps):
        res = fn()
        best = min(best, eng.last_launch_info()[1])
    return best, res


def report(name, B, ms, res, extra=None):
    st = (res["nsteps"].double() + res["nrej"].double()) if torch.is_tensor(res["nsteps"]) else (res["nsteps"] + res["nrej"]).astype(float)
    bad = int((res["status"] != 0).sum())
    row = {"config": name, "systems": B, "kernel_ms": round(ms, 3), "solves_per_s": B / ms * 1e3, "steps_per_solve": float(st.mean()),
           "failed": bad, **(extra or {})}
    rows.append(row)
    print(json.dumps(row), flush=True)


def local_case(name, model, ns, B, want, seed, groups=None):
    n, P, L = pk.local_dims(model, ns, 14)
    rng = np.random.default_rng(seed)
    params = torch.from_numpy(rng.uniform(0.05, 3.0, (B, P))).to(dev)
    y0 = torch.tensor(initial_condition(ns, model), device=dev)
    tt = torch.from_numpy(T14).to(dev)
    kw = {}
    if "score" in want or "ssr" in want:
        G = groups or 1
        kw["target"] = torch.from_numpy(rng.uniform(0.1, 2.0, (G, L))).to(dev)
        if groups:
            kw["group"] = torch.arange(B, device=dev, dtype=torch.int32) // (B // G)
    ms, res = timed(lambda: eng.solve_local_batch(model, params, y0, ns, tt, want=want, **kw))
    report(name, B, ms, res)


# cfg1: distributive 3 sites, Morris N=1000 -> 11 000 rows, Y + elementary effects
theta = np.random.default_rng(1).uniform(0.05, 3.0, 10)
prob = sensitivity.define_sensitivity_problem_ds(3, theta)
X = torch.from_numpy(sensitivity.morris_sample(prob, 1000, 400, seed=42)).to(dev)
y0 = torch.tensor(initial_condition(3, "distmod"), device=dev)
tt = torch.from_numpy(T14).to(dev)
ms, res = timed(lambda: eng.solve_local_batch("distmod", X, y0, 3, tt, want=("Y",)))
t0 = time.perf_counter(); Si = eng.morris_ee(X, res["Y"], 400, scaled=True); torch.cuda.synchronize(); ee_ms = (time.perf_counter() - t0) * 1e3
report("cfg1 distmod ns=3 Morris N=1000 (11000 rows), fused Y", X.shape[0], ms, res, {"morris_ee_wall_ms": round(ee_ms, 3)})
local_case("cfg1b distmod ns=3, 1M sets, fused score (north-star target shape)", "distmod", 3, 1 << 20 if not args.quick else 1 << 16, ("score",), 11)
local_case("cfg2 succmod ns=5, 1M sets, fused ssr+score", "succmod", 5, 1_000_000 if not args.quick else 1 << 16, ("ssr", "score"), 2)
local_case("cfg3 randmod ns=6 (65 states), 262144 sets, flat out", "randmod", 6, 262144 if not args.quick else 2048, ("flat",), 3)
local_case("cfg4 distmod ns=4, 1000 proteins x 256 starts, fused loss", "distmod", 4, 256000 if not args.quick else 25600, ("ssr",), 4, groups=1000 if not args.quick else 100)

# cfg5: global network N=120 (state_dim ~530), 16384 parameter vectors (+-5 %), fused metric + losses
s = synthetic_system(seed=5, N=120, K=40, max_sites=4, model=0)
B = 16384 if not args.quick else 592
base = s.pack_params()
P = torch.from_numpy(base[None, :] * (1.0 + 0.05 * np.random.default_rng(5).uniform(-1, 1, (B, base.size)))).to(dev)
ld = synthetic_loss_data(s, T15, seed=12)
mt = metric_time_indices(T15, T14, [4.0, 8.0, 15.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0], T14)
y0g = torch.from_numpy(s.y0()).to(dev)
ms, res = timed(lambda: simulate_batch(s, P, T15, ("metric", "loss"), y0=y0g, loss_data=ld, metric_times=mt, engine=eng), reps=2)
report("cfg5 global network N=120 K=40 state_dim=%d, fused metric+loss" % s.idx.state_dim, B, ms, res,
       {"dims": eng.global_dims(s._topo_id[eng.token])})
# the same network with the combinatorial kinetic model (one state per phosphorylation pattern, ~1000 states)
s2 = synthetic_system(seed=5, N=120, K=40, max_sites=4, model=2)
B2 = 4096 if not args.quick else 296
P2 = P[:B2].contiguous()
ld2 = synthetic_loss_data(s2, T15, seed=12)
y0g2 = torch.from_numpy(s2.y0()).to(dev)
ms, res = timed(lambda: simulate_batch(s2, P2, T15, ("metric", "loss"), y0=y0g2, loss_data=ld2, metric_times=mt, engine=eng), reps=2)
report("cfg5-comb global network, combinatorial model N=120 K=40 state_dim=%d, fused metric+loss" % s2.idx.state_dim, B2, ms, res,
       {"dims": eng.global_dims(s2._topo_id[eng.token])})
if args.out:
    json.dump({"device": eng.device_name, "rows": rows}, open(args.out, "w"), indent=1)
