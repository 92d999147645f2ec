"""Dev/validation script (GPU box, torchrun, N>=2): sharded runs agree with the single-GPU result.
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import phoskintime_b200 as pk
from phoskintime_b200 import parallel, sensitivity
from phoskintime_b200.global_model import run_sensitivity_analysis, synthetic_system
from phoskintime_b200.steady import initial_condition

local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
eng = pk.get_engine(local_rank)
run = parallel.ShardedRun(engine=eng, backend="nccl")
T14 = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])

# local Morris (cfg1 shape, smaller): every rank solves whole trajectories, Y all-gathered by NCCL inside the library
theta = np.random.default_rng(1).uniform(0.05, 3.0, 10)
prob = sensitivity.define_sensitivity_problem_ds(3, theta)
X = sensitivity.morris_sample(prob, 203, 400, seed=42)            # 203 trajectories: ragged over 2/4/8 ranks
y0 = np.asarray(initial_condition(3, "distmod"))
full = eng.solve_local_batch("distmod", X, y0, 3, T14, want=("Y",))["Y"]
lo, hi = run.bounds(X.shape[0], align=11)
mine = eng.solve_local_batch("distmod", torch.from_numpy(X[lo:hi]).cuda(), torch.from_numpy(y0).cuda(), 3,
                             torch.from_numpy(T14).cuda(), want=("Y",))["Y"]
got = run.allgather(mine, X.shape[0], align=11).cpu().numpy()
assert np.array_equal(got, full), np.abs(got - full).max()

# global Morris: sharded by whole trajectories, gathered, identical indices on every rank
s = synthetic_system(seed=11, N=10, K=5, max_sites=3, model=0)
single = run_sensitivity_analysis(s, N=4, num_levels=4, seed=3, engine=eng)
shard = run_sensitivity_analysis(s, N=4, num_levels=4, seed=3, engine=eng, sharded=run)
assert np.array_equal(single["Y"], shard["Y"]) and np.array_equal(single["mu_star"], shard["mu_star"])
# fused solve + overlapped all-gather (pk_local_solve_allgather): equal blocks per rank, result rank-major and
# identical to the plain all-gather of the same scores
Bq = 40_000
rngq = np.random.default_rng(100 + run.rank)
pq = torch.from_numpy(rngq.uniform(0.05, 3.0, (Bq, 14))).cuda()
y0q = torch.tensor(initial_condition(5, "succmod")).cuda()
tq = torch.from_numpy(T14).cuda()
tgq = torch.from_numpy(np.random.default_rng(7).random(93)).cuda()
plain = eng.solve_local_batch("succmod", pq, y0q, 5, tq, want=("score",), target=tgq)["score"]
ref = torch.empty(run.world * Bq, dtype=torch.float64, device="cuda")
eng.allgather_f64(plain, ref)
recv = torch.full((run.world * Bq,), -1.0, dtype=torch.float64, device="cuda")
fused = eng.solve_local_batch("succmod", pq, y0q, 5, tq, want=("score",), target=tgq, gather=("score", recv, 4))
assert torch.equal(fused["score"], plain) and torch.equal(recv, ref)
assert torch.equal(recv[run.rank * Bq:(run.rank + 1) * Bq], plain)
# peer-memory gather (pk_local_solve_gather_p2p): the kernel writes each score into every rank's symmetric buffer over
# NVLink; ragged shard sizes (rank r solves Bq - 1000 r systems); every rank ends up with every rank's scores
view = run.setup_p2p(Bq)
for rep in range(3):
    view.fill_(-7.0)
    torch.cuda.synchronize()
    run.barrier()
    Br = Bq - 1000 * run.rank
    r3 = eng.solve_local_batch("succmod", pq[:Br], y0q, 5, tq, want=("score", "ssr"), target=tgq, gather_p2p="score")
    assert torch.equal(r3["score"], plain[:Br])
    g = r3["gathered"]
    assert g.shape == (run.world, Bq)
    for r in range(run.world):
        assert torch.equal(g[r, :Bq - 1000 * r], ref[r * Bq:r * Bq + Bq - 1000 * r]), (rep, r)
        assert bool((g[r, Bq - 1000 * r:] == -7.0).all())
    run.barrier()
run.barrier()
print(f"rank {run.rank}/{run.world}: peer-memory gather and fused gather identical to the plain all-gather; sharded local Morris Y and global Morris indices identical to single-GPU", flush=True)
