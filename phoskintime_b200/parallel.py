"""Sharding of independent systems across the GPUs of one box, and the gather of their
per-sample results.

The reference farms single solves out to a process pool and collects them by pickling
(sensitivity/analysis.py:241-259); here every rank integrates a contiguous block of the sample
axis on its own B200 and ONE all-gather of per-sample doubles (NCCL over NVLink/NVSwitch, issued
by libphoskin_b200's `pk_allgather_f64`) reassembles the result on every rank.  No other
communication exists on this path.

`torch.distributed` is used for the rendezvous only: it carries the 128-byte NCCL unique id to
the other ranks.  With the `gloo` backend (CPU tests) the gather itself also goes through
`torch.distributed`, which exercises the same sharding / padding / reassembly logic.
"""
import os

import numpy as np


def shard_bounds(total, world, rank, align=1):
    """Contiguous block partition of range(total) in units of `align` (e.g. whole Morris
    trajectories of D+1 rows, or whole proteins): returns (lo, hi) for `rank`."""
    if total % align:
        raise ValueError("total must be a multiple of align")
    units = total // align
    base, extra = divmod(units, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo * align, hi * align


def shard_sizes(total, world, align=1):
    return [shard_bounds(total, world, r, align)[1] - shard_bounds(total, world, r, align)[0] for r in range(world)]


class ShardedRun:
    """Per-process context: rank/world from the launcher's environment, one engine on
    cuda:LOCAL_RANK when a GPU backend is in use."""

    def __init__(self, engine=None, backend=None):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.engine = engine
        self.backend = backend
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group(backend=backend or ("nccl" if engine is not None else "gloo"))
        if self.world > 1:
            self.backend = dist.get_backend()
        if engine is not None and self.world > 1:
            ids = [engine.nccl_unique_id() if self.rank == 0 else None]
            if self.backend == "nccl":
                import torch
                torch.cuda.set_device(self.local_rank)
            dist.broadcast_object_list(ids, src=0)
            engine.init_nccl(self.world, self.rank, ids[0])

    def setup_p2p(self, slots_per_rank):
        """Symmetric result buffers for `solve_local_batch(..., gather_p2p=key)`: every rank's [world, slots_per_rank]
        buffer is mapped into every other rank through CUDA IPC (handles travel through the process group)."""
        def exchange(handle):
            if self.world == 1:
                return [handle]
            out = [None] * self.world
            self.dist.all_gather_object(out, handle)
            return out
        view = self.engine.sym_setup(slots_per_rank, exchange)
        self.barrier()                 # every rank has mapped every buffer before anyone launches
        return view

    def bounds(self, total, align=1):
        return shard_bounds(total, self.world, self.rank, align)

    def allgather(self, local, total, align=1):
        """Gather per-sample values (last axes are per-sample payload) from every rank's shard
        into the full [total, ...] array, on every rank.  `local` is a torch tensor (CUDA with
        an engine, CPU under gloo) or a numpy array (CPU)."""
        import torch
        sizes = shard_sizes(total, self.world, align)
        if self.world == 1:
            return local
        is_np = isinstance(local, np.ndarray)
        loc = torch.from_numpy(np.ascontiguousarray(local)) if is_np else local.contiguous()
        host_in = not loc.is_cuda
        if host_in and self.backend == "nccl":               # host results on a GPU run: stage through this rank's GPU
            loc = loc.cuda(self.local_rank)
        inner = int(np.prod(loc.shape[1:])) if loc.dim() > 1 else 1
        pad = max(sizes)
        send = torch.zeros((pad,) + tuple(loc.shape[1:]), dtype=loc.dtype, device=loc.device)
        send[:loc.shape[0]] = loc
        recv = torch.empty((self.world * pad,) + tuple(loc.shape[1:]), dtype=loc.dtype, device=loc.device)
        if self.engine is not None and loc.is_cuda and loc.dtype == torch.float64:
            self.engine.allgather_f64(send, recv)          # NCCL inside the C-ABI library
        else:
            self.dist.all_gather_into_tensor(recv, send) if loc.is_cuda else \
                self.dist.all_gather(list(recv.view(self.world, pad, *loc.shape[1:]).unbind(0)), send)
        parts = [recv[r * pad:r * pad + sizes[r]] for r in range(self.world)]
        full = torch.cat(parts, dim=0)
        assert full.shape[0] == total and inner >= 1
        if host_in and full.is_cuda:
            full = full.cpu()
        return full.numpy() if is_np else full

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, value):
        """max of a python float over ranks (timing)."""
        if self.world == 1:
            return value
        import torch
        dev = "cuda" if self.backend == "nccl" else "cpu"
        t = torch.tensor([value], dtype=torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())
