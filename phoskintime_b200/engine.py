"""Host-side engine: one handle per GPU, batched solves through the C ABI.

`Engine.solve_local_batch` is the batched form of the reference's
`solve_ode(params, init_cond, num_psites, t)` (models/distmod.py:93-134 and siblings): B
parameter sets in, any of {sol, flat, Y, ssr, score} out.  Inputs may be numpy arrays (host
path: the library stages them to HBM) or torch CUDA tensors (device path: used in place,
outputs come back as torch tensors on the same device).  PyTorch is only the owner of device
buffers here; all arithmetic happens in libphoskin_b200.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (GLOBAL_METRIC_IDS, METHOD_IDS, MODEL_IDS, PK_DEVICE, PK_HOST, Y_METRIC_IDS, PhoskinError,
                   PkGlobalJob, PkGlobalLossData, PkGlobalTopology, PkLocalJob)

# Tolerances left at None select the LIBRARY defaults (rtol/atol <= 0 in the C ABI), which depend on the method the kernel
# uses: ROS6L (thread-per-system kernels) 2e-5/2e-11, ROS5L (dense kernel) and the global network 2e-6/2e-9 — each
# chosen from the measured error against the reference's tight solution (<= 0.15-0.27 of the 1e-6 parity bound on every
# golden and on harsh parameter draws) — DESIGN.md §2/§5.  The two constants are the ROS5L / global values.
DEFAULT_RTOL = 2e-6
DEFAULT_ATOL = 2e-9

_engines = {}


def local_dims(model, num_psites, T):
    """(n_states, n_params, flat_len) for a model — pure arithmetic, no device needed."""
    lib = _lib.load()
    n, P, L = C.c_int(), C.c_int(), C.c_int()
    _lib.check(lib.pk_local_dims(MODEL_IDS[model], int(num_psites), int(T), C.byref(n), C.byref(P), C.byref(L)))
    return n.value, P.value, L.value


def _is_torch(x):
    return type(x).__module__.split(".")[0] == "torch"


class Engine:
    """Owns one pk_handle (stream + device workspaces) on one GPU."""

    _next_token = 0

    def __init__(self, device=0):
        self.lib = _lib.load()
        self.device = int(device)
        # unique per Engine object for the life of the process: host-side caches (uploaded topologies, installed loss
        # tables) are keyed by it — id(engine) can be reused by a later Engine after this one is collected
        Engine._next_token += 1
        self.token = Engine._next_token
        h = C.c_void_p()
        _lib.check(self.lib.pk_create(self.device, C.byref(h)))
        self._h = h
        sm, khz = C.c_int(), C.c_int()
        name = C.create_string_buffer(128)
        _lib.check(self.lib.pk_device_info(self._h, C.byref(sm), C.byref(khz), name, 128))
        self.sm_count, self.clock_khz, self.device_name = sm.value, khz.value, name.value.decode()
        self.world, self.rank = 1, 0
        self._sym = None            # torch view of the symmetric buffer of the peer-memory gather

    def close(self):
        self._sym = None
        if getattr(self, "_h", None):
            self.lib.pk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------ info
    def last_launch_info(self):
        n, ms = C.c_int(), C.c_float()
        _lib.check(self.lib.pk_last_launch_info(self._h, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def region_begin(self):
        _lib.check(self.lib.pk_region_begin(self._h))

    def region_end(self):
        ms = C.c_float()
        _lib.check(self.lib.pk_region_end(self._h, C.byref(ms)))
        return ms.value

    def measure_fp64_peak(self):
        tf, ms = C.c_double(), C.c_float()
        _lib.check(self.lib.pk_measure_fp64_peak(self._h, C.byref(tf), C.byref(ms)))
        return tf.value

    # ----------------------------------------------------------------------------- solve
    def solve_local_batch(self, model, params, init_cond, num_psites, t, want=("sol", "flat"), *,
                          target=None, sigma=None, group=None, lam=0.0, y_metric="total_signal",
                          rtol=None, atol=None, max_steps=0, normalize=False, log_params=False,
                          score_weights=(1.0, 1.0, 1.0, 1.0, 1.0), out=None, counters=True, method=None, gather=None,
                          gather_p2p=None):
        """Solve B systems.  Returns a dict with the requested keys among
        sol[B,T,n], flat[B,L], Y[B], ssr[B], score[B] plus status/nsteps/nrej[B] (int32).

        params : [B,P] (numpy or torch.cuda float64, C-contiguous)
        init_cond : [n] shared, or [B,n] per system
        target/sigma/group : fused-loss inputs — target [L] or [G,L]; sigma None, [L]/[L+P] or
            [G,·]; group [B] int32 indices into G (None = all 0)
        out : optional dict of preallocated outputs (same kind as params) to fill
        counters : also return accepted/rejected step counts per system (nsteps, nrej)
        method : None (library default: 'ros6l' on the thread-per-system kernels, 'ros5l' on the dense kernel),
            'ros6l', 'ros5l' or 'rodas4' — DESIGN.md §2
        gather : (key, recv, chunks) — fuse the solve with the NCCL all-gather of the per-sample output `key`
            ('score', 'ssr' or 'Y'; torch CUDA path only): the batch is integrated in `chunks` pieces and piece c's
            gather overlaps piece c+1 (`pk_local_solve_allgather`); recv [world*B] comes back rank-major.  Every rank
            must pass the same B.
        gather_p2p : key — the same gather WITHOUT a collective: the kernel stores every finished system's `key` straight
            into all ranks' symmetric buffers over NVLink peer memory (`pk_local_solve_gather_p2p`, after `sym_setup`);
            res['gathered'] is the [world, slots] view of this rank's buffer.  Ranks may pass different B.
        """
        want = tuple(want)
        unknown = set(want) - {"sol", "flat", "Y", "ssr", "score"}
        if unknown:
            raise ValueError(f"unknown outputs {sorted(unknown)}")
        dev = _is_torch(params)
        xp = _TorchOps(params.device) if dev else _NumpyOps()
        params = xp.f64(params)
        if params.ndim == 1:
            params = params.reshape(1, -1)
        B = int(params.shape[0])
        t_arr = xp.f64(t).reshape(-1)
        T = int(t_arr.shape[0])
        n, P, L = local_dims(model, num_psites, T)
        if params.shape[1] != P:
            raise ValueError(f"{model} with {num_psites} sites takes {P} parameters, got {params.shape[1]}")
        y0 = xp.f64(init_cond)
        if y0.ndim == 1:
            if y0.shape[0] != n:
                raise ValueError(f"init_cond must have {n} entries, got {y0.shape[0]}")
            y0_stride = 0
        else:
            if tuple(y0.shape) != (B, n):
                raise ValueError(f"init_cond must be [{n}] or [{B},{n}]")
            y0_stride = n

        job = PkLocalJob()
        self.lib.pk_local_job_init(C.byref(job))
        job.model, job.n_sites, job.B, job.T = MODEL_IDS[model], int(num_psites), B, T
        job.memspace = PK_DEVICE if dev else PK_HOST
        job.params, job.y0, job.y0_stride, job.t = xp.ptr(params), xp.ptr(y0), y0_stride, xp.ptr(t_arr)
        job.rtol = 0.0 if rtol is None else float(rtol)          # 0 -> the method's library default
        job.atol = 0.0 if atol is None else float(atol)
        job.max_steps, job.normalize, job.log_params = int(max_steps), int(bool(normalize)), int(bool(log_params))
        lam_arr = None if np.ndim(lam) == 0 else xp.f64(lam).reshape(-1)      # per-group lambda [G]
        job.lam = float(lam) if lam_arr is None else 0.0
        job.method = METHOD_IDS[method]
        for i, w in enumerate(score_weights):
            job.score_w[i] = float(w)

        keep = [params, y0, t_arr]
        res = {}
        out = out or {}

        def alloc(key, shape, dtype="f64"):
            buf = out.get(key)
            if buf is None:
                buf = xp.empty(shape, dtype)
            res[key] = buf
            return xp.ptr(buf)

        if "sol" in want:
            job.out_sol = alloc("sol", (B, T, n))
        if "flat" in want:
            job.out_flat = alloc("flat", (B, L))
        if "Y" in want:
            job.y_metric = Y_METRIC_IDS[y_metric]
            job.out_Y = alloc("Y", (B,))
        if "ssr" in want or "score" in want:
            if target is None:
                raise ValueError("ssr/score need target")
            tg = xp.f64(target)
            tg = tg.reshape(1, -1) if tg.ndim == 1 else tg
            if tg.shape[1] != L:
                raise ValueError(f"target must have {L} columns")
            G = int(tg.shape[0])
            job.target, job.n_groups = xp.ptr(tg), G
            keep.append(tg)
            if sigma is not None:
                sg = xp.f64(sigma)
                sg = sg.reshape(1, -1) if sg.ndim == 1 else sg
                if sg.shape[0] != G or sg.shape[1] not in (L, L + P):
                    raise ValueError(f"sigma must be [{G},{L}] or [{G},{L + P}]")
                job.sigma, job.sigma_len = xp.ptr(sg), int(sg.shape[1])
                keep.append(sg)
            if group is not None:
                gr = xp.i32(group).reshape(-1)
                if gr.shape[0] != B:
                    raise ValueError("group must have B entries")
                job.group = xp.ptr(gr)
                keep.append(gr)
            elif G != 1:
                raise ValueError("several target rows need `group`")
            if lam_arr is not None:
                if lam_arr.shape[0] != G:
                    raise ValueError(f"per-group lam must have {G} entries")
                job.lam_group = xp.ptr(lam_arr)
                keep.append(lam_arr)
            if "ssr" in want:
                job.out_ssr = alloc("ssr", (B,))
            if "score" in want:
                job.out_score = alloc("score", (B,))
        job.out_status = alloc("status", (B,), "i32")
        if counters:
            job.out_nsteps = alloc("nsteps", (B,), "i32")
            job.out_nrej = alloc("nrej", (B,), "i32")

        if dev:
            xp.sync()          # inputs produced on torch's stream must be visible to ours
        if gather is not None and gather_p2p is not None:
            raise ValueError("gather and gather_p2p are alternatives")
        if gather is not None:
            key, recv, chunks = gather
            if not dev or key not in res or key not in ("score", "ssr", "Y"):
                raise ValueError("gather needs torch CUDA buffers and a requested per-sample output (score, ssr or Y)")
            # precondition of pk_local_solve_allgather: the same B on every rank (pad the shorter shards), and a
            # float64 landing buffer of world*B entries on this device
            import torch
            if not (_is_torch(recv) and recv.is_cuda and recv.dtype == torch.float64 and recv.is_contiguous()):
                raise ValueError("gather: recv must be a contiguous torch CUDA float64 tensor")
            if recv.device != params.device or recv.numel() < self.world * B:
                raise ValueError(f"gather: recv must live on {params.device} and hold world*B = {self.world * B} entries "
                                 "(every rank must pass the same B: pad uneven shards)")
            _lib.check(self.lib.pk_local_solve_allgather(self._h, C.byref(job), {"score": 0, "ssr": 1, "Y": 2}[key],
                                                         int(chunks), recv.data_ptr()))
            res["gathered"] = recv
        elif gather_p2p is not None:
            key = gather_p2p
            if not dev or key not in res or key not in ("score", "ssr", "Y"):
                raise ValueError("gather_p2p needs torch CUDA buffers and a requested per-sample output (score, ssr or Y)")
            if self._sym is None:
                raise PhoskinError("gather_p2p: call Engine.sym_setup (or ShardedRun.setup_p2p) first")
            _lib.check(self.lib.pk_local_solve_gather_p2p(self._h, C.byref(job), {"score": 0, "ssr": 1, "Y": 2}[key]))
            res["gathered"] = self._sym          # [world, slots_per_rank]: row r = rank r's values (first B_r entries)
        else:
            _lib.check(self.lib.pk_local_solve_batch(self._h, C.byref(job)))
        del keep
        return res

    # ------------------------------------------------------------------------------ nlls
    def nlls_local_batch(self, model, theta0, init_cond, num_psites, t, target, lb, ub, *, sigma=None, group=None,
                         lam=0.0, log_params=False, max_iter=100, ftol=1e-8, xtol=1e-8, gtol=1e-8, fd_rel=None,
                         rtol=None, atol=None, max_steps=0, method=None, score_weights=(1.0, 1.0, 1.0, 1.0, 1.0)):
        """B bounded least-squares problems in one call (`pk_local_nlls_batch`): the batched form of
        `curve_fit(model_func, time_points, target_fit, p0, bounds=free_bounds, sigma=sigma, x_scale='jac')`
        (paramest/normest.py:278-290) with the residual model of normest.py:403-423.

        theta0 [B,P] start points; target [L] or [G,L]; sigma None, [L]/[L+P] or [G,.]; group [B] rows of target;
        lb/ub [P] finite.  Returns dict(theta[B,P], cost[B] = 0.5*sum(r^2), score[B] = score_fit at the optimum,
        status[B] (1 gtol, 2 ftol, 3 xtol, 4 max_iter, -1 failed), iters[B], nfev[B])."""
        dev = _is_torch(theta0)
        xp = _TorchOps(theta0.device) if dev else _NumpyOps()
        theta = xp.f64(theta0)
        theta = (theta.reshape(1, -1) if theta.ndim == 1 else theta)
        theta = theta.clone() if dev else theta.copy()
        B = int(theta.shape[0])
        t_arr = np.ascontiguousarray(np.asarray(t, dtype=np.float64).reshape(-1))
        T = int(t_arr.shape[0])
        n, P, L = local_dims(model, num_psites, T)
        if theta.shape[1] != P:
            raise ValueError(f"{model} with {num_psites} sites takes {P} parameters, got {theta.shape[1]}")
        lb = np.ascontiguousarray(np.broadcast_to(np.asarray(lb, dtype=np.float64), (P,)))
        ub = np.ascontiguousarray(np.broadcast_to(np.asarray(ub, dtype=np.float64), (P,)))
        y0 = xp.f64(init_cond)
        if y0.ndim == 1:
            if y0.shape[0] != n:
                raise ValueError(f"init_cond must have {n} entries")
            y0_stride = 0
        else:
            if tuple(y0.shape) != (B, n):
                raise ValueError(f"init_cond must be [{n}] or [{B},{n}]")
            y0_stride = n
        tg = xp.f64(target)
        tg = tg.reshape(1, -1) if tg.ndim == 1 else tg
        if tg.shape[1] != L:
            raise ValueError(f"target must have {L} columns")
        G = int(tg.shape[0])
        job = _lib.PkNllsJob()
        self.lib.pk_nlls_job_init(C.byref(job))
        job.model, job.n_sites, job.B, job.T = MODEL_IDS[model], int(num_psites), B, T
        job.memspace = PK_DEVICE if dev else PK_HOST
        job.theta, job.y0, job.y0_stride, job.t = xp.ptr(theta), xp.ptr(y0), y0_stride, t_arr.ctypes.data
        job.lb, job.ub, job.target, job.n_groups = lb.ctypes.data, ub.ctypes.data, xp.ptr(tg), G
        keep = [theta, y0, t_arr, lb, ub, tg]
        if sigma is not None:
            sg = xp.f64(sigma)
            sg = sg.reshape(1, -1) if sg.ndim == 1 else sg
            if sg.shape[0] != G or sg.shape[1] not in (L, L + P):
                raise ValueError(f"sigma must be [{G},{L}] or [{G},{L + P}]")
            job.sigma, job.sigma_len = xp.ptr(sg), int(sg.shape[1])
            keep.append(sg)
        if group is not None:
            gr = xp.i32(group).reshape(-1)
            if gr.shape[0] != B:
                raise ValueError("group must have B entries")
            job.group = xp.ptr(gr)
            keep.append(gr)
        elif G != 1:
            raise ValueError("several target rows need `group`")
        if np.ndim(lam) == 0:
            job.lam = float(lam)
        else:                                     # per-group lambda [G] (the lambda scan as groups of one call)
            lam_arr = xp.f64(lam).reshape(-1)
            if lam_arr.shape[0] != G:
                raise ValueError(f"per-group lam must have {G} entries")
            job.lam, job.lam_group = 0.0, xp.ptr(lam_arr)
            keep.append(lam_arr)
        job.log_params, job.max_iter = int(bool(log_params)), int(max_iter)
        job.ftol, job.xtol, job.gtol = float(ftol), float(xtol), float(gtol)
        if fd_rel is not None:
            job.fd_rel = float(fd_rel)
        if rtol is not None:
            job.rtol = float(rtol)
        if atol is not None:
            job.atol = float(atol)
        job.max_steps, job.method = int(max_steps), METHOD_IDS[method]
        for i, w in enumerate(score_weights):
            job.score_w[i] = float(w)
        res = {"theta": theta, "cost": xp.empty((B,), "f64"), "score": xp.empty((B,), "f64"),
               "status": xp.empty((B,), "i32"), "iters": xp.empty((B,), "i32"), "nfev": xp.empty((B,), "i32")}
        job.out_cost, job.out_score = xp.ptr(res["cost"]), xp.ptr(res["score"])
        job.out_status, job.out_iters, job.out_nfev = xp.ptr(res["status"]), xp.ptr(res["iters"]), xp.ptr(res["nfev"])
        if dev:
            xp.sync()
        _lib.check(self.lib.pk_local_nlls_batch(self._h, C.byref(job)))
        del keep
        return res

    # ---------------------------------------------------------------------------- morris
    def morris_ee(self, X, Y, num_levels, scaled=False, want_ee=False):
        """mu, mu*, sigma (and optionally the EE matrix) from trajectories X[N(D+1),D], Y[N(D+1)]
        — SALib.analyze.morris as called at sensitivity/analysis.py:264-265 (without the
        bootstrap confidence column)."""
        dev = _is_torch(X)
        xp = _TorchOps(X.device) if dev else _NumpyOps()
        X = xp.f64(X)
        Y = xp.f64(Y).reshape(-1)
        D = int(X.shape[1])
        rows = int(X.shape[0])
        if rows % (D + 1) or Y.shape[0] != rows:
            raise ValueError("X must hold N*(D+1) rows and Y one value per row")
        N = rows // (D + 1)
        mu, mus, sig = xp.empty((D,), "f64"), xp.empty((D,), "f64"), xp.empty((D,), "f64")
        ee = xp.empty((N, D), "f64") if want_ee else None
        if dev:
            xp.sync()
        _lib.check(self.lib.pk_morris_ee(self._h, PK_DEVICE if dev else PK_HOST, xp.ptr(X), xp.ptr(Y), N, D,
                                         int(num_levels), int(bool(scaled)), xp.ptr(mu), xp.ptr(mus),
                                         xp.ptr(sig), xp.ptr(ee) if want_ee else None))
        res = {"mu": mu, "mu_star": mus, "sigma": sig}
        if want_ee:
            res["ee"] = ee
        return res

    # --------------------------------------------------------------------- global network
    def global_upload(self, sys_, force_generic=False):
        """Upload the static topology of a `GlobalSystem` (the array part of the reference's
        System.odeint_args(), global_model/network.py:508-526) once; returns the topology id.
        `force_generic` (testing): 1/True = shared-memory LU of the Schur block although the register path fits,
        2 = additionally every large per-system array in the per-CTA global scratch (the capacity fallback the
        library chooses by itself when a network does not fit into one CTA's shared memory)."""
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        arrs = {"n_sites": i32(sys_.idx.n_sites), "W_indptr": i32(sys_.W_indptr), "W_indices": i32(sys_.W_indices),
                "W_data": f64(sys_.W_data), "TF_indptr": i32(sys_.TF_indptr), "TF_indices": i32(sys_.TF_indices),
                "TF_data": f64(sys_.TF_data), "kin_grid": f64(sys_.kin_grid), "kin_Kmat": f64(sys_.kin_Kmat),
                "tf_deg": f64(sys_.tf_deg), "driver_map": i32(sys_.driver_map)}
        tp = PkGlobalTopology()
        tp.model, tp.N, tp.K, tp.n_bins = int(sys_.model), int(sys_.idx.N), int(sys_.K), int(arrs["kin_grid"].size)
        for k, a in arrs.items():
            setattr(tp, k, a.ctypes.data)
        tp.force_generic_schur = int(force_generic)
        tid = C.c_int32(-1)
        _lib.check(self.lib.pk_global_upload(self._h, C.byref(tp), C.byref(tid)))
        return tid.value

    def global_dims(self, topo):
        v = [C.c_int32() for _ in range(4)]
        _lib.check(self.lib.pk_global_dims(self._h, int(topo), *[C.byref(x) for x in v]))
        c = [C.c_int32() for _ in range(3)]
        _lib.check(self.lib.pk_global_counts(self._h, int(topo), *[C.byref(x) for x in c]))
        return {"state_dim": v[0].value, "n_params": v[1].value, "n_reg": v[2].value, "smem_bytes": v[3].value,
                "n_proteins": c[0].value, "n_kinases": c[1].value, "total_sites": c[2].value}

    def global_set_loss_data(self, topo, ld):
        """Install the observation tables of LOSS_FN (global_model/lossfn.py:113-121; built by
        global_model/cache.py:19-155) for a topology.  `prot_map` is implied by the topology."""
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        t = PkGlobalLossData()
        keep = []
        for k in ("p_prot", "t_prot", "p_rna", "t_rna", "p_pho", "s_pho", "t_pho"):
            a = i32(ld[k]); keep.append(a); setattr(t, k, a.ctypes.data)
        for k in ("obs_prot", "w_prot", "obs_rna", "w_rna", "obs_pho", "w_pho"):
            a = f64(ld[k]); keep.append(a); setattr(t, k, a.ctypes.data)
        t.n_prot, t.n_rna, t.n_pho = len(ld["p_prot"]), len(ld["p_rna"]), len(ld["p_pho"])
        if not (len(ld["t_prot"]) == len(ld["obs_prot"]) == len(ld["w_prot"]) == t.n_prot and
                len(ld["t_rna"]) == len(ld["obs_rna"]) == len(ld["w_rna"]) == t.n_rna and
                len(ld["s_pho"]) == len(ld["t_pho"]) == len(ld["obs_pho"]) == len(ld["w_pho"]) == t.n_pho):
            raise ValueError("loss tables of one modality must have equal lengths")
        t.prot_base_idx, t.rna_base_idx, t.pho_base_idx = (int(ld["prot_base_idx"]), int(ld["rna_base_idx"]),
                                                           int(ld["pho_base_idx"]))
        _lib.check(self.lib.pk_global_set_loss_data(self._h, int(topo), C.byref(t)))

    def global_set_prior(self, topo, defaults_vec):
        a = None if defaults_vec is None else np.ascontiguousarray(defaults_vec, dtype=np.float64)
        _lib.check(self.lib.pk_global_set_prior(self._h, int(topo), None if a is None else a.ctypes.data))

    def global_release(self, topo):
        _lib.check(self.lib.pk_global_release(self._h, int(topo)))

    def global_solve_batch(self, topo, params, y0, t_eval, want=("Y",), *, rtol=None, atol=None, max_steps=0,
                           theta_mode=False, loss_mode=0, metric="total_signal", metric_times=None,
                           lambdas=(1.0, 1.0, 1.0), lambda_prior=0.0, out=None):
        """Integrate B parameter vectors of one uploaded network.  Returns a dict with the requested
        keys among Y[B,T,state_dim], loss[B,3], F[B,3], metric[B], fc[B,n_fc] plus status/nsteps/nrej[B].
        fc = the fold-change table of simulate_and_measure (simulate.py:105-182): [N*len(t_prot) | N*len(t_rna) |
        total_sites*len(t_pho)], protein-major then (site,) time.

        params : [B,P] physical values (or raw theta with theta_mode=True: softplus is applied on the
                 device, global_model/params.py:106-132), order c_k|A|B|C|D|Dp|E|tf_scale
        metric_times : dict(t_prot, t_rna, t_pho index lists, prot_b, rna_b, pho_b) — which rows of
                 t_eval simulate_and_measure tabulates (global_model/simulate.py:105-182)
        """
        want = tuple(want)
        unknown = set(want) - {"Y", "loss", "F", "metric", "fc"}
        if unknown:
            raise ValueError(f"unknown outputs {sorted(unknown)}")
        dev = _is_torch(params)
        xp = _TorchOps(params.device) if dev else _NumpyOps()
        params = xp.f64(params)
        if params.ndim == 1:
            params = params.reshape(1, -1)
        B = int(params.shape[0])
        dims = self.global_dims(topo)
        n, P = dims["state_dim"], dims["n_params"]
        if params.shape[1] != P:
            raise ValueError(f"this network takes {P} parameters, got {params.shape[1]}")
        t_arr = np.ascontiguousarray(t_eval, dtype=np.float64).reshape(-1)
        T = int(t_arr.size)
        y0 = xp.f64(y0)
        if y0.ndim == 1:
            if y0.shape[0] != n:
                raise ValueError(f"y0 must have {n} entries")
            y0_stride = 0
        else:
            if tuple(y0.shape) != (B, n):
                raise ValueError(f"y0 must be [{n}] or [{B},{n}]")
            y0_stride = n
        job = PkGlobalJob()
        self.lib.pk_global_job_init(C.byref(job))
        job.topo, job.memspace, job.B, job.T = int(topo), PK_DEVICE if dev else PK_HOST, B, T
        job.theta_mode = int(bool(theta_mode))
        job.params, job.y0, job.y0_stride, job.t_eval = xp.ptr(params), xp.ptr(y0), y0_stride, t_arr.ctypes.data
        job.rtol = 0.0 if rtol is None else float(rtol)
        job.atol = 0.0 if atol is None else float(atol)
        job.max_steps, job.loss_mode = int(max_steps), int(loss_mode)
        for i in range(3):
            job.lambdas[i] = float(lambdas[i])
        job.lambda_prior = float(lambda_prior)
        res, keep = {}, [params, y0, t_arr]
        out = out or {}

        def alloc(key, shape, dtype="f64"):
            buf = out.get(key)
            if buf is None:
                buf = xp.empty(shape, dtype)
            res[key] = buf
            return xp.ptr(buf)

        if "Y" in want:
            job.out_Y = alloc("Y", (B, T, n))
        if "loss" in want:
            job.out_loss = alloc("loss", (B, 3))
        if "F" in want:
            job.out_F = alloc("F", (B, 3))
        if "metric" in want or "fc" in want:
            if metric_times is None:
                raise ValueError("metric / fc need metric_times")
            job.metric = GLOBAL_METRIC_IDS[metric]
            for name, key in (("prot", "t_prot"), ("rna", "t_rna"), ("pho", "t_pho")):
                a = np.ascontiguousarray(metric_times[key], dtype=np.int32).reshape(-1)
                keep.append(a)
                setattr(job, "n_mt_" + name, int(a.size))
                setattr(job, "mt_" + name, a.ctypes.data)
            job.mb_prot, job.mb_rna, job.mb_pho = (int(metric_times["prot_b"]), int(metric_times["rna_b"]),
                                                   int(metric_times["pho_b"]))
            if "metric" in want:
                job.out_metric = alloc("metric", (B,))
            if "fc" in want:
                n_fc = dims["n_proteins"] * (job.n_mt_prot + job.n_mt_rna) + dims["total_sites"] * job.n_mt_pho
                job.out_fc = alloc("fc", (B, n_fc))
        job.out_status = alloc("status", (B,), "i32")
        job.out_nsteps = alloc("nsteps", (B,), "i32")
        job.out_nrej = alloc("nrej", (B,), "i32")
        if dev:
            xp.sync()
        _lib.check(self.lib.pk_global_solve_batch(self._h, C.byref(job)))
        del keep
        return res

    def global_loss_batch(self, topo, Y, loss_mode=0):
        """(loss_p, loss_r, loss_ph) per trajectory for Y[B,T,state_dim] (or [T,state_dim])."""
        dev = _is_torch(Y)
        xp = _TorchOps(Y.device) if dev else _NumpyOps()
        Y = xp.f64(Y)
        if Y.ndim == 2:
            Y = Y.reshape(1, *Y.shape)
        n = self.global_dims(topo)["state_dim"]
        if Y.ndim != 3 or Y.shape[2] != n:
            raise ValueError(f"Y must be [B,T,{n}]")
        B, T = int(Y.shape[0]), int(Y.shape[1])
        out = xp.empty((B, 3), "f64")
        if dev:
            xp.sync()
        _lib.check(self.lib.pk_global_loss_batch(self._h, int(topo), PK_DEVICE if dev else PK_HOST, xp.ptr(Y), B, T,
                                                 int(loss_mode), xp.ptr(out)))
        return out

    def global_solve_custom(self, topo, params, y0, t_eval, *, rtol=None, atol=None, max_steps=0, theta_mode=False):
        """The reference's custom DOPRI5 solver (`solve_custom`, jacspeedup.py:31-67 / solvers.py:292-758) for B parameter
        vectors — `pk_global_solve_custom`.  Returns dict(Y[B,T,n], status, nsteps, nrej)."""
        dev = _is_torch(params)
        xp = _TorchOps(params.device) if dev else _NumpyOps()
        params = xp.f64(params)
        params = params.reshape(1, -1) if params.ndim == 1 else params
        B = int(params.shape[0])
        dims = self.global_dims(topo)
        n = dims["state_dim"]
        if params.shape[1] != dims["n_params"]:
            raise ValueError(f"params must have {dims['n_params']} columns")
        y0a = xp.f64(y0)
        stride = 0 if y0a.ndim == 1 else n
        if (y0a.ndim == 1 and y0a.shape[0] != n) or (y0a.ndim == 2 and tuple(y0a.shape) != (B, n)):
            raise ValueError(f"y0 must be [{n}] or [{B},{n}]")
        t_arr = np.ascontiguousarray(np.asarray(t_eval, dtype=np.float64).reshape(-1))
        T = t_arr.shape[0]
        res = {"Y": xp.empty((B, T, n), "f64"), "status": xp.empty((B,), "i32"), "nsteps": xp.empty((B,), "i32"),
               "nrej": xp.empty((B,), "i32")}
        if dev:
            xp.sync()
        _lib.check(self.lib.pk_global_solve_custom(self._h, int(topo), PK_DEVICE if dev else PK_HOST, B, xp.ptr(params),
                                                   int(bool(theta_mode)), xp.ptr(y0a), stride, t_arr.ctypes.data_as(C.c_void_p),
                                                   T, 0.0 if rtol is None else float(rtol), 0.0 if atol is None else float(atol),
                                                   int(max_steps), xp.ptr(res["Y"]), xp.ptr(res["status"]), xp.ptr(res["nsteps"]),
                                                   xp.ptr(res["nrej"])))
        return res

    def global_rhs_batch(self, topo, params, Y, t=None, *, theta_mode=False, want_jac=False, tf_inputs=None, S_all=None):
        """f(t, y) (and the analytic Jacobian df/dy) of one uploaded network for B (parameter vector, state, time)
        triples — `pk_global_rhs_batch`.  params [B,P] or [P] (broadcast), Y [B,n], t scalar or [B].  With `tf_inputs`
        [B,N] and `S_all` [B,total_sites] the call has the semantics of the reference's model_ivp closures (the kinase /
        TF matrices are not consulted).  Returns f[B,n] or (f, J[B,n,n])."""
        dev = _is_torch(Y)
        xp = _TorchOps(Y.device) if dev else _NumpyOps()
        Y = xp.f64(Y)
        Y = Y.reshape(1, -1) if Y.ndim == 1 else Y
        B, n = int(Y.shape[0]), int(Y.shape[1])
        dims = self.global_dims(topo)
        if n != dims["state_dim"]:
            raise ValueError(f"Y must have {dims['state_dim']} columns")
        params = xp.f64(params)
        if params.ndim == 1:
            params = params.reshape(1, -1)
        if params.shape[0] == 1 and B > 1:
            params = xp.f64(params.repeat(B, 1) if dev else np.repeat(params, B, axis=0))
        if tuple(params.shape) != (B, dims["n_params"]):
            raise ValueError(f"params must be [{B},{dims['n_params']}]")
        direct = tf_inputs is not None or S_all is not None
        keep = [params, Y]
        tf_p = s_p = t_p = None
        if direct:
            if tf_inputs is None or S_all is None:
                raise ValueError("tf_inputs and S_all go together")
            cnt = dims
            tf_a, s_a = xp.f64(tf_inputs), xp.f64(S_all)
            tf_a = tf_a.reshape(1, -1) if tf_a.ndim == 1 else tf_a
            s_a = s_a.reshape(1, -1) if s_a.ndim == 1 else s_a
            if tuple(tf_a.shape) != (B, cnt["n_proteins"]) or tuple(s_a.shape) != (B, cnt["total_sites"]):
                raise ValueError(f"tf_inputs must be [{B},{cnt['n_proteins']}] and S_all [{B},{cnt['total_sites']}]")
            keep += [tf_a, s_a]
            tf_p, s_p = xp.ptr(tf_a), xp.ptr(s_a)
        else:
            if t is None:
                raise ValueError("t is required")
            t_arr = np.ascontiguousarray(np.broadcast_to(np.asarray(t, dtype=np.float64), (B,)))
            keep.append(t_arr)
            t_p = t_arr.ctypes.data_as(C.c_void_p)
        f = xp.empty((B, n), "f64")
        J = xp.empty((B, n, n), "f64") if want_jac else None
        if dev:
            xp.sync()
        _lib.check(self.lib.pk_global_rhs_batch(self._h, int(topo), PK_DEVICE if dev else PK_HOST, B, xp.ptr(params),
                                                int(bool(theta_mode)), xp.ptr(Y), t_p, xp.ptr(f), xp.ptr(J) if want_jac else None,
                                                tf_p, s_p))
        del keep
        return (f, J) if want_jac else f

    # ------------------------------------------------------------------------- multi-GPU
    def init_nccl(self, world, rank, id_bytes):
        _lib.check(self.lib.pk_nccl_init(self._h, id_bytes, int(world), int(rank)))
        self.world, self.rank = int(world), int(rank)

    def nccl_unique_id(self):
        buf = C.create_string_buffer(128)
        _lib.check(self.lib.pk_nccl_unique_id(buf))
        return buf.raw

    def sym_setup(self, slots_per_rank, exchange):
        """Allocate this rank's symmetric buffer [world, slots_per_rank] (float64), exchange the CUDA IPC handles with
        `exchange(bytes) -> list of bytes, rank-major` (e.g. torch.distributed.all_gather_object) and map every peer's
        buffer.  Returns the torch view of the local buffer."""
        import torch
        buf = C.create_string_buffer(64)
        _lib.check(self.lib.pk_sym_alloc(self._h, int(slots_per_rank), buf))
        handles = exchange(buf.raw)
        if len(handles) != self.world:
            raise PhoskinError("sym_setup: one handle per rank expected")
        if self.world > 1:
            _lib.check(self.lib.pk_sym_open(self._h, b"".join(handles)))
        ptr, slots = C.c_void_p(), C.c_int64()
        _lib.check(self.lib.pk_sym_buffer(self._h, C.byref(ptr), C.byref(slots)))

        class _Raw:                    # zero-copy torch view of library-owned device memory
            __cuda_array_interface__ = {"shape": (self.world, slots.value), "typestr": "<f8", "data": (ptr.value, False),
                                        "version": 2}
        self._sym = torch.as_tensor(_Raw(), device=torch.device("cuda", self.device))
        return self._sym

    def allgather_f64(self, send, recv):
        """NCCL all-gather of torch CUDA float64 tensors (recv.numel() == world * send.numel())."""
        import torch
        torch.cuda.current_stream(send.device).synchronize()
        _lib.check(self.lib.pk_allgather_f64(self._h, send.data_ptr(), send.numel(), recv.data_ptr()))
        return recv


class _NumpyOps:
    def f64(self, x):
        return np.ascontiguousarray(x, dtype=np.float64)

    def i32(self, x):
        return np.ascontiguousarray(x, dtype=np.int32)

    def empty(self, shape, dtype):
        return np.empty(shape, dtype=np.float64 if dtype == "f64" else np.int32)

    def ptr(self, a):
        return None if a is None else a.ctypes.data


class _TorchOps:
    def __init__(self, device):
        import torch
        self.torch = torch
        self.device = device
        if device.type != "cuda":
            raise PhoskinError("torch inputs must live on a CUDA device (no CPU path)")

    def f64(self, x):
        t = self.torch.as_tensor(x, dtype=self.torch.float64, device=self.device)
        return t.contiguous()

    def i32(self, x):
        return self.torch.as_tensor(x, dtype=self.torch.int32, device=self.device).contiguous()

    def empty(self, shape, dtype):
        return self.torch.empty(shape, dtype=self.torch.float64 if dtype == "f64" else self.torch.int32,
                                device=self.device)

    def ptr(self, a):
        return None if a is None else a.data_ptr()

    def sync(self):
        self.torch.cuda.current_stream(self.device).synchronize()


def get_engine(device=None):
    """Process-wide engine per device (created lazily; raises without a GPU)."""
    if device is None:
        import os
        device = int(os.environ.get("LOCAL_RANK", "0")) if "PHOSKIN_DEVICE" not in os.environ \
            else int(os.environ["PHOSKIN_DEVICE"])
    device = int(device)
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]
