"""Host-side mirror of the reference's `global_model/network.py` argument packing.

The reference's `System` (network.py:199-526) owns the parameter arrays, the CSR buffers of the
kinase->site matrix W and the TF->gene matrix, the kinase step-input table and `odeint_args()`
— the 23-tuple that is the wire format into its Numba kernels (network.py:508-526).  Building the
topology from CSV/XLSX files (Index, buildmat, io) is out of scope: a
`GlobalSystem` is constructed from the already-indexed arrays, which is exactly what crosses the
boundary.  The combinatorial model (MODEL 2: one state per phosphorylation pattern, block
`[mRNA, mask_0 .. mask_{2^ns-1}]`, network.py:131-146) has its own 27-tuple (network.py:471-505) with the
per-bucket rate table `S_cache` (jacspeedup.py:114-145) and the hypercube transition lists (models.py:435-485).
"""
from types import SimpleNamespace

import numpy as np

MODEL_IDS = {"distributive": 0, "sequential": 1, "combinatorial": 2, "saturating": 4, "saturation": 4,
             0: 0, 1: 1, 2: 2, 4: 4}
MAX_COMB_SITES = 8          # blocks of up to 16 patterns are inverted in registers, 32..256 patterns by one warp in the L2-resident scratch (csrc/global_net.cuh)
PARAM_KEYS = ("c_k", "A_i", "B_i", "C_i", "D_i", "Dp_i", "E_i")


class GlobalSystem:
    """Parameter arrays + static topology of one coupled kinase-TF-protein network.

    State layout (network.py:28-167, models 0/1/4): for protein i the block
    `[mRNA, P0, site_1..site_ns]` starts at `offset_y[i]`; sites are numbered globally from
    `offset_s[i]`.  Flat physical parameter vector (the batch axis of `simulate_batch`):
    `[c_k (K) | A_i (N) | B_i (N) | C_i (N) | D_i (N) | Dp_i (total_sites) | E_i (N) | tf_scale]`
    — the order of `global_model/params.py:60-101`.
    """

    def __init__(self, *, n_sites, W_indptr, W_indices, W_data, TF_indptr, TF_indices, TF_data,
                 kin_grid, kin_Kmat, tf_deg, driver_map, defaults, y0=None, model=0):
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.model = MODEL_IDS[model]
        n_sites = i32(n_sites)
        N = n_sites.size
        offset_y = np.zeros(N, np.int32)
        offset_s = np.zeros(N, np.int32)
        if self.model == 2:
            if n_sites.size and int(n_sites.max()) > MAX_COMB_SITES:
                raise ValueError(f"combinatorial model: at most {MAX_COMB_SITES} sites per protein")
            n_states = (1 << n_sites).astype(np.int32)          # network.py:131-132
            block = 1 + n_states
        else:
            n_states = None
            block = 2 + n_sites
        offset_y[1:] = np.cumsum(block)[:-1]
        offset_s[1:] = np.cumsum(n_sites)[:-1]
        self.idx = SimpleNamespace(N=N, n_sites=n_sites, offset_y=offset_y, offset_s=offset_s,
                                   state_dim=int(block.sum()), total_sites=int(n_sites.sum()))
        if n_states is not None:
            self.idx.n_states = n_states
        self.W_indptr, self.W_indices, self.W_data = i32(W_indptr), i32(W_indices), f64(W_data)
        self.TF_indptr, self.TF_indices, self.TF_data = i32(TF_indptr), i32(TF_indices), f64(TF_data)
        self.n_W_rows, self.n_TF_rows = self.W_indptr.size - 1, self.TF_indptr.size - 1
        self.kin_grid, self.kin_Kmat = f64(kin_grid), f64(kin_Kmat)
        self.K = self.kin_Kmat.shape[0]
        self.tf_deg, self.driver_map = f64(tf_deg), i32(driver_map)
        if self.n_W_rows != self.idx.total_sites or self.n_TF_rows != N:
            raise ValueError("W must have one row per site and TF one row per protein")
        if self.kin_Kmat.shape[1] != self.kin_grid.size or self.tf_deg.size != N or self.driver_map.size != N:
            raise ValueError("inconsistent kinase table / tf_deg / driver_map sizes")
        self.defaults = {k: f64(defaults[k]).copy() for k in PARAM_KEYS}
        self.defaults["tf_scale"] = float(defaults["tf_scale"])
        for k in PARAM_KEYS:
            setattr(self, k, self.defaults[k].copy())
        self.tf_scale = self.defaults["tf_scale"]
        self.custom_y0 = None if y0 is None else f64(y0).copy()
        if self.model == 2:
            # network.py:278-291: work buffers, per-bucket rate table and the hypercube transition lists
            self.P_vec_work = np.zeros(self.n_TF_rows)
            self.TF_in_work = np.zeros(self.n_TF_rows)
            self.S_cache = np.zeros((self.n_W_rows, self.kin_Kmat.shape[1]))
            (self.trans_from, self.trans_to, self.trans_site, self.trans_off, self.trans_n) = comb_transitions(n_sites)
        self._topo_id = {}          # engine id -> uploaded topology id
        self._loss_key = {}         # engine id -> loss-table dict currently installed

    @classmethod
    def shell(cls, model, n_sites, offset_y=None, offset_s=None):
        """A network with the block layout only (one dummy kinase, no W / TF edges): what the loss-only and the
        direct-mode RHS calls need.  `offset_y` / `offset_s`, when given, must be the packed layout."""
        n_sites = np.ascontiguousarray(n_sites, dtype=np.int32)
        N, S = n_sites.size, int(n_sites.sum())
        z = np.zeros(N)
        sh = cls(n_sites=n_sites, W_indptr=np.zeros(S + 1, np.int32), W_indices=[], W_data=[],
                 TF_indptr=np.zeros(N + 1, np.int32), TF_indices=[], TF_data=[], kin_grid=[0.0], kin_Kmat=np.ones((1, 1)),
                 tf_deg=np.ones(N), driver_map=np.full(N, -1),
                 defaults={"c_k": [1.0], "A_i": z, "B_i": z, "C_i": z, "D_i": z, "Dp_i": np.zeros(S), "E_i": z, "tf_scale": 0.0},
                 model=model)
        if offset_y is not None and not np.array_equal(np.asarray(offset_y), sh.idx.offset_y):
            raise ValueError("offset_y must be the packed per-protein block layout (network.py:28-167)")
        if offset_s is not None and not np.array_equal(np.asarray(offset_s), sh.idx.offset_s):
            raise ValueError("offset_s must be the running sum of n_sites")
        return sh

    # ---- reference surface -----------------------------------------------------------------
    def update(self, c_k, A_i, B_i, C_i, D_i, Dp_i, E_i, tf_scale):
        """network.py:293-302 — in-place parameter write-through."""
        self.c_k[:] = c_k
        self.A_i[:] = A_i
        self.B_i[:] = B_i
        self.C_i[:] = C_i
        self.D_i[:] = D_i
        self.Dp_i[:] = Dp_i
        self.E_i[:] = E_i
        self.tf_scale = float(tf_scale)

    def y0(self):
        """network.py:421-441: custom y0 if set, else mRNA = P0 = 1 and 0.01 per site (model 2: per
        phosphorylated pattern)."""
        if self.custom_y0 is not None:
            return self.custom_y0.copy()
        y = np.zeros(self.idx.state_dim)
        for i in range(self.idx.N):
            st = self.idx.offset_y[i]
            y[st] = y[st + 1] = 1.0
            extra = self.idx.n_states[i] - 1 if self.model == 2 else self.idx.n_sites[i]
            y[st + 2:st + 2 + extra] = 0.01
        return y

    def build_S_cache(self):
        """jacspeedup.py:114-145: S_cache[site, bucket] = sum_k W[site,k] * Kmat[k,bucket] * c_k (host, once per
        parameter set; a few thousand flops of argument packing)."""
        Kc = self.kin_Kmat * self.c_k[:, None]
        for i in range(self.n_W_rows):
            q = slice(self.W_indptr[i], self.W_indptr[i + 1])
            self.S_cache[i, :] = self.W_data[q] @ Kc[self.W_indices[q], :]
        return self.S_cache

    def odeint_args(self, S_cache=None):
        """The reference's 23-tuple (network.py:508-526); model 2: the 27-tuple of network.py:471-505."""
        if self.model == 2:
            if S_cache is None:
                raise ValueError("MODEL==2 requires S_cache (total_sites x n_timebins).")
            return (self.c_k, self.A_i, self.B_i, self.C_i, self.D_i, self.Dp_i, self.E_i, float(self.tf_scale),
                    self.kin_grid, S_cache, self.TF_indptr, self.TF_indices, self.TF_data, int(self.n_TF_rows),
                    self.idx.offset_y, self.idx.offset_s, self.idx.n_sites, self.idx.n_states,
                    self.trans_from, self.trans_to, self.trans_site, self.trans_off, self.trans_n,
                    self.tf_deg, self.driver_map, self.P_vec_work, self.TF_in_work)
        return (self.c_k, self.A_i, self.B_i, self.C_i, self.D_i, self.Dp_i, self.E_i, float(self.tf_scale),
                self.kin_grid, self.kin_Kmat, self.W_indptr, self.W_indices, self.W_data, int(self.n_W_rows),
                self.TF_indptr, self.TF_indices, self.TF_data, int(self.n_TF_rows),
                self.idx.offset_y, self.idx.offset_s, self.idx.n_sites, self.tf_deg, self.driver_map)

    # ---- batch surface ---------------------------------------------------------------------
    @property
    def n_params(self):
        return self.K + 5 * self.idx.N + self.idx.total_sites + 1

    def param_slices(self):
        sizes = [self.K, self.idx.N, self.idx.N, self.idx.N, self.idx.N, self.idx.total_sites, self.idx.N, 1]
        out, o = {}, 0
        for k, s in zip(PARAM_KEYS + ("tf_scale",), sizes):
            out[k] = slice(o, o + s)
            o += s
        return out

    def pack_params(self, p=None):
        p = p or {**{k: getattr(self, k) for k in PARAM_KEYS}, "tf_scale": self.tf_scale}
        return np.concatenate([np.asarray(p[k], np.float64).ravel() for k in PARAM_KEYS] + [[float(p["tf_scale"])]])

    def unpack_params(self, vec):
        sl = self.param_slices()
        out = {k: np.array(vec[sl[k]], dtype=np.float64) for k in PARAM_KEYS}
        out["tf_scale"] = float(vec[sl["tf_scale"]][0])
        return out

    def as_dict(self):
        """Plain dict of the arrays (the form the test oracle consumes)."""
        return {"N": self.idx.N, "K": self.K, "total_sites": self.idx.total_sites, "state_dim": self.idx.state_dim,
                "n_sites": self.idx.n_sites, "offset_y": self.idx.offset_y, "offset_s": self.idx.offset_s,
                "W_indptr": self.W_indptr, "W_indices": self.W_indices, "W_data": self.W_data,
                "n_W_rows": self.n_W_rows, "TF_indptr": self.TF_indptr, "TF_indices": self.TF_indices,
                "TF_data": self.TF_data, "kin_grid": self.kin_grid, "kin_Kmat": self.kin_Kmat,
                "tf_deg": self.tf_deg, "driver_map": self.driver_map, "y0": self.y0(), "model": self.model,
                "defaults": {**{k: self.defaults[k].copy() for k in PARAM_KEYS}, "tf_scale": self.defaults["tf_scale"]}}


def comb_transitions(n_sites):
    """Forward (phosphorylation) edges of every protein's pattern hypercube, flattened — the arrays
    models.py:435-485 hands to the Numba kernel: edge m -> m | (1 << j) for every unset bit j, patterns in
    ascending order, sites in ascending order inside a pattern."""
    frm, to, site = [], [], []
    off = np.zeros(len(n_sites), np.int32)
    cnt = np.zeros(len(n_sites), np.int32)
    for i, ns in enumerate(np.asarray(n_sites, int)):
        off[i] = len(frm)
        if ns > 0:
            m = np.repeat(np.arange(1 << ns), ns)
            j = np.tile(np.arange(ns), 1 << ns)
            free = (m >> j) & 1 == 0
            frm.extend(m[free].tolist())
            to.extend((m[free] | (1 << j[free])).tolist())
            site.extend(j[free].tolist())
        cnt[i] = len(frm) - off[i]
    i32 = lambda a: np.asarray(a, dtype=np.int32)
    return i32(frm), i32(to), i32(site), off, cnt


def synthetic_system(seed=0, N=120, K=40, max_sites=4, w_density=0.08, tf_density=0.03, n_driven=None,
                     model=0, time_points=None):
    """Seeded synthetic network of the BASELINE config-5 shape: N proteins with
    1..max_sites sites each (state_dim ~ N*(2+avg sites)), K kinases, sparse W (sites x kinases) and TF
    (N x N) matrices, Kmat = exp(0.3*N(0,1)) on the 14-point protein grid, a few proteins driven by
    kinase profiles, rate constants inside the reference's bounds (config.toml:382-410)."""
    rng = np.random.default_rng(seed)
    grid = np.asarray(time_points if time_points is not None else
                      [0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
    n_sites = rng.integers(1, max_sites + 1, N).astype(np.int32)
    if N > 3:
        n_sites[rng.integers(0, N)] = 0                      # a protein without sites (orphan TF shape)
    S = int(n_sites.sum())

    def sparse_rows(rows, cols, density, at_least_one, gen):
        indptr, indices, data = [0], [], []
        for _ in range(rows):
            k = rng.binomial(cols, density)
            if at_least_one:
                k = max(1, k)
            idx = np.sort(rng.choice(cols, size=min(k, cols), replace=False))
            indices.extend(idx.tolist())
            data.extend(gen(idx.size).tolist())
            indptr.append(len(indices))
        return np.asarray(indptr, np.int32), np.asarray(indices, np.int32), np.asarray(data, np.float64)

    W = sparse_rows(S, K, w_density, True, lambda m: rng.uniform(0.1, 1.0, m))
    TF = sparse_rows(N, N, tf_density, False, lambda m: rng.uniform(0.2, 1.0, m) * rng.choice([-1.0, 1.0], m))
    tf_deg = np.ones(N)
    for i in range(N):
        a = np.abs(TF[2][TF[0][i]:TF[0][i + 1]]).sum()
        tf_deg[i] = a if a > 0 else 1.0
    Kmat = np.maximum(np.exp(0.3 * rng.standard_normal((K, grid.size))), 1e-6)
    Kmat[:, 0] = 1.0
    driver_map = np.full(N, -1, np.int32)
    n_driven = min(K, N // 6) if n_driven is None else n_driven
    for k, p in enumerate(rng.choice(N, size=n_driven, replace=False)):
        driver_map[p] = k
    defaults = {"c_k": rng.uniform(0.5, 2.0, K), "A_i": rng.uniform(0.05, 1.0, N), "B_i": rng.uniform(0.02, 0.5, N),
                "C_i": rng.uniform(0.05, 1.0, N), "D_i": rng.uniform(0.1, 0.5, N), "Dp_i": rng.uniform(0.05, 2.0, S),
                "E_i": rng.uniform(0.05, 2.0, N), "tf_scale": 3.0}
    return GlobalSystem(n_sites=n_sites, W_indptr=W[0], W_indices=W[1], W_data=W[2], TF_indptr=TF[0],
                        TF_indices=TF[1], TF_data=TF[2], kin_grid=grid, kin_Kmat=Kmat, tf_deg=tf_deg,
                        driver_map=driver_map, defaults=defaults, model=model)


def synthetic_loss_data(sys_, time_grid, seed=0, frac=0.5):
    """Index/observation tables of `cache.prepare_fast_loss_data` (cache.py:19-155) for synthetic
    fold-change data: a random subset of (protein, time), (protein, time>=4) and (protein, site, time)."""
    rng = np.random.default_rng(seed)
    tg = np.asarray(time_grid, float)
    idx = sys_.idx
    # cache.py:138-145: second column = n_states for the combinatorial model, n_sites otherwise
    prot_map = np.stack([idx.offset_y, idx.n_states if sys_.model == 2 else idx.n_sites], axis=1).astype(np.int32)
    base = {"prot_base_idx": int(np.argmin(np.abs(tg - 0.0))), "rna_base_idx": int(np.argmin(np.abs(tg - 4.0))),
            "pho_base_idx": int(np.argmin(np.abs(tg - 0.0)))}
    t_all = np.arange(tg.size)
    t_rna = t_all[tg >= 4.0]

    def pick(pairs):
        pairs = np.asarray(pairs, np.int32)
        keep = rng.random(len(pairs)) < frac
        keep[0] = True
        return pairs[keep]

    pp = pick([(p, t) for p in range(idx.N) for t in t_all])
    pr = pick([(p, t) for p in range(idx.N) for t in t_rna])
    ph = pick([(p, s, t) for p in range(idx.N) for s in range(idx.n_sites[p]) for t in t_all])
    mk = lambda m: (np.exp(0.4 * rng.standard_normal(m)), rng.uniform(0.5, 1.5, m))
    op, wp = mk(len(pp))
    orr, wr = mk(len(pr))
    oph, wph = mk(len(ph))
    return {"p_prot": pp[:, 0].copy(), "t_prot": pp[:, 1].copy(), "obs_prot": op, "w_prot": wp,
            "p_rna": pr[:, 0].copy(), "t_rna": pr[:, 1].copy(), "obs_rna": orr, "w_rna": wr,
            "p_pho": ph[:, 0].copy(), "s_pho": ph[:, 1].copy(), "t_pho": ph[:, 2].copy(), "obs_pho": oph,
            "w_pho": wph, "prot_map": prot_map, **base}
