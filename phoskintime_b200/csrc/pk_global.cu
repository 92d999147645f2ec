// Host side of the global-network entry points of include/phoskin_b200.h: topology upload (the
// static part of System.odeint_args(), global_model/network.py:508-526), loss tables
// (global_model/lossfn.py:113-121), prior centre (optproblem.py:105-114) and the batched solve
// that replaces a loop of simulate_odeint (global_model/simulate.py:34-80).
#include <algorithm>
#include <cmath>
#include <utility>

#include "pk_internal.hpp"

#include "global_net.cuh"

namespace pkh {

struct GlobalTopoHost {
    pk::GlobalTopoDev dev;
    pk::GlobalSmem sm;
    int P = 0;
    size_t smem_bytes = 0;
    int ctas_per_sm = 1;
    std::vector<void*> allocs;        // topology arrays
    std::vector<void*> loss_allocs;   // loss tables (replaced by pk_global_set_loss_data)
    void* prior_alloc = nullptr;
    std::vector<double> kin_grid;
    bool has_loss = false;
    int T_loss_max = -1;              // largest time index the loss tables reference
    size_t binv_elems = 0;            // model 2: block-inverse scratch per resident system (doubles)

    static void free_list(std::vector<void*>& v) {
        for (void* p : v) cudaFree(p);
        v.clear();
    }
    ~GlobalTopoHost() {
        free_list(allocs);
        free_list(loss_allocs);
        if (prior_alloc) cudaFree(prior_alloc);
    }
};

void release_global_topologies(pk_handle_s* h) {
    for (GlobalTopoHost* t : h->topos) delete t;
    h->topos.clear();
}

namespace {

template <class T>
cudaError_t upload(std::vector<void*>& owner, const T* src, size_t count, const T** dst) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    owner.push_back(p);
    if (count) {
        e = cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return e;
    }
    *dst = (const T*)p;
    return cudaSuccess;
}

GlobalTopoHost* topo_of(pk_handle_s* h, int id) {
    if (!h || id < 0 || id >= (int)h->topos.size()) return nullptr;
    return h->topos[id];
}

int bucket_of(double t, const std::vector<double>& g) {       // jacspeedup.py:148-172 / utils.py:210-225
    if (t <= g.front()) return 0;
    if (t >= g.back()) return (int)g.size() - 1;
    int j = (int)(std::upper_bound(g.begin(), g.end(), t) - g.begin()) - 1;
    return std::min(std::max(j, 0), (int)g.size() - 1);
}

// Shared-memory layout of one system.  `spill` = how many of the large arrays (in the order Sc, U, clo, mult, facA, w,
// arg, y, par) live in the per-CTA global scratch instead (capacity fallback, generic Schur path only): their offsets
// are OVF_BASE + offset into the scratch slice, L.ovf_total doubles per CTA.
void layout_smem(GlobalTopoHost* th, int nnzT, int force_generic, int spill) {
    const pk::GlobalTopoDev& d = th->dev;
    pk::GlobalSmem& L = th->sm;
    const int n_big = L.n_big, big_nst = L.big_nst, binv_total = L.binv_total, c16 = L.cls16, c8 = L.cls8, c4 = L.cls4;
    memset(&L, 0, sizeof(L));
    L.n_big = n_big; L.big_nst = big_nst; L.binv_total = binv_total; L.cls16 = c16; L.cls8 = c8; L.cls4 = c4;
    int o = 0, og = 0, rank = 0;
    auto take = [&](int count) {
        const int at = o;
        o += (count + 1) & ~1;
        return at;
    };
    auto take_big = [&](int count, int order) {        // order = position in the spill list
        (void)rank;
        if (order < spill) {
            const int at = pk::OVF_BASE + og;
            og += (count + 1) & ~1;
            return at;
        }
        return take(count);
    };
    const int n = d.n, N = d.N, nQ = d.nQ;
    // register-resident Gauss-Jordan path for up to 128 regulators, shared-memory LU beyond
    L.tile = (force_generic || spill > 0 || nQ > 16 * pk::GJ_MAX_TILE) ? 0 : (nQ <= 32 ? 2 : nQ <= 64 ? 4 : nQ <= 96 ? 6 : 8);
    L.par = take_big(th->P, 8);
    L.Kt = take(d.K);
    L.Sall = take(d.S);
    L.y = take_big(n, 7);
    L.arg = take_big(n, 6);
    L.U = take_big(6 * n, 1);
    L.w = take_big(n, 5);
    L.facA = take_big(n, 4);
    L.mult = take_big(n, 3);
    L.clo = take_big(n, 2);
    L.pvec = take(N);
    L.g = take(N);
    L.m = take(N);
    L.z = take(N);
    L.tfdata = take(nnzT);
    L.tfdeg = take(N);
    L.cscr = take(d.model == 2 ? 16 * pk::COMB_MAX_STATES : 0);        // even offsets: 16-byte aligned strips
    L.z2 = take(N);
    L.fz = take(N);
    L.rabs = take(N);
    if (L.tile == 0) {
        L.ld = nQ | 1;                               // odd leading dimension: conflict-free column walks
        L.Sc = take_big(nQ * L.ld, 0);
        L.idiag = take(nQ);
        L.red = take(2 * pk::GLOBAL_WARPS + nQ);
        L.perm = take((nQ + 1) / 2 + 1);
    } else {
        const int GP = 16 * L.tile;
        L.colbuf = take(2 * GP);
        L.rowbuf = take(2 * GP);
        L.bp = take(GP);
        L.partial = take(16 * (GP + 1));
        L.red = take(2 * pk::GLOBAL_WARPS);
    }
    L.ints = o;
    int io = 0;
    auto itake = [&](int count) {
        const int at = io;
        io += count;
        return at;
    };
    L.i_offy = itake(N);
    L.i_offs = itake(N);
    L.i_ns = itake(N);
    L.i_drv = itake(N);
    L.i_qpos = itake(N);
    L.i_tfptr = itake(N + 1);
    L.i_tfidx = itake(nnzT);
    L.i_qlist = itake(nQ);
    L.i_piv = itake(16 * pk::GJ_MAX_TILE);
    L.i_pinv = itake(16 * pk::GJ_MAX_TILE);
    L.i_sprot = itake(n);
    L.i_ent = itake(L.tile * L.tile * pk::GLOBAL_BLOCK / 4);    // one byte per tile slot and thread
    L.i_boff = itake(N);
    L.i_cord = itake(N);
    o += (io + 1) / 2;
    L.total = o;
    L.ovf_total = og;
    th->smem_bytes = (size_t)o * sizeof(double);
}
constexpr int N_SPILLABLE = 9;

typedef void (*global_kernel_t)(const pk::GlobalArgs);
global_kernel_t kernel_for_tile(int tile, bool comb, bool ovf) {
    if (ovf) return comb ? pk::global_net_kernel<0, true, true> : pk::global_net_kernel<0, false, true>;
    if (comb) switch (tile) {
        case 2: return pk::global_net_kernel<2, true>;
        case 4: return pk::global_net_kernel<4, true>;
        case 6: return pk::global_net_kernel<6, true>;
        case 8: return pk::global_net_kernel<8, true>;
        default: return pk::global_net_kernel<0, true>;
    }
    switch (tile) {
        case 2: return pk::global_net_kernel<2, false>;
        case 4: return pk::global_net_kernel<4, false>;
        case 6: return pk::global_net_kernel<6, false>;
        case 8: return pk::global_net_kernel<8, false>;
        default: return pk::global_net_kernel<0, false>;
    }
}

}  // namespace
}  // namespace pkh

using pkh::fail;
using pkh::GlobalTopoHost;

extern "C" {

int pk_global_upload(pk_handle_t h, const pk_global_topology* tp, int32_t* topo_id) {
    if (!h || !tp || !topo_id) return fail("pk_global_upload: null argument");
    if (tp->model != 0 && tp->model != 1 && tp->model != 2 && tp->model != 4)
        return fail("pk_global_upload: kinetic model must be 0 (distributive), 1 (sequential), 2 (combinatorial) or "
                    "4 (saturating)");
    const bool comb = tp->model == 2;
    const int N = tp->N, K = tp->K, nb = tp->n_bins;
    if (N < 1 || K < 1 || nb < 1) return fail("pk_global_upload: N, K and n_bins must be positive");
    if (!tp->n_sites || !tp->W_indptr || !tp->TF_indptr || !tp->kin_grid || !tp->kin_Kmat || !tp->tf_deg ||
        !tp->driver_map)
        return fail("pk_global_upload: missing array");
    std::vector<int> off_y(N), off_s(N);
    int n = 0, S = 0;
    size_t binv_elems = 0;            // model 2: doubles of block-inverse scratch per resident system
    int n_big = 0, big_nst = 0;       // model 2: blocks of more than 16 patterns (inverted by one warp in the scratch)
    for (int i = 0; i < N; ++i) {
        if (tp->n_sites[i] < 0) return fail("pk_global_upload: negative n_sites");
        if (comb && tp->n_sites[i] > pk::COMB_MAX_SITES)
            return fail("pk_global_upload: the combinatorial model supports at most " + std::to_string(pk::COMB_MAX_SITES) +
                        " sites per protein (2^sites pattern states each)");
        if (comb && tp->n_sites[i] > 4) {
            ++n_big;
            big_nst = std::max(big_nst, 1 << tp->n_sites[i]);
        }
        off_y[i] = n;
        off_s[i] = S;
        // network.py:144-149: combinatorial block = mRNA + 2^ns pattern states, otherwise mRNA + P0 + ns sites
        n += comb ? 1 + (1 << tp->n_sites[i]) : 2 + tp->n_sites[i];
        S += tp->n_sites[i];
        binv_elems += comb ? (size_t)1 << (2 * tp->n_sites[i]) : 0;
    }
    if (tp->W_indptr[0] != 0 || tp->TF_indptr[0] != 0) return fail("pk_global_upload: indptr must start at 0");
    for (int s = 0; s < S; ++s)
        if (tp->W_indptr[s + 1] < tp->W_indptr[s]) return fail("pk_global_upload: W_indptr not monotone");
    for (int i = 0; i < N; ++i)
        if (tp->TF_indptr[i + 1] < tp->TF_indptr[i]) return fail("pk_global_upload: TF_indptr not monotone");
    const int nnzW = tp->W_indptr[S], nnzT = tp->TF_indptr[N];
    if ((nnzW && (!tp->W_indices || !tp->W_data)) || (nnzT && (!tp->TF_indices || !tp->TF_data)))
        return fail("pk_global_upload: missing CSR indices/data");
    for (int q = 0; q < nnzW; ++q)
        if (tp->W_indices[q] < 0 || tp->W_indices[q] >= K) return fail("pk_global_upload: W column index out of range");
    for (int q = 0; q < nnzT; ++q)
        if (tp->TF_indices[q] < 0 || tp->TF_indices[q] >= N) return fail("pk_global_upload: TF column index out of range");
    for (int i = 0; i < N; ++i) {
        if (tp->driver_map[i] < -1 || tp->driver_map[i] >= K) return fail("pk_global_upload: driver_map out of range");
        if (!(tp->tf_deg[i] != 0.0)) return fail("pk_global_upload: tf_deg must be non-zero");
    }
    for (int b = 1; b < nb; ++b)
        if (!(tp->kin_grid[b] > tp->kin_grid[b - 1])) return fail("pk_global_upload: kin_grid must be strictly increasing");
    CK(cudaSetDevice(h->device));
    // The combinatorial wrapper never consults the driver map (jacspeedup.py:318-325)
    std::vector<int> drv(tp->driver_map, tp->driver_map + N);
    if (comb) std::fill(drv.begin(), drv.end(), -1);

    // canonical TF rows for the device: column indices sorted and unique (duplicates summed) — the kernel's static
    // sparsity table of the Schur block needs at most one entry per (gene, regulator)
    std::vector<int> tf_ptr(N + 1, 0), tf_idx;
    std::vector<double> tf_val;
    for (int i = 0; i < N; ++i) {
        std::vector<std::pair<int, double>> row;
        for (int q = tp->TF_indptr[i]; q < tp->TF_indptr[i + 1]; ++q) row.emplace_back(tp->TF_indices[q], tp->TF_data[q]);
        std::stable_sort(row.begin(), row.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
        for (size_t e = 0; e < row.size(); ++e) {
            if (!tf_idx.empty() && (int)tf_idx.size() > tf_ptr[i] && tf_idx.back() == row[e].first) tf_val.back() += row[e].second;
            else { tf_idx.push_back(row[e].first); tf_val.push_back(row[e].second); }
        }
        tf_ptr[i + 1] = (int)tf_idx.size();
        if (tf_ptr[i + 1] - tf_ptr[i] > 254) return fail("pk_global_upload: a gene has more than 254 distinct regulators");
    }
    const int nnzC = (int)tf_idx.size();

    // regulator set Q: non-driven proteins whose total protein enters some gene's TF input
    std::vector<int> qpos(N, -1), qlist;
    {
        std::vector<char> used(N, 0);
        for (int q = 0; q < nnzC; ++q)
            if (tf_val[q] != 0.0) used[tf_idx[q]] = 1;
        for (int i = 0; i < N; ++i)
            if (used[i] && drv[i] < 0) {
                qpos[i] = (int)qlist.size();
                qlist.push_back(i);
            }
    }

    GlobalTopoHost* th = new GlobalTopoHost();
    memset(&th->dev, 0, sizeof(th->dev));
    pk::GlobalTopoDev& d = th->dev;
    d.model = tp->model; d.N = N; d.K = K; d.nb = nb; d.n = n; d.S = S; d.nQ = (int)qlist.size();
    th->P = K + 5 * N + S + 1;
    if (binv_elems + (size_t)pk::GLOBAL_WARPS * big_nst > (size_t)1 << 30) {
        delete th;
        return fail("pk_global_upload: combinatorial pattern blocks need more than 8 GiB of scratch per resident system");
    }
    th->sm.n_big = n_big;
    th->sm.big_nst = big_nst;
    {   // size classes of the register path in the kernel's order (larger blocks first): 16, 8, 4 patterns, then 2 / 1
        int cnt[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < N; ++i)
            if (comb && tp->n_sites[i] <= 4) ++cnt[tp->n_sites[i]];
        th->sm.cls16 = n_big + cnt[4];
        th->sm.cls8 = th->sm.cls16 + cnt[3];
        th->sm.cls4 = th->sm.cls8 + cnt[2];
    }
    th->sm.binv_total = (int)binv_elems;
    th->binv_elems = binv_elems + (size_t)pk::GLOBAL_WARPS * big_nst;     // + one pivot-row snapshot strip per warp
    th->kin_grid.assign(tp->kin_grid, tp->kin_grid + nb);
    cudaError_t e = cudaSuccess;
#define UP(field, src, count)                                                     \
    if (e == cudaSuccess) e = pkh::upload(th->allocs, src, (size_t)(count), &d.field)
    UP(offset_y, off_y.data(), N);
    UP(offset_s, off_s.data(), N);
    UP(n_sites, tp->n_sites, N);
    UP(W_indptr, tp->W_indptr, S + 1);
    UP(W_indices, tp->W_indices, nnzW);
    UP(W_data, tp->W_data, nnzW);
    UP(TF_indptr, tf_ptr.data(), N + 1);
    UP(TF_indices, tf_idx.data(), nnzC);
    UP(TF_data, tf_val.data(), nnzC);
    UP(kin_grid, tp->kin_grid, nb);
    UP(kin_Kmat, tp->kin_Kmat, (size_t)K * nb);
    UP(tf_deg, tp->tf_deg, N);
    UP(driver_map, drv.data(), N);
    UP(qlist, qlist.data(), qlist.size());
    UP(qpos, qpos.data(), N);
#undef UP
    if (e != cudaSuccess) {
        delete th;
        return fail(std::string("pk_global_upload: ") + cudaGetErrorString(e));
    }
    // Capacity: when one system does not fit into a CTA's shared memory, the large arrays move one by one (Schur matrix
    // first, then the stage vectors, the block factors, the state) into a per-CTA slice of a global scratch buffer, which
    // the 126 MB L2 keeps resident; only the small per-protein arrays and the staged topology must stay on chip.
    // force_generic_schur: 1 = shared-memory LU although the register path fits, 2 = additionally every spillable array
    // in the scratch (tests compare that against the all-shared layout: same arithmetic, bit-identical results).
    int max_optin = 0;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
    const size_t cap = (size_t)max_optin - 256;       // static shared memory of the kernel (queue slot, step state)
    int spill = tp->force_generic_schur >= 2 ? pkh::N_SPILLABLE : 0;
    pkh::layout_smem(th, nnzC, tp->force_generic_schur, spill);
    while (th->smem_bytes > cap && spill < pkh::N_SPILLABLE) pkh::layout_smem(th, nnzC, tp->force_generic_schur, ++spill);
    if (th->smem_bytes > cap) {
        const size_t need = th->smem_bytes;
        delete th;
        return fail("pk_global_upload: the per-protein arrays and the staged topology alone need " + std::to_string(need) +
                    " B of shared memory per system, device offers " + std::to_string(max_optin));
    }
    h->topos.push_back(th);
    *topo_id = (int32_t)h->topos.size() - 1;
    return 0;
}

int pk_global_set_loss_data(pk_handle_t h, int32_t topo_id, const pk_global_loss_data* ld) {
    GlobalTopoHost* th = pkh::topo_of(h, topo_id);
    if (!th || !ld) return fail("pk_global_set_loss_data: bad handle, topology id or table");
    if (ld->n_prot < 0 || ld->n_rna < 0 || ld->n_pho < 0) return fail("pk_global_set_loss_data: negative count");
    CK(cudaSetDevice(h->device));
    pk::GlobalTopoDev& d = th->dev;
    // validate indices on the host: the kernel trusts them
    std::vector<int> ns(d.N);
    CK(cudaMemcpy(ns.data(), d.n_sites, d.N * sizeof(int), cudaMemcpyDeviceToHost));
    int tmax = std::max(std::max(ld->prot_base_idx, ld->rna_base_idx), ld->pho_base_idx);
    if (ld->prot_base_idx < 0 || ld->rna_base_idx < 0 || ld->pho_base_idx < 0)
        return fail("pk_global_set_loss_data: negative base index");
    for (int k = 0; k < ld->n_prot; ++k) {
        if (ld->p_prot[k] < 0 || ld->p_prot[k] >= d.N || ld->t_prot[k] < 0) return fail("pk_global_set_loss_data: protein row out of range");
        tmax = std::max(tmax, ld->t_prot[k]);
    }
    for (int k = 0; k < ld->n_rna; ++k) {
        if (ld->p_rna[k] < 0 || ld->p_rna[k] >= d.N || ld->t_rna[k] < 0) return fail("pk_global_set_loss_data: rna row out of range");
        tmax = std::max(tmax, ld->t_rna[k]);
    }
    for (int k = 0; k < ld->n_pho; ++k) {
        if (ld->p_pho[k] < 0 || ld->p_pho[k] >= d.N || ld->t_pho[k] < 0 || ld->s_pho[k] < 0 ||
            ld->s_pho[k] >= ns[ld->p_pho[k]])
            return fail("pk_global_set_loss_data: phospho row out of range");
        tmax = std::max(tmax, ld->t_pho[k]);
    }
    GlobalTopoHost::free_list(th->loss_allocs);
    th->has_loss = false;
    cudaError_t e = cudaSuccess;
#define UP(field, count)                                                          \
    if (e == cudaSuccess) e = pkh::upload(th->loss_allocs, ld->field, (size_t)(count), &d.field)
    UP(p_prot, ld->n_prot); UP(t_prot, ld->n_prot); UP(obs_prot, ld->n_prot); UP(w_prot, ld->n_prot);
    UP(p_rna, ld->n_rna); UP(t_rna, ld->n_rna); UP(obs_rna, ld->n_rna); UP(w_rna, ld->n_rna);
    UP(p_pho, ld->n_pho); UP(s_pho, ld->n_pho); UP(t_pho, ld->n_pho); UP(obs_pho, ld->n_pho); UP(w_pho, ld->n_pho);
#undef UP
    if (e != cudaSuccess) return fail(std::string("pk_global_set_loss_data: ") + cudaGetErrorString(e));
    d.n_prot = ld->n_prot; d.n_rna = ld->n_rna; d.n_pho = ld->n_pho;
    d.prot_base = ld->prot_base_idx; d.rna_base = ld->rna_base_idx; d.pho_base = ld->pho_base_idx;
    // optproblem.py:83-85: each modality's weighted sum is divided by max(1e-6, sum of its weights)
    const double* ws[3] = {ld->w_prot, ld->w_rna, ld->w_pho};
    const int cnt[3] = {ld->n_prot, ld->n_rna, ld->n_pho};
    for (int m = 0; m < 3; ++m) {
        double s = 0.0;
        for (int k = 0; k < cnt[m]; ++k) s += ws[m][k];
        d.norm[m] = 1.0 / std::max(1e-6, s);
    }
    th->T_loss_max = tmax;
    th->has_loss = true;
    return 0;
}

int pk_global_set_prior(pk_handle_t h, int32_t topo_id, const double* defaults) {
    GlobalTopoHost* th = pkh::topo_of(h, topo_id);
    if (!th) return fail("pk_global_set_prior: bad handle or topology id");
    CK(cudaSetDevice(h->device));
    if (th->prior_alloc) { cudaFree(th->prior_alloc); th->prior_alloc = nullptr; }
    th->dev.defaults = nullptr;
    if (!defaults) return 0;
    CK(cudaMalloc(&th->prior_alloc, th->P * sizeof(double)));
    CK(cudaMemcpy(th->prior_alloc, defaults, th->P * sizeof(double), cudaMemcpyHostToDevice));
    th->dev.defaults = (const double*)th->prior_alloc;
    return 0;
}

int pk_global_release(pk_handle_t h, int32_t topo_id) {
    GlobalTopoHost* th = pkh::topo_of(h, topo_id);
    if (!th) return fail("pk_global_release: bad handle or topology id");
    cudaSetDevice(h->device);
    delete th;
    h->topos[topo_id] = nullptr;      // ids of the other topologies stay valid
    return 0;
}

int pk_global_dims(pk_handle_t h, int32_t topo_id, int32_t* state_dim, int32_t* n_params, int32_t* n_reg,
                   int32_t* smem_bytes) {
    GlobalTopoHost* th = pkh::topo_of(h, topo_id);
    if (!th) return fail("pk_global_dims: bad handle or topology id");
    if (state_dim) *state_dim = th->dev.n;
    if (n_params) *n_params = th->P;
    if (n_reg) *n_reg = th->dev.nQ;
    if (smem_bytes) *smem_bytes = (int32_t)th->smem_bytes;
    return 0;
}

int pk_global_counts(pk_handle_t h, int32_t topo_id, int32_t* n_proteins, int32_t* n_kinases, int32_t* total_sites) {
    GlobalTopoHost* th = pkh::topo_of(h, topo_id);
    if (!th) return fail("pk_global_counts: bad handle or topology id");
    if (n_proteins) *n_proteins = th->dev.N;
    if (n_kinases) *n_kinases = th->dev.K;
    if (total_sites) *total_sites = th->dev.S;
    return 0;
}

void pk_global_job_init(pk_global_job* job) {
    memset(job, 0, sizeof(*job));
    job->metric = PK_GM_NONE;
    for (int i = 0; i < 3; ++i) job->lambdas[i] = 1.0;
    job->lambda_prior = 0.0;
}

int pk_sizeof_global_job(void) { return (int)sizeof(pk_global_job); }

int pk_global_solve_batch(pk_handle_t h, const pk_global_job* j) {
    if (!h || !j) return fail("null handle or job");
    GlobalTopoHost* th = pkh::topo_of(h, j->topo);
    if (!th) return fail("pk_global_solve_batch: unknown topology id");
    if (j->B < 0) return fail("B < 0");
    if (j->T < 1) return fail("T < 1");
    if (!j->params || !j->y0 || !j->t_eval) return fail("params, y0 and t_eval are required");
    for (int k = 1; k < j->T; ++k)
        if (!(j->t_eval[k] > j->t_eval[k - 1])) return fail("t_eval must be strictly increasing");
    const bool want_loss = j->out_loss || j->out_F;
    if (want_loss && !th->has_loss) return fail("out_loss/out_F need pk_global_set_loss_data");
    if (want_loss && th->T_loss_max >= j->T) return fail("loss tables reference a time index >= T");
    if (want_loss && (j->loss_mode < -1 || j->loss_mode > 7)) return fail("loss_mode out of range");
    if (j->out_metric || j->out_fc) {
        if (j->out_metric && (j->metric < 0 || j->metric > 3)) return fail("out_metric needs a valid metric id");
        if (j->n_mt_prot < 0 || j->n_mt_rna < 0 || j->n_mt_pho < 0) return fail("negative metric time count");
        const int32_t* lists[3] = {j->mt_prot, j->mt_rna, j->mt_pho};
        const int cnt[3] = {j->n_mt_prot, j->n_mt_rna, j->n_mt_pho};
        const int base[3] = {j->mb_prot, j->mb_rna, j->mb_pho};
        for (int m = 0; m < 3; ++m) {
            if (cnt[m] && !lists[m]) return fail("metric time list missing");
            if (cnt[m] && (base[m] < 0 || base[m] >= j->T)) return fail("metric base index out of range");
            for (int k = 0; k < cnt[m]; ++k)
                if (lists[m][k] < 0 || lists[m][k] >= j->T) return fail("metric time index out of range");
        }
    }
    h->last_launches = 0;
    h->last_ms = 0.f;
    if (j->B == 0) return 0;
    CK(cudaSetDevice(h->device));
    const pk::GlobalTopoDev& d = th->dev;
    const size_t B = (size_t)j->B;
    const int n = d.n, T = j->T, P = th->P;
    const bool host = j->memspace == PK_HOST;
    cudaStream_t st = h->stream;

    // stop list: outputs plus the interior kinase-grid points (the RHS jumps there)
    std::vector<double> stops(j->t_eval, j->t_eval + T);
    for (double g : th->kin_grid)
        if (g > j->t_eval[0] && g < j->t_eval[T - 1]) stops.push_back(g);
    std::sort(stops.begin(), stops.end());
    stops.erase(std::unique(stops.begin(), stops.end()), stops.end());
    const int ns = (int)stops.size();
    std::vector<int> s_out(ns, -1), s_bucket(ns, 0);
    for (int k = 0, o = 0; k < ns; ++k) {
        if (o < T && stops[k] == j->t_eval[o]) s_out[k] = o++;
        s_bucket[k] = pkh::bucket_of(k + 1 < ns ? 0.5 * (stops[k] + stops[k + 1]) : stops[k], th->kin_grid);
    }
    const bool want_fc_tab = j->out_metric || j->out_fc;
    const int n_mt = want_fc_tab ? j->n_mt_prot + j->n_mt_rna + j->n_mt_pho : 0;
    const size_t nfc = want_fc_tab ? (size_t)d.N * (j->n_mt_prot + j->n_mt_rna) + (size_t)d.S * j->n_mt_pho : 0;
    const size_t stop_bytes = (size_t)ns * sizeof(double) + (size_t)(2 * ns + n_mt) * sizeof(int);
    CK(h->g_stops.ensure(stop_bytes));
    {
        std::vector<char> blob(stop_bytes);
        memcpy(blob.data(), stops.data(), ns * sizeof(double));
        int* ip = (int*)(blob.data() + (size_t)ns * sizeof(double));
        memcpy(ip, s_out.data(), ns * sizeof(int));
        memcpy(ip + ns, s_bucket.data(), ns * sizeof(int));
        if (n_mt) {
            int* mp = ip + 2 * ns;
            if (j->n_mt_prot) memcpy(mp, j->mt_prot, j->n_mt_prot * sizeof(int));
            if (j->n_mt_rna) memcpy(mp + j->n_mt_prot, j->mt_rna, j->n_mt_rna * sizeof(int));
            if (j->n_mt_pho) memcpy(mp + j->n_mt_prot + j->n_mt_rna, j->mt_pho, j->n_mt_pho * sizeof(int));
        }
        // synchronous copy from a pageable temporary: the blob dies at the end of this scope
        CK(cudaMemcpy(h->g_stops.p, blob.data(), stop_bytes, cudaMemcpyHostToDevice));
    }

    pk::GlobalArgs a;
    memset(&a, 0, sizeof(a));
    a.tp = d;
    a.sm = th->sm;
    a.B = j->B; a.T = T; a.P = P; a.theta_mode = j->theta_mode; a.n_stops = ns;
    a.stop_t = (const double*)h->g_stops.p;
    a.stop_out = (const int*)((const char*)h->g_stops.p + (size_t)ns * sizeof(double));
    a.stop_bucket = a.stop_out + ns;
    a.mt_prot = a.stop_bucket + ns;
    a.mt_rna = a.mt_prot + j->n_mt_prot;
    a.mt_pho = a.mt_rna + j->n_mt_rna;
    a.n_mt_prot = j->n_mt_prot; a.n_mt_rna = j->n_mt_rna; a.n_mt_pho = j->n_mt_pho;
    a.mb_prot = j->mb_prot; a.mb_rna = j->mb_rna; a.mb_pho = j->mb_pho;
    // defaults: measured error against the reference's tight solution <= 0.27 of the 1e-6 parity bound on every
    // golden network (0.13 at 1e-6/1e-9, non-monotone above 2e-6) — DESIGN.md §5
    a.rtol = j->rtol > 0 ? j->rtol : 2e-6;
    a.atol = j->atol > 0 ? j->atol : 2e-9;
    a.max_steps = j->max_steps > 0 ? j->max_steps : 200000;
    a.loss_mode = j->loss_mode < 0 ? 7 : j->loss_mode;
    a.metric = j->metric;
    for (int m = 0; m < 3; ++m) a.lam[m] = j->lambdas[m];
    a.lam_prior = j->lambda_prior;
    a.y0_stride = j->y0_stride;
    a.nfc = (long long)nfc;
    {   // fixed-point Schur solves below this ||K||_inf (global_net.cuh: schur_neumann); PHOSKIN_SCHUR_ITER=0 forces the
        // exact inversion in every step (tests compare the two paths)
        const char* ev = getenv("PHOSKIN_SCHUR_ITER");
        a.schur_iter_max = ev ? atof(ev) : 0.5;
    }
    a.counter = h->counter;

    const pkh::global_kernel_t kern = pkh::kernel_for_tile(th->sm.tile, d.model == 2, th->sm.ovf_total > 0);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)th->smem_bytes));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, pk::GLOBAL_BLOCK, th->smem_bytes));
    if (per_sm < 1) return fail("pk_global_solve_batch: kernel does not fit on an SM");
    const int grid = (int)std::min<size_t>(B, (size_t)per_sm * h->sm_count);

    const size_t y0_elems = j->y0_stride ? (B - 1) * (size_t)j->y0_stride + n : (size_t)n;
    const size_t TN = (size_t)T * n;
    if (!host) {
        a.params = j->params; a.y0 = j->y0;
        a.out_Y = j->out_Y; a.out_loss = j->out_loss; a.out_F = j->out_F; a.out_metric = j->out_metric;
        a.out_status = j->out_status; a.out_nsteps = j->out_nsteps; a.out_nrej = j->out_nrej;
        a.out_fc = j->out_fc;
    } else {
#define WS(buf, need, bytes, dstfield)                                                        \
    do {                                                                                      \
        if (need) { CK(h->buf.ensure(bytes)); dstfield = (decltype(dstfield))h->buf.p; }      \
    } while (0)
        WS(g_params, true, B * P * sizeof(double), a.params);
        WS(g_y0, true, y0_elems * sizeof(double), a.y0);
        WS(g_Y, j->out_Y, B * TN * sizeof(double), a.out_Y);
        WS(g_loss, j->out_loss, B * 3 * sizeof(double), a.out_loss);
        WS(g_F, j->out_F, B * 3 * sizeof(double), a.out_F);
        WS(g_metric, j->out_metric, B * sizeof(double), a.out_metric);
        WS(g_fc, j->out_fc, B * nfc * sizeof(double), a.out_fc);
        WS(g_status, j->out_status, B * sizeof(int32_t), a.out_status);
        WS(g_nsteps, j->out_nsteps, B * sizeof(int32_t), a.out_nsteps);
        WS(g_nrej, j->out_nrej, B * sizeof(int32_t), a.out_nrej);
#undef WS
        CK(cudaMemcpyAsync((void*)a.params, j->params, B * P * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync((void*)a.y0, j->y0, y0_elems * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    if (!a.out_Y) {                                   // per-CTA trajectory slot (reused for every system of that CTA)
        CK(h->g_traj.ensure((size_t)grid * TN * sizeof(double)));
        a.traj = (double*)h->g_traj.p;
    }
    if (th->binv_elems) {                             // model 2: per-CTA block inverses (rewritten every step)
        CK(h->g_binv.ensure((size_t)grid * th->binv_elems * sizeof(double)));
        a.binv = (double*)h->g_binv.p;
        a.binv_stride = (long long)th->binv_elems;
    }
    if (th->sm.ovf_total) {                           // capacity fallback: per-CTA slices of the arrays outside shared memory
        CK(h->g_ovf.ensure((size_t)grid * th->sm.ovf_total * sizeof(double)));
        a.ovf = (double*)h->g_ovf.p;
    }
    CK(cudaMemsetAsync(h->counter, 0, sizeof(unsigned long long), st));
    CK(cudaEventRecord(h->ev0, st));
    kern<<<grid, pk::GLOBAL_BLOCK, th->smem_bytes, st>>>(a);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return fail(std::string("global_net_kernel launch: ") + cudaGetErrorString(le));
    CK(cudaEventRecord(h->ev1, st));
    h->last_launches = 1;
    if (host) {
#define BACK(dst, src, bytes)                                                                 \
    do {                                                                                      \
        if (dst) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));            \
    } while (0)
        BACK(j->out_Y, a.out_Y, B * TN * sizeof(double));
        BACK(j->out_loss, a.out_loss, B * 3 * sizeof(double));
        BACK(j->out_F, a.out_F, B * 3 * sizeof(double));
        BACK(j->out_metric, a.out_metric, B * sizeof(double));
        BACK(j->out_fc, a.out_fc, B * nfc * sizeof(double));
        BACK(j->out_status, a.out_status, B * sizeof(int32_t));
        BACK(j->out_nsteps, a.out_nsteps, B * sizeof(int32_t));
        BACK(j->out_nrej, a.out_nrej, B * sizeof(int32_t));
#undef BACK
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    return 0;
}

int pk_global_loss_batch(pk_handle_t h, int32_t topo_id, int32_t memspace, const double* Y, int64_t B, int32_t T,
                         int32_t loss_mode, double* out_loss) {
    GlobalTopoHost* th = pkh::topo_of(h, topo_id);
    if (!th) return fail("pk_global_loss_batch: bad handle or topology id");
    if (!th->has_loss) return fail("pk_global_loss_batch needs pk_global_set_loss_data");
    if (B < 0 || T < 1 || !Y || !out_loss) return fail("pk_global_loss_batch: bad arguments");
    if (th->T_loss_max >= T) return fail("loss tables reference a time index >= T");
    if (loss_mode < -1 || loss_mode > 7) return fail("loss_mode out of range");
    h->last_launches = 0;
    h->last_ms = 0.f;
    if (B == 0) return 0;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const size_t ybytes = (size_t)B * T * th->dev.n * sizeof(double), lbytes = (size_t)B * 3 * sizeof(double);
    const double* dY = Y;
    double* dL = out_loss;
    if (memspace == PK_HOST) {
        CK(h->g_Y.ensure(ybytes));
        CK(h->g_loss.ensure(lbytes));
        dY = (const double*)h->g_Y.p;
        dL = (double*)h->g_loss.p;
        CK(cudaMemcpyAsync(h->g_Y.p, Y, ybytes, cudaMemcpyHostToDevice, st));
    }
    const int grid = (int)std::min<int64_t>(B, (int64_t)h->sm_count * 8);
    CK(cudaEventRecord(h->ev0, st));
    pk::global_loss_kernel<<<grid, pk::GLOBAL_BLOCK, 0, st>>>(th->dev, dY, B, T, loss_mode < 0 ? 7 : loss_mode, dL);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return fail(std::string("global_loss_kernel launch: ") + cudaGetErrorString(le));
    CK(cudaEventRecord(h->ev1, st));
    h->last_launches = 1;
    if (memspace == PK_HOST) CK(cudaMemcpyAsync(out_loss, dL, lbytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    return 0;
}

// debug build only: per-phase cycle totals of CTA 0 (see PH() in global_net.cuh); reading clears them
int pk_global_rhs_batch(pk_handle_t h, int32_t topo_id, int32_t memspace, int64_t B, const double* params, int32_t theta_mode,
                        const double* Y, const double* t_host, double* out_f, double* out_J, const double* tf_direct,
                        const double* S_direct) {
    GlobalTopoHost* th = pkh::topo_of(h, topo_id);
    if (!th) return fail("pk_global_rhs_batch: unknown topology id");
    if ((tf_direct == nullptr) != (S_direct == nullptr)) return fail("pk_global_rhs_batch: tf_direct and S_direct go together");
    if (B < 0 || !params || !Y || (!t_host && !tf_direct) || !out_f) return fail("pk_global_rhs_batch: params, Y, t and out_f are required");
    h->last_launches = 0;
    if (B == 0) return 0;
    CK(cudaSetDevice(h->device));
    const pk::GlobalTopoDev& d = th->dev;
    const int n = d.n, P = th->P;
    const bool host = memspace == PK_HOST;
    cudaStream_t st = h->stream;
    std::vector<int> bucket((size_t)B);
    for (int64_t b = 0; b < B && t_host; ++b) bucket[(size_t)b] = pkh::bucket_of(t_host[b], th->kin_grid);   // utils.py:210-225
    pk::GlobalRjArgs a;
    memset(&a, 0, sizeof(a));
    a.tp = d; a.B = B; a.P = P;
    CK(h->g_stops.ensure((size_t)B * sizeof(int)));
    CK(cudaMemcpy(h->g_stops.p, bucket.data(), (size_t)B * sizeof(int), cudaMemcpyHostToDevice));
    a.bucket = (const int*)h->g_stops.p;
    const size_t nf = (size_t)B * n, nJ = out_J ? (size_t)B * n * n : 0;
    if (host) {
        CK(h->g_params.ensure((size_t)B * P * sizeof(double)));
        CK(h->g_y0.ensure(nf * sizeof(double)));
        CK(h->g_Y.ensure((nf + nJ) * sizeof(double)));
        CK(cudaMemcpyAsync(h->g_params.p, params, (size_t)B * P * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->g_y0.p, Y, nf * sizeof(double), cudaMemcpyHostToDevice, st));
        a.params = (const double*)h->g_params.p; a.Y = (const double*)h->g_y0.p;
        a.out_f = (double*)h->g_Y.p; a.out_J = out_J ? (double*)h->g_Y.p + nf : nullptr;
        if (tf_direct) {
            const size_t ntf = (size_t)B * d.N, nS = (size_t)B * d.S;
            CK(h->g_traj.ensure((ntf + nS) * sizeof(double)));
            CK(cudaMemcpyAsync(h->g_traj.p, tf_direct, ntf * sizeof(double), cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync((double*)h->g_traj.p + ntf, S_direct, nS * sizeof(double), cudaMemcpyHostToDevice, st));
            a.tf_direct = (const double*)h->g_traj.p;
            a.S_direct = (const double*)h->g_traj.p + ntf;
        }
    } else {
        a.params = params; a.Y = Y; a.out_f = out_f; a.out_J = out_J;
        a.tf_direct = tf_direct; a.S_direct = S_direct;
    }
    if (theta_mode) {              // raw decision vectors -> physical parameters (params.py:106-132), in a scratch copy
        const long long np_ = (long long)B * P;
        CK(h->g_F.ensure((size_t)np_ * sizeof(double)));
        pk::softplus_kernel<<<(unsigned)((np_ + 255) / 256), 256, 0, st>>>(a.params, (double*)h->g_F.p, np_);
        a.params = (const double*)h->g_F.p;
    }
    const size_t smem = ((size_t)P + d.K + d.S + 2 * (size_t)d.N + 2 * (size_t)n) * sizeof(double) + (size_t)n * sizeof(int);
    auto kern = d.model == 2 ? pk::global_rhsjac_kernel<true> : pk::global_rhsjac_kernel<false>;
    if (smem > 227 * 1024) return fail("pk_global_rhs_batch: network too large for one CTA's shared memory");
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)std::min<int64_t>(B, (int64_t)h->sm_count * 4);
    CK(cudaEventRecord(h->ev0, st));
    kern<<<grid, pk::GLOBAL_BLOCK, smem, st>>>(a);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, st));
    h->last_launches = 1;
    if (host) {
        CK(cudaMemcpyAsync(out_f, a.out_f, nf * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (out_J) CK(cudaMemcpyAsync(out_J, a.out_J, nJ * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    return 0;
}

int pk_global_solve_custom(pk_handle_t h, int32_t topo_id, int32_t memspace, int64_t B, const double* params, int32_t theta_mode,
                           const double* y0, int64_t y0_stride, const double* t_eval_host, int32_t T, double rtol, double atol,
                           int32_t max_steps, double* out_Y, int32_t* out_status, int32_t* out_nsteps, int32_t* out_nrej) {
    GlobalTopoHost* th = pkh::topo_of(h, topo_id);
    if (!th) return fail("pk_global_solve_custom: unknown topology id");
    if (B < 0 || T < 1 || !params || !y0 || !t_eval_host || !out_Y) return fail("pk_global_solve_custom: params, y0, t_eval and out_Y are required");
    for (int k = 1; k < T; ++k)
        if (!(t_eval_host[k] > t_eval_host[k - 1])) return fail("t_eval must be strictly increasing");
    h->last_launches = 0;
    if (B == 0) return 0;
    CK(cudaSetDevice(h->device));
    const pk::GlobalTopoDev& d = th->dev;
    const int n = d.n, P = th->P;
    const bool host = memspace == PK_HOST;
    cudaStream_t st = h->stream;
    pk::GlobalRkArgs a;
    memset(&a, 0, sizeof(a));
    a.tp = d; a.B = B; a.P = P; a.T = T; a.theta_mode = theta_mode;
    a.max_steps = max_steps > 0 ? max_steps : 2000000;                       // solvers.py:294
    a.rtol = rtol > 0 ? rtol : 1e-5; a.atol = atol > 0 ? atol : 1e-7;        // solvers.py:293
    a.y0_stride = y0_stride;
    CK(h->g_t.ensure((size_t)T * sizeof(double)));
    CK(cudaMemcpy(h->g_t.p, t_eval_host, (size_t)T * sizeof(double), cudaMemcpyHostToDevice));
    a.t_eval = (const double*)h->g_t.p;
    const size_t nY = (size_t)B * T * n, ny0 = y0_stride ? ((size_t)B - 1) * (size_t)y0_stride + n : (size_t)n;
    if (host) {
        CK(h->g_params.ensure((size_t)B * P * sizeof(double)));
        CK(h->g_y0.ensure(ny0 * sizeof(double)));
        CK(h->g_Y.ensure(nY * sizeof(double)));
        CK(h->g_status.ensure((size_t)B * sizeof(int)));
        CK(h->g_nsteps.ensure((size_t)B * sizeof(int)));
        CK(h->g_nrej.ensure((size_t)B * sizeof(int)));
        CK(cudaMemcpyAsync(h->g_params.p, params, (size_t)B * P * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->g_y0.p, y0, ny0 * sizeof(double), cudaMemcpyHostToDevice, st));
        a.params = (const double*)h->g_params.p; a.y0 = (const double*)h->g_y0.p; a.out_Y = (double*)h->g_Y.p;
        a.out_status = (int*)h->g_status.p; a.out_nsteps = (int*)h->g_nsteps.p; a.out_nrej = (int*)h->g_nrej.p;
    } else {
        a.params = params; a.y0 = y0; a.out_Y = out_Y;
        a.out_status = out_status; a.out_nsteps = out_nsteps; a.out_nrej = out_nrej;
    }
    const size_t smem = ((size_t)P + d.K + d.S + 2 * (size_t)d.N + 9 * (size_t)n) * sizeof(double) + (size_t)n * sizeof(int);
    auto kern = d.model == 2 ? pk::global_dopri5_kernel<true> : pk::global_dopri5_kernel<false>;
    if (smem > 227 * 1024) return fail("pk_global_solve_custom: network too large for one CTA's shared memory");
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, pk::GLOBAL_BLOCK, smem));
    const unsigned grid = (unsigned)std::min<int64_t>(B, (int64_t)h->sm_count * std::max(per_sm, 1));
    CK(cudaEventRecord(h->ev0, st));
    kern<<<grid, pk::GLOBAL_BLOCK, smem, st>>>(a);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, st));
    h->last_launches = 1;
    if (host) {
        CK(cudaMemcpyAsync(out_Y, a.out_Y, nY * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (out_status) CK(cudaMemcpyAsync(out_status, a.out_status, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (out_nsteps) CK(cudaMemcpyAsync(out_nsteps, a.out_nsteps, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (out_nrej) CK(cudaMemcpyAsync(out_nrej, a.out_nrej, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    return 0;
}

#ifdef PK_GLOBAL_TRACE
int pk_global_trace_read(unsigned long long* out16) {
    unsigned long long zero[16] = {0};
    if (cudaMemcpyFromSymbol(out16, g_phase_cycles, sizeof(zero)) != cudaSuccess) return fail("trace read failed");
    if (cudaMemcpyToSymbol(g_phase_cycles, zero, sizeof(zero)) != cudaSuccess) return fail("trace clear failed");
    return 0;
}
#endif

}  // extern "C"
