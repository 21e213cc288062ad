// vgs_nll.cu -- K2 (log-likelihood score matrix M[s][m]) and K3 (NegativeLogLikelihood combine).
//
// Reference behaviour restated (paths relative to the reference repository):
//   src/vlmc/vlmc.pyx:79-101   _log_likelihood: sum_{t >= skip} f(p_t), f(p) = -1000 if p == 0 else ln p,
//                              fp64, strictly left to right; p_t from the longest-suffix context (:122-141)
//   src/distance/negloglikelihood.pyx:17-26   D = ((M_ii - M_ij)/L + (M_jj - M_ji)/L) / 2
//
// Kernel shape.  A persistent CTA owns a contiguous run of work units (model m, up to 512 sequences).
// The model's NLL table ("blob") is staged into shared memory with 1-D bulk async copies (TMA, UBLKCP)
// completing on an mbarrier; it is re-staged only when the model changes.  Two accumulation modes:
//   VGS_NLL_SEQUENTIAL  one lane per (sequence, model); the lane walks its sequence with 128-bit loads and
//                       adds the terms strictly in order -> bit-identical to the reference's fp64 sum;
//   VGS_NLL_WARP        one warp per (sequence, model); lane l takes 64-symbol chunks l, l+32, ...;
//                       4 partial sums per chunk, fixed-order shuffle reduction (north_star's layout).
// Per scored base: one funnel shift + mask (the (order+1)-mer ending at the base is a contiguous bit
// field of the packed sequence), one 8-byte shared-memory gather, one DADD.
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "vgs_common.cuh"

#define NLL_MAX_THREADS 1024

struct NllUnit {
    int32_t model;
    int32_t list_off;  // offset into seq_list, or first sequence index when seq_list == nullptr
    int32_t count;     // <= threads per CTA
    int32_t extra;     // extra sequence index scored by lane `count` (the diagonal M[m][m]) or -1
};

struct NllParams {
    const ModelMeta *meta;
    const uint8_t *arena;
    const uint2 *hash;
    const double *logp;
    const uint32_t *seq;
    const int64_t *seq_off;
    const int64_t *seq_len;
    const int32_t *seq_list;
    const NllUnit *units;
    int n_units;
    int skip_none;
    double *M;
    int64_t ss, sm, m_col0;   // M[s * ss + (m_col0 + model) * sm]
    int *counter;      // dynamic unit scheduler (zeroed before the launch)
};

// ---- table views -----------------------------------------------------------------------------
struct DenseTable {  // tier 1
    const double *tab;
    uint32_t fmask;
    __device__ __forceinline__ void bind(const uint8_t *blob, const ModelMeta &mm) {
        tab = (const double *)blob;
        fmask = vgs_lowmask(mm.order + 1);
    }
    __device__ __forceinline__ double term(uint32_t f) const { return tab[f & fmask]; }
    // the 16 symbols of one word, in order.  MASKED: only the positions whose bit is set in `valid` (bit j = symbol
    // j of the word) are scored; the others still gather (any index is inside the table) but add +0.0, which leaves
    // the sum bit-identical -- a partial word at the start / end of a sequence costs the same as a full one and
    // does not diverge from its neighbours in the warp
    template <bool MASKED>
    __device__ __forceinline__ void word16(uint32_t cur, uint32_t prev, uint32_t valid, double &acc) const {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const double v = tab[__funnelshift_r(cur, prev, 30 - 2 * j) & fmask];
            acc += (!MASKED || ((valid >> j) & 1u)) ? v : 0.0;
        }
    }
};

template <bool WIDE>
struct TwoLevelTable {  // tier 2 (u16 map, u32 trie entries) and tier 3 (u32 / u64: >= 32767 contexts or blocks)
    const void *l1;
    const void *ext;      // extension trie: blocks of four {value, child} entries (vgs_core.cu, k_build_two_level)
    const double *rows;   // [n_ctx + 1][4]: the last row is NaN ("no context"), map entries point at it
    uint32_t d1mask;
    int d1x2;
    __device__ __forceinline__ void bind(const uint8_t *blob, const ModelMeta &mm) {
        l1 = (const void *)blob;
        const int d1 = mm.l1_depth & 0xFF;
        d1x2 = 2 * d1;
        d1mask = vgs_lowmask(d1);
        ext = (const void *)(blob + mm.deep_off);
        rows = (const double *)(blob + mm.logp_off);
    }
    // map entry of a history -> ctx_id.  A flagged entry (a longer context ends in this D1-mer) continues in the trie:
    // one load per further history symbol, until the node has no child for it.
    __device__ __forceinline__ uint32_t resolve(uint32_t e, uint32_t hist) const {
        const uint32_t flag = WIDE ? 0x80000000u : 0x8000u;
        uint32_t id = e & (flag - 1u);
        if (e & flag) {
            uint32_t blk = id;
            uint32_t h = hist >> d1x2;
#pragma unroll 1
            do {
                if (WIDE) {
                    const uint2 w = ((const uint2 *)ext)[4 * (size_t)blk + (h & 3u)];
                    id = w.x; blk = w.y;
                    if (blk == 0xFFFFFFFFu) break;
                } else {
                    const uint32_t w = ((const uint32_t *)ext)[4 * blk + (h & 3u)];
                    id = w & 0xFFFFu; blk = w >> 16;
                    if (blk == 0xFFFFu) break;
                }
                h >>= 2;
            } while (true);
        }
        return id;
    }
    __device__ __forceinline__ uint32_t entry(uint32_t hist) const {
        return WIDE ? ((const uint32_t *)l1)[hist & d1mask] : (uint32_t)((const uint16_t *)l1)[hist & d1mask];
    }
    __device__ __forceinline__ double term(uint32_t f) const {
        const uint32_t hist = f >> 2;
        return rows[4 * resolve(entry(hist), hist) + (f & 3u)];
    }
    // One fully scored word in three phases, so that the unrolled code is branch-free except for the rare trie walk:
    // (1) the 16 first-level gathers, (2) flagged entries (~0.5 % of the bases at D1 = 8, depth 9) take the trie,
    // (3) the 16 row gathers and adds, in order.
    template <bool MASKED>
    __device__ __forceinline__ void word16(uint32_t cur, uint32_t prev, uint32_t valid, double &acc) const {
        uint32_t e[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) e[j] = entry(__funnelshift_r(cur, prev, 30 - 2 * j) >> 2);
#pragma unroll
        for (int j = 0; j < 16; ++j) e[j] = resolve(e[j], __funnelshift_r(cur, prev, 30 - 2 * j) >> 2);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const double v = rows[4 * e[j] + ((cur >> (30 - 2 * j)) & 3u)];
            acc += (!MASKED || ((valid >> j) & 1u)) ? v : 0.0;   // see DenseTable::word16
        }
    }
};

// 64 symbols (one uint4) starting at position t0; `prev` is the word before the chunk.
template <class TAB, bool SEQ>
__device__ __forceinline__ void scan_chunk(const TAB &T, const uint4 q, uint32_t prev, int64_t t0, int64_t skip, int64_t L,
                                           double &acc) {
    const uint32_t ws[4] = {q.x, q.y, q.z, q.w};
    double part[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t cur = ws[k];
        const int64_t tw = t0 + 16 * k;
        // positions [max(skip, tw), min(L, tw + 16)) of this word are scored; one code path for full and partial words
        const int64_t lo64 = skip - tw, hi64 = L - tw;
        const int lo = lo64 < 0 ? 0 : (lo64 > 16 ? 16 : (int)lo64), hi = hi64 < 0 ? 0 : (hi64 > 16 ? 16 : (int)hi64);
        if (hi > lo) {
            const uint32_t valid = (0xFFFFu >> (16 - hi)) & (0xFFFFu << lo);
            T.template word16<true>(cur, prev, valid, SEQ ? acc : part[k]);
        }
        prev = cur;
    }
    if (!SEQ) acc += (part[0] + part[1]) + (part[2] + part[3]);
}

// first min(order, L) positions of log_likelihood (skip = 0): short histories, generic hash lookup
__device__ double nll_prologue(const NllParams &p, const ModelMeta &mm, const uint32_t *w, int64_t L) {
    const uint2 *slots = p.hash + mm.hash_off;
    double acc = 0.0;
    uint32_t hist = 0;
    const int64_t top = L < mm.order ? L : mm.order;
    for (int64_t t = 0; t < top; ++t) {
        const uint32_t x = (w[t >> 4] >> (30 - 2 * (int)(t & 15))) & 3u;
        const int id = vgs_longest_suffix(slots, mm.hash_mask, mm.hash_shift, mm.order, hist, (int)t);
        acc += id < 0 ? vgs_nan() : p.logp[4 * (mm.ctx_off + id) + x];
        hist = (hist << 2) | x;
    }
    return acc;
}

template <int TIER, int MODE, bool STAGED>
__global__ void __launch_bounds__(NLL_MAX_THREADS, 1) k_nll_scores(const NllParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ int s_unit;
    const int tid = threadIdx.x;
    if (STAGED && tid == 0) vgs_mbar_init(&bar, 1);
    uint32_t parity = 0;
    int cur_model = -1;
    typename std::conditional<TIER == 1, DenseTable, TwoLevelTable<TIER == 3>>::type T;
    ModelMeta mm;
    // units are handed out dynamically (they are sorted by model, so a CTA mostly keeps its staged table)
    for (;;) {
        __syncthreads();  // every lane is done with the previous unit (and with the previous table)
        if (tid == 0) s_unit = atomicAdd(p.counter, 1);
        __syncthreads();
        const int u = s_unit;
        if (u >= p.n_units) break;
        const NllUnit un = p.units[u];
        if (un.model != cur_model) {
            mm = p.meta[un.model];
            cur_model = un.model;
            const uint8_t *blob = p.arena + mm.nll_off;
            if (STAGED) {
                if (tid == 0) vgs_stage(smem, blob, (uint32_t)mm.nll_bytes, &bar);
                vgs_mbar_wait(&bar, parity);
                parity ^= 1u;
                T.bind(smem, mm);
            } else {
                T.bind(blob, mm);
            }
        }
        const int64_t skip = mm.order;
        const int n_items = un.count + (un.extra >= 0 ? 1 : 0);
        if (MODE == VGS_NLL_SEQUENTIAL) {
            if (tid < n_items) {
                const int s = tid < un.count ? (p.seq_list ? p.seq_list[un.list_off + tid] : un.list_off + tid) : un.extra;
                const uint32_t *w = p.seq + p.seq_off[s];
                const int64_t L = p.seq_len[s];
                double acc = p.skip_none ? nll_prologue(p, mm, w, L) : 0.0;
                const int64_t n_chunks = (L + 63) >> 6;
                uint32_t prev = 0;
                const uint4 *w4 = (const uint4 *)w;
                for (int64_t c = 0; c < n_chunks; ++c) {
                    const uint4 q = __ldg(w4 + c);
                    scan_chunk<decltype(T), true>(T, q, prev, c << 6, skip, L, acc);
                    prev = q.w;
                }
                p.M[(int64_t)s * p.ss + (p.m_col0 + un.model) * p.sm] = acc;
            }
        } else {
            const int lane = tid & 31, warp = tid >> 5;
            for (int it = warp; it < n_items; it += (int)(blockDim.x >> 5)) {
                const int s = it < un.count ? (p.seq_list ? p.seq_list[un.list_off + it] : un.list_off + it) : un.extra;
                const uint32_t *w = p.seq + p.seq_off[s];
                const int64_t L = p.seq_len[s];
                const int64_t n_chunks = (L + 63) >> 6;
                const uint4 *w4 = (const uint4 *)w;
                double acc = 0.0;
                if (n_chunks >= 32) {
                    for (int64_t c = lane; c < n_chunks; c += 32) {
                        const uint4 q = __ldg(w4 + c);
                        const uint32_t prev = c ? __ldg(w + 4 * c - 1) : 0u;
                        scan_chunk<decltype(T), false>(T, q, prev, c << 6, skip, L, acc);
                    }
                } else {
                    // short sequence (< 2048 symbols): one 16-symbol word per lane and step, so that all lanes work
                    const int64_t n_words = (L + 15) >> 4;
                    for (int64_t wi = lane; wi < n_words; wi += 32) {
                        const uint32_t cur = __ldg(w + wi);
                        const uint32_t prev = wi ? __ldg(w + wi - 1) : 0u;
                        const int64_t t0 = wi << 4;
                        scan_chunk<decltype(T), false>(T, make_uint4(cur, 0u, 0u, 0u), prev, t0, skip, min(L, t0 + 16), acc);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) {
                    if (p.skip_none) acc = nll_prologue(p, mm, w, L) + acc;
                    p.M[(int64_t)s * p.ss + (p.m_col0 + un.model) * p.sm] = acc;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side: build the unit list and launch per tier
// ---------------------------------------------------------------------------------------------
// threads per CTA = sequences per unit.  512 is ~1% faster than 256 when the units are full (measured,
// B200, cfg 3); 256 wastes fewer lanes when the per-model sequence lists are not multiples of 512 (multi-GPU
// tilings).  VGS_NLL_THREADS overrides for experiments.
static int g_nll_threads = 512;
static int vgs_nll_threads() { return g_nll_threads; }
static void vgs_choose_nll_threads(const std::vector<int64_t> &list_sizes) {
    const char *e = getenv("VGS_NLL_THREADS");
    if (e) {
        int t = atoi(e);
        if (t < 32) t = 32;
        if (t > NLL_MAX_THREADS) t = NLL_MAX_THREADS;
        g_nll_threads = t / 32 * 32;
        return;
    }
    double used[2] = {0, 0}, alloc[2] = {0, 0};
    const int cand[2] = {512, 256};
    for (int64_t c : list_sizes) {
        if (c <= 0) continue;
        for (int k = 0; k < 2; ++k) {
            const int64_t units = (c + cand[k] - 1) / cand[k];
            const int64_t warps = ((c + units - 1) / units + 31) / 32;  // warps a unit keeps busy
            used[k] += (double)c;
            alloc[k] += (double)(units * warps * 32);
        }
    }
    const double e512 = alloc[0] > 0 ? used[0] / alloc[0] : 1.0, e256 = alloc[1] > 0 ? used[1] / alloc[1] : 1.0;
    g_nll_threads = (e256 > e512 + 0.03) ? 256 : 512;
}

template <int TIER, int MODE, bool STAGED>
static int launch_nll(vgs_context *h, const NllParams &p, int smem_bytes, cudaStream_t s) {
    auto kern = k_nll_scores<TIER, MODE, STAGED>;
    if (STAGED) VGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int per_sm = 1;
    // warp mode: a unit's sequences are dealt to the warps one at a time, so the CTA can always be full (one CTA per
    // SM when the table is staged: all the latency hiding comes from its own warps)
    const int threads = (MODE == VGS_NLL_WARP && !getenv("VGS_NLL_THREADS")) ? NLL_MAX_THREADS : vgs_nll_threads();
    VGS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, STAGED ? smem_bytes : 0));
    if (per_sm < 1) per_sm = 1;
    int grid = std::min(p.n_units, h->sm_count * per_sm);
    VGS_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(int), s));
    vgs_prof_begin(h, VGS_PROF_NLL_SCORES, s);
    kern<<<grid, threads, STAGED ? smem_bytes : 0, s>>>(p);
    vgs_prof_end(h, VGS_PROF_NLL_SCORES, s);
    VGS_LAUNCH_CHECK(h);
    return VGS_OK;
}

template <int TIER>
static int launch_nll_tier(vgs_context *h, const NllParams &p, int mode, int64_t max_bytes, cudaStream_t s) {
    const bool staged = max_bytes + 1024 <= (int64_t)h->max_smem_optin;
    const int smem = (int)((max_bytes + 127) / 128 * 128);
    if (mode == VGS_NLL_SEQUENTIAL) {
        return staged ? launch_nll<TIER, VGS_NLL_SEQUENTIAL, true>(h, p, smem, s) : launch_nll<TIER, VGS_NLL_SEQUENTIAL, false>(h, p, 0, s);
    }
    return staged ? launch_nll<TIER, VGS_NLL_WARP, true>(h, p, smem, s) : launch_nll<TIER, VGS_NLL_WARP, false>(h, p, 0, s);
}

// seq_lists_per_block[B] = sequences to score against every model of model-block B (models
// B*block .. min(n, (B+1)*block)-1); an empty outer vector means "all sequences, contiguous".
int vgs_nll_score_lists(vgs_context *h, const std::vector<std::vector<int32_t>> &lists, int64_t block, int64_t m_begin,
                        int64_t m_end, bool add_diagonal, int skip_mode, int mode, const VgsScoreDst &dst, cudaStream_t s,
                        uint64_t cache_key) {
    if (mode == VGS_NLL_KMER) mode = VGS_NLL_WARP;  // the scan kernels have two modes
    h->last_nll_kernel = VGS_NLL_KERNEL_SCAN;
    int rc;
    if ((rc = vgs_ensure_scan_tables(h, s))) return rc;   // dense / two-level tables: built on the first scan of a model set
    size_t n_u[3] = {0, 0, 0};
    NllUnit *d_units = nullptr;
    int32_t *d_list = nullptr;
    if (WorkCache *wc = vgs_wcache_find(h, cache_key)) {
        // steady state: the unit list of this (plan, rank) is already on the device
        for (int t = 0; t < 3; ++t) n_u[t] = (size_t)wc->n[t];
        g_nll_threads = wc->threads;
        d_units = (NllUnit *)wc->dev;
        d_list = (int32_t *)((char *)wc->dev + wc->off2);
    } else {
        std::vector<NllUnit> units[3];
        std::vector<int32_t> flat;
        std::vector<int64_t> list_off(lists.size() + 1, 0);
        for (size_t b = 0; b < lists.size(); ++b) {
            list_off[b + 1] = list_off[b] + (int64_t)lists[b].size();
            flat.insert(flat.end(), lists[b].begin(), lists[b].end());
        }
        {
            std::vector<int64_t> sizes;
            for (int64_t m = m_begin; m < m_end; ++m) {
                const size_t b = (size_t)(m / block);
                if (b >= lists.size()) break;
                sizes.push_back((int64_t)lists[b].size() + (add_diagonal ? 1 : 0));
            }
            vgs_choose_nll_threads(sizes);
        }
        for (int64_t m = m_begin; m < m_end; ++m) {
            const size_t b = (size_t)(m / block);
            if (b >= lists.size()) break;
            const std::vector<int32_t> &L = lists[b];
            const int t = h->h_meta[(size_t)m].tier - 1;
            bool diag_needed = add_diagonal && m < h->n_seq && !std::binary_search(L.begin(), L.end(), (int32_t)m);
            int64_t cnt = (int64_t)L.size();
            if (cnt == 0 && !diag_needed) continue;
            // even split into ceil(cnt / threads) units
            const int64_t per_unit = vgs_nll_threads();
            int64_t nu = std::max<int64_t>(1, (cnt + (diag_needed ? 1 : 0) + per_unit - 1) / per_unit);
            for (int64_t k = 0; k < nu; ++k) {
                int64_t a = cnt * k / nu, e = cnt * (k + 1) / nu;
                NllUnit un;
                un.model = (int32_t)m;
                un.list_off = (int32_t)(list_off[b] + a);
                un.count = (int32_t)(e - a);
                un.extra = (k == nu - 1 && diag_needed) ? (int32_t)m : -1;
                units[t].push_back(un);
            }
        }
        for (int t = 0; t < 3; ++t) n_u[t] = units[t].size();
        const int64_t n_units_total = (int64_t)(n_u[0] + n_u[1] + n_u[2]);
        if (n_units_total == 0) return VGS_OK;
        const int64_t off2 = (n_units_total * (int64_t)sizeof(NllUnit) + 127) / 128 * 128;
        const int64_t bytes = off2 + (int64_t)flat.size() * 4 + 256;
        char *base = nullptr;
        if (cache_key) {
            WorkCache *wc = vgs_wcache_add(h, cache_key, bytes);
            if (wc) {
                for (int t = 0; t < 3; ++t) wc->n[t] = (int)n_u[t];
                wc->threads = g_nll_threads;
                wc->off2 = off2;
                base = (char *)wc->dev;
            }
        }
        if (!base) {
            if ((rc = vgs_ensure_scratch2(h, bytes))) return rc;
            base = (char *)h->d_scratch2;
        }
        d_units = (NllUnit *)base;
        d_list = (int32_t *)(base + off2);
        size_t at = 0;
        for (int t = 0; t < 3; ++t) {
            if (n_u[t]) VGS_CUDA(cudaMemcpyAsync(d_units + at, units[t].data(), n_u[t] * sizeof(NllUnit), cudaMemcpyHostToDevice, s));
            at += n_u[t];
        }
        if (!flat.empty()) VGS_CUDA(cudaMemcpyAsync(d_list, flat.data(), flat.size() * 4, cudaMemcpyHostToDevice, s));
    }
    if (n_u[0] + n_u[1] + n_u[2] == 0) return VGS_OK;
    // pageable sources: the copies above have consumed the host vectors when they return
    NllParams p;
    p.meta = h->d_meta; p.arena = h->d_nll; p.hash = h->d_hash; p.logp = h->d_logp;
    p.seq = h->d_seq; p.seq_off = h->d_seq_off; p.seq_len = h->d_seq_len;
    p.seq_list = d_list;
    p.skip_none = skip_mode == VGS_SKIP_NONE;
    p.M = dst.M[0]; p.ss = dst.stride_s; p.sm = dst.stride_m; p.m_col0 = dst.m_col0;
    if ((rc = vgs_reserve(h, &h->d_counters, 16))) return rc;
    if (n_u[0]) {
        p.counter = h->d_counters;
        p.units = d_units; p.n_units = (int)n_u[0];
        if ((rc = launch_nll_tier<1>(h, p, mode, h->max_nll_bytes_t1, s))) return rc;
    }
    if (n_u[1]) {
        p.counter = h->d_counters + 1;
        p.units = d_units + n_u[0]; p.n_units = (int)n_u[1];
        if ((rc = launch_nll_tier<2>(h, p, mode, h->max_nll_bytes_t2, s))) return rc;
    }
    if (n_u[2]) {
        p.counter = h->d_counters + 3;
        p.units = d_units + n_u[0] + n_u[1]; p.n_units = (int)n_u[2];
        if ((rc = launch_nll_tier<3>(h, p, mode, h->max_nll_bytes_t3, s))) return rc;
    }
    return VGS_OK;
}

extern "C" int vgs_last_nll_kernel(vgs_handle h) { return h ? h->last_nll_kernel : 0; }

static int check_nll_ready(vgs_context *h) {
    VGS_CHECK_ARG(h, "null handle");
    VGS_CHECK_ARG(h->n_models > 0, "no models loaded (vgs_load_models)");
    VGS_CHECK_ARG(h->nll_ready, "models were loaded with VGS_LOAD_SKIP_NLL: no NLL tables");
    VGS_CHECK_ARG(h->n_seq > 0, "no sequences loaded (vgs_load_sequences_* / vgs_sample_sequences)");
    return VGS_OK;
}

extern "C" int vgs_nll_scores(vgs_handle h, int64_t s_begin, int64_t s_end, int64_t m_begin, int64_t m_end, int skip_mode,
                              int mode, double *M_d, int64_t ld, void *stream) {
    int rc = check_nll_ready(h);
    if (rc) return rc;
    VGS_CHECK_ARG(0 <= s_begin && s_begin <= s_end && s_end <= h->n_seq, "vgs_nll_scores: bad sequence range");
    VGS_CHECK_ARG(0 <= m_begin && m_begin <= m_end && m_end <= h->n_models, "vgs_nll_scores: bad model range");
    VGS_CHECK_ARG(M_d && ld >= h->n_models, "vgs_nll_scores: bad output");
    VGS_CHECK_ARG(mode >= VGS_NLL_SEQUENTIAL && mode <= VGS_NLL_KMER, "vgs_nll_scores: bad mode");
    VGS_CHECK_ARG(skip_mode == VGS_SKIP_ORDER || skip_mode == VGS_SKIP_NONE, "vgs_nll_scores: bad skip mode");
    VGS_CUDA(cudaSetDevice(h->device));
    if (s_begin == s_end || m_begin == m_end) return VGS_OK;
    if (mode == VGS_NLL_KMER && skip_mode == VGS_SKIP_ORDER && vgs_kmer_applicable(h)) {
        const int64_t nb_s = (h->n_seq + 127) / 128, nb_m = (h->n_models + 127) / 128;
        std::vector<char> need((size_t)(nb_s * nb_m), 0);
        for (int64_t sb = s_begin / 128; sb <= (s_end - 1) / 128; ++sb)
            for (int64_t mb = m_begin / 128; mb <= (m_end - 1) / 128; ++mb) need[(size_t)(sb * nb_m + mb)] = 1;
        rc = vgs_kmer_scores(h, need, nb_s, nb_m, false, vgs_dst_plain(M_d, ld), (cudaStream_t)stream);
        if (rc != -1) return rc;
    }
    std::vector<int32_t> seqs((size_t)(s_end - s_begin));
    for (int64_t i = s_begin; i < s_end; ++i) seqs[(size_t)(i - s_begin)] = (int32_t)i;
    std::vector<std::vector<int32_t>> lists(1, seqs);
    return vgs_nll_score_lists(h, lists, h->n_models, m_begin, m_end, false, skip_mode, mode, vgs_dst_plain(M_d, ld), (cudaStream_t)stream);
}

extern "C" int vgs_nll_scores_h(vgs_handle h, int skip_mode, int mode, double *M_out_h) {
    int rc = check_nll_ready(h);
    if (rc) return rc;
    VGS_CHECK_ARG(M_out_h, "vgs_nll_scores_h: null output");
    VGS_CUDA(cudaSetDevice(h->device));
    if ((rc = vgs_ensure_M(h))) return rc;
    cudaStream_t s = h->own_stream;
    std::vector<int32_t> all((size_t)h->n_seq);
    for (int64_t i = 0; i < h->n_seq; ++i) all[(size_t)i] = (int32_t)i;
    std::vector<std::vector<int32_t>> lists(1, all);
    VGS_CHECK_ARG(mode >= VGS_NLL_SEQUENTIAL && mode <= VGS_NLL_KMER, "bad mode");
    rc = -1;
    if (mode == VGS_NLL_KMER && skip_mode == VGS_SKIP_ORDER && vgs_kmer_applicable(h)) {
        const int64_t nb_s = (h->n_seq + 127) / 128, nb_m = (h->n_models + 127) / 128;
        std::vector<char> need((size_t)(nb_s * nb_m), 1);
        rc = vgs_kmer_scores(h, need, nb_s, nb_m, false, vgs_dst_plain(h->d_M, h->n_models), s);
        if (rc > 0) return rc;
    }
    if (rc == -1 && (rc = vgs_nll_score_lists(h, lists, h->n_models, 0, h->n_models, false, skip_mode, mode, vgs_dst_plain(h->d_M, h->n_models), s))) return rc;
    VGS_CUDA(cudaMemcpyAsync(M_out_h, h->d_M, 8 * (size_t)(h->n_seq * h->n_models), cudaMemcpyDeviceToHost, s));
    VGS_CUDA(cudaStreamSynchronize(s));
    return VGS_OK;
}

// ---------------------------------------------------------------------------------------------
// K3: combine scores into distances, tile by tile
// ---------------------------------------------------------------------------------------------
// x / L, correctly rounded, in three fp64 operations instead of the ~30 of the general division routine: with r = RN(1 / L)
// (computed by an IEEE division on the host), q = RN(x r), rem = x - q L (exact in one FMA), RN(q + rem r) is the correctly
// rounded quotient (Markstein's theorem; it needs L's significand not to be all ones -- L is a sequence length).  The
// distances stay bit-identical to the reference's Python-float divisions; the bit-exact tests cover millions of entries.
__device__ __forceinline__ double nll_div_len(double x, double L, double rL) {
    const double q = __dmul_rn(x, rL);
    const double rem = __fma_rn(-q, L, x);
    return __fma_rn(rem, rL, q);
}

__global__ void k_nll_combine(TilePlan pl, int rank, const double *__restrict__ M, int64_t ld, double length,
                              float *__restrict__ pay32, double *__restrict__ pay64, int *__restrict__ err) {
    const int64_t q = blockIdx.x;
    const int64_t k = q * pl.world + rank;
    const int64_t T = pl.tile;
    float *o32 = pay32 + q * T * T;
    double *o64 = pay64 ? pay64 + q * T * T : nullptr;
    __shared__ int64_t sI, sJ;
    if (threadIdx.x == 0) {
        int64_t I = 0, J = 0;
        if (k < pl.n_tiles) {
            int64_t rem = k;
            while (rem >= pl.nt - I) { rem -= pl.nt - I; ++I; }
            J = I + rem;
        }
        sI = I; sJ = J;
    }
    __syncthreads();
    const int64_t I = sI, J = sJ;
    // gridDim.y blocks share one tile slot
    const int64_t per = (T * T + gridDim.y - 1) / gridDim.y;
    const int64_t e_end = min(T * T, (int64_t)(blockIdx.y + 1) * per);
    for (int64_t e = (int64_t)blockIdx.y * per + threadIdx.x; e < e_end; e += blockDim.x) {
        const int64_t i = I * T + e / T, j = J * T + e % T;
        double d = 0.0;
        if (k < pl.n_tiles && i < pl.n && j < pl.n) {
            const double mii = M[i * ld + i], mij = M[i * ld + j], mjj = M[j * ld + j], mji = M[j * ld + i];
            const double rlen = 1.0 / length;
            const double lr = nll_div_len(__dsub_rn(mii, mij), length, rlen);
            const double rl = nll_div_len(__dsub_rn(mjj, mji), length, rlen);
            d = __dmul_rn(__dadd_rn(lr, rl), 0.5);   // halving is exact
            if (d != d) atomicOr(err, 1);
        }
        o32[e] = (float)d;
        if (o64) o64[e] = d;
    }
}

// single-GPU form: every entry of the [n][n] matrix straight from M, no payload / assemble pass.  Entry (j, i) adds
// the same two quotients as (i, j) in the other order, so both triangles carry the same bits as the tiled path.
// transposed != 0: M is the transposed score matrix Mt[m][s] (what the k-mer contraction and the model-sharded path write)
__global__ void __launch_bounds__(256) k_nll_combine_full(int64_t n, const double *__restrict__ M, int64_t ld, double length,
                                                          float *__restrict__ D32, double *__restrict__ D64, int *__restrict__ err,
                                                          int transposed) {
    __shared__ double t_ji[32][33];   // M[j][i] for the tile, read row-wise (coalesced), used column-wise
    __shared__ double d_i[32], d_j[32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    vgs_dep_trigger();
    vgs_dep_wait();
    const double rlen = 1.0 / length;   // one IEEE division per thread
    const int64_t i0 = (int64_t)blockIdx.y * 32, j0 = (int64_t)blockIdx.x * 32;
    // every global load of the thread is issued before the barrier (one round trip to memory instead of two: 8.1 -> 6 us at
    // cfg 3, where the kernel is all latency)
    double direct4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = ty + 8 * q;
        const int64_t j = j0 + r, i = i0 + tx;
        t_ji[r][tx] = (j < n && i < n) ? M[j * ld + i] : 0.0;
        const int64_t i2 = i0 + r, j2 = j0 + tx;
        direct4[q] = (i2 < n && j2 < n) ? M[i2 * ld + j2] : 0.0;
    }
    if (threadIdx.x < 32) {
        const int64_t i = i0 + threadIdx.x;
        d_i[threadIdx.x] = i < n ? M[i * ld + i] : 0.0;
    } else if (threadIdx.x < 64) {
        const int64_t j = j0 + threadIdx.x - 32;
        d_j[threadIdx.x - 32] = j < n ? M[j * ld + j] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = ty + 8 * q;
        const int64_t i = i0 + r, j = j0 + tx;
        if (i >= n || j >= n) continue;
        const double direct = direct4[q], mirrored = t_ji[tx][r];
        const double mij = transposed ? mirrored : direct, mji = transposed ? direct : mirrored;
        const double lr = nll_div_len(__dsub_rn(d_i[r], mij), length, rlen);
        const double rl = nll_div_len(__dsub_rn(d_j[tx], mji), length, rlen);
        const double d = __dmul_rn(__dadd_rn(lr, rl), 0.5);   // IEEE addition commutes; halving is exact
        if (d != d) atomicOr(err, 1);
        D32[i * n + j] = (float)d;
        if (D64) D64[i * n + j] = d;
    }
}


// ---------------------------------------------------------------------------------------------
// model-sharded scores (north_star (c) for the NLL kinds): this handle holds a SHARD of the models (local model m is
// column m_col0 + m of the job) and ALL sequences.  Every sequence is scored against the shard and the result is the
// band Mt[m_col0 + m][s] of the transposed score matrix -- contiguous, so the bands of all ranks assemble the matrix with
// one all-gather and no scatter kernel; with peer-mapped destinations the int8 GEMM's epilogue stores the band straight
// into every rank's matrix over NVLink.  Returns VGS_OK, an error (> 0), or -1: the k-mer contraction does not apply
// (only when mode == VGS_NLL_KMER and the caller asked for it with kmer_only)
static int nll_scores_band_impl(vgs_context *h, int mode, int64_t m_col0, int64_t ld, double *const *dsts, int n_dst, cudaStream_t s) {
    VgsScoreDst dst;
    for (int i = 0; i < VGS_MAX_PEERS; ++i) dst.M[i] = i < n_dst ? dsts[i] : nullptr;
    dst.n_dst = n_dst; dst.stride_s = 1; dst.stride_m = ld; dst.m_col0 = m_col0;
    int rc = -1;
    if (mode == VGS_NLL_KMER && vgs_kmer_applicable(h)) {
        const int64_t nb_s = (h->n_seq + 127) / 128, nb_m = (h->n_models + 127) / 128;
        uint64_t key = vgs_mix(vgs_mix(0x42414e44ull, h->model_gen), (uint64_t)h->n_seq);
        const std::vector<char> need((size_t)(nb_s * nb_m), 1);
        rc = vgs_kmer_scores(h, need, nb_s, nb_m, false, dst, s, key);
        if (rc > 0) return rc;
    }
    const bool fused_peer_stores = rc == VGS_OK && h->last_nll_kernel == VGS_NLL_KERNEL_I8MMA;
    if (rc == -1) {
        std::vector<int32_t> all((size_t)h->n_seq);
        for (int64_t i = 0; i < h->n_seq; ++i) all[(size_t)i] = (int32_t)i;
        std::vector<std::vector<int32_t>> lists(1, all);
        VgsScoreDst local = dst;
        local.n_dst = 1;
        if ((rc = vgs_nll_score_lists(h, lists, h->n_models, 0, h->n_models, false, VGS_SKIP_ORDER, mode, local, s))) return rc;
    }
    if (n_dst > 1 && !fused_peer_stores) {
        // the scan (and the fp64 fallback) write the local band only: copy it to the peers (plain peer-to-peer copies)
        const size_t off = (size_t)(m_col0 * ld), bytes = (size_t)(h->n_models * ld) * sizeof(double);
        for (int d = 1; d < n_dst; ++d) VGS_CUDA(cudaMemcpyAsync(dsts[d] + off, dsts[0] + off, bytes, cudaMemcpyDefault, s));
    }
    return VGS_OK;
}

static int launch_nll_combine(vgs_context *h, const TilePlan &pl, int rank, int64_t length, float *pay32, double *pay64,
                              cudaStream_t s, float *D32_direct = nullptr, double *D64_direct = nullptr) {
    if (D32_direct) {
        const unsigned g = (unsigned)((pl.n + 31) / 32);
        k_nll_combine_full<<<dim3(g, g), 256, 0, s>>>(pl.n, h->d_M, h->n_models, (double)length, D32_direct, D64_direct, h->d_err, 0);
        VGS_LAUNCH_CHECK(h);
        return VGS_OK;
    }
    if (pl.slots <= 0) return VGS_OK;
    const dim3 grid((unsigned)pl.slots, (unsigned)std::max<int64_t>(1, std::min<int64_t>(32, pl.tile * pl.tile / 2048)));
    k_nll_combine<<<grid, 256, 0, s>>>(pl, rank, h->d_M, h->n_models, (double)length, pay32, pay64, h->d_err);
    VGS_LAUNCH_CHECK(h);
    return VGS_OK;
}

// D32_direct (single GPU, whole matrix): the distances go straight into the [n][n] matrix instead of a tile payload
static int nll_tiles_impl(vgs_context *h, int64_t length, int mode, const TilePlan &pl, int rank, float *pay32, double *pay64,
                          cudaStream_t s, float *D32_direct = nullptr, double *D64_direct = nullptr) {
    int rc = vgs_ensure_M(h);
    if (rc) return rc;
    // the work lists depend only on (model set, plan, rank): built once, then reused from the device
    uint64_t key = vgs_mix(0x5647534eull, h->model_gen);
    key = vgs_mix(key, (uint64_t)pl.n);
    key = vgs_mix(key, (uint64_t)pl.tile);
    key = vgs_mix(key, (uint64_t)pl.world * 1024u + (uint64_t)rank);
    if (mode == VGS_NLL_KMER && vgs_kmer_applicable(h) && D32_direct) {
        // whole matrix on this GPU: all scores as one band of the transposed score matrix (coalesced epilogue stores),
        // then the distances straight from it
        double *dsts[1] = {h->d_M};
        rc = nll_scores_band_impl(h, mode, 0, h->n_seq, dsts, 1, s);
        if (rc > 0) return rc;
        if (rc == VGS_OK) {
            const unsigned g = (unsigned)((pl.n + 31) / 32);
            vgs_launch_dep(k_nll_combine_full, dim3(g, g), dim3(256), 0, s, pl.n, (const double *)h->d_M, h->n_seq, (double)length, D32_direct,
                           D64_direct, h->d_err, 1);
            VGS_LAUNCH_CHECK(h);
            return VGS_OK;
        }
    } else if (mode == VGS_NLL_KMER && vgs_kmer_applicable(h) && pl.tile % 128 == 0) {
        // needed (sequence-block, model-block) pairs at the contraction's 128 x 128 granularity; the diagonal
        // scores M[i][i] come from k_kmer_diag (all of them, on every rank: N dot products)
        const uint64_t kkey = vgs_mix(key, 2);
        const int64_t nb = (pl.n + 127) / 128, per = pl.tile / 128;
        std::vector<char> need;
        if (!vgs_wcache_find(h, kkey)) {
            need.assign((size_t)(nb * nb), 0);
            for (int64_t k = rank; k < pl.n_tiles; k += pl.world) {
                int64_t I, J;
                vgs_tile_coords(pl, k, &I, &J);
                for (int64_t a = I * per; a < std::min(nb, (I + 1) * per); ++a)
                    for (int64_t b = J * per; b < std::min(nb, (J + 1) * per); ++b) {
                        need[(size_t)(a * nb + b)] = 1;
                        need[(size_t)(b * nb + a)] = 1;
                    }
            }
        }
        rc = vgs_kmer_scores(h, need, nb, nb, true, vgs_dst_plain(h->d_M, h->n_models), s, kkey);
        if (rc > 0) return rc;
        if (rc == VGS_OK) return launch_nll_combine(h, pl, rank, length, pay32, pay64, s, D32_direct, D64_direct);
    }
    const uint64_t skey = vgs_mix(key, 1);
    std::vector<std::vector<int32_t>> lists;
    if (!vgs_wcache_find(h, skey)) {
        // which sequence blocks does this rank need against each model block?
        lists.resize((size_t)pl.nt);
        std::vector<std::vector<char>> need((size_t)pl.nt, std::vector<char>((size_t)pl.nt, 0));
        for (int64_t k = rank; k < pl.n_tiles; k += pl.world) {
            int64_t I, J;
            vgs_tile_coords(pl, k, &I, &J);
            need[(size_t)J][(size_t)I] = 1;  // model block J scores sequence block I ...
            need[(size_t)I][(size_t)J] = 1;  // ... and model block I scores sequence block J
        }
        for (int64_t B = 0; B < pl.nt; ++B)
            for (int64_t S = 0; S < pl.nt; ++S)
                if (need[(size_t)B][(size_t)S])
                    for (int64_t x = S * pl.tile; x < std::min(pl.n, (S + 1) * pl.tile); ++x) lists[(size_t)B].push_back((int32_t)x);
    }
    // models whose block touches no owned tile need nothing; the diagonal score M[m][m] rides along
    rc = vgs_nll_score_lists(h, lists, pl.tile, 0, h->n_models, true, VGS_SKIP_ORDER, mode, vgs_dst_plain(h->d_M, h->n_models), s, skey);
    if (rc) return rc;
    return launch_nll_combine(h, pl, rank, length, pay32, pay64, s, D32_direct, D64_direct);
}

extern "C" int vgs_nll_tiles(vgs_handle h, int64_t length, int mode, int64_t tile, int rank, int world, float *payload_d,
                             void *stream) {
    int rc = check_nll_ready(h);
    if (rc) return rc;
    VGS_CHECK_ARG(h->n_seq == h->n_models, "vgs_nll_tiles: need exactly one sequence per model (sequence i sampled from model i)");
    VGS_CHECK_ARG(length > 0 && tile > 0 && world > 0 && rank >= 0 && rank < world && payload_d, "vgs_nll_tiles: bad arguments");
    VGS_CHECK_ARG(mode >= VGS_NLL_SEQUENTIAL && mode <= VGS_NLL_KMER, "vgs_nll_tiles: bad mode");
    VGS_CHECK_ARG(mode != VGS_NLL_KMER || tile % 128 == 0, "vgs_nll_tiles: the k-mer mode needs a tile edge that is a multiple of 128");
    VGS_CUDA(cudaSetDevice(h->device));
    TilePlan pl;
    vgs_make_plan(&pl, h->n_models, tile, world, 1);
    return nll_tiles_impl(h, length, mode, pl, rank, payload_d, nullptr, (cudaStream_t)stream);
}


extern "C" int vgs_nll_scores_band(vgs_handle h, int mode, int64_t m_col0, int64_t ld, double *const *Mt_dsts_h, int n_dst, void *stream) {
    int rc = check_nll_ready(h);
    if (rc) return rc;
    VGS_CHECK_ARG(mode >= VGS_NLL_SEQUENTIAL && mode <= VGS_NLL_KMER, "vgs_nll_scores_band: bad mode");
    VGS_CHECK_ARG(m_col0 >= 0 && ld >= h->n_seq && Mt_dsts_h && n_dst >= 1 && n_dst <= VGS_MAX_PEERS, "vgs_nll_scores_band: bad arguments");
    for (int d = 0; d < n_dst; ++d) VGS_CHECK_ARG(Mt_dsts_h[d] != nullptr, "vgs_nll_scores_band: null destination");
    VGS_CUDA(cudaSetDevice(h->device));
    return nll_scores_band_impl(h, mode, m_col0, ld, Mt_dsts_h, n_dst, (cudaStream_t)stream);
}

// one step of the model-sharded job in ONE call (three kernels' worth of host work, one trip through the binding): the band
// into every destination, the peer barrier (when the destinations are peer-mapped), the local combine
extern "C" int vgs_nll_sharded_step(vgs_handle h, int mode, int64_t m_col0, int64_t n_total, double *const *Mt_dsts_h, int n_dst,
                                    uint32_t *const *flags_h, int rank, int world, uint32_t epoch, int64_t length, float *D32_d,
                                    void *stream) {
    int rc = vgs_nll_scores_band(h, mode, m_col0, n_total, Mt_dsts_h, n_dst, stream);
    if (rc) return rc;
    VGS_CHECK_ARG(length > 0 && D32_d, "vgs_nll_sharded_step: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (n_dst > 1) {
        VGS_CHECK_ARG(flags_h && n_dst == world, "vgs_nll_sharded_step: peer stores need every rank's matrix and flag array");
        if ((rc = vgs_peer_barrier_impl(h, flags_h, rank, world, epoch, s))) return rc;
    }
    const unsigned g = (unsigned)((n_total + 31) / 32);
    VGS_CHECK_ARG(g <= 65535, "vgs_nll_sharded_step: matrix too large");
    vgs_launch_dep(k_nll_combine_full, dim3(g, g), dim3(256), 0, s, n_total, (const double *)Mt_dsts_h[0], n_total, (double)length, D32_d,
                   (double *)nullptr, h->d_err, 1);
    VGS_LAUNCH_CHECK(h);
    return VGS_OK;
}

extern "C" int vgs_nll_combine(vgs_handle h, int64_t n, const double *Mt_d, int64_t ld, int64_t length, float *D32_d, double *D64_d,
                               void *stream) {
    VGS_CHECK_ARG(h && n > 0 && Mt_d && ld >= n && length > 0 && D32_d, "vgs_nll_combine: bad arguments");
    VGS_CUDA(cudaSetDevice(h->device));
    const unsigned g = (unsigned)((n + 31) / 32);
    VGS_CHECK_ARG(g <= 65535, "vgs_nll_combine: matrix too large");
    k_nll_combine_full<<<dim3(g, g), 256, 0, (cudaStream_t)stream>>>(n, Mt_d, ld, (double)length, D32_d, D64_d, h->d_err, 1);
    VGS_LAUNCH_CHECK(h);
    return VGS_OK;
}

static int64_t default_tile(int64_t n) { return n <= 2048 ? 64 : (n <= 8192 ? 128 : 256); }

extern "C" int vgs_nll_matrix(vgs_handle h, int64_t length, int mode, float *D32_d, double *D64_d, void *stream) {
    int rc = check_nll_ready(h);
    if (rc) return rc;
    VGS_CHECK_ARG(h->n_seq == h->n_models, "vgs_nll_matrix: need exactly one sequence per model (sequence i sampled from model i)");
    VGS_CHECK_ARG(length > 0 && D32_d, "vgs_nll_matrix: bad arguments");
    VGS_CHECK_ARG(mode >= VGS_NLL_SEQUENTIAL && mode <= VGS_NLL_KMER, "vgs_nll_matrix: bad mode");
    VGS_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    TilePlan pl;
    vgs_make_plan(&pl, h->n_models, mode == VGS_NLL_KMER ? 128 : default_tile(h->n_models), 1, 1);
    // one GPU: scores for every block, then the distances straight into the matrix (no payload, no assemble)
    return nll_tiles_impl(h, length, mode, pl, 0, nullptr, nullptr, s, D32_d, D64_d);
}

extern "C" int vgs_nll_matrix_h(vgs_handle h, int64_t length, int mode, float *D32_out_h, double *D64_out_h) {
    int rc = check_nll_ready(h);
    if (rc) return rc;
    VGS_CHECK_ARG(D32_out_h || D64_out_h, "vgs_nll_matrix_h: no output");
    VGS_CUDA(cudaSetDevice(h->device));
    const int64_t n = h->n_models;
    if ((rc = vgs_reserve(h, &h->d_out32, n * n))) return rc;
    if (D64_out_h && (rc = vgs_reserve(h, &h->d_out64, n * n))) return rc;
    float *d32 = h->d_out32;
    double *d64 = D64_out_h ? h->d_out64 : nullptr;
    cudaStream_t s = h->own_stream;
    // the result starts for the host behind the kernels and the error word is read behind it: ONE wait for all of it (on
    // an error the caller's buffers hold an unspecified matrix and the call reports the error).  The k-mer contraction's
    // once-per-sequence-set check rides along (i8_defer_check); in the rare case it voids the pass, the pass is redone.
    for (int attempt = 0; attempt < 2; ++attempt) {
        h->i8_defer_check = attempt == 0;
        rc = vgs_nll_matrix(h, length, mode, d32, d64, s);
        h->i8_defer_check = false;
        if (!rc && D32_out_h) rc = cudaMemcpyAsync(D32_out_h, d32, 4 * (size_t)(n * n), cudaMemcpyDeviceToHost, s) == cudaSuccess ? VGS_OK : VGS_ERR_CUDA;
        if (!rc && D64_out_h) rc = cudaMemcpyAsync(D64_out_h, d64, 8 * (size_t)(n * n), cudaMemcpyDeviceToHost, s) == cudaSuccess ? VGS_OK : VGS_ERR_CUDA;
        const int rc_flag = vgs_read_error_flag(h, s);   // synchronises
        if (!rc) rc = rc_flag;
        if (!vgs_kmer_i8_settle(h)) break;
    }
    return rc;
}
