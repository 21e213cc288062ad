// vgs_estimate.cu -- K6: EstimateVLMC for all (left model, right sequence) pairs.
//
// Reference behaviour restated (src/distance/estimate.pyx, paths relative to the reference repository):
//   :53-71  _count_events: for every position t of the right model's sequence S and every suffix c of
//           S[:t] (length <= min(t, order_left)) that is a context of the LEFT model, counts[c][S[t]] += 1,
//           i.e. counts[c][x] = number of occurrences of the substring c.x in S;
//   :73-78  pseudo-counts: if a visited context has a zero count, add 1 to all four;
//   :40-51  estimated probabilities = counts / total;
//   :80-108 Pearson statistic sum (obs - exp)^2 / exp over visited contexts, over the symbols with
//           p_left > 0 except the last such symbol (scipy.stats.power_divergence, lambda_ = "pearson").
//
// Device formulation.  The counts depend on the left model only through WHICH substrings are counted, so
// per right sequence we build dense k-mer count tables for k = 1 .. K (K = max order + 1) once:
//   level K by a histogram pass over the packed sequence (atomics into L2), level k < K by summing the
//   four (k+1)-mers that extend a k-mer to the left, plus the one k-mer that starts the sequence.
// A warp then handles one (left, right) pair: lanes walk the left model's contexts (coalesced keys and
// fp64 rows), fetch the four counts of c.A, c.C, c.G, c.T with one 16-byte load, form the Pearson terms
// in fp64 and shuffle-reduce.
#include <algorithm>

#include "vgs_common.cuh"

__host__ __device__ inline int64_t level_off(int k) { return ((((int64_t)1) << (2 * k)) - 4) / 3; }  // u32 units, k >= 1

// level K histogram: grid (chunks, sequences-in-batch)
__global__ void k_est_count_top(int K, int64_t first_seq, const uint32_t *__restrict__ seq, const int64_t *__restrict__ seq_off,
                                const int64_t *__restrict__ seq_len, uint32_t *__restrict__ tables, int64_t per_seq) {
    const int64_t s = first_seq + blockIdx.y;
    const int64_t L = seq_len[s];
    const uint32_t *w = seq + seq_off[s];
    uint32_t *top = tables + (int64_t)blockIdx.y * per_seq + level_off(K);
    const uint32_t kmask = vgs_lowmask(K);
    const int64_t span = 256;  // positions per thread
    const int64_t n_threads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t a = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * span; a < L; a += n_threads * span) {
        const int64_t e = min(L, a + span);
        int64_t t = max((int64_t)0, a - (K - 1));
        uint32_t code = 0;
        for (; t < e; ++t) {
            const uint32_t x = (w[t >> 4] >> (30 - 2 * (int)(t & 15))) & 3u;
            code = ((code << 2) | x) & kmask;
            if (t >= a && t >= K - 1) atomicAdd(&top[code], 1u);
        }
    }
}

// level k from level k+1: cnt_k[c] = sum_a cnt_{k+1}[(a << 2k) | c] + [c == S[0:k]]
__global__ void k_est_count_down(int k, int64_t first_seq, const uint32_t *__restrict__ seq, const int64_t *__restrict__ seq_off,
                                 const int64_t *__restrict__ seq_len, uint32_t *__restrict__ tables, int64_t per_seq) {
    const int64_t s = first_seq + blockIdx.y;
    const int64_t L = seq_len[s];
    uint32_t *base = tables + (int64_t)blockIdx.y * per_seq;
    const uint32_t *up = base + level_off(k + 1);
    uint32_t *cur = base + level_off(k);
    const int64_t n = (int64_t)1 << (2 * k);
    uint32_t prefix = 0xFFFFFFFFu;
    if (L >= k) {
        const uint32_t *w = seq + seq_off[s];
        prefix = 0;
        for (int t = 0; t < k; ++t) prefix = (prefix << 2) | ((w[t >> 4] >> (30 - 2 * (t & 15))) & 3u);
    }
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        uint32_t v = up[c] + up[n + c] + up[2 * n + c] + up[3 * n + c];
        if ((uint32_t)c == prefix) v += 1u;
        cur[c] = v;
    }
}

// all levels of one sequence in one CTA (K <= 7: the top level fits in shared memory): histogram of the K-mers with
// shared-memory atomics, then the shorter levels by the same left-extension sums, ping-ponging between two shared
// arrays; every level is written to the global table as it is finished.  One launch instead of memset + K launches,
// and no atomics in L2.
__global__ void __launch_bounds__(512) k_est_count_fused(int K, int64_t first_seq, const uint32_t *__restrict__ seq,
                                                         const int64_t *__restrict__ seq_off, const int64_t *__restrict__ seq_len,
                                                         uint32_t *__restrict__ tables, int64_t per_seq) {
    extern __shared__ uint32_t est_sm[];
    const int64_t s = first_seq + blockIdx.x;
    const int64_t L = seq_len[s];
    const uint32_t *w = seq + seq_off[s];
    uint32_t *out = tables + (int64_t)blockIdx.x * per_seq;
    const int64_t nK = (int64_t)1 << (2 * K);
    uint32_t *up = est_sm, *cur = est_sm + nK;
    for (int64_t c = threadIdx.x; c < nK; c += blockDim.x) up[c] = 0u;
    __syncthreads();
    {
        const uint32_t kmask = vgs_lowmask(K);
        const int64_t span = (L + blockDim.x - 1) / blockDim.x;
        const int64_t a = (int64_t)threadIdx.x * span, e = min(L, a + span);
        if (a < e) {
            int64_t t = max((int64_t)0, a - (K - 1));
            uint32_t code = 0;
            for (; t < e; ++t) {
                const uint32_t x = (w[t >> 4] >> (30 - 2 * (int)(t & 15))) & 3u;
                code = ((code << 2) | x) & kmask;
                if (t >= a && t >= K - 1) atomicAdd(&up[code], 1u);
            }
        }
    }
    __syncthreads();
    for (int64_t c = threadIdx.x; c < nK; c += blockDim.x) out[level_off(K) + c] = up[c];
    for (int k = K - 1; k >= 1; --k) {
        const int64_t n = (int64_t)1 << (2 * k);
        uint32_t prefix = 0xFFFFFFFFu;
        if (L >= k) {
            prefix = 0;
            for (int t = 0; t < k; ++t) prefix = (prefix << 2) | ((w[t >> 4] >> (30 - 2 * (t & 15))) & 3u);
        }
        for (int64_t c = threadIdx.x; c < n; c += blockDim.x) {
            uint32_t v = up[c] + up[n + c] + up[2 * n + c] + up[3 * n + c];
            if ((uint32_t)c == prefix) v += 1u;
            cur[c] = v;
            out[level_off(k) + c] = v;
        }
        __syncthreads();
        uint32_t *tmp = up; up = cur; cur = tmp;
    }
}

__global__ void k_est_invp(int64_t n, const double *__restrict__ prob64, double *__restrict__ invp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const double v = prob64[i]; invp[i] = v > 0 ? __ddiv_rn(1.0, v) : 0.0; }
}

struct EstParams {
    const ModelMeta *meta;
    const uint32_t *key;
    const double *prob64;
    const double *invp;      // 1 / p per (context, symbol), 0 where p == 0 (built once per model set)
    const uint32_t *tables;  // count tables of the J-block's sequences
    int64_t per_seq;
    TilePlan pl;
    int64_t J0, J1;          // tile columns [J0, J1); every owned tile (I, J) of these columns is done by one launch
    int rank;
    float *pay32;
    double *pay64;
};

__global__ void k_est_pairs(const EstParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t T = p.pl.tile;
    const int64_t nJ = p.J1 - p.J0;
    if (warp >= p.pl.nt * nJ * T * T) return;
    const int64_t tile_idx = warp / (T * T), within = warp % (T * T);
    const int64_t I = tile_idx / nJ, J = p.J0 + tile_idx % nJ;
    const int64_t k = I * p.pl.nt + J;
    if (k % p.pl.world != p.rank) return;   // tile (I, J) belongs to another rank
    const int64_t q = k / p.pl.world;
    const int64_t ii = within / T, jj = within % T;
    const int64_t i = I * T + ii, j = J * T + jj;
    double acc = 0.0;
    if (i < p.pl.n && j < p.pl.n) {
        const ModelMeta mm = p.meta[i];
        const uint32_t *tab = p.tables + (j - p.J0 * T) * p.per_seq;
        for (int ci = lane; ci < mm.n_ctx; ci += 32) {
            const uint32_t key = __ldg(p.key + mm.ctx_off + ci);
            const int len = (31 - __clz(key)) >> 1;
            const uint32_t code = key ^ (1u << (2 * len));
            const uint4 cn = *(const uint4 *)(tab + level_off(len + 1) + 4 * (int64_t)code);
            int64_t c0 = cn.x, c1 = cn.y, c2 = cn.z, c3 = cn.w;
            int64_t tot = c0 + c1 + c2 + c3;
            if (tot != 0 && (c0 == 0 || c1 == 0 || c2 == 0 || c3 == 0)) { c0++; c1++; c2++; c3++; tot += 4; }
            if (tot > 0) {
                // (obs - exp)^2 / exp with obs = count / total: ONE division per context (1 / total) and the
                // precomputed 1 / p instead of six; every factor is within an ulp of the exact quotient, far inside
                // the 1e-9 bound on the statistic (the counts fluctuate by >= 1e-3 relative)
                const double2 *pr = (const double2 *)(p.prob64 + 4 * (mm.ctx_off + ci));
                const double2 *ir = (const double2 *)(p.invp + 4 * (mm.ctx_off + ci));
                const double2 p01 = __ldg(pr), p23 = __ldg(pr + 1), i01 = __ldg(ir), i23 = __ldg(ir + 1);
                const double pv[4] = {p01.x, p01.y, p23.x, p23.y};
                const double iv[4] = {i01.x, i01.y, i23.x, i23.y};
                const int64_t cv[4] = {c0, c1, c2, c3};
                const double rtot = __ddiv_rn(1.0, (double)tot);
                int last = -1;
#pragma unroll
                for (int x = 0; x < 4; ++x) if (pv[x] > 0) last = x;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    if (pv[x] > 0 && x != last) {
                        const double d = __dsub_rn(__dmul_rn((double)cv[x], rtot), pv[x]);
                        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(d, d), iv[x]));
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (lane == 0) {
        const int64_t e = q * T * T + within;
        p.pay32[e] = (float)acc;
        if (p.pay64) p.pay64[e] = acc;
    }
}

// K <= 7: the count table of one right sequence (<= 87 KB) is staged in shared memory (bulk async copy) and shared by
// every left model of the group: the per-context 16-byte count gather comes from shared memory instead of L2.
// CTA = (right sequence, group of EST_GROUP left models); warps take the left models in turn.
#define EST_GROUP 128
#define EST_STAGED_THREADS 512
__global__ void __launch_bounds__(EST_STAGED_THREADS) k_est_pairs_staged(const EstParams p) {
    extern __shared__ __align__(128) uint8_t est_smem[];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t T = p.pl.tile;
    const int64_t j = p.J0 * T + blockIdx.x;           // right sequence
    const int64_t i_begin = (int64_t)blockIdx.y * EST_GROUP, i_end = min(p.pl.n, i_begin + EST_GROUP);
    if (j >= p.pl.n || i_begin >= i_end) return;
    const int64_t J = j / T, jj = j % T;
    // does this rank own any tile (I, J) of the group's model blocks?
    bool any = false;
    for (int64_t I = i_begin / T; I <= (i_end - 1) / T; ++I) any = any || ((I * p.pl.nt + J) % p.pl.world) == p.rank;
    if (!any) return;
    const uint32_t bytes = (uint32_t)((p.per_seq * 4 + 15) / 16 * 16);
    if (tid == 0) vgs_mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) vgs_stage(est_smem, p.tables + (j - p.J0 * T) * p.per_seq, bytes, &bar);
    vgs_mbar_wait(&bar, 0);
    const uint32_t *tab = (const uint32_t *)est_smem;
    for (int64_t i = i_begin + warp; i < i_end; i += EST_STAGED_THREADS / 32) {
        const int64_t k = (i / T) * p.pl.nt + J;
        if (k % p.pl.world != p.rank) continue;
        const ModelMeta mm = p.meta[i];
        double acc = 0.0;
        for (int ci = lane; ci < mm.n_ctx; ci += 32) {
            const uint32_t key = __ldg(p.key + mm.ctx_off + ci);
            const int len = (31 - __clz(key)) >> 1;
            const uint32_t code = key ^ (1u << (2 * len));
            const uint4 cn = *(const uint4 *)(tab + level_off(len + 1) + 4 * (int64_t)code);
            int64_t c0 = cn.x, c1 = cn.y, c2 = cn.z, c3 = cn.w;
            int64_t tot = c0 + c1 + c2 + c3;
            if (tot != 0 && (c0 == 0 || c1 == 0 || c2 == 0 || c3 == 0)) { c0++; c1++; c2++; c3++; tot += 4; }
            if (tot > 0) {
                // (obs - exp)^2 / exp with obs = count / total: ONE division per context (1 / total) and the
                // precomputed 1 / p instead of six; every factor is within an ulp of the exact quotient, far inside
                // the 1e-9 bound on the statistic (the counts fluctuate by >= 1e-3 relative)
                const double2 *pr = (const double2 *)(p.prob64 + 4 * (mm.ctx_off + ci));
                const double2 *ir = (const double2 *)(p.invp + 4 * (mm.ctx_off + ci));
                const double2 p01 = __ldg(pr), p23 = __ldg(pr + 1), i01 = __ldg(ir), i23 = __ldg(ir + 1);
                const double pv[4] = {p01.x, p01.y, p23.x, p23.y};
                const double iv[4] = {i01.x, i01.y, i23.x, i23.y};
                const int64_t cv[4] = {c0, c1, c2, c3};
                const double rtot = __ddiv_rn(1.0, (double)tot);
                int last = -1;
#pragma unroll
                for (int x = 0; x < 4; ++x) if (pv[x] > 0) last = x;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    if (pv[x] > 0 && x != last) {
                        const double d = __dsub_rn(__dmul_rn((double)cv[x], rtot), pv[x]);
                        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(d, d), iv[x]));
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            const int64_t e = (k / p.pl.world) * T * T + (i % T) * T + jj;
            p.pay32[e] = (float)acc;
            if (p.pay64) p.pay64[e] = acc;
        }
    }
}

static int est_tiles_impl(vgs_context *h, const TilePlan &pl, int rank, float *pay32, double *pay64, cudaStream_t s) {
    const int K = h->max_order + 1;
    const int64_t per_seq = level_off(K + 1);  // u32 entries for levels 1..K
    // memory budget for the count tables: queried only when the existing scratch is too small for everything at once
    // (cudaMemGetInfo is a blocking driver call; a steady-state step must not make it)
    size_t free_b = 0, total_b = 0;
    if (per_seq * 4 * pl.tile * pl.nt + 256 > h->scratch2_bytes) VGS_CUDA(cudaMemGetInfo(&free_b, &total_b));
    // count tables for as many tile columns at a time as fit in the budget (all of them for the usual shapes)
    const int64_t budget = std::max<int64_t>(per_seq * 4 * pl.tile, std::min<int64_t>((int64_t)4 << 30, (int64_t)(free_b / 2) + h->scratch2_bytes));
    const int64_t cols_per_batch = std::max<int64_t>(1, std::min<int64_t>(pl.nt, budget / (per_seq * 4 * pl.tile)));
    const int64_t need = per_seq * 4 * pl.tile * cols_per_batch;
    if (per_seq * 4 * pl.tile > (int64_t)(free_b / 2) + h->scratch2_bytes) {
        vgs_set_error("EstimateVLMC count tables for order %d need %lld bytes per tile column; reduce the tile size", h->max_order,
                      (long long)(per_seq * 4 * pl.tile));
        return VGS_ERR_UNSUPPORTED;
    }
    int rc = vgs_ensure_scratch2(h, need + 256);
    if (rc) return rc;
    uint32_t *tables = (uint32_t *)h->d_scratch2;
    if (pl.slots > 0) VGS_CUDA(cudaMemsetAsync(pay32, 0, (size_t)(pl.slots * pl.tile * pl.tile * 4), s));
    if (pay64 && pl.slots > 0) VGS_CUDA(cudaMemsetAsync(pay64, 0, (size_t)(pl.slots * pl.tile * pl.tile * 8), s));
    if ((rc = vgs_reserve(h, &h->d_invp, h->total_ctx * 4))) return rc;
    if (h->invp_gen != h->model_gen) {
        k_est_invp<<<(unsigned)((h->total_ctx * 4 + 255) / 256), 256, 0, s>>>(h->total_ctx * 4, h->d_prob64, h->d_invp);
        VGS_LAUNCH_CHECK(h);
        h->invp_gen = h->model_gen;
    }
    const bool fused = K <= 7;
    const int fused_smem = fused ? (int)(((((int64_t)1) << (2 * K)) + (((int64_t)1) << (2 * (K - 1)))) * 4) : 0;
    if (fused) VGS_CUDA(cudaFuncSetAttribute(k_est_count_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_smem));
    for (int64_t J0 = 0; J0 < pl.nt; J0 += cols_per_batch) {
        const int64_t J1 = std::min(pl.nt, J0 + cols_per_batch);
        // owned tiles in these columns?
        bool any = false;
        for (int64_t J = J0; J < J1 && !any; ++J)
            for (int64_t I = 0; I < pl.nt && !any; ++I) any = ((I * pl.nt + J) % pl.world) == rank;
        if (!any) continue;
        const int64_t j0 = J0 * pl.tile, nj = std::min(pl.n, J1 * pl.tile) - j0;
        vgs_prof_begin(h, VGS_PROF_ESTIMATE_COUNTS, s);
        if (fused) {
            k_est_count_fused<<<(unsigned)nj, 512, fused_smem, s>>>(K, j0, h->d_seq, h->d_seq_off, h->d_seq_len, tables, per_seq);
            VGS_LAUNCH_CHECK(h);
        } else {
            for (int64_t b0 = 0; b0 < nj; b0 += 32768) {   // grid.y limit
                const int64_t nb = std::min<int64_t>(32768, nj - b0);
                uint32_t *tb = tables + b0 * per_seq;
                VGS_CUDA(cudaMemsetAsync(tb, 0, (size_t)(per_seq * 4 * nb), s));
                int64_t max_len = 1;
                for (int64_t j = j0 + b0; j < j0 + b0 + nb; ++j) max_len = std::max(max_len, h->h_seq_len[(size_t)j]);
                dim3 gtop((unsigned)std::max<int64_t>(1, std::min<int64_t>((max_len + 256 * 128 - 1) / (256 * 128), 64)), (unsigned)nb);
                k_est_count_top<<<gtop, 128, 0, s>>>(K, j0 + b0, h->d_seq, h->d_seq_off, h->d_seq_len, tb, per_seq);
                VGS_LAUNCH_CHECK(h);
                for (int k = K - 1; k >= 1; --k) {
                    const int64_t n = (int64_t)1 << (2 * k);
                    dim3 g((unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 4096)), (unsigned)nb);
                    k_est_count_down<<<g, 256, 0, s>>>(k, j0 + b0, h->d_seq, h->d_seq_off, h->d_seq_len, tb, per_seq);
                    VGS_LAUNCH_CHECK(h);
                }
            }
        }
        vgs_prof_end(h, VGS_PROF_ESTIMATE_COUNTS, s);
        {
            EstParams p;
            p.meta = h->d_meta; p.key = h->d_key; p.prob64 = h->d_prob64; p.invp = h->d_invp; p.tables = tables; p.per_seq = per_seq;
            p.pl = pl; p.J0 = J0; p.J1 = J1; p.rank = rank; p.pay32 = pay32; p.pay64 = pay64;
            const int64_t n_warps = pl.nt * (J1 - J0) * pl.tile * pl.tile;
            VGS_CHECK_ARG((n_warps * 32 + 255) / 256 < ((int64_t)1 << 31), "estimate kernel grid too large");
            const int stage_bytes = (int)((per_seq * 4 + 15) / 16 * 16);
            const int64_t n_groups = (pl.n + EST_GROUP - 1) / EST_GROUP;
            vgs_prof_begin(h, VGS_PROF_ESTIMATE_PAIRS, s);
            if (fused && stage_bytes + 2048 <= h->max_smem_optin && n_groups < 65536) {
                VGS_CUDA(cudaFuncSetAttribute(k_est_pairs_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, stage_bytes));
                k_est_pairs_staged<<<dim3((unsigned)nj, (unsigned)n_groups), EST_STAGED_THREADS, stage_bytes, s>>>(p);
            } else {
                k_est_pairs<<<(unsigned)((n_warps * 32 + 255) / 256), 256, 0, s>>>(p);
            }
            vgs_prof_end(h, VGS_PROF_ESTIMATE_PAIRS, s);
            VGS_LAUNCH_CHECK(h);
        }
    }
    return VGS_OK;
}

static int check_est(vgs_context *h) {
    VGS_CHECK_ARG(h, "null handle");
    VGS_CHECK_ARG(h->n_models > 0, "no models loaded (vgs_load_models)");
    VGS_CHECK_ARG(h->prob_ready, "models were loaded with VGS_LOAD_NLL_ONLY: no transition rows on the device");
    VGS_CHECK_ARG(h->n_seq == h->n_models, "EstimateVLMC needs exactly one sequence per model (sequence j sampled from model j)");
    if (h->max_order + 1 > 13) {
        vgs_set_error("EstimateVLMC dense counting supports order <= 12 (got %d)", h->max_order);
        return VGS_ERR_UNSUPPORTED;
    }
    return VGS_OK;
}

extern "C" int vgs_estimate_tiles(vgs_handle h, int64_t tile, int rank, int world, float *payload_d, void *stream) {
    int rc = check_est(h);
    if (rc) return rc;
    VGS_CHECK_ARG(tile > 0 && world > 0 && rank >= 0 && rank < world && payload_d, "vgs_estimate_tiles: bad arguments");
    VGS_CUDA(cudaSetDevice(h->device));
    TilePlan pl;
    vgs_make_plan(&pl, h->n_models, tile, world, 0);
    return est_tiles_impl(h, pl, rank, payload_d, nullptr, (cudaStream_t)stream);
}

extern "C" int vgs_estimate_matrix(vgs_handle h, float *D32_d, double *D64_d, void *stream) {
    int rc = check_est(h);
    if (rc) return rc;
    VGS_CHECK_ARG(D32_d, "vgs_estimate_matrix: null output");
    VGS_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    TilePlan pl;
    const int64_t n = h->n_models;
    int64_t tile = 64;
    const int K = h->max_order + 1;
    while (tile > 4 && level_off(K + 1) * 4 * tile > ((int64_t)4 << 30)) tile /= 2;  // keep the count tables under 4 GB
    vgs_make_plan(&pl, n, tile, 1, 0);
    const int64_t pay_elems = pl.slots * pl.tile * pl.tile;
    if ((rc = vgs_ensure_scratch(h, pay_elems * (D64_d ? 12 : 4) + 256))) return rc;
    float *pay32 = (float *)h->d_scratch;
    double *pay64 = D64_d ? (double *)((char *)h->d_scratch + (pay_elems * 4 + 255) / 256 * 256) : nullptr;
    if ((rc = est_tiles_impl(h, pl, 0, pay32, pay64, s))) return rc;
    return vgs_assemble_impl(h, pl, pay32, pay64, D32_d, D64_d, s);
}

extern "C" int vgs_estimate_matrix_h(vgs_handle h, float *D32_out_h, double *D64_out_h) {
    int rc = check_est(h);
    if (rc) return rc;
    VGS_CHECK_ARG(D32_out_h || D64_out_h, "vgs_estimate_matrix_h: no output");
    VGS_CUDA(cudaSetDevice(h->device));
    const int64_t n = h->n_models;
    if ((rc = vgs_reserve(h, &h->d_out32, n * n))) return rc;
    if (D64_out_h && (rc = vgs_reserve(h, &h->d_out64, n * n))) return rc;
    float *d32 = h->d_out32;
    double *d64 = D64_out_h ? h->d_out64 : nullptr;
    cudaStream_t s = h->own_stream;
    rc = vgs_estimate_matrix(h, d32, d64, s);
    if (!rc && D32_out_h) rc = cudaMemcpyAsync(D32_out_h, d32, 4 * (size_t)(n * n), cudaMemcpyDeviceToHost, s) == cudaSuccess ? VGS_OK : VGS_ERR_CUDA;
    if (!rc && D64_out_h) rc = cudaMemcpyAsync(D64_out_h, d64, 8 * (size_t)(n * n), cudaMemcpyDeviceToHost, s) == cudaSuccess ? VGS_OK : VGS_ERR_CUDA;
    cudaStreamSynchronize(s);
    return rc;
}
