// vgs_pairs_tc.cu -- K5: FrobeniusNorm (intersection) and Projection on the tensor cores, for sets whose contexts fill a
// useful part of the universe of all strings up to the deepest order (full Markov chains, cfg 2-like shallow trees).
//
// Reference (paths relative to the reference repository): src/distance/frobenius.pyx:21-53 (C = keys_l & keys_r,
// d = ||P_l - P_r||_F / sqrt(|C|), float32 rows), src/distance/projection.pyx:7-36 (zero rows for absent contexts,
// d = ||v_l - v_r||_2).  Over the dense universe U (index u = (4^len - 1) / 3 + code, column 4 u + symbol) with
// a_i = model i's rows (0 where the context is absent) and n_i = its context indicator:
//
//   Frobenius:   S_ij = sum_c [ sq_i(c) n_j(c) + sq_j(c) n_i(c) ] - 2 a_i . a_j ,   sq_i(c) = sum_x a_i(c, x)^2
//                |C|  = n_i . n_j                                  d = sqrt(S) / sqrt(|C|)
//   Projection:  S_ij = a_i . a_i + a_j . a_j - 2 a_i . a_j        d = sqrt(S)
//
// a . b - style formulas cancel catastrophically in floating point (|a|^2 ~ 10^2, S down to 10^-6 between relatives), so
// everything is done in EXACT integers on the int8 tensor-core GEMM (vgs_i8gemm.cuh): every float32 probability becomes
// q = round(p * 2^30) (p in [0, 1]) in four balanced base-256 digits, sq_i(c) = sum_x q^2 in eight, and the three
// contractions are integer GEMMs whose int32 digit-pair sums are recombined with 128-bit integers.  S is then the exact
// sum of squared differences of the QUANTISED rows -- exactly 0 for equal rows, no cancellation error at all.  What
// differs from the CUDA-core path (and the reference): the rows are rounded to 2^-31 absolute instead of being float32
// (|dq| <= 2^-31 per entry), and the difference a - b is not rounded to float32 before squaring (relative 2^-24 per
// term).  Bound, stated in the tests: |d_tc - d_oracle| <= 1.3e-7 d + 1e-9.
//
// Cost: 16 (digit pairs) x 4 |U| MACs per ordered pair for a . a (upper triangle only), 9 |U| for the masked sums --
// independent of how many contexts the models have, where the CUDA-core kernel's cost grows with them.  The dispatcher
// (vgs_pairs.cu) takes this path when sum(n_ctx) / (N |U|) >= 3 % (measured crossover), depth <= 7 and N <= 4096.
#include <stdlib.h>

#include <algorithm>

#include "vgs_i8gemm.cuh"

#define TC_SCALE 1073741824.0   // 2^30
#define TC_Q_ROWS 9             // per model in the second operand: eight digits of sq(c), then the context indicator

__device__ __forceinline__ int64_t tc_universe_index(int len, uint32_t code) { return ((((int64_t)1) << (2 * len)) - 1) / 3 + (int64_t)code; }

// one thread per context: the digit rows of its four probabilities, of their squared sum, and the indicator
__global__ void k_tc_operands(int64_t total, const uint8_t *__restrict__ len, const uint32_t *__restrict__ code,
                              const float4 *__restrict__ prob32, const int32_t *__restrict__ model_of_ctx, int64_t p_rows_pad,
                              int8_t *__restrict__ P4, int64_t q_rows_pad, int8_t *__restrict__ Q9, int64_t i_rows_pad,
                              int8_t *__restrict__ I1, int *__restrict__ flag) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= total) return;
    const int64_t m = model_of_ctx[c];
    const int64_t u = tc_universe_index(len[c], code[c]);
    const float4 r = prob32[c];
    const float pv[4] = {r.x, r.y, r.z, r.w};
    long long q[4];
    unsigned long long sq = 0;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        if (!(pv[x] >= 0.0f && pv[x] <= 1.0f)) atomicOr(flag, 1);   // not a probability: no fixed-point form (CUDA-core path)
        q[x] = __double2ll_rn((double)pv[x] * TC_SCALE);
        sq += (unsigned long long)(q[x] * q[x]);
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        uint32_t w = 0;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const long long d = ((q[x] + 128) & 255) - 128;
            w |= ((uint32_t)d & 255u) << (8 * x);
            q[x] = (q[x] - d) >> 8;
        }
        *(uint32_t *)(P4 + g8_off(m * 4 + p, 4 * u, p_rows_pad)) = w;
    }
    long long s = (long long)sq;   // < 2^63
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const long long d = ((s + 128) & 255) - 128;
        Q9[g8_off(m * TC_Q_ROWS + t, u, q_rows_pad)] = (int8_t)d;
        s = (s - d) >> 8;
    }
    Q9[g8_off(m * TC_Q_ROWS + 8, u, q_rows_pad)] = 1;
    I1[g8_off(m, u, i_rows_pad)] = 1;
}

__global__ void k_tc_model_of_ctx(int64_t n_models, const ModelMeta *__restrict__ meta, int32_t *__restrict__ model_of_ctx) {
    const ModelMeta mm = meta[blockIdx.x];
    for (int c = threadIdx.x; c < mm.n_ctx; c += blockDim.x) model_of_ctx[mm.ctx_off + c] = (int32_t)blockIdx.x;
}

// GEMM epilogue: the raw int32 digit-pair sums, row-major [A row][B row]
struct EpiRaw {
    static constexpr int CHUNK = 32;
    int32_t *out;
    int64_t ld, a_rows;
    struct State {};
    __device__ __forceinline__ State prepare(const G8Unit &, int64_t) const { return State(); }
    __device__ __forceinline__ void prefetch(const G8Unit &, int64_t, int, const State &) const {}
    __device__ __forceinline__ void unit_done() const {}
    __device__ __forceinline__ void operator()(const G8Unit &u, int64_t a_row, int c0, const int32_t (&v)[32], const State &) const {
        if (a_row >= a_rows) return;
        int4 *dst = (int4 *)(out + a_row * ld + u.b_row0 + c0);
#pragma unroll
        for (int t = 0; t < 8; ++t) dst[t] = make_int4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
    }
};

// 128-bit signed sum of v * 256^e
__device__ __forceinline__ void tc_acc(__int128 &acc, int32_t v, int e) { acc += ((__int128)v) << (8 * e); }
__device__ __forceinline__ double tc_to_double(__int128 s) {   // s >= 0, < 2^100
    const unsigned long long lo = (unsigned long long)s;
    const long long hi = (long long)(s >> 64);
    return __fma_rn((double)hi, 18446744073709551616.0, (double)lo);
}

// one thread per unordered pair (i <= j): recombine the digit-pair sums, write both triangles
__global__ void __launch_bounds__(256) k_tc_combine(int64_t n, int kind, const int32_t *__restrict__ G, int64_t g_ld,
                                                    const int32_t *__restrict__ R, int64_t r_ld, float *__restrict__ D32,
                                                    double *__restrict__ D64) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n || j < i) return;
    __int128 dot = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int4 g = *(const int4 *)(G + (4 * i + p) * g_ld + 4 * j);
        tc_acc(dot, g.x, p); tc_acc(dot, g.y, p + 1); tc_acc(dot, g.z, p + 2); tc_acc(dot, g.w, p + 3);
    }
    double d;
    if (kind == VGS_PAIR_PROJECTION) {
        __int128 nii = 0, njj = 0;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int4 a = *(const int4 *)(G + (4 * i + p) * g_ld + 4 * i), b = *(const int4 *)(G + (4 * j + p) * g_ld + 4 * j);
            tc_acc(nii, a.x, p); tc_acc(nii, a.y, p + 1); tc_acc(nii, a.z, p + 2); tc_acc(nii, a.w, p + 3);
            tc_acc(njj, b.x, p); tc_acc(njj, b.y, p + 1); tc_acc(njj, b.z, p + 2); tc_acc(njj, b.w, p + 3);
        }
        const __int128 S = nii + njj - 2 * dot;
        d = sqrt(tc_to_double(S) * (1.0 / (TC_SCALE * TC_SCALE)));
    } else {
        __int128 sij = 0, sji = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            tc_acc(sij, R[(TC_Q_ROWS * i + t) * r_ld + j], t);
            tc_acc(sji, R[(TC_Q_ROWS * j + t) * r_ld + i], t);
        }
        const int cnt = R[(TC_Q_ROWS * i + 8) * r_ld + j];
        const __int128 S = sij + sji - 2 * dot;
        // empty intersection: 0 / 0 in the reference (numpy: nan)
        d = cnt > 0 ? sqrt(tc_to_double(S) * (1.0 / (TC_SCALE * TC_SCALE))) / sqrt((double)cnt) : vgs_nan();
    }
    D32[i * n + j] = (float)d;
    D32[j * n + i] = (float)d;
    if (D64) { D64[i * n + j] = d; D64[j * n + i] = d; }
}

// unit list of one of the two GEMMs, cached on the device for the life of the model set (a steady-state call uploads nothing)
static int tc_units(vgs_context *h, uint64_t key, const std::vector<G8Rect> &rects, int kblocks, const G8Unit **d_units, int *n_units,
                    cudaStream_t s) {
    if (WorkCache *wc = vgs_wcache_find(h, key)) {
        *d_units = (const G8Unit *)wc->dev;
        *n_units = wc->n[0];
        return VGS_OK;
    }
    std::vector<G8Unit> units;
    g8_make_units(rects, kblocks, h->sm_count / 2, 1, 32, &units);
    const int64_t bytes = (int64_t)units.size() * (int64_t)sizeof(G8Unit) + 64;
    WorkCache *wc = vgs_wcache_add(h, key, bytes);
    void *dev;
    if (wc) {
        wc->n[0] = (int)units.size();
        dev = wc->dev;
    } else {
        int rc = vgs_ensure_scratch2(h, bytes);
        if (rc) return rc;
        dev = h->d_scratch2;
    }
    if (!units.empty()) VGS_CUDA(cudaMemcpyAsync(dev, units.data(), units.size() * sizeof(G8Unit), cudaMemcpyHostToDevice, s));
    *d_units = (const G8Unit *)dev;
    *n_units = (int)units.size();
    return VGS_OK;   // pageable source: the copy has consumed the vector when it returns
}

// is the set one this path covers, and does it pay?  (depth, size, density of the universe; VGS_PAIR_TC=0 / 1 forces)
bool vgs_pair_tc_wanted(vgs_context *h, int kind) {
    if (kind != VGS_PAIR_FROBENIUS_INTERSECTION && kind != VGS_PAIR_PROJECTION) return false;
    static int env_force = -2;
    if (env_force == -2) { const char *e = getenv("VGS_PAIR_TC"); env_force = e ? atoi(e) : -1; }
    const int force = h->pair_path >= 0 ? h->pair_path : env_force;   // vgs_set_pair_path wins over the environment
    if (force == 0) return false;
    if (h->max_order > VGS_UMAP_MAX_ORDER || h->n_models > 4096 || h->n_models < 2) return false;
    if (force == 1) return true;
    const double universe = (double)((((int64_t)1) << (2 * (h->max_order + 1))) - 1) / 3.0;
    return (double)h->total_ctx >= 0.03 * universe * (double)h->n_models && h->n_models >= 64;
}

// the whole [n][n] matrix on this GPU.  Returns VGS_OK, an error, or -1 = a row is not a probability (caller: CUDA cores)
int vgs_pair_tc_matrix(vgs_context *h, int kind, float *D32, double *D64, cudaStream_t s) {
    const int64_t n = h->n_models;
    const int D = h->max_order;
    const int64_t U = ((((int64_t)1) << (2 * (D + 1))) - 1) / 3;
    const int64_t K8 = (4 * U + 127) / 128 * 128, KQ8 = (U + 127) / 128 * 128;
    const int64_t p_rows = 4 * n, p_rows_pad = (p_rows + 255) / 256 * 256;
    const int64_t q_rows = TC_Q_ROWS * n, q_rows_pad = (q_rows + 255) / 256 * 256;
    const int64_t i_rows_pad = (n + 31) / 32 * 32;
    const bool frob = kind == VGS_PAIR_FROBENIUS_INTERSECTION;
    int rc;
    // operands (rebuilt per model set; kept while the set is loaded)
    if (h->tc_gen != h->model_gen) {
        if ((rc = vgs_reserve(h, &h->d_tc_P, p_rows_pad * K8))) return rc;
        if ((rc = vgs_reserve(h, &h->d_tc_Q, q_rows_pad * KQ8))) return rc;
        if ((rc = vgs_reserve(h, &h->d_tc_I, i_rows_pad * KQ8))) return rc;
        if ((rc = vgs_reserve(h, &h->d_tc_model, h->total_ctx))) return rc;
        if ((rc = vgs_reserve(h, &h->d_counters, 16))) return rc;
        VGS_CUDA(cudaMemsetAsync(h->d_tc_P, 0, (size_t)(p_rows_pad * K8), s));
        VGS_CUDA(cudaMemsetAsync(h->d_tc_Q, 0, (size_t)(q_rows_pad * KQ8), s));
        VGS_CUDA(cudaMemsetAsync(h->d_tc_I, 0, (size_t)(i_rows_pad * KQ8), s));
        VGS_CUDA(cudaMemsetAsync(h->d_counters + 11, 0, sizeof(int), s));
        k_tc_model_of_ctx<<<(unsigned)n, 128, 0, s>>>(n, h->d_meta, h->d_tc_model);
        VGS_LAUNCH_CHECK(h);
        k_tc_operands<<<(unsigned)((h->total_ctx + 255) / 256), 256, 0, s>>>(h->total_ctx, h->d_len, h->d_code, h->d_prob32, h->d_tc_model,
                                                                          p_rows_pad, h->d_tc_P, q_rows_pad, h->d_tc_Q, i_rows_pad,
                                                                          h->d_tc_I, h->d_counters + 11);
        VGS_LAUNCH_CHECK(h);
        int flag = 0;
        VGS_CUDA(cudaMemcpyAsync(&flag, h->d_counters + 11, sizeof(int), cudaMemcpyDeviceToHost, s));
        VGS_CUDA(cudaStreamSynchronize(s));
        h->tc_ok = flag == 0;
        h->tc_gen = h->model_gen;
    }
    if (!h->tc_ok) return -1;
    // digit-pair sums
    const int64_t g_ld = (p_rows + 31) / 32 * 32, r_ld = i_rows_pad;
    if ((rc = vgs_reserve(h, &h->d_tc_G, p_rows_pad * g_ld))) return rc;
    if (frob && (rc = vgs_reserve(h, &h->d_tc_R, q_rows_pad * r_ld))) return rc;
    const uint64_t key = vgs_mix(vgs_mix(0x50544355ull, h->model_gen), (uint64_t)n);
    {
        // a . a: upper triangle of 256-row (64-model) blocks
        std::vector<G8Rect> rects;
        if (!vgs_wcache_find(h, vgs_mix(key, 1)))
            for (int64_t a = 0; a * 256 < p_rows; ++a) {
                G8Rect r;
                r.a_blk = (int32_t)a; r.b_row0 = (int32_t)(a * 256); r.b_rows = (int32_t)(g_ld - a * 256);
                if (r.b_rows > 0) rects.push_back(r);
            }
        G8Args g;
        g.A = (const uint8_t *)h->d_tc_P; g.B = (const uint8_t *)h->d_tc_P; g.a_rows_pad = p_rows_pad; g.b_rows_pad = p_rows_pad;
        g.a_signed = 1; g.err = h->d_err; g.prof = nullptr; g.stages = G8_STAGES; g.debug = 0; g.gran = 32;
        if ((rc = tc_units(h, vgs_mix(key, 1), rects, (int)(K8 / G8_BK), &g.units, &g.n_units, s))) return rc;
        EpiRaw e;
        e.out = h->d_tc_G; e.ld = g_ld; e.a_rows = p_rows;
        vgs_prof_begin(h, VGS_PROF_PAIR_PARTIALS, s);
        rc = g8_launch(h, g, e, s);
        vgs_prof_end(h, VGS_PROF_PAIR_PARTIALS, s);
        if (rc) return rc;
    }
    if (frob) {
        // sq . n and n . n: every (model, model) pair, both orders
        std::vector<G8Rect> rects;
        if (!vgs_wcache_find(h, vgs_mix(key, 2)))
            for (int64_t a = 0; a * 256 < q_rows; ++a) {
                G8Rect r;
                r.a_blk = (int32_t)a; r.b_row0 = 0; r.b_rows = (int32_t)r_ld;
                rects.push_back(r);
            }
        G8Args g;
        g.A = (const uint8_t *)h->d_tc_Q; g.B = (const uint8_t *)h->d_tc_I; g.a_rows_pad = q_rows_pad; g.b_rows_pad = i_rows_pad;
        g.a_signed = 1; g.err = h->d_err; g.prof = nullptr; g.stages = G8_STAGES; g.debug = 0; g.gran = 32;
        if ((rc = tc_units(h, vgs_mix(key, 2), rects, (int)(KQ8 / G8_BK), &g.units, &g.n_units, s))) return rc;
        EpiRaw e;
        e.out = h->d_tc_R; e.ld = r_ld; e.a_rows = q_rows;
        if ((rc = g8_launch(h, g, e, s))) return rc;
    }
    k_tc_combine<<<dim3((unsigned)((n + 255) / 256), (unsigned)n), 256, 0, s>>>(n, kind, h->d_tc_G, g_ld, h->d_tc_R, r_ld, D32, D64);
    VGS_LAUNCH_CHECK(h);
    h->last_pair_kernel = 1;
    return VGS_OK;
}
