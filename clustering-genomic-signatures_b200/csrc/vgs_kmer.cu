// vgs_kmer.cu -- NLL "k-mer contraction" mode (VGS_NLL_KMER): the score matrix as an fp64 contraction.
//
// Same quantity as K2 (src/vlmc/vlmc.pyx:79-101): M[s][m] = sum_{t >= order_m} f(p_m(S_s[t] | history)).
// With D = the largest order of the set, every term with t >= D depends on S_s only through the (D+1)-mer
// ending at t, so
//
//     M[s][m] = sum_k  C[s][k] * T[m][k]  +  head(s, m)
//
//   C[s][k]  number of positions t >= D of sequence s whose (D+1)-mer is k      -- one histogram per sequence,
//                                                                                  reused for EVERY model
//   T[m][k]  = logp_m[longest_suffix_m(history of k)][symbol of k]               -- the tier-1 dense table at depth D
//   head     the terms t in [order_m, D) of models shallower than D              -- a few direct look-ups
//
// The scan kernel performs L random 8-byte shared-memory gathers per (s, m); it is bound by shared-memory
// bank conflicts (6.4 wavefronts per warp-wide LDS.64, profiles/).  The contraction replaces them by
// 4^(D+1) fused multiply-adds per (s, m) from conflict-free operand tiles, which is cheaper on B200 whenever
// 4^(D+1) is below ~10 L (64 fp64 FMA lanes per SM vs ~4.6 gathers per SM per clock).  This file holds the fp64
// forms (mma.sync DMMA and a plain-FMA kernel): the counts are exact small integers but ln p needs all 53 bits, so no
// floating-point tensor-core format applies.  The default is the exact int8 tensor-core form of vgs_kmer_i8.cu (ln p
// as fixed point in six to eight base-256 digits); the fp64 kernels remain as its fallback and cross-check.
// Summation order differs from the reference's left-to-right order (parity: <= 1e-12 relative on scores);
// the bit-exact path remains VGS_NLL_SEQUENTIAL.
#include <stdlib.h>

#include <algorithm>

#include "vgs_i8gemm.cuh"

#define KG_BLOCK 128   // scores per CTA tile edge (sequences x models)
#define KG_BK 8        // k-mers per shared-memory stage
#define KG_THREADS 256 // 16 x 16 threads, 8 x 8 scores each

// ---- uniform-depth dense tables: Tu[m][k], NaN (no matching context) replaced by 0 and flagged ------------
__global__ void k_kmer_tables(const ModelMeta *__restrict__ meta, const uint2 *__restrict__ hash, const double *__restrict__ logp,
                              int depth, int64_t stride, double *__restrict__ Tu, int *__restrict__ nan_flag) {
    const ModelMeta mm = meta[blockIdx.y];
    const uint32_t n_hist = 1u << (2 * depth);
    const uint2 *slots = hash + mm.hash_off;
    double *out = Tu + (int64_t)blockIdx.y * stride;
    for (uint32_t hc = blockIdx.x * blockDim.x + threadIdx.x; hc < n_hist; hc += gridDim.x * blockDim.x) {
        const int id = vgs_longest_suffix(slots, mm.hash_mask, mm.hash_shift, mm.order, hc, depth);
        double4 v = make_double4(0.0, 0.0, 0.0, 0.0);
        if (id >= 0) {
            const double *r = logp + 4 * (mm.ctx_off + id);
            v = make_double4(r[0], r[1], r[2], r[3]);
        } else {
            atomicOr(nan_flag, 1);
        }
        double *o = out + 4 * (size_t)hc;
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
}

// ---- per-sequence (D+1)-mer histogram, one CTA per sequence, shared-memory atomics -----------------------
__global__ void k_kmer_hist(int depth, const uint32_t *__restrict__ seq, const int64_t *__restrict__ seq_off,
                            const int64_t *__restrict__ seq_len, int64_t stride, double *__restrict__ C) {
    extern __shared__ uint32_t hist[];
    const int64_t s = blockIdx.x;
    const int64_t K = (int64_t)1 << (2 * (depth + 1));
    for (int64_t k = threadIdx.x; k < K; k += blockDim.x) hist[k] = 0u;
    __syncthreads();
    const int64_t L = seq_len[s];
    const uint32_t *w = seq + seq_off[s];
    const uint32_t kmask = vgs_lowmask(depth + 1);
    const int64_t span = (L + blockDim.x - 1) / blockDim.x;
    const int64_t a = (int64_t)threadIdx.x * span, e = min(L, a + span);
    if (a < e) {
        int64_t t = max((int64_t)0, a - depth);
        uint32_t code = 0;
        for (; t < e; ++t) {
            const uint32_t x = (w[t >> 4] >> (30 - 2 * (int)(t & 15))) & 3u;
            code = ((code << 2) | x) & kmask;
            if (t >= a && t >= depth) atomicAdd(&hist[code], 1u);
        }
    }
    __syncthreads();
    double *out = C + s * stride;
    for (int64_t k = threadIdx.x; k < stride; k += blockDim.x) out[k] = k < K ? (double)hist[k] : 0.0;
}

// ---- fp64 contraction -----------------------------------------------------------------------------------
struct KgUnit {
    int32_t sb, mb;      // 128-blocks of sequences / models
    int32_t k0, k1;      // k range [k0, k1), multiples of KG_BK
    int32_t slot;        // partial-sum slot (block pair index * split + split index)
};

struct KgParams {
    const double *C;     // [n_seq][K]
    const double *Tu;    // [n_models][K]
    const KgUnit *units;
    int n_units;
    int64_t K, n_seq, n_models;
    double *partial;     // [slots][128][128]
    int *counter;
};

__global__ void __launch_bounds__(KG_THREADS, 1) k_kmer_gemm(const KgParams p) {
    __shared__ __align__(16) double As[2][KG_BK][KG_BLOCK];
    __shared__ __align__(16) double Bs[2][KG_BK][KG_BLOCK];
    __shared__ int s_unit;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;            // 16 x 16 thread grid
    // global -> shared loader: thread t moves KG_BK/2 consecutive k of one row of each operand
    const int lrow = tid & 127, lk = (tid >> 7) * (KG_BK / 2);
    for (;;) {
        __syncthreads();
        if (tid == 0) s_unit = atomicAdd(p.counter, 1);
        __syncthreads();
        const int u = s_unit;
        if (u >= p.n_units) break;
        const KgUnit un = p.units[u];
        const int64_t srow = (int64_t)un.sb * KG_BLOCK + lrow, mrow = (int64_t)un.mb * KG_BLOCK + lrow;
        const bool s_ok = srow < p.n_seq, m_ok = mrow < p.n_models;
        const double *ga = p.C + (s_ok ? srow : 0) * p.K + lk;
        const double *gb = p.Tu + (m_ok ? mrow : 0) * p.K + lk;
        double acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
        double2 ra[KG_BK / 4], rb[KG_BK / 4];
        // prologue: first stage
        {
            const int k = un.k0;
#pragma unroll
            for (int q = 0; q < KG_BK / 4; ++q) {
                ra[q] = s_ok ? __ldg((const double2 *)(ga + k) + q) : make_double2(0.0, 0.0);
                rb[q] = m_ok ? __ldg((const double2 *)(gb + k) + q) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int q = 0; q < KG_BK / 4; ++q) {
                As[0][lk + 2 * q][lrow] = ra[q].x; As[0][lk + 2 * q + 1][lrow] = ra[q].y;
                Bs[0][lk + 2 * q][lrow] = rb[q].x; Bs[0][lk + 2 * q + 1][lrow] = rb[q].y;
            }
        }
        __syncthreads();
        int buf = 0;
        for (int k = un.k0; k < un.k1; k += KG_BK) {
            const bool more = k + KG_BK < un.k1;
            if (more) {
#pragma unroll
                for (int q = 0; q < KG_BK / 4; ++q) {
                    ra[q] = s_ok ? __ldg((const double2 *)(ga + k + KG_BK) + q) : make_double2(0.0, 0.0);
                    rb[q] = m_ok ? __ldg((const double2 *)(gb + k + KG_BK) + q) : make_double2(0.0, 0.0);
                }
            }
#pragma unroll
            for (int kk = 0; kk < KG_BK; ++kk) {
                // thread (ty, tx) owns rows {16 i2 + 2 ty + r} and columns {16 j2 + 2 tx + r}: every 128-bit shared
                // load of a quarter-warp covers 128 contiguous bytes (no bank conflicts)
                double a[8], b[8];
#pragma unroll
                for (int i2 = 0; i2 < 4; ++i2) {
                    const double2 v = *(const double2 *)&As[buf][kk][16 * i2 + 2 * ty];
                    a[2 * i2] = v.x; a[2 * i2 + 1] = v.y;
                    const double2 w = *(const double2 *)&Bs[buf][kk][16 * i2 + 2 * tx];
                    b[2 * i2] = w.x; b[2 * i2 + 1] = w.y;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = __fma_rn(a[i], b[j], acc[i][j]);
            }
            if (more) {
#pragma unroll
                for (int q = 0; q < KG_BK / 4; ++q) {
                    As[buf ^ 1][lk + 2 * q][lrow] = ra[q].x; As[buf ^ 1][lk + 2 * q + 1][lrow] = ra[q].y;
                    Bs[buf ^ 1][lk + 2 * q][lrow] = rb[q].x; Bs[buf ^ 1][lk + 2 * q + 1][lrow] = rb[q].y;
                }
            }
            __syncthreads();
            buf ^= 1;
        }
        double *out = p.partial + (int64_t)un.slot * KG_BLOCK * KG_BLOCK;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 16 * (i >> 1) + 2 * ty + (i & 1);
#pragma unroll
            for (int j2 = 0; j2 < 4; ++j2) {
                const int c = 16 * j2 + 2 * tx;
                *(double2 *)&out[r * KG_BLOCK + c] = make_double2(acc[i][2 * j2], acc[i][2 * j2 + 1]);
            }
        }
    }
}

// ---- the same contraction on the fp64 tensor-core path (mma.sync m8n8k4 f64, SASS DMMA) ---------------------
// The FMA kernel above needs 16 shared-memory doubles per 64 FMAs per lane and ends up L1/shared-bound (41 % of
// the fp64 peak measured).  With 8x8x4 fp64 MMA tiles a warp computes a 64 x 32 block from 8 + 4 fragment loads
// per k4-step (32 MMAs, 8192 FMAs): ~8x fewer shared-memory wavefronts per FMA.  Operand tiles keep the global
// layout (row-major, k contiguous) in k4-slices [slice][row][4], so a half-warp fragment load is 128 contiguous
// bytes (conflict free) and the global->shared copies are plain 16-byte cp.async (LDGSTS), 3 stages deep.
#ifndef KD_BK
#define KD_BK 16      // k-mers per pipeline stage (32 measured 2 % slower: 253 registers, same pipe utilisation)
#endif
#define KD_STAGES 3
#define KD_STAGE_DOUBLES (2 * KD_BK * KG_BLOCK)   // A tile + B tile of one stage

__device__ __forceinline__ void vgs_dmma(double &d0, double &d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void vgs_cp_async16(void *smem_dst, const void *gmem_src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(vgs_smem_addr(smem_dst)), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void vgs_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void vgs_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(KG_THREADS, 1) k_kmer_dmma(const KgParams p) {
    extern __shared__ __align__(128) double dsm[];   // [KD_STAGES][A: 4 slices x 128 rows x 4 | B: same]
    __shared__ int s_unit;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ws = (warp >> 2) * 64, wm = (warp & 3) * 32;   // warp tile origin: 64 sequences x 32 models
    const int frow = lane >> 2, fk = lane & 3;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_unit = atomicAdd(p.counter, 1);
        __syncthreads();
        const int u = s_unit;
        if (u >= p.n_units) break;
        const KgUnit un = p.units[u];
        const int n_steps = (un.k1 - un.k0) / KD_BK;

        auto issue = [&](int step) {
            double *st = dsm + (size_t)(step % KD_STAGES) * KD_STAGE_DOUBLES;
            const int64_t k = (int64_t)un.k0 + (int64_t)step * KD_BK;
            constexpr int CPR = KD_BK / 2;                // 16-byte chunks per operand row and stage
#pragma unroll
            for (int j = 0; j < KG_BLOCK * CPR / KG_THREADS; ++j) {
                const int c = tid + KG_THREADS * j;
                const int row = c / CPR, ck = c % CPR;    // chunk ck covers k = 2 ck, 2 ck + 1
                const int dst = ((ck >> 1) * KG_BLOCK + row) * 4 + (ck & 1) * 2;
                const int64_t srow = (int64_t)un.sb * KG_BLOCK + row, mrow = (int64_t)un.mb * KG_BLOCK + row;
                const bool s_ok = srow < p.n_seq, m_ok = mrow < p.n_models;
                vgs_cp_async16(st + dst, p.C + (s_ok ? srow : 0) * p.K + k + 2 * ck, s_ok ? 16 : 0);
                vgs_cp_async16(st + KD_BK * KG_BLOCK + dst, p.Tu + (m_ok ? mrow : 0) * p.K + k + 2 * ck, m_ok ? 16 : 0);
            }
        };

        double acc[8][4][2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

#pragma unroll
        for (int st = 0; st < KD_STAGES - 1; ++st) {
            if (st < n_steps) issue(st);
            vgs_cp_async_commit();
        }
        for (int step = 0; step < n_steps; ++step) {
            vgs_cp_async_wait<KD_STAGES - 2>();
            __syncthreads();
            if (step + KD_STAGES - 1 < n_steps) issue(step + KD_STAGES - 1);
            vgs_cp_async_commit();
            const double *A = dsm + (size_t)(step % KD_STAGES) * KD_STAGE_DOUBLES;
            const double *B = A + KD_BK * KG_BLOCK;
#pragma unroll
            for (int sl = 0; sl < KD_BK / 4; ++sl) {
                double a[8], b[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = A[(sl * KG_BLOCK + ws + 8 * i + frow) * 4 + fk];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = B[(sl * KG_BLOCK + wm + 8 * j + frow) * 4 + fk];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) vgs_dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
        vgs_cp_async_wait<0>();
        double *out = p.partial + (int64_t)un.slot * KG_BLOCK * KG_BLOCK;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = ws + 8 * i + frow, c = wm + 8 * j + 2 * fk;
                *(double2 *)&out[r * KG_BLOCK + c] = make_double2(acc[i][j][0], acc[i][j][1]);
            }
    }
}

// ---- epilogue: fixed-order sum of the k-splits + head terms of models shallower than D ---------------------
struct KgEpi {
    const double *partial;
    const int32_t *pairs;   // [n_pairs][2] (sb, mb)
    int n_pairs, split, depth, skip_diag;
    const ModelMeta *meta;
    const uint2 *hash;
    const double *logp;
    const uint32_t *seq;
    const int64_t *seq_off, *seq_len;
    int64_t n_seq, n_models, ss, sm, m_col0;   // M[s * ss + (m_col0 + m) * sm]
    double *M;
};

// terms t in [order_m, D) of a model shallower than D (positions the (D+1)-mer histogram does not cover)
__device__ __forceinline__ double kmer_head(const KgEpi &p, const ModelMeta &mm, int64_t s) {
    if (mm.order >= p.depth) return 0.0;
    const uint2 *slots = p.hash + mm.hash_off;
    const uint32_t omask = vgs_lowmask(mm.order);
    const uint32_t *w = p.seq + p.seq_off[s];
    const int64_t L = p.seq_len[s];
    const int64_t top = L < p.depth ? L : p.depth;
    uint32_t hist = 0;
    double head = 0.0;
    for (int64_t t = 0; t < top; ++t) {
        const uint32_t x = (w[t >> 4] >> (30 - 2 * (int)(t & 15))) & 3u;
        if (t >= mm.order) {
            const int id = vgs_longest_suffix(slots, mm.hash_mask, mm.hash_shift, mm.order, hist & omask, mm.order);
            head += id < 0 ? vgs_nan() : p.logp[4 * (mm.ctx_off + id) + x];
        }
        hist = (hist << 2) | x;
    }
    return head;
}

// diagonal scores M[m][m] as plain dot products (one warp each, fixed summation order).  The distance tiles need
// M[i][i] of every row and column they touch; computing those few entries here instead of whole diagonal blocks
// keeps a rank's work equal to its share, and makes M[i][i] the same number whatever the sharding.
__global__ void k_kmer_diag(const KgEpi p, const double *__restrict__ C, const double *__restrict__ Tu, int64_t K) {
    __shared__ double part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t m = blockIdx.x;   // one 256-thread block per model: 2 K doubles streamed with 16-byte loads
    const double *c = C + (p.m_col0 + m) * K, *t = Tu + m * K;
    double a0 = 0.0, a1 = 0.0;
    for (int64_t k = 2 * threadIdx.x; k < K; k += 2 * blockDim.x) {
        const double2 cv = *(const double2 *)(c + k), tv = *(const double2 *)(t + k);
        a0 = __fma_rn(cv.x, tv.x, a0);
        a1 = __fma_rn(cv.y, tv.y, a1);
    }
    double v = a0 + a1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) part[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
        p.M[(p.m_col0 + m) * p.ss + (p.m_col0 + m) * p.sm] = kmer_head(p, p.meta[m], p.m_col0 + m) + s;
    }
}

__global__ void k_kmer_epilogue(const KgEpi p) {
    const int pair = blockIdx.x;
    const int e_begin = blockIdx.y * (KG_BLOCK * KG_BLOCK / 8), e_end = e_begin + KG_BLOCK * KG_BLOCK / 8;
    const int sb = p.pairs[2 * pair], mb = p.pairs[2 * pair + 1];
    for (int e = e_begin + threadIdx.x; e < e_end; e += blockDim.x) {
        const int r = e / KG_BLOCK, c = e % KG_BLOCK;
        const int64_t s = (int64_t)sb * KG_BLOCK + r, m = (int64_t)mb * KG_BLOCK + c;
        if (s >= p.n_seq || m >= p.n_models) continue;
        if (p.skip_diag && s == p.m_col0 + m) continue;  // written by k_kmer_diag
        double v = 0.0;
        for (int q = 0; q < p.split; ++q) v += p.partial[((int64_t)pair * p.split + q) * KG_BLOCK * KG_BLOCK + e];
        p.M[s * p.ss + (p.m_col0 + m) * p.sm] = kmer_head(p, p.meta[m], s) + v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------
bool vgs_kmer_applicable_orders(vgs_context *h) { return h->max_nll_bytes_t2 == 0 && h->max_nll_bytes_t3 == 0 && h->max_order <= VGS_DENSE_MAX_ORDER; }
bool vgs_kmer_applicable(vgs_context *h) {
    if (h->max_nll_bytes_t2 > 0 || h->max_nll_bytes_t3 > 0 || h->max_order > VGS_DENSE_MAX_ORDER) return false;
    return true;
}

static int ensure_kmer_tables(vgs_context *h, cudaStream_t s) {
    const int D = h->max_order;
    const int64_t K = std::max<int64_t>(KD_BK, (int64_t)1 << (2 * (D + 1)));  // row stride, padded to one MMA k-stage
    if (h->kmer_tables_ready) return VGS_OK;
    int rc;
    if ((rc = vgs_ensure_exact_logp(h, s))) return rc;
    if ((rc = vgs_reserve(h, &h->d_kmer_T, h->n_models * K))) return rc;
    if (K != ((int64_t)1 << (2 * (D + 1)))) VGS_CUDA(cudaMemsetAsync(h->d_kmer_T, 0, sizeof(double) * (size_t)(h->n_models * K), s));
    if ((rc = vgs_reserve(h, &h->d_counters, 16))) return rc;
    VGS_CUDA(cudaMemsetAsync(h->d_counters + 8, 0, sizeof(int), s));
    const uint32_t n_hist = 1u << (2 * D);
    for (int64_t m0 = 0; m0 < h->n_models; m0 += 65535) {
        dim3 grid((unsigned)std::min<uint32_t>((n_hist + 255) / 256, 16), (unsigned)std::min<int64_t>(h->n_models - m0, 65535));
        k_kmer_tables<<<grid, 256, 0, s>>>(h->d_meta + m0, h->d_hash, h->d_logp, D, K, h->d_kmer_T + m0 * K, h->d_counters + 8);
        VGS_LAUNCH_CHECK(h);
    }
    int flag = 0;
    VGS_CUDA(cudaMemcpyAsync(&flag, h->d_counters + 8, sizeof(int), cudaMemcpyDeviceToHost, s));
    VGS_CUDA(cudaStreamSynchronize(s));
    h->kmer_has_nan = flag != 0;
    h->kmer_tables_ready = true;
    return VGS_OK;
}

// the exact int8 tensor-core contraction (vgs_kmer_i8.cu) for the needed block pairs: per 256-sequence block the runs
// of needed 128-model blocks become rectangles, cut into GEMM units (cached on the device per cache_key).
// Returns VGS_OK, an error, or -1 = not applicable (caller continues with the fp64 path or the scan)
static int kmer_i8_path(vgs_context *h, const std::vector<char> &need, int64_t nb_s, int64_t nb_m, bool diag_separately,
                        const VgsScoreDst &dst, cudaStream_t s, uint64_t cache_key) {
    int rc = vgs_kmer_i8_prepare_tables(h, s);
    if (rc != VGS_OK) return rc;
    const int nd = h->i8_nd, gran = vgs_kmer_i8_gran(h);   // digit rows per model; granule of the unit widths
    const uint64_t key = cache_key ? vgs_mix(cache_key, 0x6938 + (uint64_t)nd) : 0;
    I8Work w;
    bool diag_covered = false;
    if (WorkCache *wc = vgs_wcache_find(h, key)) {
        w.n_units = wc->n[0]; w.n_rects = wc->n[1]; diag_covered = wc->n[2] != 0; w.split = wc->split; w.max_rect_models = wc->n[3];
        w.d_units = (const G8Unit *)wc->dev;
        w.d_rects = (const G8Rect *)((char *)wc->dev + wc->off2);
    } else {
        std::vector<G8Rect> rects;
        const int64_t nb_a = (nb_s + 1) / 2;   // 256-sequence blocks
        const int64_t rows_total = (h->n_models * nd + gran - 1) / gran * gran;
        for (int64_t a = 0; a < nb_a; ++a) {
            int64_t mb = 0;
            while (mb < nb_m) {
                auto needed = [&](int64_t b) {
                    return need[(size_t)(2 * a * nb_m + b)] || (2 * a + 1 < nb_s && need[(size_t)((2 * a + 1) * nb_m + b)]);
                };
                if (!needed(mb)) { ++mb; continue; }
                int64_t e = mb;
                while (e < nb_m && needed(e)) ++e;
                G8Rect r;
                r.a_blk = (int32_t)a;
                r.b_row0 = (int32_t)(mb * 128 * nd);   // 128 models x nd digit rows: a whole number of granules
                r.b_rows = (int32_t)(std::min<int64_t>(e * 128 * nd, rows_total) - r.b_row0);
                if (r.b_rows > 0) rects.push_back(r);
                mb = e;
            }
        }
        diag_covered = true;
        for (int64_t a = 0; a < std::min(nb_s, nb_m); ++a) diag_covered = diag_covered && need[(size_t)(a * nb_m + a)];
        std::vector<G8Unit> units;
        const int kblocks = (int)(std::max<int64_t>(G8_BK, (int64_t)1 << (2 * (h->max_order + 1))) / G8_BK);
        static int max_split = -1;
        if (max_split < 0) { const char *e = getenv("VGS_I8_SPLITK"); max_split = e ? std::max(1, atoi(e)) : 8; }
        // the int32 partial-sum buffer of the k-split stays below 1 GB
        const int ms = (h->n_seq * h->n_models * 4 * nd * max_split <= ((int64_t)2 << 30)) ? max_split : 1;
        w.split = g8_make_units(rects, kblocks, h->sm_count / 2, ms, gran, &units);
        w.n_units = (int)units.size();
        w.n_rects = (int)rects.size();
        w.max_rect_models = 0;
        for (const G8Rect &r : rects) w.max_rect_models = std::max(w.max_rect_models, (r.b_rows + nd - 1) / nd);
        const int64_t off2 = ((int64_t)units.size() * (int64_t)sizeof(G8Unit) + 127) / 128 * 128;
        const int64_t bytes = off2 + (int64_t)rects.size() * (int64_t)sizeof(G8Rect) + 256;
        char *base = nullptr;
        if (key) {
            WorkCache *wc = vgs_wcache_add(h, key, bytes);
            if (wc) {
                wc->n[0] = w.n_units; wc->n[1] = w.n_rects; wc->n[2] = diag_covered ? 1 : 0; wc->n[3] = w.max_rect_models; wc->split = w.split; wc->off2 = off2;
                base = (char *)wc->dev;
            }
        }
        if (!base) {
            if ((rc = vgs_ensure_scratch2(h, bytes))) return rc;
            base = (char *)h->d_scratch2;
        }
        w.d_units = (const G8Unit *)base;
        w.d_rects = (const G8Rect *)(base + off2);
        if (!units.empty()) VGS_CUDA(cudaMemcpyAsync(base, units.data(), units.size() * sizeof(G8Unit), cudaMemcpyHostToDevice, s));
        if (!rects.empty()) VGS_CUDA(cudaMemcpyAsync(base + off2, rects.data(), rects.size() * sizeof(G8Rect), cudaMemcpyHostToDevice, s));
        // pageable sources: the copies above have consumed the host vectors when they return
    }
    if (w.n_units == 0 && !diag_separately) return VGS_OK;
    // exact integers: the GEMM's own M[i][i] has the same bits as the dot-product kernel's, so that kernel is skipped
    // whenever the rectangles already cover the diagonal blocks (always on one GPU and in model-sharded runs)
    return vgs_kmer_i8_scores(h, w, diag_separately && !diag_covered, dst, s);
}

// scores for the needed (sequence-block, model-block) pairs at 128-granularity; need[sb * nb_m + mb] != 0
int vgs_kmer_scores(vgs_context *h, const std::vector<char> &need, int64_t nb_s, int64_t nb_m, bool diag_separately,
                    const VgsScoreDst &dst, cudaStream_t s, uint64_t cache_key) {
    int rc;
    // 1. the exact int8 tensor-core path when the set qualifies; VGS_KMER_GEMM=dmma|fma keeps fp64
    {
        static int want_i8 = -1;
        if (want_i8 < 0) { const char *e = getenv("VGS_KMER_GEMM"); want_i8 = (e && (e[0] == 'd' || e[0] == 'f')) ? 0 : 1; }
        if (want_i8) {
            rc = kmer_i8_path(h, need, nb_s, nb_m, diag_separately, dst, s, cache_key);
            if (rc != -1) return rc;
            if (!h->i8_tables_ok) return -1;   // a model without root context: the scan raises the reference's error
        }
    }
    // 2. fp64 path (more than 64 frequent k-mers in a sequence, or forced): dense fp64 tables + fp64 histograms
    rc = ensure_kmer_tables(h, s);
    if (rc) return rc;
    if (h->kmer_has_nan) return -1;  // caller falls back to the scan kernel
    double *M = dst.M[0];
    const int D = h->max_order;
    const int64_t K = std::max<int64_t>(KD_BK, (int64_t)1 << (2 * (D + 1)));
    int n_pairs = 0, split = 1, n_units = 0;
    KgUnit *d_units = nullptr;
    int32_t *d_pairs = nullptr;
    if (WorkCache *wc = vgs_wcache_find(h, cache_key)) {
        n_pairs = wc->n[0]; n_units = wc->n[1]; split = wc->split;
        d_units = (KgUnit *)wc->dev;
        d_pairs = (int32_t *)((char *)wc->dev + wc->off2);
    } else {
        std::vector<int32_t> pairs;
        for (int64_t sb = 0; sb < nb_s; ++sb)
            for (int64_t mb = 0; mb < nb_m; ++mb)
                if (need[(size_t)(sb * nb_m + mb)]) { pairs.push_back((int32_t)sb); pairs.push_back((int32_t)mb); }
        n_pairs = (int)(pairs.size() / 2);
        if (n_pairs > 0) {
            const int64_t ksteps = K / KD_BK;   // unit boundaries are multiples of one MMA stage
            // split K so that the FULL problem (every block pair) gives >= ~8 units per SM.  The split depends only on the
            // problem shape, never on which block pairs this call computes, so every entry is summed in the same order
            // however the tile space is sharded (1/2/4/8-rank matrices stay bit-identical).
            split = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(16, ksteps),
                                                                          (8 * 148 + nb_s * nb_m - 1) / (nb_s * nb_m)));
            std::vector<KgUnit> units;
            for (int pi = 0; pi < n_pairs; ++pi)
                for (int q = 0; q < split; ++q) {
                    KgUnit un;
                    un.sb = pairs[2 * pi]; un.mb = pairs[2 * pi + 1];
                    un.k0 = (int32_t)(ksteps * q / split * KD_BK);
                    un.k1 = (int32_t)(ksteps * (q + 1) / split * KD_BK);
                    un.slot = pi * split + q;
                    units.push_back(un);
                }
            n_units = (int)units.size();
            const int64_t off2 = ((int64_t)units.size() * (int64_t)sizeof(KgUnit) + 127) / 128 * 128;
            const int64_t meta_bytes = off2 + (int64_t)pairs.size() * 4 + 256;
            char *base = nullptr;
            if (cache_key) {
                WorkCache *wc = vgs_wcache_add(h, cache_key, meta_bytes);
                if (wc) {
                    wc->n[0] = n_pairs; wc->n[1] = n_units; wc->split = split; wc->off2 = off2;
                    base = (char *)wc->dev;
                }
            }
            if (!base) {
                if ((rc = vgs_ensure_scratch2(h, meta_bytes))) return rc;
                base = (char *)h->d_scratch2;
            }
            d_units = (KgUnit *)base;
            d_pairs = (int32_t *)(base + off2);
            VGS_CUDA(cudaMemcpyAsync(d_units, units.data(), units.size() * sizeof(KgUnit), cudaMemcpyHostToDevice, s));
            VGS_CUDA(cudaMemcpyAsync(d_pairs, pairs.data(), pairs.size() * 4, cudaMemcpyHostToDevice, s));
        }
    }
    if (n_pairs == 0 && !diag_separately) return VGS_OK;
    // fp64 histograms of the resident sequences (a pass over the packed sequences; part of every call unless the caller
    // pinned them with VGS_KMER_KEEP_COUNTS=1)
    static int keep_counts = -1;
    if (keep_counts < 0) { const char *e = getenv("VGS_KMER_KEEP_COUNTS"); keep_counts = (e && e[0] == '1') ? 1 : 0; }
    if (!h->kmer_counts_ready || !keep_counts) {
        if ((rc = vgs_reserve(h, &h->d_kmer_C, h->n_seq * K))) return rc;
        const int smem = (int)(((int64_t)1 << (2 * (D + 1))) * 4);
        VGS_CUDA(cudaFuncSetAttribute(k_kmer_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k_kmer_hist<<<(unsigned)h->n_seq, 256, smem, s>>>(D, h->d_seq, h->d_seq_off, h->d_seq_len, K, h->d_kmer_C);
        VGS_LAUNCH_CHECK(h);
        h->kmer_counts_ready = true;
    }
    if (n_pairs > 0) {
        const int64_t part_elems = (int64_t)n_pairs * split * KG_BLOCK * KG_BLOCK;
        if ((rc = vgs_reserve(h, &h->d_kmer_partial, part_elems))) return rc;
        if ((rc = vgs_reserve(h, &h->d_counters, 16))) return rc;
        VGS_CUDA(cudaMemsetAsync(h->d_counters + 2, 0, sizeof(int), s));
        KgParams p;
        p.C = h->d_kmer_C; p.Tu = h->d_kmer_T; p.units = d_units; p.n_units = n_units;
        p.K = K; p.n_seq = h->n_seq; p.n_models = h->n_models; p.partial = h->d_kmer_partial; p.counter = h->d_counters + 2;
        const int grid = std::min<int>(n_units, h->sm_count);
        static int use_fma = -1;
        if (use_fma < 0) { const char *e = getenv("VGS_KMER_GEMM"); use_fma = (e && e[0] == 'f') ? 1 : 0; }
        vgs_prof_begin(h, VGS_PROF_NLL_SCORES, s);
        h->last_nll_kernel = use_fma ? VGS_NLL_KERNEL_FMA : VGS_NLL_KERNEL_DMMA;
        if (use_fma) {
            k_kmer_gemm<<<grid, KG_THREADS, 0, s>>>(p);
        } else {
            const int smem = KD_STAGES * KD_STAGE_DOUBLES * 8;
            VGS_CUDA(cudaFuncSetAttribute(k_kmer_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            k_kmer_dmma<<<grid, KG_THREADS, smem, s>>>(p);
        }
        vgs_prof_end(h, VGS_PROF_NLL_SCORES, s);
        VGS_LAUNCH_CHECK(h);
    }
    KgEpi e;
    e.partial = h->d_kmer_partial; e.pairs = d_pairs; e.n_pairs = n_pairs; e.split = split; e.depth = D;
    e.meta = h->d_meta; e.hash = h->d_hash; e.logp = h->d_logp; e.seq = h->d_seq; e.seq_off = h->d_seq_off; e.seq_len = h->d_seq_len;
    e.n_seq = h->n_seq; e.n_models = h->n_models; e.ss = dst.stride_s; e.sm = dst.stride_m; e.m_col0 = dst.m_col0; e.M = M;
    e.skip_diag = diag_separately ? 1 : 0;
    if (n_pairs > 0) {
        k_kmer_epilogue<<<dim3((unsigned)n_pairs, 8), 256, 0, s>>>(e);
        VGS_LAUNCH_CHECK(h);
    }
    if (diag_separately) {
        const int64_t nd = std::min(h->n_seq, h->n_models);
        k_kmer_diag<<<(unsigned)nd, 256, 0, s>>>(e, h->d_kmer_C, h->d_kmer_T, K);
        VGS_LAUNCH_CHECK(h);
    }
    return VGS_OK;
}
