// vgs_kmer_i8.cu -- the NLL k-mer contraction as an EXACT integer GEMM on the 5th-generation tensor cores
// (tcgen05.mma kind::i8, CTA pairs, accumulators in tensor memory; GEMM core: vgs_i8gemm.cuh).
//
// Same quantity as vgs_kmer.cu:  M[s][m] = sum_k C[s][k] * T[m][k] + head(s, m)   (src/vlmc/vlmc.pyx:79-101 restated
// as a contraction).  ln p needs 53 bits, so no floating-point tensor-core format can carry T --
// but the contraction is a sum of (small integer) x (fixed-point number), and that is exact in integer arithmetic:
//
//   T[m][k]  ->  q = round(T * 2^51)  (|T| <= 1000 < 2^10, so |q| < 2^61), written as eight balanced base-256 digits
//                q = sum_p d_p * 256^p,  d_p in [-128, 127]                      -- int8 "B" operand, 8 rows per model
//   C[s][k]  ->  the low byte of the k-mer count                                   -- uint8 "A" operand
//   S_p[s][m] = sum_k (C[s][k] & 255) * d_p[m][k]   int32, exact: |S_p| <= 16384 * 255 * 128 < 2^31
//               + 256 * (C[s][k] >> 8) * d_p[m][k] over the few k-mers of a sequence that occur more than 255 times
//                 (listed by the histogram kernel; added as int64 where the digit sums are folded)
//   M[s][m]  = 2^-51 * sum_p S_p * 256^p   (two int64 halves, ONE fp64 rounding)  + head(s, m)
//
// The only error is the 2^-52 quantisation of each ln p: |dM| <= L * 2^-52 ~ 2e-12 absolute on |M| ~ 1e4, tighter than
// any fp64 summation order -- and every (s, m) entry is the same bits whatever the tiling, the k-split or the number of
// GPUs (integer sums do not depend on the order), so sharded matrices stay identical and duplicates score exactly
// alike.  Sets with more than 64 frequent k-mers in one sequence (or with a model without root context) take the fp64
// MMA path of vgs_kmer.cu.
//
// The digit operand is built straight from the models (hash + host-libm ln p): no fp64 table, no per-model dense NLL
// table is needed for this mode.  Both operands live in HBM in the GEMM core's slab layout (vgs_i8gemm.cuh).
#include <stdlib.h>

#include <algorithm>

#include "vgs_i8gemm.cuh"

#define I8_SCALE_BITS 51
#define I8_MAX_OVER 64       // per sequence: k-mers whose count exceeds one byte

// the eight exact digit sums of one (sequence, model) -> fp64
__device__ __forceinline__ double i8_fold(const long long *S) {
    const long long lo = S[0] + S[1] * 256 + S[2] * 65536 + S[3] * 16777216;
    const long long hi = S[4] + S[5] * 256 + S[6] * 65536 + S[7] * 16777216;
    const double v = __fma_rn((double)hi, 4294967296.0, (double)lo);
    return v * (1.0 / 2251799813685248.0);   // 2^-51, exact
}

// fixed-point form of one table entry (the same rounding everywhere)
__device__ __forceinline__ long long i8_quantise(double t) { return __double2ll_rn(t * 2251799813685248.0); }

// what the contraction needs to know about the models and sequences besides the two operands
struct I8Side {
    const ModelMeta *meta;
    const uint2 *hash;
    const double *logp;
    const uint32_t *seq;
    const int64_t *seq_off, *seq_len;
    const int32_t *over_cnt;   // [n_seq] frequent k-mers per sequence (nullptr: the set has none)
    const uint2 *over_list;    // [n_seq][I8_MAX_OVER] {k, count >> 8}
    const int8_t *digits;      // the digit operand (slab layout): the frequent k-mers' digits are read back from it
    int64_t d_rows_pad;
    int depth;                 // D = largest order of the set; k-mers are (D + 1)-mers
    int any_shallow;           // some model is shallower than D (only then does anything look at the per-model metadata)
};

// adds 256 * hi * d_p(T[m][k]) for the frequent k-mers of sequence s.  The digits come straight from the GEMM's own digit
// operand (eight independent one-byte loads per k-mer, from lines the tile has just streamed through L2) -- the same
// integers the tensor cores multiplied, so the sum stays exact.  (A first version re-derived them through the model's
// hash: seven dependent cache misses per model, 100 us for ONE sequence with ONE frequent k-mer at cfg 3.)
__device__ __forceinline__ void i8_add_frequent(long long *S, const I8Side &sd, int64_t m, int64_t s, int n_over) {
    const uint2 *list = sd.over_list + s * I8_MAX_OVER;
    for (int o = 0; o < n_over; ++o) {
        const uint2 e = list[o];
        const long long w = 256ll * (long long)e.y;
#pragma unroll
        for (int p = 0; p < 8; ++p) S[p] += w * (long long)sd.digits[g8_off(m * 8 + p, e.x, sd.d_rows_pad)];
    }
}

// terms t in [order_m, D) of a model shallower than D (positions the (D+1)-mer histogram does not cover)
__device__ __forceinline__ double i8_head(const I8Side &sd, const ModelMeta &mm, int64_t s) {
    if (mm.order >= sd.depth) return 0.0;
    const uint2 *slots = sd.hash + mm.hash_off;
    const uint32_t *w = sd.seq + sd.seq_off[s];
    const int64_t L = sd.seq_len[s];
    const int64_t top = L < sd.depth ? L : sd.depth;
    const uint32_t omask = vgs_lowmask(mm.order);
    uint32_t hist = 0;
    double head = 0.0;
    for (int64_t t = 0; t < top; ++t) {
        const uint32_t x = (w[t >> 4] >> (30 - 2 * (int)(t & 15))) & 3u;
        if (t >= mm.order) {
            const int id = vgs_longest_suffix(slots, mm.hash_mask, mm.hash_shift, mm.order, hist & omask, mm.order);
            head += id < 0 ? vgs_nan() : sd.logp[4 * (mm.ctx_off + id) + x];
        }
        hist = (hist << 2) | x;
    }
    return head;
}

// ---- operand builders ----------------------------------------------------------------------------------------
// digits[slab(m * 8 + p, k)] = p-th balanced base-256 digit of round(ln p_m(k) * 2^51); one thread per 16 k-mers
// (= 4 histories x 4 symbols) of one model.  flag bit 0: an entry is not a log-probability (|T| > 1000 or NaN);
// bit 1: a history has no context at all (model without root context: the scan raises the reference's RuntimeError)
__global__ void k_i8_digits(int64_t m_base, I8Side sd, int64_t K, int64_t K8, int64_t rows_pad, int8_t *__restrict__ digits,
                            int *__restrict__ flag) {
    const int64_t m = m_base + blockIdx.y;
    const ModelMeta mm = sd.meta[m];
    const uint2 *slots = sd.hash + mm.hash_off;
    for (int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ch < (K8 >> 4); ch += (int64_t)gridDim.x * blockDim.x) {
        long long q[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t hc = 4 * ch + i;   // history code
            double4 v = make_double4(0.0, 0.0, 0.0, 0.0);
            if (16 * ch + 4 * i < K) {
                const int id = vgs_longest_suffix(slots, mm.hash_mask, mm.hash_shift, mm.order, (uint32_t)hc, sd.depth);
                if (id >= 0) {
                    const double *r = sd.logp + 4 * (mm.ctx_off + id);
                    v = make_double4(r[0], r[1], r[2], r[3]);
                } else {
                    atomicOr(flag, 2);
                }
            }
            const double t4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                if (!(fabs(t4[x]) <= 1000.0)) { atomicOr(flag, 1); q[4 * i + x] = 0; }
                else q[4 * i + x] = i8_quantise(t4[x]);
            }
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const long long d = ((q[j] + 128) & 255) - 128;
                w[j >> 2] |= ((uint32_t)d & 255u) << (8 * (j & 3));
                q[j] = (q[j] - d) >> 8;
            }
            *(uint4 *)(digits + g8_off(m * 8 + p, 16 * ch, rows_pad)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

// per-sequence (D+1)-mer histogram, low bytes as the uint8 operand (one CTA per sequence); k-mers counted more than 255
// times go to the sequence's list {k, count >> 8}; more than I8_MAX_OVER of them raise flag bit 1; any at all: bit 0.
// PACK16: two 16-bit counters per shared-memory word (every sequence shorter than 65536): half the shared memory,
// twice the resident CTAs
template <bool PACK16>
__global__ void k_i8_hist(int depth, const uint32_t *__restrict__ seq, const int64_t *__restrict__ seq_off,
                          const int64_t *__restrict__ seq_len, int64_t K8, int64_t rows_pad, uint8_t *__restrict__ C8,
                          int32_t *__restrict__ over_cnt, uint2 *__restrict__ over_list, int *__restrict__ flag) {
    extern __shared__ uint32_t hist[];
    __shared__ int n_over;
    if (threadIdx.x == 0) n_over = 0;
    const int64_t s = blockIdx.x;
    const int64_t K = (int64_t)1 << (2 * (depth + 1));
    const int64_t n_words = (K < 16 ? 16 : K) / (PACK16 ? 2 : 1);   // at least one 16-count chunk
    for (int64_t k = threadIdx.x; k < n_words; k += blockDim.x) hist[k] = 0u;
    __syncthreads();
    const int64_t L = seq_len[s];
    const uint32_t *w = seq + seq_off[s];
    const uint32_t kmask = vgs_lowmask(depth + 1);
    // thread t takes the 16-symbol words t, t + blockDim, ...: the k-mers ENDING in that word (the word before it
    // supplies their first symbols), one coalesced 4-byte load per word.  Words that lie wholly inside [depth, L) -- all
    // but the first and the last -- skip the per-symbol range tests (the kernel is bound by instruction issue, ncu: ALU
    // pipe 64 %, 46 instructions per symbol before this split)
    const int64_t n_w = (L + 15) >> 4;
    for (int64_t wi = threadIdx.x; wi < n_w; wi += blockDim.x) {
        const uint32_t cur = __ldg(w + wi);
        const uint32_t prev = wi ? __ldg(w + wi - 1) : 0u;
        const int64_t t0 = wi << 4;
        if (t0 >= depth && t0 + 16 <= L) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t code = __funnelshift_r(cur, prev, 30 - 2 * j) & kmask;
                if (PACK16) atomicAdd(&hist[code >> 1], 1u << ((code & 1u) << 4));
                else atomicAdd(&hist[code], 1u);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int64_t t = t0 + j;
                if (t < L && t >= depth) {
                    const uint32_t code = __funnelshift_r(cur, prev, 30 - 2 * j) & kmask;
                    if (PACK16) atomicAdd(&hist[code >> 1], 1u << ((code & 1u) << 4));
                    else atomicAdd(&hist[code], 1u);
                }
            }
        }
    }
    __syncthreads();
    // one 16-byte k-chunk (16 counts) per thread and store, slab layout (K8 is a multiple of 128).  PACK16: the low bytes
    // of four 16-bit counters come out of two words with one byte permute, their high bytes with another -- non-zero only
    // for the rare counts above 255
    const int64_t n_chunks = K8 >> 4;
    for (int64_t ch = threadIdx.x; ch < n_chunks; ch += blockDim.x) {
        uint32_t o[4] = {0u, 0u, 0u, 0u};
        if (16 * ch < K) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (PACK16) {
                    const uint2 c2 = *(const uint2 *)&hist[8 * ch + 2 * q];
                    o[q] = __byte_perm(c2.x, c2.y, 0x6420);
                    const uint32_t high = __byte_perm(c2.x, c2.y, 0x7531);
                    if (high) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t hi = (high >> (8 * j)) & 255u;
                            if (hi) {
                                const int idx = atomicAdd(&n_over, 1);
                                if (idx < I8_MAX_OVER) over_list[s * I8_MAX_OVER + idx] = make_uint2((uint32_t)(16 * ch + 4 * q + j), hi);
                            }
                        }
                    }
                } else {
                    const uint4 c4 = *(const uint4 *)&hist[16 * ch + 4 * q];
                    const uint32_t cs[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (cs[j] > 255u) {
                            const int idx = atomicAdd(&n_over, 1);
                            if (idx < I8_MAX_OVER) over_list[s * I8_MAX_OVER + idx] = make_uint2((uint32_t)(16 * ch + 4 * q + j), cs[j] >> 8);
                        }
                        o[q] |= (cs[j] & 255u) << (8 * j);
                    }
                }
            }
        }
        *(uint4 *)(C8 + g8_off(s, 16 * ch, rows_pad)) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        over_cnt[s] = n_over < I8_MAX_OVER ? n_over : I8_MAX_OVER;
        if (n_over > 0) atomicOr(flag, n_over > I8_MAX_OVER ? 3 : 1);
    }
}

// ---- epilogue of the GEMM: 32 columns = 4 models x 8 digits of one sequence ------------------------------------
struct I8Out {
    double *M[VGS_MAX_PEERS];   // destination score matrices: this GPU's and (peer-mapped) every other rank's
    int n_dst;
    int64_t ss, sm;             // strides of the sequence / model index
    int64_t m_col0;             // column of local model 0 (model shards: local model m is column m_col0 + m)
    int64_t n_seq, n_models;    // local operand extents
    int skip_diag;
    int32_t *P;                 // k-split launches: exact int32 digit sums of every part, [part][model][digit][sequence]
    int64_t p_ld;               // sequences per row of P (padded to 32)
    int n_parts;
};

struct EpiNll {
    I8Out o;
    I8Side sd;
    __device__ __forceinline__ int frequent_count(int64_t s) const { return sd.over_cnt ? sd.over_cnt[s] : 0; }
    // S = the complete exact digit sums of (s, m)
    __device__ __forceinline__ void store_sums(int64_t s, int64_t m, const long long *S) const {
        double v = i8_fold(S);
        if (sd.any_shallow) v = i8_head(sd, sd.meta[m], s) + v;   // IEEE addition commutes: same bits as head + fold
        const int64_t at = s * o.ss + (o.m_col0 + m) * o.sm;
        for (int d = 0; d < o.n_dst; ++d) o.M[d][at] = v;
    }
    __device__ __forceinline__ void store(int64_t s, int64_t m, long long *S, int n_over) const {
        if (n_over) i8_add_frequent(S, sd, m, s, n_over);
        store_sums(s, m, S);
    }
    // peer-mapped destinations: the band must have left this GPU before the rank signals the peer barrier
    __device__ __forceinline__ void unit_done() const {
        if (o.n_dst > 1) __threadfence_system();
    }
    // whole warp, while the unit's MMAs still run: pull the frequent-k-mer lists and the digit lines operator() will read
    // for this chunk into L2 (prefetch only: no registers held, nothing depends on it)
    __device__ __forceinline__ void prefetch(const G8Unit &u, int64_t s, int c0) const {
        if (u.partial || !sd.over_cnt) return;
        const int lane = threadIdx.x & 31;
        const int n_over = s < o.n_seq ? sd.over_cnt[s] : 0;
        unsigned pending = __ballot_sync(0xffffffffu, n_over > 0);
        while (pending) {
            const int r = __ffs(pending) - 1;
            pending &= pending - 1;
            const long long s_r = __shfl_sync(0xffffffffu, (long long)s, r);
            const int n_r = __shfl_sync(0xffffffffu, n_over, r);
            const int64_t row = (int64_t)u.b_row0 + c0 + lane;   // digit row this lane reads in operator()
            const uint2 *list = sd.over_list + s_r * I8_MAX_OVER;
            for (int oi = 0; oi < n_r; ++oi) {
                const uint2 e = list[oi];
                asm volatile("prefetch.global.L2 [%0];" ::"l"(sd.digits + g8_off(row, e.x, sd.d_rows_pad)));
            }
        }
    }
    // called by all 32 lanes of an epilogue warp together (lane = sequence row s, v = 32 columns = 4 models x 8 digits)
    __device__ __forceinline__ void operator()(const G8Unit &u, int64_t s, int c0, const int32_t (&v)[32]) const {
        const bool live = s < o.n_seq;
        const int64_t m0 = (int64_t)(u.b_row0 + c0) >> 3;
        if (u.partial) {
            // part (u.partial - 1) of a k-split: its digit sums go to the part's own slab with plain stores, coalesced over
            // the warp's 32 sequences (a first version added them into one buffer with integer atomics: 8 M atomics at
            // 2 GPUs cost 110 us); k_i8_fold adds the parts -- integers, so the split changes no bit of the result
            if (!live) return;
            int32_t *dst = o.P + ((int64_t)(u.partial - 1) * o.n_models + m0) * 8 * o.p_ld + s;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (m0 + q >= o.n_models) break;
#pragma unroll
                for (int d = 0; d < 8; ++d) dst[(8 * q + d) * o.p_ld] = v[8 * q + d];
            }
            return;
        }
        // Rows with frequent k-mers are finished by the whole warp: lane l fetches digit l & 7 of model m0 + (l >> 3) --
        // exactly the column the row's owner holds in v[l] -- for every frequent k-mer of the row, and the owner picks
        // the 32 sums up by shuffle.  (Left to the owner alone, ONE such row at cfg 3 -- 112 dependent-latency loads --
        // added 18 us to the whole kernel.)
        const int lane = threadIdx.x & 31;
        const int n_over = live ? frequent_count(s) : 0;
        unsigned pending = __ballot_sync(0xffffffffu, n_over > 0);
        while (pending) {
            const int r = __ffs(pending) - 1;
            pending &= pending - 1;
            const long long s_r = __shfl_sync(0xffffffffu, (long long)s, r);
            const int n_r = __shfl_sync(0xffffffffu, n_over, r);
            const int64_t m_l = m0 + (lane >> 3);
            long long F = 0;
            if (m_l < o.n_models) {
                const uint2 *list = sd.over_list + s_r * I8_MAX_OVER;
                for (int oi = 0; oi < n_r; ++oi) {
                    const uint2 e = list[oi];
                    F += 256ll * (long long)e.y * (long long)sd.digits[g8_off(m_l * 8 + (lane & 7), e.x, sd.d_rows_pad)];
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                long long S[8];
#pragma unroll
                for (int d = 0; d < 8; ++d) S[d] = (long long)v[8 * q + d] + __shfl_sync(0xffffffffu, F, 8 * q + d);
                const int64_t m = m0 + q;
                if (lane == r && m < o.n_models && !(o.skip_diag && s == o.m_col0 + m)) store_sums(s, m, S);
            }
        }
        if (!live || n_over) return;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t m = m0 + q;
            if (m >= o.n_models) break;
            if (o.skip_diag && s == o.m_col0 + m) continue;
            long long S[8];
#pragma unroll
            for (int d = 0; d < 8; ++d) S[d] = (long long)v[8 * q + d];
            store_sums(s, m, S);
        }
    }
};

// k-split launches: the parts' digit sums -> M for every entry of the computed rectangles (thread = sequence: coalesced)
__global__ void __launch_bounds__(256) k_i8_fold(const EpiNll e, const G8Rect *__restrict__ rects) {
    const G8Rect r = rects[blockIdx.z];
    const int64_t s = (int64_t)r.a_blk * 256 + threadIdx.x;
    const int64_t m = (r.b_row0 >> 3) + blockIdx.x;
    if (s >= e.o.n_seq || m >= ((int64_t)(r.b_row0 + r.b_rows) >> 3) || m >= e.o.n_models) return;
    if (e.o.skip_diag && s == e.o.m_col0 + m) return;
    long long S[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = 0; p < e.o.n_parts; ++p) {
        const int32_t *src = e.o.P + ((int64_t)p * e.o.n_models + m) * 8 * e.o.p_ld + s;
#pragma unroll
        for (int d = 0; d < 8; ++d) S[d] += (long long)src[d * e.o.p_ld];
    }
    e.store(s, m, S, e.frequent_count(s));
}

// diagonal scores M[m][m] with the same arithmetic (one CTA per model): identical bits to the GEMM's entry
__global__ void __launch_bounds__(256) k_i8_diag(const EpiNll e, const uint8_t *__restrict__ C8, int64_t c_rows_pad,
                                                 const int8_t *__restrict__ digits, int64_t d_rows_pad, int64_t K8) {
    __shared__ int part[8][8];
    const int64_t m = blockIdx.x, sq = e.o.m_col0 + m;   // sequence sq was sampled from (local) model m
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int64_t n_chunks = K8 >> 4;
    const int64_t dr = m * 8;
    for (int64_t ch = threadIdx.x; ch < n_chunks; ch += blockDim.x) {
        const uint4 cv = *(const uint4 *)(C8 + g8_off(sq, 16 * ch, c_rows_pad));
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint4 dv = *(const uint4 *)((const uint8_t *)digits + g8_off(dr + q, 16 * ch, d_rows_pad));
            asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[q]) : "r"(cv.x), "r"(dv.x));
            asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[q]) : "r"(cv.y), "r"(dv.y));
            asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[q]) : "r"(cv.z), "r"(dv.z));
            asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[q]) : "r"(cv.w), "r"(dv.w));
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        if (lane == 0) part[warp][q] = acc[q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long S[8];
        for (int q = 0; q < 8; ++q) {
            int t = 0;
            for (int w2 = 0; w2 < 8; ++w2) t += part[w2][q];
            S[q] = (long long)t;
        }
        e.store(sq, m, S, e.frequent_count(sq));
    }
}

// ---------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------
static int64_t i8_k8(vgs_context *h) { return std::max<int64_t>(G8_BK, (int64_t)1 << (2 * (h->max_order + 1))); }
static int64_t i8_digit_rows_pad(vgs_context *h) { return (h->n_models * 8 + 255) / 256 * 256; }
static int64_t i8_count_rows_pad(vgs_context *h) { return (h->n_seq + 255) / 256 * 256; }

static I8Side i8_side(vgs_context *h, bool with_over) {
    I8Side sd;
    sd.meta = h->d_meta; sd.hash = h->d_hash; sd.logp = h->d_logp;
    sd.seq = h->d_seq; sd.seq_off = h->d_seq_off; sd.seq_len = h->d_seq_len;
    sd.over_cnt = with_over ? (const int32_t *)h->d_i8_over : nullptr;
    sd.over_list = with_over ? (const uint2 *)(h->d_i8_over + ((h->n_seq + 1) / 2) * 2) : nullptr;
    sd.digits = h->d_i8_digits; sd.d_rows_pad = i8_digit_rows_pad(h);
    sd.depth = h->max_order;
    sd.any_shallow = h->min_order < h->max_order ? 1 : 0;
    return sd;
}

// the digit operand of the current model set: begin (allocate, clear padding, clear the flag) -> range(s) -> end (after a
// stream synchronisation: is it usable?).  vgs_load_models drives the three steps itself, block by block behind the ln p
// upload; vgs_kmer_i8_prepare_tables does them in one go for sets loaded with VGS_LOAD_LAZY_TABLES.
int vgs_kmer_i8_digits_begin(vgs_context *h, cudaStream_t s) {
    const int64_t K = (int64_t)1 << (2 * (h->max_order + 1));
    const int64_t K8 = i8_k8(h);
    int rc;
    const int64_t rows_pad = i8_digit_rows_pad(h);
    if ((rc = vgs_reserve(h, &h->d_i8_digits, rows_pad * K8))) return rc;
    if (rows_pad != h->n_models * 8 || K8 != K)   // rows beyond the set / columns beyond 4^(D+1): zero
        VGS_CUDA(cudaMemsetAsync(h->d_i8_digits, 0, (size_t)(rows_pad * K8), s));
    if ((rc = vgs_reserve(h, &h->d_counters, 16))) return rc;
    VGS_CUDA(cudaMemsetAsync(h->d_counters + 9, 0, sizeof(int), s));
    h->i8_tables_gen = ~0ull;
    return VGS_OK;
}

int vgs_kmer_i8_digits_range(vgs_context *h, int64_t m_begin, int64_t m_end, cudaStream_t s) {
    const int64_t K = (int64_t)1 << (2 * (h->max_order + 1));
    const int64_t K8 = i8_k8(h);
    const I8Side sd = i8_side(h, false);
    for (int64_t m0 = m_begin; m0 < m_end; m0 += 65535) {
        dim3 grid((unsigned)std::min<int64_t>(((K8 >> 4) + 127) / 128, 64), (unsigned)std::min<int64_t>(m_end - m0, 65535));
        k_i8_digits<<<grid, 128, 0, s>>>(m0, sd, K, K8, i8_digit_rows_pad(h), h->d_i8_digits, h->d_counters + 9);
        VGS_LAUNCH_CHECK(h);
    }
    return VGS_OK;
}

int vgs_kmer_i8_digits_end(vgs_context *h, cudaStream_t s) {
    int flag = 0;
    VGS_CUDA(cudaMemcpyAsync(&flag, h->d_counters + 9, sizeof(int), cudaMemcpyDeviceToHost, s));
    VGS_CUDA(cudaStreamSynchronize(s));
    h->i8_tables_ok = flag == 0;
    h->i8_tables_gen = h->model_gen;
    return VGS_OK;
}

// the digit operand when the load did not build it; returns VGS_OK, an error, or -1 = not applicable (a model without root
// context, or a table entry that is no log-probability)
int vgs_kmer_i8_prepare_tables(vgs_context *h, cudaStream_t s) {
    if (h->i8_tables_gen == h->model_gen) return h->i8_tables_ok ? VGS_OK : -1;
    int rc;
    if ((rc = vgs_kmer_i8_digits_begin(h, s))) return rc;
    if ((rc = vgs_kmer_i8_digits_range(h, 0, h->n_models, s))) return rc;
    if ((rc = vgs_kmer_i8_digits_end(h, s))) return rc;
    return h->i8_tables_ok ? VGS_OK : -1;
}

// scores for the given rectangles (256-sequence block x model range); d_rects / d_units live on the device (cached by the
// caller).  Returns VGS_OK, an error, or -1 when a sequence has too many frequent k-mers (caller uses the fp64 path).
int vgs_kmer_i8_scores(vgs_context *h, const I8Work &w, bool diag_separately, const VgsScoreDst &dst, cudaStream_t s) {
    const int D = h->max_order;
    const int64_t K8 = i8_k8(h);
    int rc;
    const int64_t c_rows_pad = i8_count_rows_pad(h), d_rows_pad = i8_digit_rows_pad(h);
    if ((rc = vgs_reserve(h, &h->d_i8_C, c_rows_pad * K8))) return rc;
    if ((rc = vgs_reserve(h, &h->d_i8_over, h->n_seq * (I8_MAX_OVER * 2 + 1) + 2))) return rc;
    int32_t *over_cnt = (int32_t *)h->d_i8_over;
    uint2 *over_list = (uint2 *)(h->d_i8_over + ((h->n_seq + 1) / 2) * 2);
    if ((rc = vgs_reserve(h, &h->d_counters, 16))) return rc;
    // histograms of the resident sequences: part of every call (a pass over the packed sequences)
    int64_t max_len = 0;
    for (int64_t i = 0; i < h->n_seq; ++i) max_len = std::max(max_len, h->h_seq_len[(size_t)i]);
    const bool pack16 = max_len < 65536;
    const int hist_threads = 256;
    const int hist_smem = (int)(std::max<int64_t>(16, (int64_t)1 << (2 * (D + 1))) * (pack16 ? 2 : 4));
    if (h->i8_counts_state == 0) {
        VGS_CUDA(cudaFuncSetAttribute(k_i8_hist<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_smem));
        VGS_CUDA(cudaFuncSetAttribute(k_i8_hist<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_smem));
        VGS_CUDA(cudaMemsetAsync(h->d_counters + 10, 0, sizeof(int), s));
        if (c_rows_pad != h->n_seq) VGS_CUDA(cudaMemsetAsync(h->d_i8_C, 0, (size_t)(c_rows_pad * K8), s));   // rows beyond the set
    }
    if (pack16)
        k_i8_hist<true><<<(unsigned)h->n_seq, hist_threads, hist_smem, s>>>(D, h->d_seq, h->d_seq_off, h->d_seq_len, K8, c_rows_pad, h->d_i8_C,
                                                                   over_cnt, over_list, h->d_counters + 10);
    else
        k_i8_hist<false><<<(unsigned)h->n_seq, hist_threads, hist_smem, s>>>(D, h->d_seq, h->d_seq_off, h->d_seq_len, K8, c_rows_pad, h->d_i8_C,
                                                                    over_cnt, over_list, h->d_counters + 10);
    VGS_LAUNCH_CHECK(h);
    if (h->i8_counts_state == 0) {
        // first call on this sequence set: any frequent k-mers, and few enough per sequence?  (one synchronisation per
        // loaded set; the histograms of later calls are the same numbers)
        int flag = 0;
        VGS_CUDA(cudaMemcpyAsync(&flag, h->d_counters + 10, sizeof(int), cudaMemcpyDeviceToHost, s));
        VGS_CUDA(cudaStreamSynchronize(s));
        h->i8_counts_state = (flag & 2) ? 2 : ((flag & 1) ? 3 : 1);   // 1 = no frequent k-mers at all, 3 = some, 2 = too many
    }
    if (h->i8_counts_state == 2) return -1;
    EpiNll e;
    e.sd = i8_side(h, h->i8_counts_state == 3);
    for (int d = 0; d < VGS_MAX_PEERS; ++d) e.o.M[d] = d < dst.n_dst ? dst.M[d] : nullptr;
    e.o.n_dst = dst.n_dst; e.o.ss = dst.stride_s; e.o.sm = dst.stride_m; e.o.m_col0 = dst.m_col0;
    e.o.n_seq = h->n_seq; e.o.n_models = h->n_models; e.o.skip_diag = diag_separately ? 1 : 0;
    e.o.P = nullptr; e.o.p_ld = 0; e.o.n_parts = 1;
    if (w.n_units > 0) {
        e.o.p_ld = (h->n_seq + 31) / 32 * 32;
        e.o.n_parts = w.split;
        if (w.split > 1) {
            if ((rc = vgs_reserve(h, &h->d_i8_P, (int64_t)w.split * h->n_models * 8 * e.o.p_ld))) return rc;
            e.o.P = h->d_i8_P;
        }
        G8Args g;
        g.A = h->d_i8_C; g.B = (const uint8_t *)h->d_i8_digits; g.a_rows_pad = c_rows_pad; g.b_rows_pad = d_rows_pad;
        g.units = w.d_units; g.n_units = w.n_units; g.a_signed = 0; g.err = h->d_err; g.prof = nullptr;
        h->last_nll_kernel = VGS_NLL_KERNEL_I8MMA;
        vgs_prof_begin(h, VGS_PROF_NLL_SCORES, s);
        rc = g8_launch(h, g, e, s);
        vgs_prof_end(h, VGS_PROF_NLL_SCORES, s);
        if (rc) return rc;
        if (w.split > 1) {
            // widest rectangle in models (grid.x covers it; narrower rectangles leave early)
            k_i8_fold<<<dim3((unsigned)w.max_rect_models, 1, (unsigned)w.n_rects), 256, 0, s>>>(e, w.d_rects);
            VGS_LAUNCH_CHECK(h);
        }
    }
    if (diag_separately) {
        const int64_t nd = std::min(h->n_seq, h->n_models);
        k_i8_diag<<<(unsigned)nd, 256, 0, s>>>(e, h->d_i8_C, c_rows_pad, h->d_i8_digits, d_rows_pad, K8);
        VGS_LAUNCH_CHECK(h);
    }
    return VGS_OK;
}
