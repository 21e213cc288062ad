// vgs_kmer_i8.cu -- the NLL k-mer contraction as an EXACT integer GEMM on the 5th-generation tensor cores
// (tcgen05.mma kind::i8, CTA pairs, accumulators in tensor memory; GEMM core: vgs_i8gemm.cuh).
//
// Same quantity as vgs_kmer.cu:  M[s][m] = sum_k C[s][k] * T[m][k] + head(s, m)   (src/vlmc/vlmc.pyx:79-101 restated
// as a contraction).  ln p needs 53 bits, so no floating-point tensor-core format can carry T --
// but the contraction is a sum of (small integer) x (fixed-point number), and that is exact in integer arithmetic:
//
//   T[m][k]  ->  q = round(T * 2^f), written as ND balanced base-256 digits
//                q = sum_p d_p * 256^p,  d_p in [-128, 127]                      -- int8 "B" operand, ND rows per model
//   C[s][k]  ->  the low byte of the k-mer count                                   -- uint8 "A" operand
//   S_p[s][m] = sum_k (C[s][k] & 255) * d_p[m][k]   int32, exact: |S_p| <= 16384 * 255 * 128 < 2^31
//               + 256 * (C[s][k] >> 8) * d_p[m][k] over the few k-mers of a sequence that occur more than 255 times
//                 (listed by the histogram kernel; added as int64 where the digit sums are folded)
//   M[s][m]  = 2^-f * sum_p S_p * 256^p   (two int64 halves)  + head(s, m)
//
// f = 43 fraction bits by default (vgs_set_kmer_fraction_bits / VGS_I8_FRAC_BITS: 35..51).  The GEMM's work is
// proportional to ND, and ND follows from the range of the set's ln p: six digits hold every |T| < 16 at f = 43 (every
// transition probability above 1.1e-7), seven hold the whole legal range |T| <= 1000 including the -1000 of p = 0.  The
// operand is built optimistically with the smaller ND; the builder flags a value that does not fit and the operand is
// rebuilt once with the larger one (q itself does not depend on ND: results are the same bits either way, and on every
// rank of a model-sharded job whatever ND its shard needed).
// The only error is the quantisation of each ln p, |dT| <= 2^-(f+1): |dM| <= L * 2^-44 = 5.7e-10 absolute on |M| ~ 1.4e4
// at cfg 3 (4e-14 relative; the fp64 scan's own rounding noise on the same sum is ~1e-10), |d distance| <= 2^-43 =
// 1.1e-13 -- and every (s, m) entry is the same bits whatever the tiling, the k-split or the number of GPUs (integer sums
// do not depend on the order), so sharded matrices stay identical and duplicates score exactly alike.  Sets with more
// than 64 frequent k-mers in one sequence (or with a model without root context) take the fp64 MMA path of vgs_kmer.cu.
//
// The digit operand is built straight from the models (hash + host-libm ln p): no fp64 table, no per-model dense NLL
// table is needed for this mode.  Both operands live in HBM in the GEMM core's slab layout (vgs_i8gemm.cuh).
#include <stdlib.h>

#include <algorithm>

#include "vgs_i8gemm.cuh"

#define I8_MAX_OVER 64       // per sequence: k-mers whose count exceeds one byte

// the ND exact digit sums of one (sequence, model) -> fp64 (inv_scale = 2^-f, exact)
template <int ND>
__device__ __forceinline__ double i8_fold(const long long *S, double inv_scale) {
    const long long lo = S[0] + S[1] * 256 + S[2] * 65536 + S[3] * 16777216;
    long long hi = 0;
#pragma unroll
    for (int p = ND - 1; p >= 4; --p) hi = hi * 256 + S[p];
    const double v = __fma_rn((double)hi, 4294967296.0, (double)lo);
    return v * inv_scale;
}

// fixed-point form of one table entry (the same rounding everywhere; scale = 2^f)
__device__ __forceinline__ long long i8_quantise(double t, double scale) { return __double2ll_rn(t * scale); }

// unit granule of the GEMM for ND digit rows per model: a multiple of 32 (MMA / bulk-copy alignment) and of the
// functor's chunk (4 models = 4 ND columns)
static inline int i8_gran(int nd) { return nd == 8 ? 32 : (nd == 6 ? 96 : 224); }
int vgs_kmer_i8_gran(vgs_context *h) { return i8_gran(h->i8_nd); }

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
    double scale, inv_scale;   // 2^f, 2^-f
    int nd;                    // digit rows per model
    int depth;                 // D = largest order of the set; k-mers are (D + 1)-mers
    int any_shallow;           // some model is shallower than D (only then does anything look at the per-model metadata)
};

// adds 256 * hi * d_p(T[m][k]) for the frequent k-mers of sequence s.  The digits come straight from the GEMM's own digit
// operand (one independent one-byte load per digit and k-mer, from lines the tile has just streamed through L2) -- the same
// integers the tensor cores multiplied, so the sum stays exact.  (A first version re-derived them through the model's
// hash: seven dependent cache misses per model, 100 us for ONE sequence with ONE frequent k-mer at cfg 3.)
template <int ND>
__device__ __forceinline__ void i8_add_frequent(long long *S, const I8Side &sd, int64_t m, int64_t s, int n_over) {
    const uint2 *list = sd.over_list + s * I8_MAX_OVER;
    for (int o = 0; o < n_over; ++o) {
        const uint2 e = list[o];
        const long long w = 256ll * (long long)e.y;
#pragma unroll
        for (int p = 0; p < ND; ++p) S[p] += w * (long long)sd.digits[g8_off(m * ND + p, e.x, sd.d_rows_pad)];
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
// digits[slab(m * ND + p, k)] = p-th balanced base-256 digit of round(ln p_m(k) * 2^f).  One CTA per model.
// First the context of EVERY history up to the set's depth D, level by level in shared memory: a history of length k is its
// own context if the model has it (one hash probe), otherwise it inherits the context of its suffix of length k - 1 -- the
// longest-suffix rule (src/vlmc/vlmc.pyx:131-141) with (4^(D+1) - 1) / 3 probes per model instead of a walk of up to D + 1
// probes for each of the 4^D histories (1000 models of depth 6: 150 -> 35 us).  Then one thread per 16 k-mers (= 4 histories
// x 4 symbols) quantises and splits the digits.
// flag bit 0: an entry is not a log-probability (|T| > 1000 or NaN); bit 1: a history has no context at all (model without
// root context: the scan raises the reference's RuntimeError); bit 2: a value does not fit ND digits (the host rebuilds the
// operand with more)
// the context of every history of length <= D of one model, level by level in shared memory (whole CTA; ends with a barrier)
__device__ __forceinline__ void i8_contexts_by_level(uint16_t *ctx_of, const ModelMeta &mm, const uint2 *slots, int D) {
    if (threadIdx.x == 0) {
        const int id = vgs_hash_find(slots, mm.hash_mask, mm.hash_shift, vgs_key(0, 0u));
        ctx_of[0] = id < 0 ? (uint16_t)0xFFFFu : (uint16_t)id;
    }
    __syncthreads();
    for (int k = 1; k <= D; ++k) {
        const uint32_t n_k = 1u << (2 * k), off_k = (n_k - 1u) / 3u, off_prev = (n_k / 4u - 1u) / 3u, mask_prev = n_k / 4u - 1u;
        for (uint32_t hc = threadIdx.x; hc < n_k; hc += blockDim.x) {
            const int id = k <= mm.order ? vgs_hash_find(slots, mm.hash_mask, mm.hash_shift, vgs_key(k, hc)) : -1;
            ctx_of[off_k + hc] = id >= 0 ? (uint16_t)id : ctx_of[off_prev + (hc & mask_prev)];   // newest symbol in the low bits
        }
        __syncthreads();
    }
}

// The context half of the digit builder on its own: needs the hashes only, not ln p -- vgs_load_models runs it while the
// transition rows are still crossing PCIe (VGS_LOAD_DEVICE_LOG) and hands the result (ctx_pre[m][4^D], the deepest level)
// to k_i8_digits, which is then a pure streaming pass.
__global__ void __launch_bounds__(256) k_i8_contexts(int64_t m_base, I8Side sd, uint16_t *__restrict__ ctx_pre) {
    extern __shared__ uint16_t ctx_of[];
    const int64_t m = m_base + blockIdx.x;
    const ModelMeta mm = sd.meta[m];
    const int D = sd.depth;
    i8_contexts_by_level(ctx_of, mm, sd.hash + mm.hash_off, D);
    const uint32_t n_D = 1u << (2 * D);
    const uint16_t *ctx_D = ctx_of + ((n_D - 1u) / 3u);
    uint16_t *out = ctx_pre + m * (int64_t)n_D;
    for (uint32_t hc = threadIdx.x; hc < n_D; hc += blockDim.x) out[hc] = ctx_D[hc];
}

__global__ void __launch_bounds__(256) k_i8_digits(int64_t m_base, I8Side sd, int64_t K, int64_t K8, int64_t rows_pad,
                                                   int8_t *__restrict__ digits, int *__restrict__ flag,
                                                   const uint16_t *__restrict__ ctx_pre) {
    extern __shared__ uint16_t ctx_of[];   // levels 0 .. D one after the other: level k starts at (4^k - 1) / 3
    const int64_t m = m_base + blockIdx.x;
    const ModelMeta mm = sd.meta[m];
    const int D = sd.depth;
    if (!ctx_pre) i8_contexts_by_level(ctx_of, mm, sd.hash + mm.hash_off, D);
    const uint16_t *ctx_D = ctx_pre ? ctx_pre + m * ((int64_t)1 << (2 * D)) : ctx_of + (((1u << (2 * D)) - 1u) / 3u);
    for (int64_t ch = threadIdx.x; ch < (K8 >> 4); ch += blockDim.x) {
        long long q[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t hc = 4 * ch + i;   // history code
            double4 v = make_double4(0.0, 0.0, 0.0, 0.0);
            if (16 * ch + 4 * i < K) {
                const uint32_t id = ctx_D[hc];
                if (id != 0xFFFFu) {
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
                else q[4 * i + x] = i8_quantise(t4[x], sd.scale);
            }
        }
        // balanced base-256 digits of all 16 values at once: adding 0x80 to every digit position (with the carries of an
        // ordinary 64-bit addition) turns the balanced digits into the sum's plain bytes minus 128, i.e. its bytes with the
        // top bit flipped.  (Digit by digit in 64-bit arithmetic this loop was 3/4 of the kernel's instructions.)
        const unsigned long long bias = sd.nd >= 8 ? 0x8080808080808080ull : ((0x8080808080808080ull) >> (8 * (8 - sd.nd)));
        uint32_t lo[16], hi[16];
        long long left = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const long long biased = q[j] + (long long)bias;            // |q| < 2^61: no overflow
            if (sd.nd < 8) left |= biased >> (8 * sd.nd);              // what does not fit nd digits (0 when all of it does)
            const unsigned long long dg = (unsigned long long)biased ^ bias;
            lo[j] = (uint32_t)dg;
            hi[j] = (uint32_t)(dg >> 32);
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            if (p < sd.nd) {
                uint32_t w[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t *src = p < 4 ? lo : hi;
                    const uint32_t sel = (uint32_t)((p & 3) | (((p & 3) + 4) << 4));   // byte p of either operand
                    const uint32_t t0 = __byte_perm(src[4 * g], src[4 * g + 1], sel), t1 = __byte_perm(src[4 * g + 2], src[4 * g + 3], sel);
                    w[g] = __byte_perm(t0, t1, 0x5410);
                }
                *(uint4 *)(digits + g8_off(m * sd.nd + p, 16 * ch, rows_pad)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        if (left) atomicOr(flag, 4);
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
    vgs_dep_trigger();
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

// ---- epilogue of the GEMM: one chunk = 4 models x ND digits of one sequence ---------------------------------------
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

// exact digit sums held as int32 (the tensor cores' own output) -> fp64: IMAD.WIDE chains
template <int ND>
__device__ __forceinline__ double i8_fold32(const int32_t *a, double inv_scale) {
    long long lo = (long long)a[0] + (long long)a[1] * 256 + (long long)a[2] * 65536 + (long long)a[3] * 16777216;
    long long hi = (long long)a[ND - 1];
#pragma unroll
    for (int p = ND - 2; p >= 4; --p) hi = hi * 256 + (long long)a[p];
    return __fma_rn((double)hi, 4294967296.0, (double)lo) * inv_scale;
}

// One definition of a score for every path (GEMM epilogue, k-split fold, diagonal kernel, any tiling, any rank count):
//     M[s][m] = ( head(s, m) + fold(S) ) + fold(F)
// S = the tensor cores' digit sums over the low count bytes; head = the first positions of a model shallower than the set's
// depth (only such models have one); F = 256 * sum over the sequence's frequent k-mers of (count >> 8) * digit (only
// sequences that have such k-mers), each folded exactly and added in this order.
// The GEMM's epilogue writes fold(S) and nothing else: it is exposed time on every CTA pair.  The two rare terms are added
// afterwards by k_i8_heads and k_i8_frequent, launched only for sets that have shallow models / frequent k-mers (the
// k-split fold and the diagonal kernel add them themselves).  An earlier version did both inside the epilogue: ONE sequence
// with ONE k-mer counted 260 times then held the CTA pairs of its 256-sequence block back by 8 of the kernel's 92 us (two
// dependent cache misses per chunk), and the inlined head walk doubled the kernel's code.
template <int ND>
struct EpiNll {
    static constexpr int CHUNK = 4 * ND;
    I8Out o;
    I8Side sd;
    struct State {};
    __device__ __forceinline__ int frequent_count(int64_t s) const { return sd.over_cnt ? sd.over_cnt[s] : 0; }
    __device__ __forceinline__ State prepare(const G8Unit &, int64_t) const { return State(); }
    __device__ __forceinline__ void prefetch(const G8Unit &, int64_t, int, const State &) const {}
    __device__ __forceinline__ void put(int64_t s, int64_t m, double v) const {
        const int64_t at = s * o.ss + (o.m_col0 + m) * o.sm;
        o.M[0][at] = v;
        for (int d = 1; d < o.n_dst; ++d) o.M[d][at] = v;
    }
    __device__ __forceinline__ bool shallow(int64_t m) const { return sd.any_shallow && sd.meta[m].order < sd.depth; }
    __device__ __forceinline__ double frequent_term(int64_t s, int64_t m, int n_over) const {
        long long F[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) F[d] = 0;
        i8_add_frequent<ND>(F, sd, m, s, n_over);
        return i8_fold<ND>(F, sd.inv_scale);
    }
    // fold / diagonal kernels (one entry per thread): S = the complete digit sums over the low count bytes
    __device__ __forceinline__ void store(int64_t s, int64_t m, const long long *S, int n_over) const {
        double v = i8_fold<ND>(S, sd.inv_scale);
        if (shallow(m)) v = i8_head(sd, sd.meta[m], s) + v;
        if (n_over) v = v + frequent_term(s, m, n_over);
        put(s, m, v);
    }
    // peer-mapped destinations: the band must have left this GPU before the rank signals the peer barrier
    __device__ __forceinline__ void unit_done() const {
        if (o.n_dst > 1) __threadfence_system();
    }
    // lane = sequence row s, v = CHUNK columns = 4 models x ND digits
    __device__ __forceinline__ void operator()(const G8Unit &u, int64_t s, int c0, const int32_t (&v)[CHUNK], const State &) const {
        if (s >= o.n_seq) return;
        const int64_t m0 = (int64_t)(u.b_row0 + c0) / ND;
        if (u.partial) {
            // part (u.partial - 1) of a k-split: its digit sums go to the part's own slab with plain stores, coalesced over
            // the warp's 32 sequences (a first version added them into one buffer with integer atomics: 8 M atomics at
            // 2 GPUs cost 110 us); k_i8_fold adds the parts -- integers, so the split changes no bit of the result
            int32_t *dst = o.P + ((int64_t)(u.partial - 1) * o.n_models + m0) * ND * o.p_ld + s;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (m0 + q >= o.n_models) break;
#pragma unroll
                for (int d = 0; d < ND; ++d) dst[(ND * q + d) * o.p_ld] = v[ND * q + d];
            }
            return;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t m = m0 + q;
            if (m >= o.n_models) break;
            if (o.skip_diag && s == o.m_col0 + m) continue;
            put(s, m, i8_fold32<ND>(&v[ND * q], sd.inv_scale));
        }
    }
};

// k-split launches: the parts' digit sums -> M for every entry of the computed rectangles (thread = sequence: coalesced)
template <int ND>
__global__ void __launch_bounds__(256) k_i8_fold(const EpiNll<ND> e, const G8Rect *__restrict__ rects) {
    vgs_dep_trigger();
    vgs_dep_wait();
    const G8Rect r = rects[blockIdx.z];
    const int64_t s = (int64_t)r.a_blk * 256 + threadIdx.x;
    const int64_t m = r.b_row0 / ND + blockIdx.x;
    if (s >= e.o.n_seq || m >= (int64_t)(r.b_row0 + r.b_rows) / ND || m >= e.o.n_models) return;
    if (e.o.skip_diag && s == e.o.m_col0 + m) return;
    long long S[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) S[d] = 0;
    for (int p = 0; p < e.o.n_parts; ++p) {
        const int32_t *src = e.o.P + ((int64_t)p * e.o.n_models + m) * ND * e.o.p_ld + s;
#pragma unroll
        for (int d = 0; d < ND; ++d) S[d] += (long long)src[d * e.o.p_ld];
    }
    e.store(s, m, S, e.frequent_count(s));
}

// after an unsplit GEMM, sets with models shallower than the set's depth: M = head + M for every entry of those models that
// the launch's rectangles cover (grid.x = model of the rectangle, grid.z = rectangle, thread = sequence)
template <int ND>
__global__ void __launch_bounds__(256) k_i8_heads(const EpiNll<ND> e, const G8Rect *__restrict__ rects) {
    vgs_dep_trigger();
    vgs_dep_wait();
    const G8Rect r = rects[blockIdx.z];
    const int64_t s = (int64_t)r.a_blk * 256 + threadIdx.x;
    const int64_t m = r.b_row0 / ND + blockIdx.x;
    if (m >= (int64_t)(r.b_row0 + r.b_rows) / ND || m >= e.o.n_models || !e.shallow(m)) return;
    if (s >= e.o.n_seq || (e.o.skip_diag && s == e.o.m_col0 + m)) return;
    const int64_t at = s * e.o.ss + (e.o.m_col0 + m) * e.o.sm;
    e.put(s, m, i8_head(e.sd, e.sd.meta[m], s) + e.o.M[0][at]);
    if (e.o.n_dst > 1) __threadfence_system();
}

// after an unsplit GEMM (and k_i8_heads): the frequent-k-mer term of the sequences that have one (a handful), for every
// entry the launch's rectangles cover.  One thread per sequence looks at its count; a sequence that has frequent k-mers is
// then served by the CTAs of its grid row (grid.y x 128 threads = the models of the widest rectangle, one entry per thread:
// a single CTA looping over 1000 models took 12 us of dependent cache misses).  The entry was written by this GPU earlier
// on the stream: read locally, written to every destination.
template <int ND>
__global__ void __launch_bounds__(128) k_i8_frequent(const EpiNll<ND> e, const G8Rect *__restrict__ rects, int n_rects) {
    __shared__ int todo[128];
    __shared__ int n_todo;
    vgs_dep_trigger();
    vgs_dep_wait();
    if (threadIdx.x == 0) n_todo = 0;
    __syncthreads();
    const int64_t s_mine = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s_mine < e.o.n_seq && e.sd.over_cnt[s_mine] > 0) todo[atomicAdd(&n_todo, 1)] = (int)threadIdx.x;
    __syncthreads();
    for (int t = 0; t < n_todo; ++t) {
        const int64_t s = (int64_t)blockIdx.x * blockDim.x + todo[t];
        const int n_over = e.sd.over_cnt[s];
        for (int ri = 0; ri < n_rects; ++ri) {
            const G8Rect r = rects[ri];
            if ((int64_t)r.a_blk != (s >> 8)) continue;
            const int64_t m_end = min((long long)e.o.n_models, (long long)((r.b_row0 + r.b_rows) / ND));
            const int64_t m = r.b_row0 / ND + (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
            if (m >= m_end || (e.o.skip_diag && s == e.o.m_col0 + m)) continue;
            const int64_t at = s * e.o.ss + (e.o.m_col0 + m) * e.o.sm;
            e.put(s, m, e.o.M[0][at] + e.frequent_term(s, m, n_over));
        }
    }
    if (e.o.n_dst > 1) __threadfence_system();
}

// diagonal scores M[m][m] with the same arithmetic (one CTA per model): identical bits to the GEMM's entry
template <int ND>
__global__ void __launch_bounds__(256) k_i8_diag(const EpiNll<ND> e, const uint8_t *__restrict__ C8, int64_t c_rows_pad,
                                                 const int8_t *__restrict__ digits, int64_t d_rows_pad, int64_t K8) {
    __shared__ int part[8][ND];
    vgs_dep_trigger();
    vgs_dep_wait();
    const int64_t m = blockIdx.x, sq = e.o.m_col0 + m;   // sequence sq was sampled from (local) model m
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int acc[ND];
#pragma unroll
    for (int q = 0; q < ND; ++q) acc[q] = 0;
    const int64_t n_chunks = K8 >> 4;
    const int64_t dr = m * ND;
    for (int64_t ch = threadIdx.x; ch < n_chunks; ch += blockDim.x) {
        const uint4 cv = *(const uint4 *)(C8 + g8_off(sq, 16 * ch, c_rows_pad));
#pragma unroll
        for (int q = 0; q < ND; ++q) {
            const uint4 dv = *(const uint4 *)((const uint8_t *)digits + g8_off(dr + q, 16 * ch, d_rows_pad));
            asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[q]) : "r"(cv.x), "r"(dv.x));
            asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[q]) : "r"(cv.y), "r"(dv.y));
            asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[q]) : "r"(cv.z), "r"(dv.z));
            asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[q]) : "r"(cv.w), "r"(dv.w));
        }
    }
#pragma unroll
    for (int q = 0; q < ND; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        if (lane == 0) part[warp][q] = acc[q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long S[ND];
        for (int q = 0; q < ND; ++q) {
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
// digit rows: the models' ND rows each, padded to whole granules (the rectangles end there) and then to 256
static int64_t i8_digit_rows_pad(vgs_context *h) {
    const int64_t gran = i8_gran(h->i8_nd);
    return ((h->n_models * h->i8_nd + gran - 1) / gran * gran + 255) / 256 * 256;
}
static int64_t i8_count_rows_pad(vgs_context *h) { return (h->n_seq + 255) / 256 * 256; }
// digits that hold |T| < 2^int_bits at f fraction bits (sign + integer + fraction bits); the kernels exist for 6, 7, 8
static int i8_digits_for(int frac_bits, int int_bits) { return std::max(6, std::min(8, (frac_bits + int_bits + 1 + 7) / 8)); }

static I8Side i8_side(vgs_context *h, bool with_over) {
    I8Side sd;
    sd.meta = h->d_meta; sd.hash = h->d_hash; sd.logp = h->logp_device ? h->d_logp_dev : h->d_logp;
    sd.seq = h->d_seq; sd.seq_off = h->d_seq_off; sd.seq_len = h->d_seq_len;
    sd.over_cnt = with_over ? (const int32_t *)h->d_i8_over : nullptr;
    sd.over_list = with_over ? (const uint2 *)(h->d_i8_over + ((h->n_seq + 1) / 2) * 2) : nullptr;
    sd.digits = h->d_i8_digits; sd.d_rows_pad = i8_digit_rows_pad(h);
    sd.scale = ldexp(1.0, h->i8_frac_bits); sd.inv_scale = ldexp(1.0, -h->i8_frac_bits);
    sd.nd = h->i8_nd;
    sd.depth = h->max_order;
    sd.any_shallow = h->min_order < h->max_order ? 1 : 0;
    return sd;
}

// the digit operand of the current model set: begin (choose ND, allocate, clear padding, clear the flag) -> range(s) ->
// end (after a stream synchronisation: is it usable?  did every value fit?).  vgs_load_models drives the three steps
// itself, block by block behind the ln p upload; vgs_kmer_i8_prepare_tables does them in one go for sets loaded with
// VGS_LOAD_LAZY_TABLES.
static int i8_digits_begin_nd(vgs_context *h, int nd, cudaStream_t s) {
    h->i8_nd = nd;
    const int64_t K = (int64_t)1 << (2 * (h->max_order + 1));
    const int64_t K8 = i8_k8(h);
    int rc;
    const int64_t rows_pad = i8_digit_rows_pad(h);
    if ((rc = vgs_reserve(h, &h->d_i8_digits, rows_pad * K8))) return rc;
    // zero what the builder does not write: the rows beyond the set (per 128-byte k-block slab the rows lie one after the
    // other, so they are one strided region) -- the builder itself writes zeros for the columns beyond 4^(D+1)
    if (rows_pad != h->n_models * nd)
        VGS_CUDA(cudaMemset2DAsync(h->d_i8_digits + h->n_models * nd * G8_BK, (size_t)(rows_pad * G8_BK), 0,
                                   (size_t)((rows_pad - h->n_models * nd) * G8_BK), (size_t)(K8 / G8_BK), s));
    if ((rc = vgs_reserve(h, &h->d_counters, 16))) return rc;
    VGS_CUDA(cudaMemsetAsync(h->d_counters + 9, 0, sizeof(int), s));
    h->i8_tables_gen = ~0ull;
    return VGS_OK;
}

// wide == false: the digits that hold |ln p| < vgs_kmer_i8_narrow_limit (16 at f = 43: every probability above 1.1e-7);
// wide: the digits of the full legal range |ln p| <= 1000 (the -1000 of p = 0 included).  A narrow build whose values turn
// out not to fit is rebuilt wide by vgs_kmer_i8_digits_end.
int vgs_kmer_i8_digits_begin(vgs_context *h, bool wide, cudaStream_t s) {
    return i8_digits_begin_nd(h, i8_digits_for(h->i8_frac_bits, wide ? 10 : 4), s);
}
double vgs_kmer_i8_narrow_limit(vgs_context *h) { return ldexp(1.0, 8 * i8_digits_for(h->i8_frac_bits, 4) - 1 - h->i8_frac_bits); }

// does any ln p of the set lie outside the narrow range?  (lazy builds: one pass over the resident ln p decides the digit
// count before the operand is written)
__global__ void k_i8_range(int64_t n, const double *__restrict__ logp, double limit, int *__restrict__ flag) {
    bool out = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out |= !(fabs(logp[i]) < limit);
    if (__any_sync(0xffffffffu, out) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

int vgs_kmer_i8_digits_range(vgs_context *h, int64_t m_begin, int64_t m_end, cudaStream_t s) {
    const int64_t K = (int64_t)1 << (2 * (h->max_order + 1));
    const int64_t K8 = i8_k8(h);
    const I8Side sd = i8_side(h, false);
    const int smem = (int)(((((int64_t)1 << (2 * (h->max_order + 1))) - 1) / 3) * sizeof(uint16_t));   // <= 10.7 KB at depth 6
    // contexts found ahead of time for exactly this model set (vgs_kmer_i8_contexts_ahead): the builder only streams
    const uint16_t *pre = h->i8_ctx_gen == h->model_gen ? h->d_i8_ctx : nullptr;
    if (m_end > m_begin)
        k_i8_digits<<<(unsigned)(m_end - m_begin), 256, pre ? 0 : smem, s>>>(m_begin, sd, K, K8, i8_digit_rows_pad(h), h->d_i8_digits,
                                                                             h->d_counters + 9, pre);
    VGS_LAUNCH_CHECK(h);
    return VGS_OK;
}

// the context lookups of the digit builder, ahead of ln p (needs meta + hashes of the current model set on the device)
int vgs_kmer_i8_contexts_ahead(vgs_context *h, cudaStream_t s) {
    const int D = h->max_order;
    int rc;
    if ((rc = vgs_reserve(h, &h->d_i8_ctx, h->n_models * ((int64_t)1 << (2 * D))))) return rc;
    const I8Side sd = i8_side(h, false);
    const int smem = (int)(((((int64_t)1 << (2 * (D + 1))) - 1) / 3) * sizeof(uint16_t));
    k_i8_contexts<<<(unsigned)h->n_models, 256, smem, s>>>(0, sd, h->d_i8_ctx);
    VGS_LAUNCH_CHECK(h);
    h->i8_ctx_gen = h->model_gen;
    return VGS_OK;
}

int vgs_kmer_i8_digits_end(vgs_context *h, cudaStream_t s) {
    int flag = 0;
    VGS_CUDA(cudaMemcpyAsync(&flag, h->d_counters + 9, sizeof(int), cudaMemcpyDeviceToHost, s));
    VGS_CUDA(cudaStreamSynchronize(s));
    const int nd_full = i8_digits_for(h->i8_frac_bits, 10);
    if ((flag & 4) && !(flag & 3) && h->i8_nd < nd_full) {
        // a value beyond the optimistic range (a probability below e^-16, or the -1000 of p = 0): once more, with the digits
        // of the whole legal range
        int rc;
        if ((rc = i8_digits_begin_nd(h, nd_full, s))) return rc;
        if ((rc = vgs_kmer_i8_digits_range(h, 0, h->n_models, s))) return rc;
        VGS_CUDA(cudaMemcpyAsync(&flag, h->d_counters + 9, sizeof(int), cudaMemcpyDeviceToHost, s));
        VGS_CUDA(cudaStreamSynchronize(s));
    }
    h->i8_tables_ok = flag == 0;
    h->i8_tables_gen = h->model_gen;
    return VGS_OK;
}

// the digit operand when the load did not build it; returns VGS_OK, an error, or -1 = not applicable (a model without root
// context, or a table entry that is no log-probability)
int vgs_kmer_i8_prepare_tables(vgs_context *h, cudaStream_t s) {
    if (h->i8_tables_gen == h->model_gen) return h->i8_tables_ok ? VGS_OK : -1;
    int rc;
    if ((rc = vgs_reserve(h, &h->d_counters, 16))) return rc;
    int wide = 0;
    VGS_CUDA(cudaMemsetAsync(h->d_counters + 11, 0, sizeof(int), s));
    k_i8_range<<<std::min(4 * h->sm_count, (int)((h->total_ctx * 4 + 255) / 256)), 256, 0, s>>>(h->total_ctx * 4, h->logp_device ? h->d_logp_dev : h->d_logp, vgs_kmer_i8_narrow_limit(h), h->d_counters + 11);
    VGS_LAUNCH_CHECK(h);
    VGS_CUDA(cudaMemcpyAsync(&wide, h->d_counters + 11, sizeof(int), cudaMemcpyDeviceToHost, s));
    VGS_CUDA(cudaStreamSynchronize(s));
    if ((rc = vgs_kmer_i8_digits_begin(h, wide != 0, s))) return rc;
    if ((rc = vgs_kmer_i8_digits_range(h, 0, h->n_models, s))) return rc;
    if ((rc = vgs_kmer_i8_digits_end(h, s))) return rc;
    return h->i8_tables_ok ? VGS_OK : -1;
}

template <int ND>
static int i8_launch(vgs_context *h, const I8Work &w, bool diag_separately, const VgsScoreDst &dst, int64_t K8, int64_t c_rows_pad,
                     int64_t d_rows_pad, cudaStream_t s) {
    int rc;
    EpiNll<ND> e;
    e.sd = i8_side(h, h->i8_counts_state == 3);
    for (int d = 0; d < VGS_MAX_PEERS; ++d) e.o.M[d] = d < dst.n_dst ? dst.M[d] : nullptr;
    e.o.n_dst = dst.n_dst; e.o.ss = dst.stride_s; e.o.sm = dst.stride_m; e.o.m_col0 = dst.m_col0;
    e.o.n_seq = h->n_seq; e.o.n_models = h->n_models; e.o.skip_diag = diag_separately ? 1 : 0;
    e.o.P = nullptr; e.o.p_ld = 0; e.o.n_parts = 1;
    if (w.n_units > 0) {
        e.o.p_ld = (h->n_seq + 31) / 32 * 32;
        e.o.n_parts = w.split;
        if (w.split > 1) {
            if ((rc = vgs_reserve(h, &h->d_i8_P, (int64_t)w.split * h->n_models * ND * e.o.p_ld))) return rc;
            e.o.P = h->d_i8_P;
        }
        G8Args g;
        g.A = h->d_i8_C; g.B = (const uint8_t *)h->d_i8_digits; g.a_rows_pad = c_rows_pad; g.b_rows_pad = d_rows_pad;
        g.units = w.d_units; g.n_units = w.n_units; g.a_signed = 0; g.err = h->d_err; g.prof = nullptr; g.gran = i8_gran(ND);
        h->last_nll_kernel = VGS_NLL_KERNEL_I8MMA;
        vgs_prof_begin(h, VGS_PROF_NLL_SCORES, s);
        rc = g8_launch(h, g, e, s);
        vgs_prof_end(h, VGS_PROF_NLL_SCORES, s);
        if (rc) return rc;
        if (w.split > 1) {
            // widest rectangle in models (grid.x covers it; narrower rectangles leave early)
            vgs_launch_dep(k_i8_fold<ND>, dim3((unsigned)w.max_rect_models, 1, (unsigned)w.n_rects), dim3(256), 0, s, e, w.d_rects);
            VGS_LAUNCH_CHECK(h);
        } else {
            if (e.sd.any_shallow) {
                vgs_launch_dep(k_i8_heads<ND>, dim3((unsigned)w.max_rect_models, 1, (unsigned)w.n_rects), dim3(256), 0, s, e, w.d_rects);
                VGS_LAUNCH_CHECK(h);
            }
            if (h->i8_counts_state == 3) {
                vgs_launch_dep(k_i8_frequent<ND>, dim3((unsigned)((h->n_seq + 127) / 128), (unsigned)((w.max_rect_models + 127) / 128)), dim3(128), 0, s, e,
                               w.d_rects, w.n_rects);
                VGS_LAUNCH_CHECK(h);
            }
        }
    }
    if (diag_separately) {
        const int64_t nd = std::min(h->n_seq, h->n_models);
        vgs_launch_dep(k_i8_diag<ND>, dim3((unsigned)nd), dim3(256), 0, s, e, (const uint8_t *)h->d_i8_C, c_rows_pad, (const int8_t *)h->d_i8_digits,
                       d_rows_pad, K8);
        VGS_LAUNCH_CHECK(h);
    }
    return VGS_OK;
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
    bool deferred = false;
    if (h->i8_counts_state == 0) {
        // first call on this sequence set: any frequent k-mers, and few enough per sequence?  (once per loaded set; the
        // histograms of later calls are the same numbers)
        if (h->i8_defer_check) {
            // a host-buffer entry point waits for its result anyway: no wait here.  The chain is launched as for a set
            // WITH frequent k-mers (their kernel finds what there is), the flag travels behind it, and the caller settles
            // the state after its own synchronisation (vgs_kmer_i8_settle: "too many" voids the pass and it is redone)
            if (!h->h_i8_flag) VGS_CUDA(cudaHostAlloc((void **)&h->h_i8_flag, sizeof(int), cudaHostAllocDefault));
            h->i8_counts_state = 3;
            deferred = true;
        } else {
            int flag = 0;
            VGS_CUDA(cudaMemcpyAsync(&flag, h->d_counters + 10, sizeof(int), cudaMemcpyDeviceToHost, s));
            VGS_CUDA(cudaStreamSynchronize(s));
            h->i8_counts_state = (flag & 2) ? 2 : ((flag & 1) ? 3 : 1);   // 1 = no frequent k-mers at all, 3 = some, 2 = too many
        }
    }
    if (h->i8_counts_state == 2) return -1;
    switch (h->i8_nd) {
    case 6: rc = i8_launch<6>(h, w, diag_separately, dst, K8, c_rows_pad, d_rows_pad, s); break;
    case 7: rc = i8_launch<7>(h, w, diag_separately, dst, K8, c_rows_pad, d_rows_pad, s); break;
    default: rc = i8_launch<8>(h, w, diag_separately, dst, K8, c_rows_pad, d_rows_pad, s); break;
    }
    if (deferred) {
        // (behind the whole chain: a copy between the histogram and the GEMM would end their dependent-launch overlap)
        cudaMemcpyAsync(h->h_i8_flag, h->d_counters + 10, sizeof(int), cudaMemcpyDeviceToHost, s);
        h->i8_check_pending = true;
    }
    return rc;
}

// after the caller's synchronisation of a pass that ran with i8_defer_check: the frequent-k-mer state of the sequence set.
// Returns true when the pass is void (a sequence with more frequent k-mers than the side list holds) and must be redone.
bool vgs_kmer_i8_settle(vgs_context *h) {
    if (!h->i8_check_pending) return false;
    h->i8_check_pending = false;
    if (h->i8_counts_state == 0) return false;   // the sets changed in between: nothing to settle
    const int flag = *h->h_i8_flag;
    h->i8_counts_state = (flag & 2) ? 2 : ((flag & 1) ? 3 : 1);
    return h->i8_counts_state == 2;
}

// fraction bits f of the fixed-point ln p of the k-mer contraction (35..51; default 43): |d score| <= L * 2^-(f+1),
// |d distance| <= 2^-f.  Takes effect with the next operand build (the current one is dropped).
extern "C" int vgs_set_kmer_fraction_bits(vgs_handle h, int bits) {
    VGS_CHECK_ARG(h && bits >= 35 && bits <= 51, "vgs_set_kmer_fraction_bits: 35 .. 51");
    if (bits != h->i8_frac_bits) {
        h->i8_frac_bits = bits;
        h->i8_tables_gen = ~0ull;
        vgs_wcache_invalidate(h);
    }
    return VGS_OK;
}

// the operand as built for the current model set: digit rows per model (0: not built) and fraction bits
extern "C" int vgs_kmer_digit_info(vgs_handle h, int *digits_out, int *fraction_bits_out) {
    VGS_CHECK_ARG(h, "null handle");
    if (digits_out) *digits_out = (h->i8_tables_gen == h->model_gen && h->i8_tables_ok) ? h->i8_nd : 0;
    if (fraction_bits_out) *fraction_bits_out = h->i8_frac_bits;
    return VGS_OK;
}
