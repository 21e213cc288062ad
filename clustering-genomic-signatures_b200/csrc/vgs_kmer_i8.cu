// vgs_kmer_i8.cu -- the NLL k-mer contraction as an EXACT integer GEMM on the 5th-generation tensor cores
// (tcgen05.mma kind::i8, accumulators in tensor memory).
//
// Same quantity as vgs_kmer.cu:  M[s][m] = sum_k C[s][k] * T[m][k] + head(s, m)   (src/vlmc/vlmc.pyx:79-101 restated
// as a contraction).  ln p needs 53 bits, so no floating-point tensor-core format can carry T --
// but the contraction is a sum of (small integer) x (fixed-point number), and that is exact in integer arithmetic:
//
//   T[m][k]  ->  q = round(T * 2^51)  (|T| <= 1000 < 2^10, so |q| < 2^61), written as eight balanced base-256 digits
//                q = sum_p d_p * 256^p,  d_p in [-128, 127]                      -- int8 "B" operand, 8 rows per model
//   C[s][k]  ->  the low byte of the k-mer count                                   -- uint8 "A" operand
//   S_p[s][m] = sum_k (C[s][k] & 255) * d_p[m][k]   int32, exact: |S_p| <= 16384 * 255 * 128 < 2^31
//   M[s][m]  = 2^-51 * sum_p S_p * 256^p   (two int64 halves, one fp64 rounding each)
//               + the same fold of 256 * (C[s][k] >> 8) * d_p[m][k] over the few k-mers of a sequence that occur more
//                 than 255 times (listed by the histogram kernel, added by k_i8_frequent: exact integers again)
//
// The only error is the 2^-52 quantisation of each ln p: |dM| <= L * 2^-52 ~ 2e-12 absolute on |M| ~ 1e4, tighter than
// any fp64 summation order -- and every (s, m) entry is the same bits whatever the tiling or k-split, so 1/2/4/8-rank matrices
// stay identical and duplicates score exactly alike.  Sets with more than 64 such frequent k-mers in one sequence (or
// without a root context) take the fp64 MMA path of vgs_kmer.cu.
//
// Layout.  Both operands live in HBM already as the shared-memory image of the canonical K-major SWIZZLE_128B UMMA layout, tiled as
// [128-row block][128-k-mer block][128 rows x 128 bytes, 128-byte swizzled]: a pipeline stage (128 rows x 128 k-mers) of either operand is
// ONE contiguous 16 KB block, moved by a single 1-D bulk async copy (TMA) that completes on the stage's mbarrier.
// Kernel (template <NACC, STAGES>).  CTA tile = 128 sequences x NACC accumulators of 128 operand rows (16 models x 8
// digits each), all fed by the same count tile; two CTAs per SM.  <2, 2> (128 x 256 tiles, 48 KB stages) when that
// still fills the CTA slots, <1, 3> (twice as many tiles of half the latency) for multi-GPU shares and small sets.  One thread drives the main loop
// asynchronously: it re-arms a free stage (expect_tx + one bulk copy per operand tile), waits for the next full
// stage, issues the tcgen05.mma (M128 N128 K32) of that stage and commits them to the stage's "empty" mbarrier.  The
// other threads sleep at a CTA barrier until the last commit has landed; then the four warps read their 32 TMEM
// lanes with tcgen05.ld, fold the eight digit sums of each model and write M; the second CTA of the SM keeps the
// tensor pipe busy meanwhile.
// Measured on cfg 3 (GEMM ms; accumulators x stages x CTAs/SM): 2x2x2 0.134, 1x3x2 0.151, 4x2x1 0.168, 2x4x1 0.214, and
// a 2 x 2 tile (two count tiles x two digit tiles, half the operand traffic) 0.154: every variant issues one
// M128 N128 K32 MMA per ~148 cycles per SM, i.e. the kernel sits at the single-CTA (cta_group::1) MMA rate, about half
// of the int8 peak; operand traffic, shared-memory swizzle (none vs 128 B: same time) and pipeline depth are not the
// limit.  The next step is cta_group::2 (CTA pairs, M = 256).
#include <stdlib.h>

#include <algorithm>

#include "vgs_common.cuh"

#define I8_TM 128
#define I8_TN 128
#define I8_MODELS_PER_TILE (I8_TN / 8)
#define I8_BK 128            // k-mers (= bytes per operand row) per stage = one 128-byte swizzle atom
#define I8_CHUNKS (I8_BK / 16)
#ifndef I8_CTAS_PER_SM
#define I8_CTAS_PER_SM 2     // measured (cfg 3, GEMM ms): NACC x STAGES x CTAs = 2x2x2 0.134, 1x3x2 0.151, 4x2x1 0.168, 2x4x1 0.214
#endif
#define I8_THREADS 128
#define I8_SBO 1024                                   // bytes between consecutive 8-row groups (8 rows x 128 bytes)
#define I8_OPERAND_BYTES (I8_TM * I8_BK)              // one operand tile of one stage: 128 rows x 128 bytes = 16 KB
#define I8_SCALE_BITS 51
#define I8_MAX_OVER 64       // per sequence: k-mers whose count exceeds one byte (their high part is added in the epilogue)

// byte offset of element (row, k) in the tiled operand layout: [row >> 7][k >> 7] tiles of 128 rows x 128 bytes, each
// the shared-memory image of the K-major SWIZZLE_128B canonical layout (row r at r * 128, its 16-byte chunk c stored
// at chunk position c ^ (r & 7)): the tensor core reads eight rows of one k-chunk from eight different bank groups
__host__ __device__ __forceinline__ int64_t i8_tiled(int64_t row, int64_t k, int64_t K8) {
    return (((row >> 7) * (K8 >> 7) + (k >> 7)) << 14) + ((row & 127) << 7) + ((((k >> 4) & 7) ^ (row & 7)) << 4) + (k & 15);
}

// ---- PTX wrappers --------------------------------------------------------------------------------------------
__device__ __forceinline__ void i8_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void i8_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, 128-byte swizzle (cute::UMMA::SmemDescriptor): start address >> 4, leading
// byte offset (unused by swizzled K-major layouts: 1), stride byte offset (between 8-row groups) >> 4, version 1 =
// Blackwell, layout type 2 = SWIZZLE_128B.  The four K = 32 slices of a 128-byte row are reached by advancing the
// start address by 32 bytes (the tile is 1024-byte aligned, so the swizzle phase is that of the address bits).
__device__ __forceinline__ uint64_t i8_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((I8_SBO >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = S32, A = unsigned 8 bit, B = signed 8 bit, both K-major,
// N = 128, M = 128, dense, no saturation
__device__ __forceinline__ uint32_t i8_instr_desc() {
    return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(I8_TN >> 3) << 17) | ((uint32_t)(I8_TM >> 4) << 24);
}
__device__ __forceinline__ void i8_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void i8_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(vgs_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void i8_tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// plain spin on an mbarrier phase (unique labels: this is inlined at several call sites)
__device__ __forceinline__ void i8_mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(vgs_smem_addr(bar)), "r"(parity)
            : "memory");
    }
}

// the eight exact digit sums of one (sequence, model) -> fp64
__device__ __forceinline__ double i8_fold(const long long *S) {
    const long long lo = S[0] + S[1] * 256 + S[2] * 65536 + S[3] * 16777216;
    const long long hi = S[4] + S[5] * 256 + S[6] * 65536 + S[7] * 16777216;
    const double v = __fma_rn((double)hi, 4294967296.0, (double)lo);
    return v * (1.0 / 2251799813685248.0);   // 2^-51, exact
}

// fixed-point form of one table entry and its digits (the same rounding everywhere)
__device__ __forceinline__ long long i8_quantise(double t) { return __double2ll_rn(t * 2251799813685248.0); }

// adds 256 * hi * digits(T[m][k]) for the frequent k-mers of sequence s
__device__ __forceinline__ void i8_add_frequent(long long *S, const uint2 *__restrict__ list, int n_over, const double *__restrict__ Trow) {
    for (int o = 0; o < n_over; ++o) {
        const uint2 e = list[o];
        long long q = i8_quantise(Trow[e.x]);
        const long long w = 256ll * (long long)e.y;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const long long d = ((q + 128) & 255) - 128;
            S[p] += w * d;
            q = (q - d) >> 8;
        }
    }
}

// ---- operand builders ----------------------------------------------------------------------------------------
// digits[tiled(m * 8 + p, k)] = p-th balanced base-256 digit of round(T[m][k] * 2^51)
__global__ void k_i8_digits(int64_t m_base, int64_t K, int64_t K8, const double *__restrict__ T, int8_t *__restrict__ digits,
                            int *__restrict__ flag) {
    const int64_t m = m_base + blockIdx.y;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < K8; k += (int64_t)gridDim.x * blockDim.x) {
        long long q = 0;
        if (k < K) {
            const double t = T[m * K + k];
            if (!(fabs(t) <= 1000.0)) atomicOr(flag, 1);   // not a log-probability (or NaN): no fixed-point form
            else q = i8_quantise(t);
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const long long d = ((q + 128) & 255) - 128;
            digits[i8_tiled(m * 8 + p, k, K8)] = (int8_t)d;
            q = (q - d) >> 8;
        }
    }
}

// per-sequence (D+1)-mer histogram, low bytes as the uint8 operand (one CTA per sequence); k-mers counted more than 255
// times go to the sequence's list {k, count >> 8}; more than I8_MAX_OVER of them raise the flag
// PACK16: two 16-bit counters per shared-memory word (every sequence shorter than 65536): half the shared memory,
// twice the resident CTAs
template <bool PACK16>
__global__ void k_i8_hist(int depth, const uint32_t *__restrict__ seq, const int64_t *__restrict__ seq_off,
                          const int64_t *__restrict__ seq_len, int64_t K8, uint8_t *__restrict__ C8, int32_t *__restrict__ over_cnt,
                          uint2 *__restrict__ over_list, int *__restrict__ flag) {
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
    const int64_t span = (L + blockDim.x - 1) / blockDim.x;
    const int64_t a = (int64_t)threadIdx.x * span, e = min(L, a + span);
    if (a < e) {
        int64_t t = max((int64_t)0, a - depth);
        uint32_t code = 0;
        for (; t < e; ++t) {
            const uint32_t x = (w[t >> 4] >> (30 - 2 * (int)(t & 15))) & 3u;
            code = ((code << 2) | x) & kmask;
            if (t >= a && t >= depth) {
                if (PACK16) atomicAdd(&hist[code >> 1], (code & 1u) ? 0x10000u : 1u);
                else atomicAdd(&hist[code], 1u);
            }
        }
    }
    __syncthreads();
    // one 16-byte k-chunk (16 counts) per thread and store, tiled layout (K8 is a multiple of 128)
    const int64_t n_chunks = K8 >> 4;
    for (int64_t ch = threadIdx.x; ch < n_chunks; ch += blockDim.x) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (16 * ch < K) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t cs[4];
                if (PACK16) {
                    const uint2 c2 = *(const uint2 *)&hist[8 * ch + 2 * q];
                    cs[0] = c2.x & 0xFFFFu; cs[1] = c2.x >> 16; cs[2] = c2.y & 0xFFFFu; cs[3] = c2.y >> 16;
                } else {
                    const uint4 c4 = *(const uint4 *)&hist[16 * ch + 4 * q];
                    cs[0] = c4.x; cs[1] = c4.y; cs[2] = c4.z; cs[3] = c4.w;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (cs[j] > 255u) {
                        const int idx = atomicAdd(&n_over, 1);
                        if (idx < I8_MAX_OVER) over_list[s * I8_MAX_OVER + idx] = make_uint2((uint32_t)(16 * ch + 4 * q + j), cs[j] >> 8);
                    }
                    w[q] |= (cs[j] & 255u) << (8 * j);
                }
            }
        }
        *(uint4 *)(C8 + i8_tiled(s, 16 * ch, K8)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        over_cnt[s] = n_over < I8_MAX_OVER ? n_over : I8_MAX_OVER;
        if (n_over > I8_MAX_OVER) atomicOr(flag, 2);
    }
}

struct I8Params {
    const uint8_t *C8;      // [n_seq][K8]
    const int8_t *digits;   // [n_models * 8][K8]
    const int32_t *over_cnt;  // [n_seq] frequent k-mers per sequence
    const uint2 *over_list;   // [n_seq][I8_MAX_OVER] {k, count >> 8}
    const double *T;          // [n_models][K] fp64 tables (source of the digits; read for the frequent k-mers)
    int64_t K;
    const int32_t *pairs;   // [n_pairs][2] (sequence 128-block, model 128-block)
    int n_pairs;
    int split;              // k-split: each tile's K range is cut into `split` units (1 = none)
    int32_t *P;             // split > 1: exact int32 digit sums [n_seq][n_models * 8], added atomically, zero between calls
    int64_t K8, n_seq, n_models, ld;
    double *M;
    int skip_diag, depth;
    const ModelMeta *meta;
    const uint8_t *arena;
    const uint32_t *seq;
    const int64_t *seq_off, *seq_len;
};

// terms t in [order_m, D) of a model shallower than D (positions the (D+1)-mer histogram does not cover)
__device__ __forceinline__ double i8_head(const I8Params &p, const ModelMeta &mm, int64_t s) {
    if (mm.order >= p.depth) return 0.0;
    const double *tab = (const double *)(p.arena + mm.nll_off);
    const uint32_t fmask = vgs_lowmask(mm.order + 1);
    const uint32_t *w = p.seq + p.seq_off[s];
    const int64_t L = p.seq_len[s];
    const int64_t top = L < p.depth ? L : p.depth;
    uint32_t code = 0;
    double head = 0.0;
    for (int64_t t = 0; t < top; ++t) {
        const uint32_t x = (w[t >> 4] >> (30 - 2 * (int)(t & 15))) & 3u;
        code = (code << 2) | x;
        if (t >= mm.order) head += tab[code & fmask];
    }
    return head;
}

template <int NACC, int STAGES>
__global__ void __launch_bounds__(I8_THREADS, I8_CTAS_PER_SM) k_kmer_i8mma(const I8Params p) {
    constexpr int STAGE_BYTES = (1 + NACC) * I8_OPERAND_BYTES;   // the count tile + one digit tile per accumulator
    constexpr int TMEM_COLS = NACC * I8_TN;
    extern __shared__ __align__(1024) uint8_t i8_smem[];
    __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = vgs_smem_addr(i8_smem);

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { vgs_mbar_init(&full[i], 1); vgs_mbar_init(&empty[i], 1); }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(vgs_smem_addr(&tmem_base_slot)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    i8_tc_fence_before();
    __syncthreads();
    i8_tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t idesc = i8_instr_desc();
    const int n_steps_total = (int)(p.K8 / I8_BK);
    const int64_t tiles_per_row_block = p.K8 >> 7;       // 16 KB tiles per 128-row block
    // a 128 x 128-model block pair = 8 sub-tiles of 16 models; a CTA tile takes NACC of them (one accumulator each,
    // all fed by the same count tile).  Consecutive MMAs go to different accumulators: four independent accumulation
    // chains keep the tensor pipe busy where a single chain waits for its own previous MMA.
    const int tiles_per_pair = 8 / NACC;
    const int n_units = p.n_pairs * tiles_per_pair * p.split;
    int64_t g0 = 0;                           // global stage counter (continues across tiles: mbarrier parities)

    // k-split (launches with fewer tiles than CTA slots: multi-GPU shares, small sets): a unit is one part of a
    // tile's K range; its digit sums are added to P with integer atomics -- exact and order-independent, so the
    // result has the same bits as the unsplit tile -- and k_i8_fold turns P into M afterwards.
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int tile = unit / p.split, part = unit % p.split;
        const int step_begin = (int)((int64_t)n_steps_total * part / p.split);
        const int n_steps = (int)((int64_t)n_steps_total * (part + 1) / p.split) - step_begin;
        const int pair = tile / tiles_per_pair;
        const int64_t sb = p.pairs[2 * pair];
        const int64_t s0 = sb * 128;
        const int64_t m0 = (int64_t)p.pairs[2 * pair + 1] * 128 + (int64_t)(tile % tiles_per_pair) * (NACC * I8_MODELS_PER_TILE);
        if (m0 >= p.n_models) continue;                  // the last 128-model block of the set may be partly empty
        // sub-tiles that hold models (the rest of the last block is skipped: no copies, no MMAs, no output)
        int n_acc = (int)((p.n_models - m0 + I8_MODELS_PER_TILE - 1) / I8_MODELS_PER_TILE);
        if (n_acc > NACC) n_acc = NACC;
        const int64_t nb = m0 / I8_MODELS_PER_TILE;      // first 128-row block of the digit operand (16 models x 8 digits)

        if (tid == 0) {
            const uint8_t *a_src = p.C8 + ((sb * tiles_per_row_block) << 14);
            const uint8_t *b_src = (const uint8_t *)p.digits + ((nb * tiles_per_row_block) << 14);
            const size_t b_block_stride = (size_t)tiles_per_row_block << 14;
            auto load = [&](int64_t g, int step) {
                const int stage = (int)(g % STAGES);
                if (g >= STAGES) i8_mbar_wait(&empty[stage], (uint32_t)((g / STAGES - 1) & 1));   // MMAs of the slot's last use
                uint8_t *sa = i8_smem + (size_t)stage * STAGE_BYTES;
                vgs_mbar_expect_tx(&full[stage], (uint32_t)((1 + n_acc) * I8_OPERAND_BYTES));
                vgs_bulk_g2s(sa, a_src + (size_t)(step_begin + step) * I8_OPERAND_BYTES, I8_OPERAND_BYTES, &full[stage]);
                for (int b = 0; b < n_acc; ++b)
                    vgs_bulk_g2s(sa + (size_t)(1 + b) * I8_OPERAND_BYTES, b_src + b * b_block_stride + (size_t)(step_begin + step) * I8_OPERAND_BYTES,
                                 I8_OPERAND_BYTES, &full[stage]);
            };
            for (int i = 0; i < STAGES - 1 && i < n_steps; ++i) load(g0 + i, i);
            for (int step = 0; step < n_steps; ++step) {
                const int nxt = step + STAGES - 1;
                if (nxt < n_steps) load(g0 + nxt, nxt);   // refill first: the copies get the longest head start
                const int64_t g = g0 + step;
                const int stage = (int)(g % STAGES);
                i8_mbar_wait(&full[stage], (uint32_t)((g / STAGES) & 1));
                i8_tc_fence_after();
                const uint32_t sa = smem0 + (uint32_t)stage * STAGE_BYTES;
#pragma unroll
                for (int j = 0; j < I8_BK / 32; ++j) {
                    const uint64_t da = i8_smem_desc(sa + (uint32_t)(32 * j));
#pragma unroll
                    for (int b = 0; b < NACC; ++b) {
                        if (b < n_acc) {
                            const uint64_t db = i8_smem_desc(sa + (uint32_t)(1 + b) * I8_OPERAND_BYTES + (uint32_t)(32 * j));
                            i8_mma(tmem_base + (uint32_t)(b * I8_TN), da, db, idesc, (step > 0 || j > 0) ? 1u : 0u);
                        }
                    }
                }
                i8_commit(&empty[stage]);
            }
            // all MMAs of this tile done?
            const int64_t g_last = g0 + n_steps - 1;
            i8_mbar_wait(&empty[(int)(g_last % STAGES)], (uint32_t)((g_last / STAGES) & 1));
        }
        __syncthreads();                       // the other threads sleep here during the main loop
        i8_tc_fence_after();
        // epilogue: thread = one sequence row (TMEM lane), 16 columns (two models) at a time
        const int64_t s = s0 + 32 * warp + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
#pragma unroll 1
        for (int j = 0; j < n_acc * (I8_TN / 16); ++j) {
            int32_t v[16];
            i8_tmem_ld16(taddr + (uint32_t)(16 * j), v);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int64_t m = m0 + 2 * j + q;
                if (s >= p.n_seq || m >= p.n_models) continue;
                if (p.split > 1) {
                    int32_t *dst = p.P + (s * p.n_models + m) * 8;
#pragma unroll
                    for (int d = 0; d < 8; ++d) atomicAdd(dst + d, v[8 * q + d]);
                } else if (!(p.skip_diag && s == m)) {
                    long long S[8];
#pragma unroll
                    for (int d = 0; d < 8; ++d) S[d] = (long long)v[8 * q + d];
                    p.M[s * p.ld + m] = i8_head(p, p.meta[m], s) + i8_fold(S);   // + k_i8_frequent for rows that need it
                }
            }
        }
        i8_tc_fence_before();
        __syncthreads();                       // TMEM is free for the next tile's first (overwriting) MMAs
        g0 += n_steps;
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// k-split launches: P -> M for every entry of the computed block pairs, and P back to zero for the next call
__global__ void __launch_bounds__(128) k_i8_fold(const I8Params p) {
    const int pair = blockIdx.x >> 7, r = blockIdx.x & 127;
    const int64_t s = (int64_t)p.pairs[2 * pair] * 128 + r;
    const int64_t m = (int64_t)p.pairs[2 * pair + 1] * 128 + threadIdx.x;
    if (s >= p.n_seq || m >= p.n_models) return;
    int4 *src = (int4 *)(p.P + (s * p.n_models + m) * 8);
    const int4 lo = src[0], hi = src[1];
    src[0] = make_int4(0, 0, 0, 0);
    src[1] = make_int4(0, 0, 0, 0);
    if (p.skip_diag && s == m) return;
    const long long S[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    p.M[s * p.ld + m] = i8_head(p, p.meta[m], s) + i8_fold(S);
}

// second term of the sequences that have frequent k-mers: M[s][m] += fold(256 * sum_o hi_o * digits(T[m][k_o])).
// One CTA per (block pair, sequence row); rows without frequent k-mers leave at once.  The term is an exact integer
// sum folded once, so M[s][m] = fold(low bytes) + fold(high parts) is the same bits wherever it is computed.
__global__ void __launch_bounds__(128) k_i8_frequent(const I8Params p) {
    const int pair = blockIdx.x >> 7, r = blockIdx.x & 127;
    const int64_t s = (int64_t)p.pairs[2 * pair] * 128 + r;
    if (s >= p.n_seq) return;
    const int n_over = p.over_cnt[s];
    if (n_over == 0) return;
    const int64_t m = (int64_t)p.pairs[2 * pair + 1] * 128 + threadIdx.x;
    if (m >= p.n_models || (p.skip_diag && s == m)) return;
    long long S[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    i8_add_frequent(S, p.over_list + s * I8_MAX_OVER, n_over, p.T + m * p.K);
    p.M[s * p.ld + m] += i8_fold(S);
}

// diagonal scores M[m][m] with the same arithmetic (one CTA per model): identical bits to the GEMM's entry
__global__ void __launch_bounds__(256) k_i8_diag(const I8Params p) {
    __shared__ int part[8][8];
    __shared__ unsigned long long corr[8];
    const int64_t m = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 8) corr[threadIdx.x] = 0ull;
    int acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // one 16-byte k-chunk per thread and step: the chunk of the count row and of the eight digit rows (consecutive rows
    // of one 128-row block: 8 x 16 contiguous bytes)
    const int64_t n_chunks = p.K8 >> 4;
    const int64_t dr = m * 8;
    for (int64_t ch = threadIdx.x; ch < n_chunks; ch += blockDim.x) {
        const uint4 cv = *(const uint4 *)(p.C8 + i8_tiled(m, 16 * ch, p.K8));
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint4 dv = *(const uint4 *)((const uint8_t *)p.digits + i8_tiled(dr + q, 16 * ch, p.K8));
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
    // frequent k-mers of this sequence: exact integer sums, any order
    const int n_over = p.over_cnt[m];
    if ((int)threadIdx.x < n_over) {
        long long S[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        i8_add_frequent(S, p.over_list + m * I8_MAX_OVER + threadIdx.x, 1, p.T + m * p.K);
#pragma unroll
        for (int q = 0; q < 8; ++q) atomicAdd(&corr[q], (unsigned long long)S[q]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long S[8], Sc[8];
        for (int q = 0; q < 8; ++q) {
            int t = 0;
            for (int w2 = 0; w2 < 8; ++w2) t += part[w2][q];
            S[q] = (long long)t;
            Sc[q] = (long long)corr[q];
        }
        double v = i8_head(p, p.meta[m], m) + i8_fold(S);
        if (n_over) v += i8_fold(Sc);
        p.M[m * p.ld + m] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------
static int64_t i8_k8(vgs_context *h) { return std::max<int64_t>(I8_BK, (int64_t)1 << (2 * (h->max_order + 1))); }

// builds (once per model set) the digit operand from the fp64 tables; returns VGS_OK, an error, or -1 = not applicable
int vgs_kmer_i8_prepare_tables(vgs_context *h, const double *T, int64_t K, cudaStream_t s) {
    if (h->i8_tables_gen == h->model_gen) return h->i8_tables_ok ? VGS_OK : -1;
    const int64_t K8 = i8_k8(h);
    int rc;
    const int64_t digit_rows = (h->n_models * 8 + 127) / 128 * 128;
    if ((rc = vgs_reserve(h, &h->d_i8_digits, digit_rows * K8))) return rc;
    VGS_CUDA(cudaMemsetAsync(h->d_i8_digits, 0, (size_t)(digit_rows * K8), s));   // rows of the last block beyond the set
    if ((rc = vgs_reserve(h, &h->d_counters, 16))) return rc;
    VGS_CUDA(cudaMemsetAsync(h->d_counters + 9, 0, sizeof(int), s));
    for (int64_t m0 = 0; m0 < h->n_models; m0 += 65535) {
        dim3 grid((unsigned)std::min<int64_t>((K8 + 255) / 256, 64), (unsigned)std::min<int64_t>(h->n_models - m0, 65535));
        k_i8_digits<<<grid, 256, 0, s>>>(m0, K, K8, T, h->d_i8_digits, h->d_counters + 9);
        VGS_LAUNCH_CHECK(h);
    }
    int flag = 0;
    VGS_CUDA(cudaMemcpyAsync(&flag, h->d_counters + 9, sizeof(int), cudaMemcpyDeviceToHost, s));
    VGS_CUDA(cudaStreamSynchronize(s));
    h->i8_tables_ok = flag == 0;
    h->i8_tables_gen = h->model_gen;
    return h->i8_tables_ok ? VGS_OK : -1;
}

// scores for the given (sequence-block, model-block) pairs at 128 granularity (d_pairs on the device).
// Returns VGS_OK, an error, or -1 when a sequence has too many frequent k-mers (caller uses the fp64 path).
int vgs_kmer_i8_scores(vgs_context *h, const int32_t *d_pairs, int n_pairs, bool diag_separately, const double *T, int64_t K,
                       double *M, int64_t ld, cudaStream_t s) {
    const int D = h->max_order;
    const int64_t K8 = i8_k8(h);
    int rc;
    const int64_t count_rows = (h->n_seq + 127) / 128 * 128;
    if ((rc = vgs_reserve(h, &h->d_i8_C, count_rows * K8))) return rc;
    if ((rc = vgs_reserve(h, &h->d_i8_over, h->n_seq * (I8_MAX_OVER * 2 + 1) + 2))) return rc;
    int32_t *over_cnt = (int32_t *)h->d_i8_over;
    uint2 *over_list = (uint2 *)(h->d_i8_over + ((h->n_seq + 1) / 2) * 2);
    if ((rc = vgs_reserve(h, &h->d_counters, 16))) return rc;
    // histograms of the resident sequences: part of every call (a pass over the packed sequences)
    int64_t max_len = 0;
    for (int64_t i = 0; i < h->n_seq; ++i) max_len = std::max(max_len, h->h_seq_len[(size_t)i]);
    const bool pack16 = max_len < 65536;
    const int hist_threads = max_len >= 8192 ? 512 : 256;   // shorter per-thread symbol chains on long sequences
    const int hist_smem = (int)(std::max<int64_t>(16, (int64_t)1 << (2 * (D + 1))) * (pack16 ? 2 : 4));
    VGS_CUDA(cudaFuncSetAttribute(k_i8_hist<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_smem));
    VGS_CUDA(cudaFuncSetAttribute(k_i8_hist<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_smem));
    if (h->i8_counts_state == 0) {
        VGS_CUDA(cudaMemsetAsync(h->d_counters + 10, 0, sizeof(int), s));
        VGS_CUDA(cudaMemsetAsync(h->d_i8_C, 0, (size_t)(count_rows * K8), s));   // rows of the last block beyond the set
    }
    if (pack16)
        k_i8_hist<true><<<(unsigned)h->n_seq, hist_threads, hist_smem, s>>>(D, h->d_seq, h->d_seq_off, h->d_seq_len, K8, h->d_i8_C, over_cnt,
                                                                   over_list, h->d_counters + 10);
    else
        k_i8_hist<false><<<(unsigned)h->n_seq, hist_threads, hist_smem, s>>>(D, h->d_seq, h->d_seq_off, h->d_seq_len, K8, h->d_i8_C, over_cnt,
                                                                    over_list, h->d_counters + 10);
    VGS_LAUNCH_CHECK(h);
    if (h->i8_counts_state == 0) {
        // first call on this sequence set: few enough frequent k-mers per sequence?  (one synchronisation per loaded set)
        int flag = 0;
        VGS_CUDA(cudaMemcpyAsync(&flag, h->d_counters + 10, sizeof(int), cudaMemcpyDeviceToHost, s));
        VGS_CUDA(cudaStreamSynchronize(s));
        h->i8_counts_state = flag ? 2 : 1;
    }
    if (h->i8_counts_state == 2) return -1;
    I8Params p;
    p.C8 = h->d_i8_C; p.digits = h->d_i8_digits; p.pairs = d_pairs; p.n_pairs = n_pairs;
    p.over_cnt = over_cnt; p.over_list = over_list; p.T = T; p.K = K;
    p.K8 = K8; p.n_seq = h->n_seq; p.n_models = h->n_models; p.ld = ld; p.M = M;
    p.skip_diag = diag_separately ? 1 : 0; p.depth = D;
    p.meta = h->d_meta; p.arena = h->d_nll; p.seq = h->d_seq; p.seq_off = h->d_seq_off; p.seq_len = h->d_seq_len;
    if (n_pairs > 0) {
        // two accumulators per CTA (128 x 256 tiles, less operand traffic) when that still fills the CTA slots, else
        // one (twice as many tiles of half the latency: multi-GPU shares, small sets)
        const int slots = I8_CTAS_PER_SM * h->sm_count;
        const bool wide = (int64_t)n_pairs * 4 * 4 >= (int64_t)slots * 3;
        // fewer 128 x 128 tiles than CTA slots: cut every tile's K range so that the units fill the slots (at most 4
        // parts, at least 8 pipeline steps each); needs the int32 sum buffer, kept below 1 GB
        p.split = 1;
        p.P = nullptr;
        static int want_split = -1;
        if (want_split < 0) { const char *e = getenv("VGS_I8_SPLITK"); want_split = (e && e[0] == '0') ? 0 : 1; }
        const int64_t p_elems = h->n_seq * h->n_models * 8;
        if (!wide && want_split && n_pairs * 8 * 2 <= slots && p_elems * 4 <= ((int64_t)1 << 30)) {
            int split = std::min(4, slots / (n_pairs * 8));
            split = (int)std::min<int64_t>(split, std::max<int64_t>(1, (K8 / I8_BK) / 8));
            if (split > 1) {
                int32_t *before = h->d_i8_P;
                if ((rc = vgs_reserve(h, &h->d_i8_P, p_elems))) return rc;
                if (h->d_i8_P != before || h->i8_P_elems != p_elems) {   // fresh or re-shaped buffer: all zero once
                    VGS_CUDA(cudaMemsetAsync(h->d_i8_P, 0, (size_t)p_elems * 4, s));
                    h->i8_P_elems = p_elems;
                }
                p.split = split;
                p.P = h->d_i8_P;
            }
        }
        h->last_nll_kernel = VGS_NLL_KERNEL_I8MMA;
        vgs_prof_begin(h, VGS_PROF_NLL_SCORES, s);
        if (wide) {
            const int smem = 2 * (1 + 2) * I8_OPERAND_BYTES;
            VGS_CUDA(cudaFuncSetAttribute(k_kmer_i8mma<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            k_kmer_i8mma<2, 2><<<std::min(n_pairs * 4, slots), I8_THREADS, smem, s>>>(p);
        } else {
            const int smem = 3 * (1 + 1) * I8_OPERAND_BYTES;
            VGS_CUDA(cudaFuncSetAttribute(k_kmer_i8mma<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            k_kmer_i8mma<1, 3><<<std::min(n_pairs * 8 * p.split, slots), I8_THREADS, smem, s>>>(p);
        }
        vgs_prof_end(h, VGS_PROF_NLL_SCORES, s);
        VGS_LAUNCH_CHECK(h);
        if (p.split > 1) {
            k_i8_fold<<<(unsigned)(n_pairs * 128), 128, 0, s>>>(p);
            VGS_LAUNCH_CHECK(h);
        }
        k_i8_frequent<<<(unsigned)(n_pairs * 128), 128, 0, s>>>(p);
        VGS_LAUNCH_CHECK(h);
    }
    if (diag_separately) {
        const int64_t nd = std::min(h->n_seq, h->n_models);
        k_i8_diag<<<(unsigned)nd, 256, 0, s>>>(p);
        VGS_LAUNCH_CHECK(h);
    }
    return VGS_OK;
}
