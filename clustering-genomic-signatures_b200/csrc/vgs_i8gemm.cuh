// vgs_i8gemm.cuh -- exact integer GEMM core on the 5th-generation tensor cores, CTA pairs (tcgen05.mma cta_group::2
// kind::i8, M = 256, accumulators in tensor memory), shared by the NLL k-mer contraction (vgs_kmer_i8.cu) and the dense
// Frobenius / Projection path (vgs_pairs_tc.cu).
//
//   S[r][c] = sum_k A[r][k] * B[c][k]      A: unsigned or signed 8 bit, B: signed 8 bit, S: exact int32
//
// Why this shape.  The first version (cta_group::1, 128 x 256 CTA tiles, two CTAs per SM) moved 0.0117 operand bytes per
// MAC out of L2 and ran at the L2 -> SM delivery rate (2.1 GB per cfg-3 launch = 13.6 TB/s, 44 % of the int8 MMA rate).
// A CTA pair on a 256 x N tile needs (256 + N) / (256 N) bytes per MAC (0.0078 at N = 256) and each SM stages only its
// own half of either operand.
//
// Operand layout in HBM ("slab" layout): for every block of 128 consecutive k (= 128 bytes per row) all rows are stored
// one after the other as the shared-memory image of the canonical K-major SWIZZLE_128B UMMA layout:
//     offset(row, k) = ((k / 128) * rows_pad + row) * 128 + (((k / 16) % 8) ^ (row % 8)) * 16 + k % 16
// so ANY run of rows that starts at a multiple of 8 is, per k-block, one contiguous piece that a single 1-D bulk async
// copy (TMA, SASS UBLKCP) drops into a 1024-byte aligned stage buffer -- no tensor maps, no layout work in the kernel,
// and the tile width N can be chosen per launch (multiples of 32) so that the tile count divides the CTA pairs evenly.
//
// Kernel: one CTA pair (cluster of 2) per SM pair, persistent over a static list of units (tile x k range); 576 threads:
//   warp 0   producer: per stage one expect_tx + two bulk copies (its 128 A rows, its half of the B rows)
//   warp 1   leader CTA: the one thread that issues every tcgen05.mma of the pair and commits them to the stage's
//            "empty" barriers of BOTH CTAs (multicast commit); peer CTA: relays "my half of the stage has landed" to
//            the leader's "full" barrier (relaxed remote mbarrier arrive: the 1-D bulk copy can only signal a barrier
//            of the CTA it writes to)
//   warps 2-17 epilogue: tcgen05.ld of their 32 TMEM lanes, 32 columns at a time, handed to the caller's functor (four
//            warps per lane quadrant take every fourth chunk).  A unit is up to 512 columns wide = both halves of tensor
//            memory, fed by one A tile (see the kernel's comment).
#pragma once
#include <stdio.h>
#include <stdlib.h>

#include "vgs_common.cuh"

#define G8_BK 128                 // k per pipeline stage = bytes per operand row per stage = one swizzle atom
#define G8_STAGES 4               // stages at the widest units (4 x 48 KB = 192 KB of the 227 KB a CTA can have)
#define G8_MAX_STAGES 8           // barrier slots: narrower granules / units leave room for more, smaller stages
#define G8_EPI_WARPS 16           // four per TMEM lane quadrant
#define G8_THREADS (32 * (2 + G8_EPI_WARPS))
#define G8_A_ROWS 128             // A rows per CTA (256 per pair)
#define G8_MAX_N 256              // B rows per MMA / accumulator (128 per CTA)
#define G8_UNIT_MAX_N 512         // B rows per unit: two accumulators
#define G8_A_BYTES (G8_A_ROWS * G8_BK)
#define G8_B_BYTES ((G8_MAX_N / 2) * G8_BK)                    // one accumulator's B half-tile per stage and CTA
#define G8_STAGE_BYTES (G8_A_BYTES + 2 * G8_B_BYTES)           // 48 KB
#define G8_SMEM_BYTES (G8_STAGES * G8_STAGE_BYTES)
#define G8_SMEM_LIMIT (226 * 1024)   // dynamic shared memory the kernel may ask for (227 KB per CTA minus its static 1 KB)
#define G8_TMEM_COLS 512
#define G8_WATCHDOG_CYCLES (1ll << 32)   // ~2 s: a lost barrier signal traps instead of hanging the GPU

struct G8Unit {
    int32_t a_blk;    // 256-row block of A
    int32_t b_row0;   // first B row, multiple of 8 (16 for the peer's half to stay 8-aligned: n / 2 is a multiple of 16)
    int32_t n;        // B rows of this unit: multiple of the launch's granule (32 unless the functor's chunks are narrower),
                      // <= 512 (more than one accumulator's worth: two accumulators, see g8_first_half)
    int32_t kb0, kbn; // k range in 128-blocks
    int32_t partial;  // != 0: this unit is part (partial - 1) of a k-split (the functor stores its sums to the part's slab)
};

struct G8Args {
    const uint8_t *A, *B;
    int64_t a_rows_pad, b_rows_pad;   // rows per k-block slab (multiples of 256 / 32)
    const G8Unit *units;
    int n_units;
    int a_signed;                     // A operand: 0 = unsigned 8 bit, 1 = signed
    int *err;                         // device error word (bit 2: watchdog)
    int debug;                        // developer knob VGS_G8_DEBUG: bit 0 = skip the epilogue functor (timing experiments only)
    int stages;                       // pipeline stages in use (set by g8_launch; tuning knob VGS_G8_STAGES caps it)
    int b_bytes;                      // bytes of one accumulator's B half-tile in a stage = (largest accumulator / 2) rows x 128
    int stage_bytes;                  // A tile + two B half-tiles: 16 KB + 2 b_bytes (a multiple of 1024)
    int gran;                         // granule of the unit widths (multiple of 32 and of the functor's chunk width)
    long long *prof;                  // optional [n_clusters][8] cycle counters of the leader CTA (VGS_G8_PROF=1), else nullptr
};

// a unit wider than one accumulator (the largest multiple of the granule <= 256 B rows) is computed as two MMAs per
// k-slice: the first g8_first_half(n, gran) rows, then the rest -- both multiples of the granule, so that no chunk of the
// epilogue functor straddles the two accumulators
__host__ __device__ __forceinline__ int g8_max_acc(int gran) { return (G8_MAX_N / gran) * gran; }
__host__ __device__ __forceinline__ int g8_first_half(int n, int gran) {
    const int max_acc = g8_max_acc(gran);
    if (n <= max_acc) return n;
    const int half = ((n / 2 + gran - 1) / gran) * gran;
    return half < max_acc ? half : max_acc;
}
__host__ __device__ __forceinline__ int64_t g8_off(int64_t row, int64_t k, int64_t rows_pad) {
    return (((k >> 7) * rows_pad + row) << 7) + ((((k >> 4) & 7) ^ (row & 7)) << 4) + (k & 15);
}

#ifdef __CUDACC__
// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ void g8_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void g8_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t g8_cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void g8_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4, leading byte
// offset 1 (unused by swizzled K-major layouts), stride byte offset 1024 >> 4 (between 8-row groups), version 1
// (Blackwell), layout type 2 (SWIZZLE_128B).  The four K = 32 slices of a 128-byte row: start address + 32 j.
__device__ __forceinline__ uint64_t g8_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 at bit 4), A format at bit 7 (0 unsigned / 1 signed 8
// bit), B signed (1 at bit 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t g8_idesc(int m, int n, int a_signed) {
    return (2u << 4) | ((uint32_t)(a_signed ? 1 : 0) << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ uint64_t g8_desc_join(uint32_t lo, uint32_t hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
__device__ __forceinline__ void g8_mma2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void g8_mma1(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (once) on the mbarrier at this shared-memory offset in BOTH CTAs of the pair when every tcgen05.mma issued so
// far by this thread has completed
__device__ __forceinline__ void g8_commit2(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     vgs_smem_addr(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void g8_commit1(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(vgs_smem_addr(bar)) : "memory");
}
// arrive on the mbarrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void g8_remote_arrive(uint64_t *bar, uint32_t rank) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
        "}\n" ::"r"(vgs_smem_addr(bar)),
        "r"(rank)
        : "memory");
}
// relay arrival: "the bulk copies into my shared memory have completed" -- the data itself was made visible by the copy's
// own complete_tx, so the arrival needs no release fence (a release.cluster here costs ~1000 cycles per stage and made
// the relay the bottleneck of the whole pipeline)
__device__ __forceinline__ void g8_remote_arrive_relaxed(uint64_t *bar, uint32_t rank) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t"
        "}\n" ::"r"(vgs_smem_addr(bar)),
        "r"(rank)
        : "memory");
}
__device__ __forceinline__ void g8_local_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(vgs_smem_addr(bar)) : "memory");
}
// wait for a phase of an mbarrier.  CLUSTER: acquire at cluster scope (the arrival is a release from the other CTA's
// threads); otherwise the default CTA-scope acquire (arrivals from this CTA's threads, from TMA completions and from
// tcgen05.commit, and relaxed relay arrivals that carry no data of their own).  A signal that never comes sets the
// error word and traps -- a failed launch, not a hung GPU.
template <bool CLUSTER>
__device__ __forceinline__ void g8_wait_t(uint64_t *bar, uint32_t parity, int *err) {
    uint32_t done = 0;
    long long t0 = 0;
    int spins = 0;
    while (true) {
        if (CLUSTER)
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t"
                "}\n"
                : "=r"(done)
                : "r"(vgs_smem_addr(bar)), "r"(parity)
                : "memory");
        else
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t"
                "}\n"
                : "=r"(done)
                : "r"(vgs_smem_addr(bar)), "r"(parity)
                : "memory");
        if (done) break;
        if (++spins == 1024) t0 = clock64();
        if (spins > 1024 && (spins & 1023) == 0 && clock64() - t0 > G8_WATCHDOG_CYCLES) {
            if (err) atomicOr(err, 4);
            __threadfence_system();
            __trap();
        }
    }
}
__device__ __forceinline__ void g8_wait(uint64_t *bar, uint32_t parity, int *err) { g8_wait_t<false>(bar, parity, err); }
__device__ __forceinline__ void g8_wait_cluster(uint64_t *bar, uint32_t parity, int *err) { g8_wait_t<true>(bar, parity, err); }
// this warp's 32 TMEM lanes x W consecutive columns -> registers (no wait: g8_tmem_ld ends a chunk's loads with one)
__device__ __forceinline__ void g8_tmem_ld_x32(uint32_t taddr, int32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void g8_tmem_ld_x8(uint32_t taddr, int32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void g8_tmem_ld_x4(uint32_t taddr, int32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
// W columns (multiple of 4, <= 32) starting at taddr: one x32, or 8- and 4-column pieces, then ONE wait
template <int W>
__device__ __forceinline__ void g8_tmem_ld(uint32_t taddr, int32_t (&v)[W]) {
    static_assert(W % 4 == 0 && W >= 4 && W <= 32, "chunk width");
    if (W == 32) {
        g8_tmem_ld_x32(taddr, v);
    } else {
#pragma unroll
        for (int c = 0; c + 8 <= W; c += 8) g8_tmem_ld_x8(taddr + (uint32_t)c, v + c);
        if (W % 8) g8_tmem_ld_x4(taddr + (uint32_t)(W - 4), v + (W - 4));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- the kernel ---------------------------------------------------------------------------------------------
// Epi: struct with   static constexpr int CHUNK = columns per operator() call (multiple of 4, <= 32; the launch's granule is a
//                                         multiple of it)
//                    struct State;  __device__ State prepare(const G8Unit &u, int64_t a_row) const   (whole warp, while the unit's
//                                         MMAs still run: every global load the chunks' operator() would otherwise wait for -- the
//                                         epilogue is exposed time, a dependent L2 round trip per chunk in it cost 13 of 93 us)
//                    __device__ void prefetch(const G8Unit &u, int64_t a_row, int col0, const State &) const   (whole warp, before
//                                         the accumulator is ready: warm the lines the chunk's operator() will read)
//                    __device__ void operator()(const G8Unit &u, int64_t a_row, int col0, const int32_t (&v)[CHUNK], const State &) const
//                    __device__ void unit_done() const      (after a thread's last chunk of a unit)
// operator() is called once per thread (= A row a_row) and CHUNK-column chunk: columns col0 .. col0 + CHUNK - 1 of the unit
// = B rows u.b_row0 + col0 ...
//
// A unit is up to 512 B rows wide: two accumulators (TMEM columns 0.. and 256..) fed by the SAME A tile, i.e. a 256 x 512
// output tile per CTA pair.  The kernel is bound by what L2 can deliver to the SMs (measured: time per k-block follows the
// operand bytes per k-block, ~37 B/clk/SM, whatever the pipeline depth), so the tile is made as large as tensor memory
// allows: (256 + n) / (256 n) operand bytes per MAC, 0.0059 at n = 512 against 0.0078 at n = 256.  The price: both
// accumulators are busy until the unit's epilogue has run, so the epilogue is not overlapped with the next unit's MMAs;
// it is spread over sixteen warps (four per TMEM lane quadrant) to keep it short.
template <class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G8_THREADS, 1) k_g8_gemm(const G8Args g, const Epi epi) {
    extern __shared__ __align__(1024) uint8_t g8_smem[];
    __shared__ __align__(8) uint64_t full[G8_MAX_STAGES], empty[G8_MAX_STAGES], tmem_full, tmem_empty;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = g8_cta_rank();
    const bool leader = rank == 0;
    const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const uint32_t smem0 = vgs_smem_addr(g8_smem);
    const int NS = g.stages;
    unsigned long long ns_entry = 0;
    if (g.prof && tid == 0 && leader) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_entry));
    if (smem0 & 1023u) __trap();   // SWIZZLE_128B operand tiles need 1024-byte aligned stage buffers

    if (tid == 0) {
        for (int i = 0; i < G8_MAX_STAGES; ++i) {
            // leader: its own copies (one expect_tx arrival + the bytes) and the peer's relay arrival; peer: its own copies
            vgs_mbar_init(&full[i], leader ? 2 : 1);
            vgs_mbar_init(&empty[i], 1);
        }
        vgs_mbar_init(&tmem_full, 1);
        vgs_mbar_init(&tmem_empty, 2 * G8_EPI_WARPS);   // the epilogue warps of both CTAs of the pair
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(vgs_smem_addr(&tmem_base_slot)),
                     "r"((uint32_t)G8_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    g8_fence_before();
    __syncthreads();
    g8_cluster_sync();       // barriers of both CTAs initialised, both allocations done
    g8_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    vgs_dep_trigger();       // the next kernel on the stream may be staged ...
    vgs_dep_wait();          // ... and everything above overlapped the tail of the previous one (whose output the operands are)

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            int st = 0;            // pipeline stage and the parity of its current use (counters, no divisions: this thread
            uint32_t ph = 0;       // and the MMA issuer are the kernel's critical path)
            long long t_wait = 0;
            const long long t_begin = clock64();
            for (int ui = cluster; ui < g.n_units; ui += n_clusters) {
                const G8Unit u = g.units[ui];
                if (u.n == 0) continue;
                const int n0 = g8_first_half(u.n, g.gran), n1 = u.n - n0;
                const uint32_t b0_bytes = (uint32_t)(n0 >> 1) * G8_BK, b1_bytes = (uint32_t)(n1 >> 1) * G8_BK;
                const uint8_t *a_src = g.A + (((int64_t)u.kb0 * g.a_rows_pad + (int64_t)u.a_blk * 256 + rank * G8_A_ROWS) << 7);
                const uint8_t *b0_src = g.B + (((int64_t)u.kb0 * g.b_rows_pad + u.b_row0 + rank * (n0 >> 1)) << 7);
                const uint8_t *b1_src = g.B + (((int64_t)u.kb0 * g.b_rows_pad + u.b_row0 + n0 + rank * (n1 >> 1)) << 7);
                const int64_t a_step = g.a_rows_pad << 7, b_step = g.b_rows_pad << 7;
                for (int kb = 0; kb < u.kbn; ++kb) {
                    const long long t0 = g.prof ? clock64() : 0;
                    g8_wait(&empty[st], ph ^ 1u, g.err);
                    if (g.prof) t_wait += clock64() - t0;
                    uint8_t *dst = g8_smem + (size_t)st * g.stage_bytes;
                    vgs_mbar_expect_tx(&full[st], G8_A_BYTES + b0_bytes + b1_bytes);
                    vgs_bulk_g2s(dst, a_src + kb * a_step, G8_A_BYTES, &full[st]);
                    vgs_bulk_g2s(dst + G8_A_BYTES, b0_src + kb * b_step, b0_bytes, &full[st]);
                    if (n1) vgs_bulk_g2s(dst + G8_A_BYTES + g.b_bytes, b1_src + kb * b_step, b1_bytes, &full[st]);
                    if (++st == NS) { st = 0; ph ^= 1u; }
                }
            }
            if (g.prof && leader) {
                g.prof[cluster * 8 + 5] = t_wait;                 // producer: waiting for a free stage
                g.prof[cluster * 8 + 6] = clock64() - t_begin;    // producer: first to last copy issued
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            if (leader) {
                // ===== MMA issuer =====
                int it = 0;
                long long t_full = 0, t_tmem = 0, t_issue = 0;
                const uint32_t desc_lo0 = ((smem0 >> 4) & 0x3FFFu) | (1u << 16);
                const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO, version 1 (bit 46), SWIZZLE_128B (bit 61)
                const long long t_begin = clock64();
                unsigned long long ns_begin = 0;
                if (g.prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_begin));
                for (int ui = cluster; ui < g.n_units; ui += n_clusters) {
                    const G8Unit u = g.units[ui];
                    if (u.n == 0) continue;   // padding of the dealt unit list (every role skips it the same way)
                    const int n0 = g8_first_half(u.n, g.gran), n1 = u.n - n0;
                    long long t0 = g.prof ? clock64() : 0;
                    g8_wait_cluster(&tmem_empty, (uint32_t)((it & 1) ^ 1), g.err);   // the previous unit's epilogues (both CTAs) are done
                    if (g.prof) t_tmem += clock64() - t0;
                    g8_fence_after();
                    const uint32_t idesc0 = g8_idesc(256, n0, g.a_signed), idesc1 = g8_idesc(256, n1 ? n1 : 32, g.a_signed);
                    for (int kb = 0; kb < u.kbn; ++kb) {
                        if (g.prof) t0 = clock64();
                        g8_wait(&full[st], ph, g.err);   // both halves of the stage have landed
                        if (g.prof) { const long long t1 = clock64(); t_full += t1 - t0; t0 = t1; }
                        g8_fence_after();
                        // descriptors: constant high word, low word = (address >> 4) | leading-byte-offset field
                        const uint32_t a_lo = desc_lo0 + (uint32_t)st * (uint32_t)(g.stage_bytes >> 4);
#pragma unroll
                        for (int j = 0; j < G8_BK / 32; ++j) {
                            const uint64_t da = g8_desc_join(a_lo + 2u * j, desc_hi);
                            const uint32_t acc = (kb > 0 || j > 0) ? 1u : 0u;
                            g8_mma2(tmem_base, da, g8_desc_join(a_lo + (G8_A_BYTES >> 4) + 2u * j, desc_hi), idesc0, acc);
                            if (n1) g8_mma2(tmem_base + G8_MAX_N, da, g8_desc_join(a_lo + (uint32_t)((G8_A_BYTES + g.b_bytes) >> 4) + 2u * j, desc_hi), idesc1, acc);
                        }
                        g8_commit2(&empty[st]);
                        if (++st == NS) { st = 0; ph ^= 1u; }
                        if (g.prof) t_issue += clock64() - t0;
                    }
                    g8_commit2(&tmem_full);
                    ++it;
                }
                if (g.prof) {
                    long long *o = g.prof + cluster * 8;
                    unsigned long long ns_end;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_end));
                    o[0] = t_full; o[1] = (long long)(ns_end - ns_begin); o[2] = t_tmem; o[3] = t_issue; o[4] = clock64() - t_begin;
                    g.prof[n_clusters * 8 + cluster * 4 + 1] = (long long)ns_begin;
                    g.prof[n_clusters * 8 + cluster * 4 + 2] = (long long)ns_end;
                }
            } else {
                // ===== relay: my half of the stage has landed =====
                for (int ui = cluster; ui < g.n_units; ui += n_clusters) {
                    const G8Unit u = g.units[ui];
                    if (u.n == 0) continue;
                    for (int kb = 0; kb < u.kbn; ++kb) {
                        g8_wait(&full[st], ph, g.err);
                        g8_remote_arrive_relaxed(&full[st], 0);
                        if (++st == NS) { st = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else {
        // ===== epilogue: warp w reads TMEM lanes 32 (w % 4) .. + 31; the two warps of a quadrant take alternate chunks =====
        const int quad = warp & 3, part = (warp - 2) >> 2;   // G8_EPI_WARPS / 4 warps share a quadrant: chunk i goes to warp i % that
        int it = 0;
        long long t_epi = 0;
        for (int ui = cluster; ui < g.n_units; ui += n_clusters) {
            const G8Unit u = g.units[ui];
            if (u.n == 0) continue;
            const int n0 = g8_first_half(u.n, g.gran);
            const int64_t a_row = (int64_t)u.a_blk * 256 + rank * G8_A_ROWS + quad * 32 + lane;
            // the epilogue warps idle while the unit's MMAs run: let the functor load / warm whatever its chunks will read
            const typename Epi::State est = epi.prepare(u, a_row);
            constexpr int CH = Epi::CHUNK, CH_STEP = (G8_EPI_WARPS / 4) * Epi::CHUNK;
            for (int c0 = CH * part; c0 < u.n; c0 += CH_STEP) epi.prefetch(u, a_row, c0, est);
            g8_wait(&tmem_full, (uint32_t)(it & 1), g.err);
            const long long t_e0 = g.prof ? clock64() : 0;
            g8_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
            for (int c0 = CH * part; c0 < u.n; c0 += CH_STEP) {
                int32_t v[CH];
                // unit column c0 lives in accumulator 0 (TMEM column c0) or 1 (TMEM column 256 + c0 - n0)
                g8_tmem_ld<CH>(taddr + (uint32_t)(c0 < n0 ? c0 : G8_MAX_N + c0 - n0), v);
                if (!(g.debug & 1)) epi(u, a_row, c0, v, est);
            }
            epi.unit_done();
            g8_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) g8_local_arrive(&tmem_empty);
                else g8_remote_arrive(&tmem_empty, 0);
            }
            if (g.prof) t_epi += clock64() - t_e0;
            ++it;
        }
        if (g.prof && leader && warp == 2 && lane == 0) g.prof[cluster * 8 + 7] = t_epi;   // epilogue work of one warp
    }
    // teardown: nobody leaves (or frees tensor memory) while the pair may still touch this CTA's shared / tensor memory
    g8_fence_before();
    __syncthreads();
    g8_cluster_sync();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)G8_TMEM_COLS) : "memory");
    }
    if (g.prof && tid == 0 && leader) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        g.prof[n_clusters * 8 + cluster * 4 + 0] = (long long)ns_entry;
        g.prof[n_clusters * 8 + cluster * 4 + 3] = (long long)ns;
    }
}

template <class Epi>
static int g8_launch(vgs_context *h, const G8Args &g, const Epi &epi, cudaStream_t s) {
    if (g.n_units <= 0) return VGS_OK;
    auto kern = k_g8_gemm<Epi>;
    static bool attr_set[64] = {false};   // per device (one entry per template instance): not on every launch
    if (!attr_set[h->device & 63]) {
        VGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G8_SMEM_LIMIT));
        attr_set[h->device & 63] = true;
    }
    const int n_clusters = std::max(1, std::min(g.n_units, h->sm_count / 2));
    static int want_prof = -1, n_stages = -1, dbg = -1;
    if (want_prof < 0) { const char *e = getenv("VGS_G8_PROF"); want_prof = (e && e[0] == '1') ? 1 : 0; }
    if (n_stages < 0) { const char *e = getenv("VGS_G8_STAGES"); n_stages = e ? std::max(1, std::min(G8_MAX_STAGES, atoi(e))) : G8_MAX_STAGES; }
    if (dbg < 0) { const char *e = getenv("VGS_G8_DEBUG"); dbg = e ? atoi(e) : 0; }
    G8Args ga = g;
    ga.prof = nullptr;
    ga.debug = dbg;
    if (ga.gran < 32 || ga.gran % 32 || ga.gran % Epi::CHUNK || ga.gran > G8_MAX_N) {
        vgs_set_error("int8 GEMM: unit granule %d does not fit the functor's %d-column chunks", ga.gran, Epi::CHUNK);
        return VGS_ERR_INVALID;
    }
    // stage = this CTA's A rows + its halves of two accumulators' B rows; an accumulator is at most g8_max_acc(gran) rows
    // wide, so granules that do not divide 256 leave room for a deeper pipeline (the kernel is bound by how many operand
    // bytes each SM keeps in flight against the L2 round trip: 5 x 40 KB instead of 4 x 48 KB)
    ga.b_bytes = (g8_max_acc(ga.gran) / 2) * G8_BK;
    ga.stage_bytes = G8_A_BYTES + 2 * ga.b_bytes;
    ga.stages = std::max(1, std::min(std::min(n_stages, G8_MAX_STAGES), G8_SMEM_LIMIT / ga.stage_bytes));
    if (ga.stage_bytes % 1024) {
        vgs_set_error("int8 GEMM: stage of %d bytes is not 1024-byte aligned", ga.stage_bytes);
        return VGS_ERR_INVALID;
    }
    long long *d_prof = nullptr;
    if (want_prof) {
        if (cudaMalloc((void **)&d_prof, (size_t)n_clusters * 96) == cudaSuccess) {
            cudaMemsetAsync(d_prof, 0, (size_t)n_clusters * 96, s);
            ga.prof = d_prof;
        }
    }
    vgs_launch_dep(kern, dim3(2 * n_clusters), dim3(G8_THREADS), (size_t)ga.stages * ga.stage_bytes, s, ga, epi);
    VGS_LAUNCH_CHECK(h);
    if (d_prof) {
        // developer aid (VGS_G8_PROF=1): where the leader CTA's main-loop threads spend their cycles, mean (max) over the
        // CTA pairs, and the kernel's timeline
        std::vector<long long> hp((size_t)n_clusters * 12);
        cudaStreamSynchronize(s);
        cudaMemcpy(hp.data(), d_prof, hp.size() * 8, cudaMemcpyDeviceToHost);
        cudaFree(d_prof);
        double m[8] = {0, 0, 0, 0, 0, 0, 0, 0}, mx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < n_clusters; ++c)
            for (int k = 0; k < 8; ++k) {
                m[k] += (double)hp[(size_t)c * 8 + k] / n_clusters;
                mx[k] = std::max(mx[k], (double)hp[(size_t)c * 8 + k]);
            }
        fprintf(stderr, "[g8 prof] units %d pairs %d | MMA thread cycles mean (max): wait full %.0f (%.0f), wait tmem %.0f (%.0f), issue %.0f "
                        "(%.0f), total %.0f (%.0f) = %.1f us at %.0f MHz | producer: wait empty %.0f (%.0f), total %.0f (%.0f) | epilogue warp %.0f (%.0f)\n",
                g.n_units, n_clusters, m[0], mx[0], m[2], mx[2], m[3], mx[3], m[4], mx[4], m[1] / 1e3, m[1] > 0 ? m[4] / m[1] * 1e3 : 0.0, m[5], mx[5],
                m[6], mx[6], m[7], mx[7]);
        const long long *ts = hp.data() + (size_t)n_clusters * 8;
        long long e0 = ts[0], e1 = ts[0], l0 = ts[1], l1 = ts[1], f0 = ts[2], f1 = ts[2], x0 = ts[3], x1 = ts[3];
        for (int c = 0; c < n_clusters; ++c) {
            e0 = std::min(e0, ts[4 * c]); e1 = std::max(e1, ts[4 * c]);
            l0 = std::min(l0, ts[4 * c + 1]); l1 = std::max(l1, ts[4 * c + 1]);
            f0 = std::min(f0, ts[4 * c + 2]); f1 = std::max(f1, ts[4 * c + 2]);
            x0 = std::min(x0, ts[4 * c + 3]); x1 = std::max(x1, ts[4 * c + 3]);
        }
        fprintf(stderr, "[g8 prof] timeline (us after the first CTA entered): entry ..%.1f | main loop starts %.1f..%.1f | main loop ends %.1f..%.1f | "
                        "exit %.1f..%.1f\n", (e1 - e0) / 1e3, (l0 - e0) / 1e3, (l1 - e0) / 1e3, (f0 - e0) / 1e3, (f1 - e0) / 1e3, (x0 - e0) / 1e3, (x1 - e0) / 1e3);
        for (int c = 0; c < n_clusters; ++c)
            if (ts[4 * c + 3] - ts[4 * c + 2] > 10000)
                fprintf(stderr, "[g8 prof]   pair %d: main loop end %.1f, exit %.1f\n", c, (ts[4 * c + 2] - e0) / 1e3, (ts[4 * c + 3] - e0) / 1e3);
    }
    return VGS_OK;
}
#endif  // __CUDACC__

// ---- host: cutting rectangles of the output into units -------------------------------------------------------
struct G8Rect {
    int32_t a_blk;          // 256-row block of A
    int32_t b_row0, b_rows; // B row range: multiples of the granule
};
// Units for `rects` over k blocks [0, kblocks): tile widths are chosen so that the unit count is (close to) a multiple of
// the CTA pairs and as wide as that allows (operand bytes per MAC), largest first; when the tiles do not fill the machine
// the k range is cut too (at most `max_split` parts, at least 8 k-blocks each) and the units are marked partial.  Returns
// the split used.
// `gran`: granule of the unit widths and of the rectangles' row ranges (32, or the multiple of 32 the functor needs).
int g8_make_units(const std::vector<G8Rect> &rects, int kblocks, int n_clusters, int max_split, int gran, std::vector<G8Unit> *out);
