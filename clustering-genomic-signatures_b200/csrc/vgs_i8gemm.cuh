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
// Kernel: one CTA pair (cluster of 2) per SM pair, persistent over a static list of units (tile x k range); 192 threads:
//   warp 0   producer: per stage one expect_tx + two bulk copies (its 128 A rows, its half of the B rows)
//   warp 1   leader CTA: the one thread that issues every tcgen05.mma of the pair and commits them to the stage's
//            "empty" barriers of BOTH CTAs (multicast commit); peer CTA: forwards "my half of the stage has landed" to
//            the leader (remote mbarrier arrive)
//   warps 2-5 epilogue: tcgen05.ld of their 32 TMEM lanes, 32 columns at a time, handed to the caller's functor.  The
//            accumulator is double buffered (2 x 256 of the 512 TMEM columns), so the MMAs of unit t+1 run while unit t
//            is folded and stored.
#pragma once
#include "vgs_common.cuh"

#define G8_BK 128                 // k per pipeline stage = bytes per operand row per stage = one swizzle atom
#define G8_STAGES 6
#define G8_THREADS 192
#define G8_A_ROWS 128             // A rows per CTA (256 per pair)
#define G8_MAX_N 256              // B rows per unit (128 per CTA)
#define G8_A_BYTES (G8_A_ROWS * G8_BK)
#define G8_STAGE_BYTES (G8_A_BYTES + (G8_MAX_N / 2) * G8_BK)   // 32 KB
#define G8_SMEM_BYTES (G8_STAGES * G8_STAGE_BYTES)
#define G8_TMEM_COLS 512
#define G8_WATCHDOG_CYCLES (1ll << 32)   // ~2 s: a lost barrier signal traps instead of hanging the GPU

struct G8Unit {
    int32_t a_blk;    // 256-row block of A
    int32_t b_row0;   // first B row, multiple of 8 (16 for the peer's half to stay 8-aligned: n / 2 is a multiple of 16)
    int32_t n;        // B rows of this unit: multiple of 32, <= 256
    int32_t kb0, kbn; // k range in 128-blocks
    int32_t partial;  // != 0: this unit covers part of the k range (the functor accumulates atomically)
};

struct G8Args {
    const uint8_t *A, *B;
    int64_t a_rows_pad, b_rows_pad;   // rows per k-block slab (multiples of 256 / 32)
    const G8Unit *units;
    int n_units;
    int a_signed;                     // A operand: 0 = unsigned 8 bit, 1 = signed
    int *err;                         // device error word (bit 2: watchdog)
};

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
__device__ __forceinline__ void g8_local_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(vgs_smem_addr(bar)) : "memory");
}
// wait for a phase of an mbarrier (cluster-scope acquire: the signal may come from the other CTA); a signal that never
// comes sets the error word and traps -- a failed launch, not a hung GPU
__device__ __forceinline__ void g8_wait(uint64_t *bar, uint32_t parity, int *err) {
    uint32_t done = 0;
    long long t0 = 0;
    int spins = 0;
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
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
__device__ __forceinline__ void g8_tmem_ld32(uint32_t taddr, int32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- the kernel ---------------------------------------------------------------------------------------------
// Epi: struct with   __device__ void operator()(const G8Unit &u, int64_t a_row, int col0, const int32_t (&v)[32]) const
//                    __device__ void unit_done() const      (after a thread's last chunk of a unit)
// operator() is called once per thread (= A row a_row) and 32-column chunk col0 (columns col0 .. col0 + 31 = B rows u.b_row0 + col0 ...)
template <class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G8_THREADS, 1) k_g8_gemm(const G8Args g, const Epi epi) {
    extern __shared__ __align__(1024) uint8_t g8_smem[];
    __shared__ __align__(8) uint64_t full[G8_STAGES], empty[G8_STAGES], peer_full[G8_STAGES], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = g8_cta_rank();
    const bool leader = rank == 0;
    const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const uint32_t smem0 = vgs_smem_addr(g8_smem);
    if (smem0 & 1023u) __trap();   // SWIZZLE_128B operand tiles need 1024-byte aligned stage buffers

    if (tid == 0) {
        for (int i = 0; i < G8_STAGES; ++i) {
            vgs_mbar_init(&full[i], 1);
            vgs_mbar_init(&empty[i], 1);
            vgs_mbar_init(&peer_full[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            vgs_mbar_init(&tmem_full[i], 1);
            vgs_mbar_init(&tmem_empty[i], 8);   // four epilogue warps in each CTA of the pair
        }
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

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            int64_t gs = 0;   // stage use counter
            for (int ui = cluster; ui < g.n_units; ui += n_clusters) {
                const G8Unit u = g.units[ui];
                const uint32_t b_bytes = (uint32_t)(u.n >> 1) * G8_BK;
                const uint8_t *a_src = g.A + (((int64_t)u.kb0 * g.a_rows_pad + (int64_t)u.a_blk * 256 + rank * G8_A_ROWS) << 7);
                const uint8_t *b_src = g.B + (((int64_t)u.kb0 * g.b_rows_pad + u.b_row0 + rank * (u.n >> 1)) << 7);
                const int64_t a_step = g.a_rows_pad << 7, b_step = g.b_rows_pad << 7;
                for (int kb = 0; kb < u.kbn; ++kb, ++gs) {
                    const int st = (int)(gs % G8_STAGES);
                    g8_wait(&empty[st], (uint32_t)(((gs / G8_STAGES) & 1) ^ 1), g.err);
                    uint8_t *dst = g8_smem + (size_t)st * G8_STAGE_BYTES;
                    vgs_mbar_expect_tx(&full[st], G8_A_BYTES + b_bytes);
                    vgs_bulk_g2s(dst, a_src + kb * a_step, G8_A_BYTES, &full[st]);
                    vgs_bulk_g2s(dst + G8_A_BYTES, b_src + kb * b_step, b_bytes, &full[st]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int64_t gs = 0;
            if (leader) {
                // ===== MMA issuer =====
                int it = 0;
                for (int ui = cluster; ui < g.n_units; ui += n_clusters, ++it) {
                    const G8Unit u = g.units[ui];
                    if (u.n == 0) continue;   // padding of the dealt unit list (every role skips it the same way)
                    const int acc = it & 1;
                    g8_wait(&tmem_empty[acc], (uint32_t)(((it >> 1) & 1) ^ 1), g.err);   // epilogues of both CTAs have drained it
                    g8_fence_after();
                    const uint32_t idesc = g8_idesc(256, u.n, g.a_signed);
                    const uint32_t d_addr = tmem_base + (uint32_t)(acc * G8_MAX_N);
                    for (int kb = 0; kb < u.kbn; ++kb, ++gs) {
                        const int st = (int)(gs % G8_STAGES);
                        const uint32_t ph = (uint32_t)((gs / G8_STAGES) & 1);
                        g8_wait(&full[st], ph, g.err);
                        g8_wait(&peer_full[st], ph, g.err);
                        g8_fence_after();
                        const uint32_t sa = smem0 + (uint32_t)st * G8_STAGE_BYTES;
#pragma unroll
                        for (int j = 0; j < G8_BK / 32; ++j)
                            g8_mma2(d_addr, g8_smem_desc(sa + 32u * j), g8_smem_desc(sa + G8_A_BYTES + 32u * j), idesc, (kb > 0 || j > 0) ? 1u : 0u);
                        g8_commit2(&empty[st]);
                    }
                    g8_commit2(&tmem_full[acc]);
                }
            } else {
                // ===== forwarder: my half of the stage has landed =====
                for (int ui = cluster; ui < g.n_units; ui += n_clusters) {
                    const int kbn = g.units[ui].kbn;
                    for (int kb = 0; kb < kbn; ++kb, ++gs) {
                        const int st = (int)(gs % G8_STAGES);
                        g8_wait(&full[st], (uint32_t)((gs / G8_STAGES) & 1), g.err);
                        g8_remote_arrive(&peer_full[st], 0);
                    }
                }
            }
        }
    } else {
        // ===== epilogue: warp w reads TMEM lanes 32 (w % 4) .. + 31 =====
        const int quad = warp & 3;
        int it = 0;
        for (int ui = cluster; ui < g.n_units; ui += n_clusters, ++it) {
            const G8Unit u = g.units[ui];
            if (u.n == 0) continue;
            const int acc = it & 1;
            g8_wait(&tmem_full[acc], (uint32_t)((it >> 1) & 1), g.err);
            g8_fence_after();
            const int64_t a_row = (int64_t)u.a_blk * 256 + rank * G8_A_ROWS + quad * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * G8_MAX_N);
#pragma unroll 1
            for (int c0 = 0; c0 < u.n; c0 += 32) {
                int32_t v[32];
                g8_tmem_ld32(taddr + (uint32_t)c0, v);
                epi(u, a_row, c0, v);
            }
            epi.unit_done();
            g8_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) g8_local_arrive(&tmem_empty[acc]);
                else g8_remote_arrive(&tmem_empty[acc], 0);
            }
        }
    }
    // teardown: nobody leaves (or frees tensor memory) while the pair may still touch this CTA's shared / tensor memory
    g8_fence_before();
    __syncthreads();
    g8_cluster_sync();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)G8_TMEM_COLS) : "memory");
    }
}

template <class Epi>
static int g8_launch(vgs_context *h, const G8Args &g, const Epi &epi, cudaStream_t s) {
    if (g.n_units <= 0) return VGS_OK;
    auto kern = k_g8_gemm<Epi>;
    VGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G8_SMEM_BYTES));
    const int n_clusters = std::max(1, std::min(g.n_units, h->sm_count / 2));
    kern<<<2 * n_clusters, G8_THREADS, G8_SMEM_BYTES, s>>>(g, epi);
    VGS_LAUNCH_CHECK(h);
    return VGS_OK;
}
#endif  // __CUDACC__

// ---- host: cutting rectangles of the output into units -------------------------------------------------------
struct G8Rect {
    int32_t a_blk;          // 256-row block of A
    int32_t b_row0, b_rows; // B row range: b_row0 multiple of 32; b_rows multiple of 32
};
// Units for `rects` over k blocks [0, kblocks): tile widths are chosen so that the unit count is (close to) a multiple of
// the CTA pairs, largest first; when the tiles do not fill the machine the k range is cut too (at most `max_split`
// parts, at least 8 k-blocks each) and the units are marked partial.  Returns the split used.
int g8_make_units(const std::vector<G8Rect> &rects, int kblocks, int n_clusters, int max_split, std::vector<G8Unit> *out);
