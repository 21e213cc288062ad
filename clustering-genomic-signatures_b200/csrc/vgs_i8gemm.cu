// vgs_i8gemm.cu -- host side of the int8 GEMM core (unit lists) and the live tensor-pipe peak probe.
#include <algorithm>
#include <vector>

#include "vgs_i8gemm.cuh"

// number of units a rectangle list gives at tile width N
static int64_t g8_count(const std::vector<G8Rect> &rects, int N) {
    int64_t t = 0;
    for (const G8Rect &r : rects) t += (r.b_rows + N - 1) / N;
    return t;
}

// widest tile a rectangle list gives when cut into pieces of at most N rows, and the number of pieces
static void g8_shape(const std::vector<G8Rect> &rects, int N, int gran, int64_t *count, int *widest) {
    *count = 0;
    *widest = 0;
    for (const G8Rect &r : rects) {
        const int n_t = (r.b_rows + N - 1) / N, quanta = r.b_rows / gran;
        *count += n_t;
        *widest = std::max(*widest, gran * ((quanta + n_t - 1) / n_t));
    }
}

int g8_make_units(const std::vector<G8Rect> &rects, int kblocks, int n_clusters, int max_split, int gran, std::vector<G8Unit> *out) {
    out->clear();
    if (rects.empty() || kblocks <= 0) return 1;
    if (n_clusters < 1) n_clusters = 1;
    const int unit_max = 2 * g8_max_acc(gran);   // two accumulators, each a whole number of granules
    int N = unit_max, split = 1;
    static int force_n = -1;   // tuning knob: VGS_G8_FORCE_N = tile width
    if (force_n < 0) { const char *e = getenv("VGS_G8_FORCE_N"); force_n = e ? atoi(e) : 0; }
    if (force_n >= gran && force_n <= unit_max) {
        N = force_n / gran * gran;
    } else {
        // Pick the tile width and the k-split by a small cost model of the kernel (cycles per CTA pair), measured on cfg 3:
        // a k-block of a w-wide unit costs max(MMA time 2 w, operand delivery 1.73 (256 + w)) + barrier overhead; a unit
        // adds its epilogue (~30 cycles per column with eight warps); a k-split adds the fold kernel.
        // Fewer, wider tiles move fewer operand bytes; more units fill the CTA pairs evenly -- the model arbitrates.
        double best = 1e300;
        for (int cand = gran; cand <= unit_max; cand += gran) {
            int64_t count;
            int w;
            g8_shape(rects, cand, gran, &count, &w);
            for (int sp = 1; sp <= max_split; ++sp) {
                if (sp > 1 && kblocks / sp < 8) break;
                const int64_t units = count * sp, waves = (units + n_clusters - 1) / n_clusters;
                const int kb_per = (kblocks + sp - 1) / sp;
                const double t_kb = std::max(2.0 * w, 1.73 * (256 + w)) + 60.0;
                // a k-split writes 4 bytes per digit sum instead of one fp64 per 8, and runs the fold kernel afterwards
                const double t = (double)waves * (kb_per * t_kb + (sp > 1 ? 45.0 : 30.0) * w + 2500.0) + (sp > 1 ? 14000.0 : 0.0);
                if (t < best * 0.999 || (t < best * 1.001 && cand > N)) { best = t; N = cand; split = sp; }
            }
        }
    }
    for (const G8Rect &r : rects) {
        const int n_t = (r.b_rows + N - 1) / N;
        const int quanta = r.b_rows / gran, q = quanta / n_t, rem = quanta % n_t;
        int row = r.b_row0;
        for (int t = 0; t < n_t; ++t) {
            const int n = gran * (q + (t < rem ? 1 : 0));
            for (int p = 0; p < split; ++p) {
                G8Unit u;
                u.a_blk = r.a_blk; u.b_row0 = row; u.n = n;
                u.kb0 = (int)((int64_t)kblocks * p / split);
                u.kbn = (int)((int64_t)kblocks * (p + 1) / split) - u.kb0;
                u.partial = split > 1 ? p + 1 : 0;   // 1-based part number
                if (n > 0 && u.kbn > 0) out->push_back(u);
            }
            row += n;
        }
    }
    // largest first, dealt to the CTA pairs in snake order (pair c runs units c, c + n_clusters, ...): the pairs that got
    // the largest units of one wave get the smallest of the next
    // ... and, among equals, units that read the same B rows next to each other: they run at the same time on neighbouring
    // CTA pairs and walk k in step, so a B line is fetched from HBM once and served to the others from L2 (rectangle-major
    // order re-read the digit operand once per 256-row A block: 4 x 131 MB at cfg 3)
    std::stable_sort(out->begin(), out->end(), [](const G8Unit &a, const G8Unit &b) {
        const int64_t wa = (int64_t)a.n * a.kbn, wb = (int64_t)b.n * b.kbn;
        if (wa != wb) return wa > wb;
        if (a.kb0 != b.kb0) return a.kb0 < b.kb0;
        if (a.b_row0 != b.b_row0) return a.b_row0 < b.b_row0;
        return a.a_blk < b.a_blk;
    });
    const size_t nu = out->size();
    const size_t waves = (nu + (size_t)n_clusters - 1) / (size_t)n_clusters;
    G8Unit none;
    none.a_blk = 0; none.b_row0 = 0; none.n = 0; none.kb0 = 0; none.kbn = 0; none.partial = 0;
    std::vector<G8Unit> dealt(waves * (size_t)n_clusters, none);
    for (size_t i = 0; i < nu; ++i) {
        const size_t wv = i / (size_t)n_clusters, idx = i % (size_t)n_clusters;
        const size_t c = (wv & 1) ? (size_t)n_clusters - 1 - idx : idx;
        dealt[wv * (size_t)n_clusters + c] = (*out)[i];
    }
    while (!dealt.empty() && dealt.back().n == 0) dealt.pop_back();
    out->swap(dealt);
    return split;
}

extern "C" int vgs_g8_plan(const int32_t *rects_h, int n_rects, int kblocks, int n_cta_pairs, int max_split, int32_t *units_out_h,
                           int capacity, int *n_units_out, int *split_out) {
    return vgs_g8_plan_gran(rects_h, n_rects, kblocks, n_cta_pairs, max_split, 32, units_out_h, capacity, n_units_out, split_out);
}

extern "C" int vgs_g8_plan_gran(const int32_t *rects_h, int n_rects, int kblocks, int n_cta_pairs, int max_split, int gran,
                                int32_t *units_out_h, int capacity, int *n_units_out, int *split_out) {
    VGS_CHECK_ARG(rects_h && n_rects >= 0 && kblocks > 0 && n_cta_pairs > 0 && max_split >= 1 && n_units_out, "vgs_g8_plan: bad arguments");
    VGS_CHECK_ARG(gran >= 32 && gran % 32 == 0 && gran <= G8_MAX_N, "vgs_g8_plan: the granule is a multiple of 32, at most 256");
    std::vector<G8Rect> rects;
    for (int i = 0; i < n_rects; ++i) {
        G8Rect r;
        r.a_blk = rects_h[3 * i]; r.b_row0 = rects_h[3 * i + 1]; r.b_rows = rects_h[3 * i + 2];
        VGS_CHECK_ARG(r.a_blk >= 0 && r.b_row0 >= 0 && r.b_row0 % gran == 0 && r.b_rows > 0 && r.b_rows % gran == 0, "vgs_g8_plan: rows must be multiples of the granule");
        rects.push_back(r);
    }
    std::vector<G8Unit> units;
    const int split = g8_make_units(rects, kblocks, n_cta_pairs, max_split, gran, &units);
    if (split_out) *split_out = split;
    *n_units_out = (int)units.size();
    if (units_out_h) {
        VGS_CHECK_ARG((int)units.size() <= capacity, "vgs_g8_plan: output capacity too small");
        for (size_t i = 0; i < units.size(); ++i) {
            const G8Unit &u = units[i];
            int32_t *o = units_out_h + 6 * i;
            o[0] = u.a_blk; o[1] = u.b_row0; o[2] = u.n; o[3] = u.kb0; o[4] = u.kbn; o[5] = u.partial;
        }
    }
    return VGS_OK;
}

// ---- live tensor-pipe peak ------------------------------------------------------------------------------------
// The roofline of the int8 contraction needs the int8 MMA rate of THIS GPU at the clocks it reaches in a short burst, and
// there is no library int8 GEMM to measure it with.  The probe issues the kernel's own instruction (tcgen05.mma
// cta_group::2 kind::i8, M 256, N 256, K 32) back to back from resident shared memory -- no global loads, no epilogue --
// on every SM pair, and reports executed int8 ops per second: the tensor pipe's ceiling for this instruction shape.
// stage = 128 A rows + up to 256 B rows (a single CTA holds all 256 rows of B, a CTA of a pair its half) of 128 bytes
#define PK_STAGE_BYTES (G8_A_BYTES + G8_MAX_N * G8_BK)
#define PK_SMEM_BYTES (4 * PK_STAGE_BYTES)   // 192 KB (two stages used): one CTA per SM, like the GEMM
template <int CG>
__global__ void __launch_bounds__(128, 1) k_g8_peak(int iters, int n_cols, int *err) {
    extern __shared__ __align__(1024) uint8_t pk_smem[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { vgs_mbar_init(&bar[0], 1); vgs_mbar_init(&bar[1], 1); }
    // operands: zeros (the tensor pipe does not look at the values)
    for (int i = threadIdx.x; i < 2 * PK_STAGE_BYTES / 16; i += blockDim.x) ((uint4 *)pk_smem)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 0) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(vgs_smem_addr(&tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(vgs_smem_addr(&tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy zero fill -> visible to the tensor core's reads
    g8_fence_before();
    __syncthreads();
    if (CG == 2) g8_cluster_sync();
    g8_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const bool issuer = threadIdx.x == 32 && (CG == 1 || g8_cta_rank() == 0);
    if (issuer) {
        const uint32_t sa = vgs_smem_addr(pk_smem);
        const uint32_t idesc = g8_idesc(CG == 2 ? 256 : 128, n_cols, 0);
        for (int it = 0; it < iters; ++it) {
            const uint32_t st = sa + (uint32_t)(it & 1) * PK_STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint64_t da = g8_smem_desc(st + 32u * j), db = g8_smem_desc(st + G8_A_BYTES + 32u * j);
                if (CG == 2) g8_mma2(tmem_base + (uint32_t)((it & 1) * 256), da, db, idesc, 1u);
                else g8_mma1(tmem_base + (uint32_t)((it & 1) * 256), da, db, idesc, 1u);
            }
        }
        if (CG == 2) g8_commit2(&bar[0]);
        else g8_commit1(&bar[0]);
    }
    if (threadIdx.x == 32) g8_wait(&bar[0], 0, err);
    g8_fence_before();
    __syncthreads();
    if (CG == 2) g8_cluster_sync();
    if (warp == 0) {
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

extern "C" int vgs_tensor_peak_i8(vgs_handle h, int cta_group, double *tops_out, double *ms_out) {
    return vgs_tensor_rate_i8(h, cta_group, 256, tops_out, ms_out);
}

// the same probe at another instruction width N (multiple of 16, <= 256): how the tensor pipe's time per MMA depends on N
extern "C" int vgs_tensor_rate_i8(vgs_handle h, int cta_group, int n_cols, double *tops_out, double *ms_out) {
    VGS_CHECK_ARG(h && (cta_group == 1 || cta_group == 2) && tops_out, "vgs_tensor_rate_i8: bad arguments");
    VGS_CHECK_ARG(n_cols >= 16 && n_cols <= 256 && n_cols % 16 == 0, "vgs_tensor_rate_i8: N is a multiple of 16, at most 256");
    VGS_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->own_stream;
    const int iters = 2048;   // x 4 MMAs of 256 x 256 x 32 (pair) / 128 x 256 x 32 (single CTA)
    const int smem = PK_SMEM_BYTES;
    const int n_cta = cta_group == 2 ? (h->sm_count / 2) * 2 : h->sm_count;
    cudaEvent_t e0, e1;
    VGS_CUDA(cudaEventCreate(&e0));
    VGS_CUDA(cudaEventCreate(&e1));
    float best = 0.f;
    for (int rep = 0; rep < 4; ++rep) {   // first repetition = warm-up
        VGS_CUDA(cudaEventRecord(e0, s));
        if (cta_group == 2) {
            VGS_CUDA(cudaFuncSetAttribute(k_g8_peak<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)n_cta); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            VGS_CUDA(cudaLaunchKernelEx(&cfg, k_g8_peak<2>, iters, n_cols, h->d_err));
        } else {
            VGS_CUDA(cudaFuncSetAttribute(k_g8_peak<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            k_g8_peak<1><<<n_cta, 128, smem, s>>>(iters, n_cols, h->d_err);
        }
        VGS_LAUNCH_CHECK(h);
        VGS_CUDA(cudaEventRecord(e1, s));
        VGS_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        VGS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && (best == 0.f || ms < best)) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double units = cta_group == 2 ? (double)(n_cta / 2) : (double)n_cta;
    const double ops = 2.0 * units * (double)iters * 4.0 * (cta_group == 2 ? 256.0 : 128.0) * (double)n_cols * 32.0;
    *tops_out = ops / ((double)best * 1e-3) / 1e12;
    if (ms_out) *ms_out = (double)best;
    return VGS_OK;
}
