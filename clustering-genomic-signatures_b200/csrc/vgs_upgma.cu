// vgs_upgma.cu -- row f1: average-link clustering of the float32 [N][N] distance matrix on the device,
// reproducing src/clustering/average_link_clustering.pyx:21-146 (+ graph_based_clustering.pyx:67-89) exactly.
//
// The reference keeps a heap and a dict per cluster (O(N^2) Python tuples: ~160 GB and hours at N = 20 000,
// SURVEY.md §3.4).  Here the state is one fp64 [N][N] matrix plus a nearest-neighbour cache per cluster row:
//   step = (1) argmin over the cache   (2) Lance-Williams update of one row and one column
//          (3) bookkeeping             (4) refresh of the cache rows that pointed at the merged clusters.
// Semantics kept bit for bit (checked in tests/ against the reference's own golden runs):
//   - _find_min_edge (:62-80): clusters are scanned in dict insertion order (singletons in index order, every merged
//     cluster appended at the end); a cluster's candidate is its heap top = min of (distance, key tuple) -- keys are
//     sorted member tuples of disjoint clusters, so the tuple order is the order of the smallest member; the global
//     candidate is replaced only on strict '<' (:75), i.e. the first cluster in dict order wins ties;
//   - _new_dist (:119-122) under numpy's NEP-50 promotion as it runs today: entries between two never-merged
//     singletons are numpy.float32, entries produced by a merge are Python floats; int * float32 -> float32,
//     float32 + pyfloat -> float32 (pyfloat cast first), float32 / int -> float32, pyfloat arithmetic -> float64;
//   - other rows are updated from their own distances to left/right (_merge_weights :101-117), the merged row from
//     left's and right's rows (_merge_heaps :124-146).
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "vgs_common.cuh"

struct AlState {
    double *D;        // [n][n]
    int64_t n;
    int *size, *minmem, *okey, *nn_j, *parent;
    unsigned char *alive, *single;
    double *nn_d;
    int *sel;         // [0] = left, [1] = right, [2] = done flag, [3] = merges recorded
    double *merges;
};

__device__ __forceinline__ double al_new_dist(int nl, double ld, bool l32, int nr, double rd, bool r32) {
    if (l32 || r32) {
        const float a = l32 ? __fmul_rn((float)nl, (float)ld) : (float)__dmul_rn((double)nl, ld);
        const float b = r32 ? __fmul_rn((float)nr, (float)rd) : (float)__dmul_rn((double)nr, rd);
        return (double)__fdiv_rn(__fadd_rn(a, b), (float)(nl + nr));
    }
    return __ddiv_rn(__dadd_rn(__dmul_rn((double)nl, ld), __dmul_rn((double)nr, rd)), (double)(nl + nr));
}

__global__ void k_al_init(AlState s, const float *__restrict__ D32) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < s.n * s.n) s.D[e] = (double)D32[e];
    if (e < s.n) {
        s.size[e] = 1; s.minmem[e] = (int)e; s.okey[e] = (int)e; s.parent[e] = (int)e;
        s.alive[e] = 1; s.single[e] = 1;
    }
    if (e == 0) { s.sel[0] = -1; s.sel[1] = -1; s.sel[2] = 0; s.sel[3] = 0; }
}

// nearest neighbour of row x: min over alive j != x of (D[x][j], minmem[j]); one block per row
__device__ void al_rescan_row(const AlState &s, int x) {
    __shared__ double sd[128];
    __shared__ int sj[128], sm[128];
    double bd = INFINITY;
    int bj = -1, bm = 0x7fffffff;
    const double *row = s.D + (int64_t)x * s.n;
    for (int j = threadIdx.x; j < s.n; j += blockDim.x) {
        if (!s.alive[j] || j == x) continue;
        const double d = row[j];
        const int m = s.minmem[j];
        if (bj < 0 || d < bd || (d == bd && m < bm)) { bd = d; bj = j; bm = m; }
    }
    sd[threadIdx.x] = bd; sj[threadIdx.x] = bj; sm[threadIdx.x] = bm;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double d2 = sd[threadIdx.x + o];
            const int j2 = sj[threadIdx.x + o], m2 = sm[threadIdx.x + o];
            if (j2 >= 0 && (sj[threadIdx.x] < 0 || d2 < sd[threadIdx.x] || (d2 == sd[threadIdx.x] && m2 < sm[threadIdx.x]))) {
                sd[threadIdx.x] = d2; sj[threadIdx.x] = j2; sm[threadIdx.x] = m2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { s.nn_d[x] = sd[0]; s.nn_j[x] = sj[0]; }
}

__global__ void k_al_rescan_all(AlState s) { al_rescan_row(s, blockIdx.x); }

// (1) first cluster in dict order whose candidate distance is the strict minimum
__global__ void k_al_argmin(AlState s) {
    __shared__ double sd[1024];
    __shared__ int si[1024], sk[1024];
    if (s.sel[2]) return;
    double bd = INFINITY;
    int bi = -1, bk = 0x7fffffff;
    for (int i = threadIdx.x; i < s.n; i += blockDim.x) {
        if (!s.alive[i] || s.nn_j[i] < 0) continue;
        const double d = s.nn_d[i];
        const int k = s.okey[i];
        if (bi < 0 || d < bd || (d == bd && k < bk)) { bd = d; bi = i; bk = k; }
    }
    sd[threadIdx.x] = bd; si[threadIdx.x] = bi; sk[threadIdx.x] = bk;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double d2 = sd[threadIdx.x + o];
            const int i2 = si[threadIdx.x + o], k2 = sk[threadIdx.x + o];
            if (i2 >= 0 && (si[threadIdx.x] < 0 || d2 < sd[threadIdx.x] || (d2 == sd[threadIdx.x] && k2 < sk[threadIdx.x]))) {
                sd[threadIdx.x] = d2; si[threadIdx.x] = i2; sk[threadIdx.x] = k2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (si[0] < 0 || !(sd[0] < INFINITY)) {  // no edge left (reference: min_cluster stays (-1,), loop breaks)
            s.sel[2] = 1;
        } else {
            s.sel[0] = si[0];
            s.sel[1] = s.nn_j[si[0]];
            s.merges[s.sel[3]] = sd[0];
        }
    }
}

// (2) Lance-Williams update: column `left` of every other row, and row `left`
__global__ void k_al_update(AlState s) {
    if (s.sel[2]) return;
    const int l = s.sel[0], r = s.sel[1];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= s.n || !s.alive[x] || x == l || x == r) return;
    const int nl = s.size[l], nr = s.size[r];
    const bool xl32 = s.single[x] && s.single[l], xr32 = s.single[x] && s.single[r];
    const int64_t n = s.n;
    const double col = al_new_dist(nl, s.D[x * n + l], xl32, nr, s.D[x * n + r], xr32);
    const double row = al_new_dist(nl, s.D[(int64_t)l * n + x], xl32, nr, s.D[(int64_t)r * n + x], xr32);
    s.D[x * n + l] = col;
    s.D[(int64_t)l * n + x] = row;
}

// (3) bookkeeping of the merge
__global__ void k_al_meta(AlState s, int step) {
    if (s.sel[2]) return;
    const int l = s.sel[0], r = s.sel[1];
    s.alive[r] = 0;
    s.single[l] = 0;
    s.size[l] += s.size[r];
    if (s.minmem[r] < s.minmem[l]) s.minmem[l] = s.minmem[r];
    s.okey[l] = (int)s.n + step;
    s.parent[r] = l;
    s.sel[3] += 1;
}

// (4) refresh the nearest-neighbour cache: full rescan for the merged row and for rows that pointed at left/right,
//     a single comparison for every other row
__global__ void k_al_refresh(AlState s) {
    if (s.sel[2]) return;
    const int l = s.sel[0], r = s.sel[1];
    const int x = blockIdx.x;
    if (!s.alive[x]) return;
    const int cur = s.nn_j[x];
    if (x == l || cur < 0 || cur == l || cur == r) {
        al_rescan_row(s, x);
    } else if (threadIdx.x == 0) {
        const double d = s.D[(int64_t)x * s.n + l];
        if (d < s.nn_d[x] || (d == s.nn_d[x] && s.minmem[l] < s.minmem[cur])) { s.nn_d[x] = d; s.nn_j[x] = l; }
    }
}


// ---------------------------------------------------------------------------------------------------------
// persistent variant: the whole merge loop in ONE cooperative kernel, one grid barrier per merge.
//
// Rows are dealt to the CTAs (row x belongs to CTA x % gridDim.x).  Every CTA keeps its own copy of the small per-cluster
// state (size, smallest member, alive / never-merged flags: 5 bytes per cluster) in shared memory and applies every
// merge to it, so the only cross-CTA traffic per merge is one candidate record per CTA:
//   A(t)  CTA b posts (a) the best (distance, dict key) among the nearest-neighbour caches of its rows and (b) its
//         best candidate for the nearest neighbour of the row merged in step t-1 (whose new entries D[l][x] it wrote
//         itself for its own x) -- grid barrier --
//   B(t)  every CTA reduces all records identically: first (b) -> the merged row's neighbour, then the global argmin
//         -> (left, right); it updates D[x][left], D[left][x] for its rows x, refreshes their caches (one comparison,
//         or a block-wide rescan of the row when the cache pointed at left / right), applies the merge to its state.
// Every D entry a CTA reads in step t was written by itself or before the barrier of step t (D and the candidate
// records are read with ld.global.cg: no stale L1 lines).  Same arithmetic, same
// tie-breaking and therefore the same merges as the multi-kernel path above (kept for n too large for the shared state).
// ---------------------------------------------------------------------------------------------------------
#define ALP_THREADS 1024

struct AlCand {
    double d;
    int key, i, j, pad;
};

__device__ __forceinline__ bool al_better(const AlCand &a, const AlCand &b) {  // is a strictly preferable to b?
    if (a.i < 0) return false;
    if (b.i < 0) return true;
    return a.d < b.d || (a.d == b.d && a.key < b.key);
}
__device__ __forceinline__ AlCand al_shfl_xor(const AlCand &c, int o) {
    AlCand r;
    r.d = __shfl_xor_sync(0xffffffffu, c.d, o);
    r.key = __shfl_xor_sync(0xffffffffu, c.key, o);
    r.i = __shfl_xor_sync(0xffffffffu, c.i, o);
    r.j = __shfl_xor_sync(0xffffffffu, c.j, o);
    r.pad = 0;
    return r;
}
// block-wide argmin; every thread returns the winner.  Ties on (d, key) cannot occur between distinct candidates
// (keys are unique), so the reduction order does not matter.
__device__ AlCand al_block_best(AlCand c, AlCand *scratch /* [33] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const AlCand other = al_shfl_xor(c, o);
        if (al_better(other, c)) c = other;
    }
    __syncthreads();
    if (lane == 0) scratch[warp] = c;
    __syncthreads();
    if (warp == 0) {
        AlCand v;
        v.d = 0.0; v.key = 0; v.i = -1; v.j = -1; v.pad = 0;
        if (lane < (int)(blockDim.x >> 5)) v = scratch[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const AlCand other = al_shfl_xor(v, o);
            if (al_better(other, v)) v = other;
        }
        if (lane == 0) scratch[32] = v;
    }
    __syncthreads();
    return scratch[32];
}

__device__ __forceinline__ void al_grid_barrier(unsigned int *counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(ALP_THREADS, 1)
k_al_persistent(AlState s, int steps, AlCand *cand /* [2][gridDim.x][2] */, unsigned int *barrier_ctr) {
    extern __shared__ __align__(16) unsigned char al_smem[];
    const int n = (int)s.n, nb = (int)gridDim.x, b = (int)blockIdx.x, tid = (int)threadIdx.x;
    unsigned short *size = (unsigned short *)al_smem;
    unsigned short *minmem = size + n;
    unsigned char *flags = (unsigned char *)(minmem + n);            // bit 0 alive, bit 1 never merged
    int *rescan = (int *)(al_smem + (((size_t)5 * n + 15) & ~(size_t)15));
    __shared__ AlCand scratch[33];
    __shared__ int n_rescan;
    for (int i = tid; i < n; i += ALP_THREADS) { size[i] = 1; minmem[i] = (unsigned short)i; flags[i] = 3; }
    __syncthreads();
    int pend = -1;
    unsigned int bar_target = 0;
    int done_steps = 0;
    for (int t = 0; t < steps; ++t) {
        // ---- A: this CTA's candidates
        AlCand ca, cb;
        ca.d = 0.0; ca.key = 0; ca.i = -1; ca.j = -1; ca.pad = 0;
        cb = ca;
        for (int x = b + tid * nb; x < n; x += ALP_THREADS * nb) {
            if (!(flags[x] & 1) || x == pend) continue;
            const int j = s.nn_j[x];
            if (j >= 0) {
                AlCand c;
                c.d = s.nn_d[x]; c.key = s.okey[x]; c.i = x; c.j = j; c.pad = 0;
                if (al_better(c, ca)) ca = c;
            }
            if (pend >= 0) {
                AlCand c;
                c.d = __ldcg(&s.D[(int64_t)pend * n + x]); c.key = minmem[x]; c.i = x; c.j = x; c.pad = 0;
                if (al_better(c, cb)) cb = c;
            }
        }
        ca = al_block_best(ca, scratch);
        cb = al_block_best(cb, scratch);
        AlCand *post = cand + ((size_t)(t & 1) * nb + b) * 2;
        if (tid == 0) { post[0] = ca; post[1] = cb; }
        bar_target += (unsigned int)nb;
        al_grid_barrier(barrier_ctr, bar_target);
        // ---- B: identical reduction in every CTA
        const AlCand *all = cand + (size_t)(t & 1) * nb * 2;
        AlCand ga, gb;
        ga.d = 0.0; ga.key = 0; ga.i = -1; ga.j = -1; ga.pad = 0;
        gb = ga;
        for (int q = tid; q < nb; q += ALP_THREADS) {
            AlCand a2, b2;
            a2.d = __ldcg(&all[2 * q].d); a2.key = __ldcg(&all[2 * q].key); a2.i = __ldcg(&all[2 * q].i); a2.j = __ldcg(&all[2 * q].j); a2.pad = 0;
            b2.d = __ldcg(&all[2 * q + 1].d); b2.key = __ldcg(&all[2 * q + 1].key); b2.i = __ldcg(&all[2 * q + 1].i); b2.j = __ldcg(&all[2 * q + 1].j); b2.pad = 0;
            if (al_better(a2, ga)) ga = a2;
            if (al_better(b2, gb)) gb = b2;
        }
        gb = al_block_best(gb, scratch);
        ga = al_block_best(ga, scratch);
        if (pend >= 0) {
            // the row merged in the previous step: its neighbour, and its own candidacy (dict key n + t - 1)
            if (pend % nb == b && tid == 0) { s.nn_d[pend] = gb.i >= 0 ? gb.d : INFINITY; s.nn_j[pend] = gb.i >= 0 ? gb.i : -1; }
            if (gb.i >= 0) {
                AlCand c;
                c.d = gb.d; c.key = n + t - 1; c.i = pend; c.j = gb.i; c.pad = 0;
                if (al_better(c, ga)) ga = c;
            }
        }
        if (ga.i < 0 || !(ga.d < INFINITY)) break;  // no edge left (the reference's loop breaks)
        const int l = ga.i, r = ga.j;
        if (b == 0 && tid == 0) s.merges[t] = ga.d;
        done_steps = t + 1;
        if (tid == 0) n_rescan = 0;
        __syncthreads();
        const int nl = size[l], nr = size[r];
        const bool l_single = (flags[l] & 2) != 0, r_single = (flags[r] & 2) != 0;
        const int new_min = minmem[l] < minmem[r] ? minmem[l] : minmem[r];
        for (int x = b + tid * nb; x < n; x += ALP_THREADS * nb) {
            if (!(flags[x] & 1) || x == l || x == r) continue;
            const bool x_single = (flags[x] & 2) != 0;
            const bool xl32 = x_single && l_single, xr32 = x_single && r_single;
            const double col = al_new_dist(nl, __ldcg(&s.D[(int64_t)x * n + l]), xl32, nr, __ldcg(&s.D[(int64_t)x * n + r]), xr32);
            const double row = al_new_dist(nl, __ldcg(&s.D[(int64_t)l * n + x]), xl32, nr, __ldcg(&s.D[(int64_t)r * n + x]), xr32);
            s.D[(int64_t)x * n + l] = col;
            s.D[(int64_t)l * n + x] = row;
            const int cur = s.nn_j[x];
            if (cur < 0 || cur == l || cur == r) {
                rescan[atomicAdd(&n_rescan, 1)] = x;
            } else {
                const double nd = s.nn_d[x];
                if (col < nd || (col == nd && new_min < (int)minmem[cur])) { s.nn_d[x] = col; s.nn_j[x] = l; }
            }
        }
        __syncthreads();
        if (tid == 0) {
            flags[r] &= (unsigned char)~1;
            flags[l] &= (unsigned char)~2;
            size[l] = (unsigned short)(nl + nr);
            minmem[l] = (unsigned short)new_min;
            if (l % nb == b) s.okey[l] = n + t;
            if (r % nb == b) s.parent[r] = l;
        }
        __syncthreads();
        // block-wide rescans of the rows whose cache pointed at the merged clusters
        const int nres = n_rescan;
        for (int q = 0; q < nres; ++q) {
            const int x = rescan[q];
            const double *row = s.D + (int64_t)x * n;
            AlCand c;
            c.d = 0.0; c.key = 0; c.i = -1; c.j = -1; c.pad = 0;
            for (int j = tid; j < n; j += ALP_THREADS) {
                if (!(flags[j] & 1) || j == x) continue;
                AlCand v;
                v.d = __ldcg(row + j); v.key = minmem[j]; v.i = j; v.j = j; v.pad = 0;
                if (al_better(v, c)) c = v;
            }
            c = al_block_best(c, scratch);
            if (tid == 0) { s.nn_d[x] = c.i >= 0 ? c.d : INFINITY; s.nn_j[x] = c.i; }
        }
        __syncthreads();
        pend = l;
    }
    if (b == 0) {
        __syncthreads();
        for (int i = tid; i < n; i += ALP_THREADS) s.minmem[i] = minmem[i];
        if (tid == 0) s.sel[3] = done_steps;
    }
}

extern "C" int vgs_average_link(vgs_handle h, int64_t n, const float *D32_d, int64_t k, int32_t *labels_out_h,
                                double *merge_dist_out_h, int64_t *n_merges_out) {
    VGS_CHECK_ARG(h && n > 0 && D32_d && labels_out_h, "vgs_average_link: bad arguments");
    VGS_CHECK_ARG(k >= 1 && k <= n, "vgs_average_link: k must be in [1, n]");
    VGS_CHECK_ARG(n < 65536, "vgs_average_link: n must be below 65536");
    VGS_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->own_stream;
    const int64_t bytes = n * n * 8 + n * (5 * 4 + 2 + 8) + (n + 8) * 8 + 4096;
    int rc = vgs_ensure_scratch2(h, bytes);
    if (rc) return rc;
    char *p = (char *)h->d_scratch2;
    AlState s;
    s.n = n;
    s.D = (double *)p; p += n * n * 8;
    s.nn_d = (double *)p; p += n * 8;
    s.merges = (double *)p; p += (n + 8) * 8;
    s.size = (int *)p; p += n * 4;
    s.minmem = (int *)p; p += n * 4;
    s.okey = (int *)p; p += n * 4;
    s.nn_j = (int *)p; p += n * 4;
    s.parent = (int *)p; p += n * 4;
    s.sel = (int *)p; p += 64;
    s.alive = (unsigned char *)p; p += n;
    s.single = (unsigned char *)p; p += n;
    k_al_init<<<(unsigned)((n * n + 255) / 256), 256, 0, st>>>(s, D32_d);
    VGS_LAUNCH_CHECK(h);
    k_al_rescan_all<<<(unsigned)n, 128, 0, st>>>(s);
    VGS_LAUNCH_CHECK(h);
    const int64_t steps = n - k;
    // persistent path: needs the per-cluster state (5 bytes each) + a rescan list in shared memory and a co-resident grid
    bool persistent = false;
    {
        static int force_multi = -1;
        if (force_multi < 0) { const char *e = getenv("VGS_UPGMA_MULTI_KERNEL"); force_multi = (e && e[0] == '1') ? 1 : 0; }
        const int nb = h->sm_count;
        const int64_t rows_per_cta = (n + nb - 1) / nb;
        const int64_t smem = ((5 * n + 15) & ~(int64_t)15) + rows_per_cta * 4 + 16;
        if (!force_multi && steps > 0 && smem + 2048 <= (int64_t)h->max_smem_optin) {
            int coop = 0, per_sm = 0;
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
            if (cudaFuncSetAttribute(k_al_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_al_persistent, ALP_THREADS, (size_t)smem) == cudaSuccess &&
                coop && per_sm >= 1) {
                AlCand *d_cand = nullptr;
                unsigned int *d_bar = nullptr;
                if ((rc = vgs_ensure_scratch(h, (int64_t)nb * 4 * sizeof(AlCand) + 256))) return rc;
                d_bar = (unsigned int *)h->d_scratch;
                d_cand = (AlCand *)((char *)h->d_scratch + 256);
                VGS_CUDA(cudaMemsetAsync(d_bar, 0, 256, st));
                int steps_i = (int)steps;
                void *args[] = {(void *)&s, (void *)&steps_i, (void *)&d_cand, (void *)&d_bar};
                VGS_CUDA(cudaLaunchCooperativeKernel((const void *)k_al_persistent, dim3((unsigned)nb), dim3(ALP_THREADS), args, (size_t)smem, st));
                h->launches++;
                persistent = true;
            } else {
                cudaGetLastError();
            }
        }
    }
    for (int64_t step = 0; step < steps && !persistent; ++step) {
        k_al_argmin<<<1, 1024, 0, st>>>(s);
        k_al_update<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s);
        k_al_meta<<<1, 1, 0, st>>>(s, (int)step);
        k_al_refresh<<<(unsigned)n, 128, 0, st>>>(s);
        h->launches += 4;
        if ((step & 1023) == 1023) VGS_CUDA(cudaGetLastError());
    }
    VGS_CUDA(cudaGetLastError());
    std::vector<int> parent((size_t)n), minmem((size_t)n), sel(16);
    VGS_CUDA(cudaMemcpyAsync(parent.data(), s.parent, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    VGS_CUDA(cudaMemcpyAsync(minmem.data(), s.minmem, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    VGS_CUDA(cudaMemcpyAsync(sel.data(), s.sel, 16, cudaMemcpyDeviceToHost, st));
    VGS_CUDA(cudaStreamSynchronize(st));
    const int64_t n_merges = sel[3];
    if (merge_dist_out_h && n_merges > 0)
        VGS_CUDA(cudaMemcpy(merge_dist_out_h, s.merges, 8 * (size_t)n_merges, cudaMemcpyDeviceToHost));
    if (n_merges_out) *n_merges_out = n_merges;
    // labels: clusters numbered by their smallest member
    std::vector<int> root((size_t)n), remap((size_t)n, -1);
    for (int64_t i = 0; i < n; ++i) {
        int x = (int)i;
        while (parent[(size_t)x] != x) x = parent[(size_t)x];
        root[(size_t)i] = minmem[(size_t)x];
    }
    int next = 0;
    for (int64_t i = 0; i < n; ++i)
        if (root[(size_t)i] == (int)i) remap[(size_t)i] = next++;
    for (int64_t i = 0; i < n; ++i) labels_out_h[i] = remap[(size_t)root[(size_t)i]];
    return VGS_OK;
}
