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
    for (int64_t step = 0; step < steps; ++step) {
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
