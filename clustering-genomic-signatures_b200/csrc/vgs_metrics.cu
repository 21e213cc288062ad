// vgs_metrics.cu -- row f4: the O(N^2)/O(N^2 log N) passes the reference makes over the finished distance matrix.
//
// Reference behaviour restated (paths relative to the reference repository):
//   src/test_distance_function.py:75-79, :96-104 + src/util/distance_metrics.py:1-60
//        for every model i: distances to ALL models (itself included) sorted ascending with Python's stable
//        `sorted` (ties keep list order); k = number of models that share i's taxon;
//        procent_of_taxonomy_in_top = (# of the first k entries that share the taxon) / k           (:43-54)
//        average_distance_to_taxonomy = mean |d| over the entries that share the taxon              (:57-60)
//        average_distance = mean d over the whole row                                                (:11)
//   src/clustering/clustering_metrics.pyx:29-84   silhouette: a_i = mean distance to the other members of
//        i's cluster (0 for a singleton), b_i = smallest mean distance to another cluster, both returned
//        through a C float; s_i = (b_i - a_i) / max(b_i, a_i) in double
//   src/clustering/clustering_metrics.pyx:168-185 average_distance_std: per cluster numpy std (population,
//        fp64) of d(v1, v2) over ordered pairs v1 != v2, 0 for a singleton
//
// No sort is needed: the rank of entry j in the stable order is #{l : d_l < d_j} + #{l < j : d_l == d_j}, and only
// the ranks of the same-taxon entries matter.  One CTA per row, one warp per same-taxon entry, the row stays in
// L1/L2.  Sums are fp64 in a fixed order (the reference sums in sorted / set order, which is not reproducible).
#include <math.h>

#include <algorithm>
#include <numeric>

#include "vgs_common.cuh"

#define MET_THREADS 256
#define MET_WARPS (MET_THREADS / 32)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// perm[t][n]: model indices grouped by taxon; seg_begin/seg_count[t][i]: the group of row i inside perm[t]
template <typename T>
__global__ void __launch_bounds__(MET_THREADS) k_retrieval(int64_t n, const T *__restrict__ D, int n_tax,
                                                           const int32_t *__restrict__ perm, const int32_t *__restrict__ seg_begin,
                                                           const int32_t *__restrict__ seg_count, double *__restrict__ in_top,
                                                           double *__restrict__ dist_to, double *__restrict__ avg) {
    __shared__ double s_sum[MET_WARPS];
    __shared__ int s_cnt[MET_WARPS];
    const int64_t i = blockIdx.x;
    const T *row = D + i * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // mean of the row (fixed order: thread-strided partials, warp tree, warps in order)
    {
        double a = 0.0;
        for (int64_t l = threadIdx.x; l < n; l += MET_THREADS) a += (double)row[l];
        a = warp_sum(a);
        if (lane == 0) s_sum[warp] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < MET_WARPS; ++w) t += s_sum[w];
            avg[i] = t / (double)n;
        }
        __syncthreads();
    }
    for (int t = 0; t < n_tax; ++t) {
        const int k = seg_count[(int64_t)t * n + i];
        const int32_t *members = perm + (int64_t)t * n + seg_begin[(int64_t)t * n + i];
        int hits = 0;
        double asum = 0.0;
        for (int m = warp; m < k; m += MET_WARPS) {
            const int j = members[m];
            const T dj = row[j];
            int c = 0;
            for (int64_t l = lane; l < n; l += 32) {
                const T dl = row[l];
                c += (dl < dj || (dl == dj && l < j)) ? 1 : 0;
            }
            c = warp_sum_i(c);
            hits += c < k ? 1 : 0;
            asum += fabs((double)dj);
        }
        if (lane == 0) { s_cnt[warp] = hits; s_sum[warp] = asum; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int hsum = 0;
            double a = 0.0;
            for (int w = 0; w < MET_WARPS; ++w) { hsum += s_cnt[w]; a += s_sum[w]; }
            in_top[(int64_t)t * n + i] = (double)hsum / (double)k;
            dist_to[(int64_t)t * n + i] = a / (double)k;
        }
        __syncthreads();
    }
}

// clusters as segments of perm (cluster c = perm[off[c] .. off[c+1])); label[i] = cluster of i
__global__ void __launch_bounds__(MET_THREADS) k_silhouette(int64_t n, const float *__restrict__ D, int64_t n_clusters,
                                                            const int32_t *__restrict__ perm, const int32_t *__restrict__ off,
                                                            const int32_t *__restrict__ label, double *__restrict__ a_out,
                                                            double *__restrict__ b_out, double *__restrict__ s_out) {
    __shared__ float s_min[MET_WARPS];
    __shared__ float s_own;
    const int64_t i = blockIdx.x;
    const float *row = D + i * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int own = label[i];
    float best = INFINITY;
    for (int64_t c = warp; c < n_clusters; c += MET_WARPS) {
        const int b = off[c], e = off[c + 1];
        double sum = 0.0;
        for (int q = b + lane; q < e; q += 32) {
            const int j = perm[q];
            if (j != (int)i) sum += (double)row[j];
        }
        sum = warp_sum(sum);
        // the reference accumulates numpy float32 and divides in float32 (clustering_metrics.pyx:60-84)
        const float total = (float)sum;
        if (c == own) {
            const int others = e - b - 1;
            if (lane == 0) s_own = others == 0 ? 0.0f : total / (float)others;
        } else {
            const float mean = total / (float)(e - b);
            if (mean < best) best = mean;
        }
    }
    if (lane == 0) s_min[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        float b = INFINITY;
        for (int w = 0; w < MET_WARPS; ++w) b = s_min[w] < b ? s_min[w] : b;
        const double a = (double)s_own, bd = (double)b;
        a_out[i] = a;
        b_out[i] = bd;
        const double mx = bd > a ? bd : a;   // Python max(b, a): returns b unless a > b
        s_out[i] = (bd - a) / mx;            // inf/inf (one cluster) and 0/0 give NaN
    }
}

// population standard deviation of the off-diagonal distances inside each cluster, two passes, fp64
template <typename T>
__global__ void __launch_bounds__(MET_THREADS) k_cluster_std(int64_t n, const T *__restrict__ D, const int32_t *__restrict__ perm,
                                                             const int32_t *__restrict__ off, double *__restrict__ std_out) {
    __shared__ double s_part[MET_WARPS];
    __shared__ double s_mean;
    const int c = blockIdx.x;
    const int b = off[c], e = off[c + 1];
    const int64_t sz = e - b, m = sz * (sz - 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (m < 1) {
        if (threadIdx.x == 0) std_out[c] = 0.0;
        return;
    }
    for (int pass = 0; pass < 2; ++pass) {
        const double mean = pass ? s_mean : 0.0;
        double acc = 0.0;
        for (int64_t e2 = threadIdx.x; e2 < sz * sz; e2 += MET_THREADS) {
            const int64_t x = e2 / sz, y = e2 % sz;
            if (x == y) continue;
            const double v = (double)D[(int64_t)perm[b + x] * n + perm[b + y]];
            acc += pass ? (v - mean) * (v - mean) : v;
        }
        acc = warp_sum(acc);
        if (lane == 0) s_part[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < MET_WARPS; ++w) t += s_part[w];
            if (pass == 0) s_mean = t / (double)m;
            else std_out[c] = sqrt(t / (double)m);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------
// stable grouping of 0..n-1 by label: perm, and per element the begin/count of its group inside perm
static void group_by_label(const int32_t *label, int64_t n, std::vector<int32_t> &perm, std::vector<int32_t> &begin,
                           std::vector<int32_t> &count) {
    perm.resize((size_t)n);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return label[a] < label[b]; });
    begin.assign((size_t)n, 0);
    count.assign((size_t)n, 0);
    for (int64_t q = 0; q < n;) {
        int64_t r = q;
        while (r < n && label[perm[(size_t)r]] == label[perm[(size_t)q]]) ++r;
        for (int64_t x = q; x < r; ++x) { begin[(size_t)perm[(size_t)x]] = (int32_t)q; count[(size_t)perm[(size_t)x]] = (int32_t)(r - q); }
        q = r;
    }
}

extern "C" int vgs_retrieval_metrics(vgs_handle h, int64_t n, const float *D32_d, const double *D64_d, int n_tax,
                                     const int32_t *labels_h, double *in_top_out_h, double *dist_to_out_h, double *avg_out_h) {
    VGS_CHECK_ARG(h && n > 0 && (D32_d || D64_d) && n_tax >= 0 && n < ((int64_t)1 << 31), "vgs_retrieval_metrics: bad arguments");
    VGS_CHECK_ARG(n_tax == 0 || (labels_h && in_top_out_h && dist_to_out_h), "vgs_retrieval_metrics: null label / output arrays");
    VGS_CHECK_ARG(avg_out_h, "vgs_retrieval_metrics: null output");
    VGS_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->own_stream;
    const int64_t tn = (int64_t)n_tax * n;
    std::vector<int32_t> lists((size_t)(3 * tn) + 1);
    for (int t = 0; t < n_tax; ++t) {
        std::vector<int32_t> perm, begin, count;
        group_by_label(labels_h + (int64_t)t * n, n, perm, begin, count);
        std::copy(perm.begin(), perm.end(), lists.begin() + (size_t)(t * n));
        std::copy(begin.begin(), begin.end(), lists.begin() + (size_t)(tn + t * n));
        std::copy(count.begin(), count.end(), lists.begin() + (size_t)(2 * tn + t * n));
    }
    const int64_t list_bytes = (3 * tn * 4 + 255) / 256 * 256, out_elems = 2 * tn + n;
    int rc = vgs_ensure_scratch2(h, list_bytes + out_elems * 8 + 256);
    if (rc) return rc;
    int32_t *d_lists = (int32_t *)h->d_scratch2;
    double *d_out = (double *)((char *)h->d_scratch2 + list_bytes);
    if (tn > 0) VGS_CUDA(cudaMemcpyAsync(d_lists, lists.data(), (size_t)(3 * tn * 4), cudaMemcpyHostToDevice, st));
    if (D64_d)
        k_retrieval<double><<<(unsigned)n, MET_THREADS, 0, st>>>(n, D64_d, n_tax, d_lists, d_lists + tn, d_lists + 2 * tn, d_out,
                                                                 d_out + tn, d_out + 2 * tn);
    else
        k_retrieval<float><<<(unsigned)n, MET_THREADS, 0, st>>>(n, D32_d, n_tax, d_lists, d_lists + tn, d_lists + 2 * tn, d_out,
                                                                d_out + tn, d_out + 2 * tn);
    VGS_LAUNCH_CHECK(h);
    if (tn > 0) {
        VGS_CUDA(cudaMemcpyAsync(in_top_out_h, d_out, (size_t)(tn * 8), cudaMemcpyDeviceToHost, st));
        VGS_CUDA(cudaMemcpyAsync(dist_to_out_h, d_out + tn, (size_t)(tn * 8), cudaMemcpyDeviceToHost, st));
    }
    VGS_CUDA(cudaMemcpyAsync(avg_out_h, d_out + 2 * tn, (size_t)(n * 8), cudaMemcpyDeviceToHost, st));
    VGS_CUDA(cudaStreamSynchronize(st));
    return VGS_OK;
}

// clusters -> (perm, off[k+1], dense label) with clusters numbered in order of first appearance of their label
static int64_t cluster_segments(const int32_t *cluster, int64_t n, std::vector<int32_t> &perm, std::vector<int32_t> &off,
                                std::vector<int32_t> &dense) {
    std::vector<int32_t> begin, count;
    group_by_label(cluster, n, perm, begin, count);
    off.clear();
    dense.assign((size_t)n, 0);
    int32_t c = -1;
    for (int64_t q = 0; q < n; ++q) {
        if (q == 0 || cluster[perm[(size_t)q]] != cluster[perm[(size_t)q - 1]]) { off.push_back((int32_t)q); ++c; }
        dense[(size_t)perm[(size_t)q]] = c;
    }
    off.push_back((int32_t)n);
    return (int64_t)off.size() - 1;
}

extern "C" int vgs_silhouette(vgs_handle h, int64_t n, const float *D32_d, const int32_t *cluster_h, double *a_out_h,
                              double *b_out_h, double *s_out_h) {
    VGS_CHECK_ARG(h && n > 0 && D32_d && cluster_h && s_out_h && n < ((int64_t)1 << 31), "vgs_silhouette: bad arguments");
    VGS_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->own_stream;
    std::vector<int32_t> perm, off, dense;
    const int64_t k = cluster_segments(cluster_h, n, perm, off, dense);
    const int64_t list_bytes = ((2 * n + k + 1) * 4 + 255) / 256 * 256;
    int rc = vgs_ensure_scratch2(h, list_bytes + 3 * n * 8 + 256);
    if (rc) return rc;
    int32_t *d_perm = (int32_t *)h->d_scratch2, *d_label = d_perm + n, *d_off = d_label + n;
    double *d_out = (double *)((char *)h->d_scratch2 + list_bytes);
    VGS_CUDA(cudaMemcpyAsync(d_perm, perm.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    VGS_CUDA(cudaMemcpyAsync(d_label, dense.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    VGS_CUDA(cudaMemcpyAsync(d_off, off.data(), (size_t)(k + 1) * 4, cudaMemcpyHostToDevice, st));
    k_silhouette<<<(unsigned)n, MET_THREADS, 0, st>>>(n, D32_d, k, d_perm, d_off, d_label, d_out, d_out + n, d_out + 2 * n);
    VGS_LAUNCH_CHECK(h);
    if (a_out_h) VGS_CUDA(cudaMemcpyAsync(a_out_h, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    if (b_out_h) VGS_CUDA(cudaMemcpyAsync(b_out_h, d_out + n, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    VGS_CUDA(cudaMemcpyAsync(s_out_h, d_out + 2 * n, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    VGS_CUDA(cudaStreamSynchronize(st));
    return VGS_OK;
}

extern "C" int vgs_cluster_distance_std(vgs_handle h, int64_t n, const float *D32_d, const double *D64_d, const int32_t *cluster_h,
                                        int64_t capacity, double *std_out_h, int64_t *n_clusters_out) {
    VGS_CHECK_ARG(h && n > 0 && (D32_d || D64_d) && cluster_h && std_out_h && n < ((int64_t)1 << 31), "vgs_cluster_distance_std: bad arguments");
    VGS_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->own_stream;
    std::vector<int32_t> perm, off, dense;
    const int64_t k = cluster_segments(cluster_h, n, perm, off, dense);
    if (n_clusters_out) *n_clusters_out = k;
    VGS_CHECK_ARG(capacity >= k, "vgs_cluster_distance_std: output capacity below the number of clusters");
    const int64_t list_bytes = ((n + k + 1) * 4 + 255) / 256 * 256;
    int rc = vgs_ensure_scratch2(h, list_bytes + k * 8 + 256);
    if (rc) return rc;
    int32_t *d_perm = (int32_t *)h->d_scratch2, *d_off = d_perm + n;
    double *d_out = (double *)((char *)h->d_scratch2 + list_bytes);
    VGS_CUDA(cudaMemcpyAsync(d_perm, perm.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    VGS_CUDA(cudaMemcpyAsync(d_off, off.data(), (size_t)(k + 1) * 4, cudaMemcpyHostToDevice, st));
    if (D64_d) k_cluster_std<double><<<(unsigned)k, MET_THREADS, 0, st>>>(n, D64_d, d_perm, d_off, d_out);
    else k_cluster_std<float><<<(unsigned)k, MET_THREADS, 0, st>>>(n, D32_d, d_perm, d_off, d_out);
    VGS_LAUNCH_CHECK(h);
    VGS_CUDA(cudaMemcpyAsync(std_out_h, d_out, (size_t)k * 8, cudaMemcpyDeviceToHost, st));
    VGS_CUDA(cudaStreamSynchronize(st));
    return VGS_OK;
}
