// vgs_matrix.cu -- K7: tile plan, payload -> [N][N] scatter/mirror, and the reference's matrix hand-off
// layouts (src/clustering/util.pyx:6-33: float32 [N*N][3] rows (i, j, d) and float32 [N][N]).
#include "vgs_common.cuh"

void vgs_make_plan(TilePlan *p, int64_t n, int64_t tile, int world, int symmetric) {
    p->n = n;
    p->tile = tile;
    p->nt = (n + tile - 1) / tile;
    p->world = world;
    p->symmetric = symmetric;
    p->n_tiles = symmetric ? p->nt * (p->nt + 1) / 2 : p->nt * p->nt;
    p->slots = (p->n_tiles + world - 1) / world;
}

void vgs_tile_coords(const TilePlan &p, int64_t k, int64_t *I, int64_t *J) {
    if (!p.symmetric) {
        *I = k / p.nt;
        *J = k % p.nt;
        return;
    }
    int64_t i = 0, rem = k;
    while (rem >= p.nt - i) { rem -= p.nt - i; ++i; }
    *I = i;
    *J = i + rem;
}

extern "C" int vgs_tile_plan(int64_t n, int64_t tile, int world, int symmetric, int64_t *n_tiles, int64_t *slots_per_rank) {
    VGS_CHECK_ARG(n > 0 && tile > 0 && world > 0, "vgs_tile_plan: bad arguments");
    TilePlan p;
    vgs_make_plan(&p, n, tile, world, symmetric);
    if (n_tiles) *n_tiles = p.n_tiles;
    if (slots_per_rank) *slots_per_rank = p.slots;
    return VGS_OK;
}

template <typename T>
__global__ void k_assemble(TilePlan pl, const T *__restrict__ gathered, T *__restrict__ D) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (j >= pl.n) return;
    const int64_t Tl = pl.tile;
    int64_t I = i / Tl, J = j / Tl, ii = i % Tl, jj = j % Tl, k;
    if (pl.symmetric) {
        if (I > J) {  // mirror: D[i][j] = D[j][i] lives in tile (J, I)
            int64_t t = I; I = J; J = t;
            t = ii; ii = jj; jj = t;
        }
        k = vgs_tri_index(I, J, pl.nt);
    } else {
        k = I * pl.nt + J;
    }
    const int64_t owner = k % pl.world, slot = k / pl.world;
    D[i * pl.n + j] = gathered[((owner * pl.slots + slot) * Tl + ii) * Tl + jj];
}

int vgs_assemble_impl(vgs_context *h, const TilePlan &p, const float *g32, const double *g64, float *D32, double *D64,
                      cudaStream_t s) {
    dim3 block(256), grid((unsigned)((p.n + 255) / 256), (unsigned)p.n);
    VGS_CHECK_ARG(p.n <= 65535 * 16, "matrix too large for the assemble grid");
    if (p.n > 65535) {
        vgs_set_error("vgs_assemble: n > 65535 is not supported");
        return VGS_ERR_UNSUPPORTED;
    }
    if (D32 && g32) {
        k_assemble<float><<<grid, block, 0, s>>>(p, g32, D32);
        VGS_LAUNCH_CHECK(h);
    }
    if (D64 && g64) {
        k_assemble<double><<<grid, block, 0, s>>>(p, g64, D64);
        VGS_LAUNCH_CHECK(h);
    }
    return VGS_OK;
}

extern "C" int vgs_assemble(vgs_handle h, int64_t n, int64_t tile, int world, int symmetric, const float *gathered_d,
                            float *D32_d, void *stream) {
    VGS_CHECK_ARG(h && n > 0 && tile > 0 && world > 0 && gathered_d && D32_d, "vgs_assemble: bad arguments");
    VGS_CUDA(cudaSetDevice(h->device));
    TilePlan p;
    vgs_make_plan(&p, n, tile, world, symmetric);
    return vgs_assemble_impl(h, p, gathered_d, nullptr, D32_d, nullptr, (cudaStream_t)stream);
}

// float32 [N*N][3] rows (float(i), float(j), d) row-major in (i, j)  (src/clustering/util.pyx:12-20)
__global__ void k_triples(int64_t n, const float *__restrict__ D, float *__restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * n) return;
    out[3 * e + 0] = (float)(e / n);
    out[3 * e + 1] = (float)(e % n);
    out[3 * e + 2] = D[e];
}

extern "C" int vgs_matrix_to_triples(vgs_handle h, int64_t n, const float *D32_d, float *triples_d, void *stream) {
    VGS_CHECK_ARG(h && n > 0 && D32_d && triples_d, "vgs_matrix_to_triples: bad arguments");
    VGS_CUDA(cudaSetDevice(h->device));
    k_triples<<<(unsigned)((n * n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, D32_d, triples_d);
    VGS_LAUNCH_CHECK(h);
    return VGS_OK;
}
