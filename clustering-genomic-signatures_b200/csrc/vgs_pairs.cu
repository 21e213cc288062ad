// vgs_pairs.cu -- K4: FrobeniusNorm (intersection / union) and Projection for all pairs.
//
// Reference behaviour restated (paths relative to the reference repository):
//   src/distance/frobenius.pyx:21-53   C = keys_l ∩ keys_r (default) or ∪; a context missing from a model
//        takes the row of its longest suffix present there (:47-50); rows are float32; the difference is a
//        float32 subtraction; d = ||P_l - P_r||_F / sqrt(|C|)
//   src/distance/projection.pyx:7-36   every model embedded over the union universe with zero rows for
//        absent contexts; d = ||v_l - v_r||_2 (not normalised)
// Here the float32 subtraction is kept and the sum of squares is accumulated in fp64 (each square of a
// float32 difference is exact in fp64), so results agree with the oracle to fp64 rounding and with the
// reference's float32 BLAS dot to ~1e-7 relative (its own rounding).
//
// Kernel shape.  One CTA = (tile slot, direction, group of JG "staged" models y).  The staged models'
// exact-context hashes and fp32 rows are copied into shared memory with bulk async copies (TMA, UBLKCP);
// each warp then streams the contexts of one "streamed" model x at a time (coalesced key + float4 row
// loads, 32 contexts per step) and probes all JG staged hashes with the same registers.  Directed
// partial results go to scratch:  S[x][y] = sum over shared contexts of |a_x - a_y|^2,
// U[x][y] = sum over contexts of x missing from y (vs. y's fallback row, or vs. zero for Projection),
// C[x][y] = number of shared contexts.  A second small kernel combines the two directions per entry.
#include <algorithm>

#include "vgs_common.cuh"

#define PAIR_THREADS 256
#define PAIR_MAX_JG 8   // staged models per CTA, shallow sets (universe maps)
#define PAIR_DEEP_JG 4  // staged models per CTA, deep sets (perfect hashes)

struct PairParams {
    const ModelMeta *meta;
    const uint32_t *key;
    const float4 *prob32;
    const uint2 *hash;
    const uint16_t *umap;
    int64_t umap_stride;
    const uint8_t *ph;      // perfect-hash arena (deep sets)
    TilePlan pl;
    int rank, kind, jg, n_groups, n_dir;
    double *S, *U, *V;      // V: Projection only (squares of the staged model's rows at the shared contexts)
    int32_t *C;
    const double *norm2;    // Projection only: sum of squares of every fp32 row entry, per model
    int *err;
};

__device__ __forceinline__ void tile_of_slot(const TilePlan &pl, int64_t k, int64_t *I, int64_t *J) {
    if (!pl.symmetric) { *I = k / pl.nt; *J = k % pl.nt; return; }
    int64_t i = 0, rem = k;
    while (rem >= pl.nt - i) { rem -= pl.nt - i; ++i; }
    *I = i; *J = i + rem;
}

__device__ __forceinline__ double sq4(const float4 a, const float4 b) {
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;  // float32 subtraction (frobenius.pyx:34)
    // each product of two float32 values is exact in fp64, so the fused form rounds exactly like mul-then-add
    double r = __dmul_rn((double)dx, (double)dx);
    r = __fma_rn((double)dy, (double)dy, r);
    r = __fma_rn((double)dz, (double)dz, r);
    r = __fma_rn((double)dw, (double)dw, r);
    return r;
}

__device__ __forceinline__ double norm4(const float4 a) {
    double r = __dmul_rn((double)a.x, (double)a.x);
    r = __fma_rn((double)a.y, (double)a.y, r);
    r = __fma_rn((double)a.z, (double)a.z, r);
    r = __fma_rn((double)a.w, (double)a.w, r);
    return r;
}

// per-model sum of squares in exactly the order k_pair_partials accumulates a streamed model (lane-strided partial
// sums, xor-shuffle tree): when every context of a model is shared with its partner, the two sums are equal bit for bit
__global__ void k_pair_norms(int64_t n_models, const ModelMeta *__restrict__ meta, const float4 *__restrict__ prob32,
                             double *__restrict__ norm2) {
    const int lane = threadIdx.x & 31;
    const int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= n_models) return;
    const ModelMeta mm = meta[m];
    double a = 0.0;
    for (int ci = lane; ci < mm.n_ctx; ci += 32) a += norm4(__ldg(prob32 + mm.ctx_off + ci));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) norm2[m] = a;
}

// UMAP = true : shallow sets; each staged model carries a dense map over every string of length <= max order
//               (one u16 gather gives presence AND the longest-suffix fallback row: no probing loop, no divergence)
// UMAP = false: deep sets; exact-context hash probes, and for the union a probe per shorter suffix
// KIND: Projection needs one direction only: ||v_x - v_y||^2 = S + (|v_x|^2 - A_x) + (|v_y|^2 - A_y) with S the squared
// differences and A_x, A_y the squares of x's and y's rows over the SHARED contexts (an absent context is a zero row)
// THREADS: one CTA per SM (the staged tables fill shared memory), so the warps that hide the gather latency must all
// come from that CTA: 1024 threads where the accumulators leave 64 registers per thread, 512 otherwise.
template <bool STAGED, bool UMAP, int MAXJ, int KIND, int THREADS>
__global__ void __launch_bounds__(THREADS) k_pair_partials(const PairParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ int64_t sI, sJ;
    __shared__ ModelMeta ymeta[MAXJ];
    __shared__ uint32_t yoff_h[MAXJ], yoff_r[MAXJ];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t w = blockIdx.x;
    const int g = (int)(w % p.n_groups); w /= p.n_groups;
    const int dir = (int)(w % p.n_dir);
    const int64_t q = w / p.n_dir;
    const int64_t k = q * p.pl.world + p.rank;
    if (k >= p.pl.n_tiles) return;
    if (tid == 0) {
        int64_t I, J;
        tile_of_slot(p.pl, k, &I, &J);
        sI = I; sJ = J;
        if (STAGED) vgs_mbar_init(&bar, 1);
    }
    __syncthreads();
    const int64_t T = p.pl.tile;
    const int64_t I = sI, J = sJ;
    if (dir == 1 && I == J) return;  // a diagonal tile holds every ordered pair already
    const int64_t Xb = dir ? J : I, Yb = dir ? I : J;
    const int64_t y0 = Yb * T + (int64_t)g * p.jg;
    const int64_t y_end = min(p.pl.n, (Yb + 1) * T);
    const int ny = (int)min((int64_t)p.jg, y_end - y0);
    if (ny <= 0) return;
    const int64_t x0 = Xb * T;
    const int nx = (int)(min(p.pl.n, (Xb + 1) * T) - x0);

    if (tid < ny) ymeta[tid] = p.meta[y0 + tid];
    __syncthreads();
    const uint32_t map_bytes = UMAP ? (uint32_t)p.umap_stride * 2u : 0u;
    if (STAGED) {
        if (tid == 0) {
            uint32_t off = 0, total = 0;
            for (int yy = 0; yy < ny; ++yy) {
                const uint32_t hb = UMAP ? map_bytes : (uint32_t)ymeta[yy].ph_bytes, rb = (uint32_t)ymeta[yy].n_ctx * 16u;
                yoff_h[yy] = off; off += hb;
                yoff_r[yy] = off; off += rb;
                total += hb + rb;
            }
            vgs_mbar_expect_tx(&bar, total);
            for (int yy = 0; yy < ny; ++yy) {
                if (UMAP) vgs_copy_pieces(smem + yoff_h[yy], p.umap + (y0 + yy) * p.umap_stride, map_bytes, &bar);
                else vgs_copy_pieces(smem + yoff_h[yy], p.ph + ymeta[yy].ph_off, (uint32_t)ymeta[yy].ph_bytes, &bar);
                vgs_copy_pieces(smem + yoff_r[yy], p.prob32 + ymeta[yy].ctx_off, (uint32_t)ymeta[yy].n_ctx * 16u, &bar);
            }
        }
        __syncthreads();
        vgs_mbar_wait(&bar, 0);
    }

    const void *yh[MAXJ];
    const float4 *yr[MAXJ];
#pragma unroll
    for (int yy = 0; yy < MAXJ; ++yy) {
        if (yy < ny) {
            if (STAGED) yh[yy] = smem + yoff_h[yy];
            else yh[yy] = UMAP ? (const void *)(p.umap + (y0 + yy) * p.umap_stride) : (const void *)(p.ph + ymeta[yy].ph_off);
            yr[yy] = STAGED ? (const float4 *)(smem + yoff_r[yy]) : p.prob32 + ymeta[yy].ctx_off;
        } else {
            yh[yy] = nullptr; yr[yy] = nullptr;
        }
    }

    // per staged model, in registers for the whole CTA lifetime: perfect-hash multipliers and shifts
    uint32_t ya2[MAXJ], ya3[MAXJ], ysh[MAXJ];   // ysh = (32 - lgr) | (32 - lgcap) << 8 | (lgr << 16)
    bool all_ph = true;
    if (!UMAP) {
#pragma unroll
        for (int yy = 0; yy < MAXJ; ++yy) {
            ya2[yy] = 0u; ya3[yy] = 0u; ysh[yy] = 0u;
            if (yy < ny) {
                const int seed = ymeta[yy].ph_seed;
                all_ph = all_ph && seed >= 0;
                ya2[yy] = vgs_ph_a2(seed); ya3[yy] = vgs_ph_a3(seed);
                ysh[yy] = (uint32_t)(32 - ymeta[yy].ph_lgr) | ((uint32_t)ymeta[yy].hash_shift << 8) | ((uint32_t)ymeta[yy].ph_lgr << 16);
            }
        }
    }
    const int64_t part_base = ((q * 2 + dir) * T) * T;
    constexpr int kind = KIND;
    for (int xx = warp; xx < nx; xx += THREADS / 32) {
        const ModelMeta xm = p.meta[x0 + xx];
        double s[MAXJ], u[MAXJ], v[MAXJ];   // v is dead (and removed by the compiler) unless KIND is Projection
        int c[MAXJ];
#pragma unroll
        for (int yy = 0; yy < MAXJ; ++yy) { s[yy] = 0.0; u[yy] = 0.0; v[yy] = 0.0; c[yy] = 0; }
        for (int ci = lane; ci < xm.n_ctx; ci += 32) {
            const uint32_t key = __ldg(p.key + xm.ctx_off + ci);
            const float4 px = __ldg(p.prob32 + xm.ctx_off + ci);
            const int len = (31 - __clz(key)) >> 1;
            const uint32_t code = key ^ (1u << (2 * len));
            if (UMAP) {
                const uint32_t uidx = (((1u << (2 * len)) - 1u) / 3u) + code;
#pragma unroll
                for (int yy = 0; yy < MAXJ; ++yy) {
                    if (yy >= ny) break;
                    const uint32_t e = ((const uint16_t *)yh[yy])[uidx];
                    const uint32_t id = e & 0x7FFFu;
                    const bool exact = (e & 0x8000u) != 0u;
                    if (kind == VGS_PAIR_FROBENIUS_INTERSECTION) {
                        if (exact) { s[yy] += sq4(px, yr[yy][id]); c[yy] += 1; }
                    } else if (kind == VGS_PAIR_FROBENIUS_UNION) {
                        if (id == VGS_UMAP_NONE) { atomicOr(p.err, 1); continue; }
                        const double v = sq4(px, yr[yy][id]);
                        if (exact) { s[yy] += v; c[yy] += 1; } else { u[yy] += v; }
                    } else {
                        if (exact) {
                            const float4 py = yr[yy][id];
                            s[yy] += sq4(px, py); u[yy] += norm4(px); v[yy] += norm4(py); c[yy] += 1;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int yy = 0; yy < MAXJ; ++yy) {
                    if (yy >= ny) break;
                    // exact-context probe: perfect hash (two dependent loads, no loop); a model whose build failed
                    // (never seen so far) falls back to the probing hash in global memory
                    const uint8_t *blob = (const uint8_t *)yh[yy];
                    int id;
                    if (all_ph) {
                        const uint32_t sh = ysh[yy];
                        const uint32_t dsp = ((const uint16_t *)blob)[(key * 0x9E3779B1u) >> (sh & 31u)];
                        const uint32_t csh = (sh >> 8) & 63u;
                        const uint32_t h2 = (key * ya2[yy]) >> csh, h3 = ((key * ya3[yy]) >> csh) | 1u;
                        const uint2 e = ((const uint2 *)(blob + (((2u << (sh >> 16)) + 15u) & ~15u)))[(h2 + dsp * h3) & (0xFFFFFFFFu >> csh)];
                        id = e.x == key ? (int)e.y : -1;
                    } else {
                        const int seed = ymeta[yy].ph_seed;
                        id = seed >= 0 ? vgs_ph_find(blob, ymeta[yy].ph_lgr, 32 - ymeta[yy].hash_shift, seed, key)
                                       : vgs_hash_find(p.hash + ymeta[yy].hash_off, ymeta[yy].hash_mask, ymeta[yy].hash_shift, key);
                    }
                    if (id >= 0) {
                        const float4 py = yr[yy][id];
                        s[yy] += sq4(px, py);
                        c[yy] += 1;
                        if (kind == VGS_PAIR_PROJECTION) { u[yy] += norm4(px); v[yy] += norm4(py); }
                    } else if (kind == VGS_PAIR_FROBENIUS_UNION) {
                        int top = len - 1 < ymeta[yy].order ? len - 1 : ymeta[yy].order;
                        for (int ln = top; ln >= 0 && id < 0; --ln) {
                            const uint32_t k2 = vgs_key(ln, code & vgs_lowmask(ln));
                            const int seed = ymeta[yy].ph_seed;
                            id = seed >= 0 ? vgs_ph_find(blob, ymeta[yy].ph_lgr, 32 - ymeta[yy].hash_shift, seed, k2)
                                           : vgs_hash_find(p.hash + ymeta[yy].hash_off, ymeta[yy].hash_mask, ymeta[yy].hash_shift, k2);
                        }
                        if (id < 0) atomicOr(p.err, 1);
                        else u[yy] += sq4(px, yr[yy][id]);
                    }
                }
            }
        }
#pragma unroll
        for (int yy = 0; yy < MAXJ; ++yy) {
            if (yy >= ny) break;
            double sv = s[yy], uv = u[yy], vv = v[yy];
            int cv = c[yy];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sv += __shfl_xor_sync(0xffffffffu, sv, o);
                uv += __shfl_xor_sync(0xffffffffu, uv, o);
                if (kind == VGS_PAIR_PROJECTION) vv += __shfl_xor_sync(0xffffffffu, vv, o);
                cv += __shfl_xor_sync(0xffffffffu, cv, o);
            }
            if (lane == 0) {
                const int64_t e = part_base + (int64_t)xx * T + (g * p.jg + yy);
                p.S[e] = sv; p.U[e] = uv; p.C[e] = cv;
                if (kind == VGS_PAIR_PROJECTION) p.V[e] = vv;
            }
        }
    }
}

__global__ void k_pair_combine(PairParams p, float *__restrict__ pay32, double *__restrict__ pay64) {
    const int64_t q = blockIdx.x;
    const int64_t k = q * p.pl.world + p.rank;
    const int64_t T = p.pl.tile;
    __shared__ int64_t sI, sJ;
    if (threadIdx.x == 0) {
        int64_t I = 0, J = 0;
        if (k < p.pl.n_tiles) tile_of_slot(p.pl, k, &I, &J);
        sI = I; sJ = J;
    }
    __syncthreads();
    const int64_t I = sI, J = sJ;
    const double *S0 = p.S + (q * 2) * T * T, *U0 = p.U + (q * 2) * T * T, *U1 = p.U + (q * 2 + 1) * T * T;
    const int32_t *C0 = p.C + (q * 2) * T * T;
    for (int64_t e = threadIdx.x; e < T * T; e += blockDim.x) {
        const int64_t ii = e / T, jj = e % T, i = I * T + ii, j = J * T + jj;
        double d = 0.0;
        if (k < p.pl.n_tiles && i < p.pl.n && j < p.pl.n) {
            // canonical orientation (a, b): inside a diagonal tile both (i, j) and (j, i) were accumulated, in
            // different context orders; reading the a <= b copy for both makes the matrix exactly symmetric
            int64_t a = ii, b = jj;
            if (I == J && ii > jj) { a = jj; b = ii; }
            const int64_t ab = a * T + b, ba = b * T + a;
            const double s = S0[ab];
            const int cnt = C0[ab];
            if (p.kind == VGS_PAIR_FROBENIUS_INTERSECTION) {
                d = cnt > 0 ? sqrt(s) / sqrt((double)cnt) : vgs_nan();  // empty intersection: 0/0 in the reference
            } else if (p.kind == VGS_PAIR_PROJECTION) {
                // x = the streamed model of the canonical orientation (tile row a), y = the staged one (column b).
                // The unshared squares come from the norms; exactly zero when every context is shared, otherwise at
                // least one whole probability row (>= 0.25), far above the cancellation error of the subtraction.
                const int64_t mx = I * T + a, my = J * T + b;
                const double ux = cnt == p.meta[mx].n_ctx ? 0.0 : fmax(0.0, p.norm2[mx] - U0[ab]);
                const double uy = cnt == p.meta[my].n_ctx ? 0.0 : fmax(0.0, p.norm2[my] - p.V[(q * 2) * T * T + ab]);
                d = sqrt(s + ux + uy);
            } else {
                const double u_ij = U0[ab];
                const double u_ji = (I == J) ? U0[ba] : U1[ba];
                const double ssq = s + u_ij + u_ji;
                const int n_union = p.meta[i].n_ctx + p.meta[j].n_ctx - cnt;
                d = sqrt(ssq) / sqrt((double)n_union);
            }
        }
        pay32[q * T * T + e] = (float)d;
        if (pay64) pay64[q * T * T + e] = d;
    }
}

template <int KIND>
constexpr int pair_threads() { return KIND == VGS_PAIR_FROBENIUS_INTERSECTION ? 1024 : 512; }

template <int KIND>
static int launch_pair(const PairParams &p, bool staged, bool umap, int64_t n_cta, int smem, cudaStream_t s) {
    constexpr int TH = pair_threads<KIND>();
    if (staged && umap) {
        VGS_CUDA(cudaFuncSetAttribute(k_pair_partials<true, true, PAIR_MAX_JG, KIND, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k_pair_partials<true, true, PAIR_MAX_JG, KIND, TH><<<(unsigned)n_cta, TH, smem, s>>>(p);
    } else if (staged) {
        VGS_CUDA(cudaFuncSetAttribute(k_pair_partials<true, false, PAIR_DEEP_JG, KIND, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k_pair_partials<true, false, PAIR_DEEP_JG, KIND, TH><<<(unsigned)n_cta, TH, smem, s>>>(p);
    } else if (umap) {
        k_pair_partials<false, true, PAIR_MAX_JG, KIND, 256><<<(unsigned)n_cta, 256, 0, s>>>(p);
    } else {
        k_pair_partials<false, false, PAIR_DEEP_JG, KIND, 256><<<(unsigned)n_cta, 256, 0, s>>>(p);
    }
    return VGS_OK;
}

static int pair_tiles_impl(vgs_context *h, int kind, const TilePlan &pl, int rank, float *pay32, double *pay64, cudaStream_t s) {
    PairParams p;
    p.meta = h->d_meta; p.key = h->d_key; p.prob32 = h->d_prob32; p.hash = h->d_hash;
    p.pl = pl; p.rank = rank; p.kind = kind; p.err = h->d_err;
    p.n_dir = kind == VGS_PAIR_FROBENIUS_UNION ? 2 : 1;
    // how many staged models fit?
    {
        int rcu = vgs_ensure_umap(h, s);   // shallow sets: dense universe maps, built on first use
        if (rcu) return rcu;
    }
    const bool umap = h->umap_stride > 0;
    p.umap = h->d_umap; p.umap_stride = h->umap_stride;
    if (!umap) {
        int rc0 = vgs_ensure_ph(h, s);
        if (rc0) return rc0;
    }
    p.ph = h->d_ph;
    const int64_t per_model_raw = umap ? h->umap_stride * 2 + (int64_t)h->max_n_ctx * 16 : h->max_ph_bytes + (int64_t)h->max_n_ctx * 16;
    const int64_t per_model = (per_model_raw + 127) / 128 * 128;
    const int64_t cap = (int64_t)h->max_smem_optin - 2048;
    bool staged = per_model <= cap;
    int jg = 4;
    if (staged) {
        // one CTA per SM with up to 1024 threads: as many staged models as fit (each streamed context is then probed
        // against all of them from the same registers)
        const int64_t max_jg = umap ? PAIR_MAX_JG : PAIR_DEEP_JG;
        jg = (int)std::min<int64_t>(max_jg, cap / per_model);
        if (jg < 1) jg = 1;
    }
    jg = (int)std::min<int64_t>(jg, pl.tile);
    if (!umap) jg = std::min(jg, PAIR_DEEP_JG);
    p.jg = jg;
    p.n_groups = (int)((pl.tile + jg - 1) / jg);
    const int64_t part = pl.slots * 2 * pl.tile * pl.tile;
    int rc = vgs_ensure_scratch2(h, part * 28 + h->n_models * 8 + 1024);
    if (rc) return rc;
    p.S = (double *)h->d_scratch2;
    p.U = p.S + part;
    p.V = p.U + part;
    double *norm2 = p.V + part;
    p.norm2 = norm2;
    p.C = (int32_t *)(norm2 + h->n_models);
    if (pl.slots == 0) return VGS_OK;
    VGS_CUDA(cudaMemsetAsync(h->d_scratch2, 0, (size_t)(part * 28 + h->n_models * 8), s));
    if (kind == VGS_PAIR_PROJECTION) {
        k_pair_norms<<<(unsigned)((h->n_models * 32 + 255) / 256), 256, 0, s>>>(h->n_models, h->d_meta, h->d_prob32, norm2);
        VGS_LAUNCH_CHECK(h);
    }
    const int64_t n_cta = pl.slots * p.n_dir * p.n_groups;
    VGS_CHECK_ARG(n_cta < ((int64_t)1 << 31), "pair kernel grid too large");
    const int smem = staged ? (int)(per_model * jg) : 0;
    vgs_prof_begin(h, VGS_PROF_PAIR_PARTIALS, s);
    if (kind == VGS_PAIR_FROBENIUS_INTERSECTION) rc = launch_pair<VGS_PAIR_FROBENIUS_INTERSECTION>(p, staged, umap, n_cta, smem, s);
    else if (kind == VGS_PAIR_FROBENIUS_UNION) rc = launch_pair<VGS_PAIR_FROBENIUS_UNION>(p, staged, umap, n_cta, smem, s);
    else rc = launch_pair<VGS_PAIR_PROJECTION>(p, staged, umap, n_cta, smem, s);
    if (rc) return rc;
    vgs_prof_end(h, VGS_PROF_PAIR_PARTIALS, s);
    VGS_LAUNCH_CHECK(h);
    k_pair_combine<<<(unsigned)pl.slots, 256, 0, s>>>(p, pay32, pay64);
    VGS_LAUNCH_CHECK(h);
    return VGS_OK;
}

static int check_pair_args(vgs_context *h, int kind) {
    VGS_CHECK_ARG(h, "null handle");
    VGS_CHECK_ARG(h->n_models > 0, "no models loaded (vgs_load_models)");
    VGS_CHECK_ARG(h->prob_ready, "models were loaded with VGS_LOAD_NLL_ONLY: no transition rows on the device");
    VGS_CHECK_ARG(kind == VGS_PAIR_FROBENIUS_INTERSECTION || kind == VGS_PAIR_FROBENIUS_UNION || kind == VGS_PAIR_PROJECTION,
                  "bad pair-distance kind");
    return VGS_OK;
}

extern "C" int vgs_last_pair_kernel(vgs_handle h) { return h ? h->last_pair_kernel : 0; }
// which kernels vgs_pair_matrix would use for `kind` on the loaded set (the tile entry points always use the CUDA cores)
extern "C" int vgs_pair_kernel_choice(vgs_handle h, int kind) {
    if (!h || h->n_models <= 0 || !h->prob_ready) return 0;
    if (!vgs_pair_tc_wanted(h, kind)) return 0;
    return (h->tc_gen == h->model_gen && !h->tc_ok) ? 0 : 1;
}
extern "C" int vgs_set_pair_path(vgs_handle h, int path) {
    VGS_CHECK_ARG(h && path >= -1 && path <= 1, "vgs_set_pair_path: path is -1 (by density), 0 (CUDA cores) or 1 (tensor cores)");
    h->pair_path = path;
    return VGS_OK;
}

extern "C" int vgs_pair_tiles(vgs_handle h, int kind, int64_t tile, int rank, int world, float *payload_d, void *stream) {
    int rc = check_pair_args(h, kind);
    if (rc) return rc;
    VGS_CHECK_ARG(tile > 0 && world > 0 && rank >= 0 && rank < world && payload_d, "vgs_pair_tiles: bad arguments");
    VGS_CUDA(cudaSetDevice(h->device));
    TilePlan pl;
    vgs_make_plan(&pl, h->n_models, tile, world, 1);
    return pair_tiles_impl(h, kind, pl, rank, payload_d, nullptr, (cudaStream_t)stream);
}

extern "C" int vgs_pair_matrix(vgs_handle h, int kind, float *D32_d, double *D64_d, void *stream) {
    int rc = check_pair_args(h, kind);
    if (rc) return rc;
    VGS_CHECK_ARG(D32_d, "vgs_pair_matrix: null output");
    VGS_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    // dense enough universes (full chains, shallow trees): exact integer GEMMs on the tensor cores (vgs_pairs_tc.cu)
    if (vgs_pair_tc_wanted(h, kind)) {
        rc = vgs_pair_tc_matrix(h, kind, D32_d, D64_d, s);
        if (rc != -1) return rc;
    }
    h->last_pair_kernel = 0;
    TilePlan pl;
    const int64_t n = h->n_models;
    // tile edge: 64/128 normally; a few huge models (tables larger than shared memory, read from L2/HBM) need many
    // more CTAs than pairs/4096, so the tile shrinks until the grid covers the SMs a few times
    int64_t tile = n <= 2048 ? 64 : 128;
    if ((rc = vgs_ensure_umap(h, s))) return rc;
    {
        const int64_t per_model = h->umap_stride > 0 ? h->umap_stride * 2 + (int64_t)h->max_n_ctx * 16
                                                     : h->max_pair_bytes + (int64_t)h->max_n_ctx * 16;
        if (per_model > (int64_t)h->max_smem_optin - 2048)
            while (tile > 4 && ((n + tile - 1) / tile) * ((n + tile - 1) / tile + 1) / 2 * ((tile + 3) / 4) < 4 * h->sm_count) tile /= 2;
    }
    vgs_make_plan(&pl, n, tile, 1, 1);
    const int64_t pay_elems = pl.slots * pl.tile * pl.tile;
    if ((rc = vgs_ensure_scratch(h, pay_elems * (D64_d ? 12 : 4) + 256))) return rc;
    float *pay32 = (float *)h->d_scratch;
    double *pay64 = D64_d ? (double *)((char *)h->d_scratch + (pay_elems * 4 + 255) / 256 * 256) : nullptr;
    if ((rc = pair_tiles_impl(h, kind, pl, 0, pay32, pay64, s))) return rc;
    return vgs_assemble_impl(h, pl, pay32, pay64, D32_d, D64_d, s);
}

extern "C" int vgs_pair_matrix_h(vgs_handle h, int kind, float *D32_out_h, double *D64_out_h) {
    int rc = check_pair_args(h, kind);
    if (rc) return rc;
    VGS_CHECK_ARG(D32_out_h || D64_out_h, "vgs_pair_matrix_h: no output");
    VGS_CUDA(cudaSetDevice(h->device));
    const int64_t n = h->n_models;
    if ((rc = vgs_reserve(h, &h->d_out32, n * n))) return rc;
    if (D64_out_h && (rc = vgs_reserve(h, &h->d_out64, n * n))) return rc;
    float *d32 = h->d_out32;
    double *d64 = D64_out_h ? h->d_out64 : nullptr;
    cudaStream_t s = h->own_stream;
    rc = vgs_pair_matrix(h, kind, d32, d64, s);
    // the result starts for the host behind the kernels and the error word is read behind it: ONE wait for all three (on
    // an error the caller's buffers hold an unspecified matrix and the call reports the error)
    if (!rc && D32_out_h) rc = cudaMemcpyAsync(D32_out_h, d32, 4 * (size_t)(n * n), cudaMemcpyDeviceToHost, s) == cudaSuccess ? VGS_OK : VGS_ERR_CUDA;
    if (!rc && D64_out_h) rc = cudaMemcpyAsync(D64_out_h, d64, 8 * (size_t)(n * n), cudaMemcpyDeviceToHost, s) == cudaSuccess ? VGS_OK : VGS_ERR_CUDA;
    const int rc_flag = vgs_read_error_flag(h, s);   // synchronises
    return rc ? rc : rc_flag;
}
