// vgs_common.cuh -- internal declarations shared by the libvgs.so translation units.
//
// Device data layout (all per handle, resident in HBM for the life of the model set):
//   key[total]     u32   (1 << 2*len) | code, dict order            (ctx_id = index - ctx_off[m])
//   prob32[total]  float4 transition rows cast to fp32               (Frobenius / Projection)
//   prob64[total*4] f64  transition rows                             (Estimate, sampler)
//   logp[total*4]  f64   host-libm ln p, -1000 where p == 0          (NLL)
//   hash[...]      uint2 {key, ctx_id} open addressing, load <= 0.5  (every exact-context probe)
//   nll blob arena: per model either
//       tier 1 (order <= 6): dense table  f64[4^(order+1)]  indexed by (history << 2 | symbol)
//       tier 2 / 3 (order > 6): [first-level map over D1-mers | extension trie blocks (len > D1) | f64 logp rows + NaN row]
//   sequences: 2-bit packed, 16 symbols / u32, first symbol in the top bits, 16-byte aligned starts
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/vgs.h"

#define VGS_MAX_ORDER 15
#define VGS_DENSE_MAX_ORDER 6   // tier 1 limit: 4^7 * 8 B = 128 KB of shared memory
#define VGS_L1_DEPTH 6          // tier 2 first-level map: minimum depth (and the depth of 32-bit maps)
#define VGS_L1_MAX_DEPTH 8      // 4^8 u16 entries = 128 KB
#define VGS_T2_SMEM_BUDGET (220 * 1024)
#define VGS_NO_CTX 0x7FFFFFFFu
#define VGS_UMAP_MAX_ORDER 7    // dense "universe" maps for the pair kernels: (4^8 - 1) / 3 = 21845 u16 entries
#define VGS_UMAP_NONE 0x7FFFu

struct ModelMeta {
    int64_t ctx_off;      // first context of this model in key/prob arrays
    int64_t hash_off;     // first slot of this model's hash
    int64_t nll_off;      // byte offset of the NLL blob in the arena (256-byte aligned)
    int32_t n_ctx;
    int32_t order;
    uint32_t hash_mask;   // capacity - 1
    int32_t hash_shift;   // 32 - log2(capacity)
    int32_t tier;         // 1 dense, 2 two-level (u16 map), 3 two-level (u32 map, >= 32767 contexts)
    int32_t nll_bytes;    // blob size, multiple of 16
    uint32_t deep_mask;   // tier 2 / 3: capacity of the extension trie in blocks
    int32_t deep_shift;
    int32_t deep_off;     // byte offset of the extension trie inside the blob
    int32_t logp_off;     // byte offset of the logp rows inside the blob
    // perfect hash (hash-and-displace) over the same contexts: exactly two dependent loads per exact-context
    // probe, no probing loop -> no warp divergence (pair kernels on deep sets)
    int64_t ph_off;       // byte offset in the ph arena: [u16 disp[2^ph_lgr] (16-byte padded) | uint2 table[capacity]]
    int32_t ph_lgr;       // log2(number of displacement buckets)
    int32_t ph_seed;      // which multiplier pair built it (-1: build failed, use the probing hash)
    int32_t ph_bytes;     // blob size (multiple of 16)
    int32_t l1_depth;     // tier 2: low byte = depth D1 of the first-level map, bit 8 = 32-bit entries
};

// device-resident work lists (scan units, contraction units) of a (plan, rank) are cached across calls so that a
// steady-state step issues no host->device list upload and no host-side list building
struct WorkCache {
    uint64_t key = 0;
    void *dev = nullptr;
    int64_t bytes = 0;
    int n[4] = {0, 0, 0, 0};   // entry counts (per tier / units, pairs)
    int split = 1, threads = 0;
    int64_t off2 = 0;          // byte offset of the second array inside dev
};

// where a score launch writes: M[d][s * stride_s + (m_col0 + m) * stride_m] for every destination d (this GPU's matrix
// and, on the peer-store path, the mapped matrices of the other ranks)
#define VGS_MAX_PEERS 8
struct VgsScoreDst {
    double *M[VGS_MAX_PEERS];
    int n_dst;
    int64_t stride_s, stride_m, m_col0;
};
inline VgsScoreDst vgs_dst_plain(double *M, int64_t ld) {
    VgsScoreDst d;
    for (int i = 0; i < VGS_MAX_PEERS; ++i) d.M[i] = nullptr;
    d.M[0] = M; d.n_dst = 1; d.stride_s = ld; d.stride_m = 1; d.m_col0 = 0;
    return d;
}
struct G8Unit;
struct G8Rect;
struct I8Work {   // device-resident unit / rectangle lists of one int8 contraction launch
    const G8Unit *d_units;
    int n_units;
    const G8Rect *d_rects;
    int n_rects;
    int split;
    int max_rect_models;   // models of the widest rectangle
};

struct vgs_context {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    int max_smem_optin = 0;
    int64_t launches = 0;

    // models
    int64_t n_models = 0, total_ctx = 0;
    std::vector<int64_t> h_ctx_off;
    std::vector<ModelMeta> h_meta;
    double *h_logp = nullptr;   // pinned staging of the host-computed ln p table
    int64_t h_logp_cap = 0;
    bool nll_ready = false;     // false when loaded with VGS_LOAD_SKIP_NLL
    bool prob_ready = false;    // false when loaded with VGS_LOAD_NLL_ONLY (no transition rows on the device)
    // derived tables are built on first use by the kernels that need them (a k-mer-mode NLL run needs none of these)
    int dense_state = 0;        // scan tables (tier 1 dense / tier 2-3 two-level): 0 = not built, 1 = built
    bool umap_tried = false;    // universe maps of the pair kernels
    void *pool = nullptr;       // host thread pool (VgsPool, vgs_core.cu)
    std::vector<WorkCache> wcache;
    uint64_t model_gen = 0;     // bumped by every vgs_load_models
    int max_order = 0, min_order = 0;
    int64_t max_nll_bytes_t1 = 0, max_nll_bytes_t2 = 0, max_nll_bytes_t3 = 0;
    int64_t max_pair_bytes = 0;  // hash + fp32 rows of the largest model
    ModelMeta *d_meta = nullptr;
    uint32_t *d_key = nullptr;
    uint8_t *d_len = nullptr;
    uint32_t *d_code = nullptr;
    float4 *d_prob32 = nullptr;
    double *d_prob64 = nullptr;
    double *d_logp = nullptr;
    double *d_logp_dev = nullptr;   // VGS_LOAD_DEVICE_LOG: the device's own ln p (read by the k-mer contraction only)
    bool logp_exact = true;         // d_logp holds libm's ln p (false after a VGS_LOAD_DEVICE_LOG load until a bit-exact mode asks)
    bool logp_device = false;       // the current set was loaded with VGS_LOAD_DEVICE_LOG
    double *d_invp = nullptr;   // Estimate: 1 / p, built on first use per model set
    uint64_t invp_gen = ~0ull;
    uint2 *d_hash = nullptr;
    int64_t hash_slots = 0;
    uint8_t *d_nll = nullptr;
    int64_t nll_bytes = 0;
    // pair kernels, shallow sets (max_order <= VGS_UMAP_MAX_ORDER): per model a dense map over EVERY string of
    // length <= max_order (index (4^len - 1) / 3 + code) -> u16: low 15 bits = ctx_id of the longest suffix that is
    // a context (VGS_UMAP_NONE if none), bit 15 = the string itself is a context
    uint8_t *d_ph = nullptr;     // perfect-hash arena
    int64_t ph_bytes = 0;
    int64_t max_ph_bytes = 0;    // largest per-model blob
    bool ph_ready = false;
    uint16_t *d_umap = nullptr;
    int64_t umap_stride = 0;     // entries per model (multiple of 8), 0 = not built
    int max_n_ctx = 0;

    // sequences
    int64_t n_seq = 0, total_symbols = 0, total_words = 0;
    std::vector<int64_t> h_seq_off, h_seq_len;
    uint32_t *d_seq = nullptr;
    int64_t *d_seq_off = nullptr, *d_seq_len = nullptr;

    // scratch
    double *d_M = nullptr;      // [n_seq][n_models] score matrix
    int64_t M_elems = 0;
    void *d_scratch = nullptr;  // generic scratch (pair partials, payloads, count tables)
    int64_t scratch_bytes = 0;
    void *d_scratch2 = nullptr;
    int64_t scratch2_bytes = 0;
    int *d_err = nullptr;       // device error flag (bit 0: no context, bit 1: bad symbol)
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // vgs_load_models: the bulk upload of the transition rows, beside the table builds
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    void *h_meta_pin = nullptr;           // pinned copy of the per-model metadata (goes between the two halves of the rows on the copy stream)
    int64_t h_meta_pin_bytes = 0;

    // capacity of every handle-owned device buffer (keyed by the address of the pointer member):
    // buffers are reused across vgs_load_* calls and only re-allocated when they must grow
    std::map<void *, int64_t> caps;
    int *d_counters = nullptr;  // dynamic-scheduler counters
    // k-mer contraction mode (vgs_kmer.cu)
    double *d_kmer_T = nullptr, *d_kmer_C = nullptr, *d_kmer_partial = nullptr;
    bool kmer_tables_ready = false, kmer_counts_ready = false, kmer_has_nan = false;
    // exact int8 tensor-core contraction (vgs_kmer_i8.cu)
    int8_t *d_i8_digits = nullptr;   // [n_models * i8_nd][K8] balanced base-256 digits of round(ln p * 2^i8_frac_bits)
    int i8_frac_bits = 43;           // fraction bits of the fixed-point ln p (vgs_set_kmer_fraction_bits, VGS_I8_FRAC_BITS)
    int i8_nd = 8;                   // digit rows per model of the operand as built (6, 7 or 8)
    uint8_t *d_i8_C = nullptr;       // [n_seq][K8] k-mer counts (low bytes)
    int32_t *d_i8_over = nullptr;    // [n_seq] counts of frequent k-mers | [n_seq][64] {k, count >> 8}
    int32_t *d_i8_P = nullptr;       // k-split launches: exact int32 digit sums [part][n_models * i8_nd][n_seq padded]
    uint64_t i8_tables_gen = ~0ull;
    bool i8_tables_ok = false;
    bool i8_guess_wide = false;      // device-log loads: digit range the previous set of this handle needed (the next build's guess)
    uint16_t *d_i8_ctx = nullptr;    // [n_models][4^D] context of every depth-D history, found ahead of ln p by the load
    uint64_t i8_ctx_gen = ~0ull;     // model set d_i8_ctx belongs to
    int last_nll_kernel = 0;         // VGS_NLL_KERNEL_* of the most recent score launch
    bool i8_defer_check = false;     // set by a host-buffer entry point around its pass: the frequent-k-mer check needs no wait of its own
    bool i8_check_pending = false;   // ... and its flag is on the way to h_i8_flag (vgs_kmer_i8_settle)
    int *h_i8_flag = nullptr;        // pinned
    int i8_counts_state = 0;         // 0 = not checked for this sequence set, 1 = no frequent k-mers, 3 = some, 2 = too many
    // tensor-core Frobenius / Projection (vgs_pairs_tc.cu): digit operands over the dense universe, raw digit-pair sums
    int8_t *d_tc_P = nullptr, *d_tc_Q = nullptr, *d_tc_I = nullptr;
    int32_t *d_tc_model = nullptr, *d_tc_G = nullptr, *d_tc_R = nullptr;
    uint64_t tc_gen = ~0ull;
    bool tc_ok = false;
    int last_pair_kernel = 0;        // 0 = CUDA-core pair kernels, 1 = tensor-core path
    int pair_path = -1;              // vgs_set_pair_path: -1 = by density, 0 = CUDA cores, 1 = tensor cores where they apply
    float *d_out32 = nullptr;   // staging for the *_h entry points
    double *d_out64 = nullptr;

    // optional per-kernel event timing (vgs_profile_*)
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof[4];
};

// RAII-less helpers: bracket a kernel launch with events when profiling is on
inline void vgs_prof_begin(vgs_context *h, int id, cudaStream_t s) {
    if (!h->profile) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, s);
    h->prof[id].push_back({a, b});
}
inline void vgs_prof_end(vgs_context *h, int id, cudaStream_t s) {
    if (!h->profile || h->prof[id].empty()) return;
    cudaEventRecord(h->prof[id].back().second, s);
}

// ---- error plumbing ---------------------------------------------------------------------------
void vgs_set_error(const char *fmt, ...);
int vgs_cuda_fail(cudaError_t e, const char *what, const char *file, int line);
#define VGS_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess) return vgs_cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)
#define VGS_CHECK_ARG(cond, msg)            \
    do {                                    \
        if (!(cond)) {                      \
            vgs_set_error("%s", msg);       \
            return VGS_ERR_INVALID;         \
        }                                   \
    } while (0)
#define VGS_LAUNCH_CHECK(h)                                                          \
    do {                                                                             \
        (h)->launches++;                                                             \
        cudaError_t _e = cudaGetLastError();                                         \
        if (_e != cudaSuccess) return vgs_cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

// ---- programmatic dependent launch (Hopper+): a kernel launched through vgs_launch_dep may begin (its CTAs scheduled, its
// prologue run) while the previous kernel on the stream is still finishing; it must call vgs_dep_wait() before it touches
// anything that kernel wrote.  Kernels call vgs_dep_trigger() early so that THEIR successor can be staged.  Both are no-ops
// in an ordinary launch.  cfg 3's step is four short kernels: the launch gaps were 9 of its 120 us.  VGS_PDL=0 disables.
#ifdef __CUDACC__
__device__ __forceinline__ void vgs_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void vgs_dep_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool vgs_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline void vgs_launch_dep(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = vgs_pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);   // errors surface in the caller's VGS_LAUNCH_CHECK
}

int vgs_ensure_scratch(vgs_context *h, int64_t bytes);
int vgs_ensure_scratch2(vgs_context *h, int64_t bytes);
int vgs_ensure_M(vgs_context *h);
int vgs_read_error_flag(vgs_context *h, cudaStream_t s);  // syncs s, returns a vgs_status
template <typename T>
int vgs_dev_alloc(T **p, int64_t n) {
    if (n <= 0) n = 1;
    cudaError_t e = cudaMalloc((void **)p, (size_t)n * sizeof(T));
    if (e != cudaSuccess) {
        vgs_set_error("cudaMalloc of %lld bytes failed: %s", (long long)(n * (int64_t)sizeof(T)), cudaGetErrorString(e));
        *p = nullptr;
        return VGS_ERR_MEMORY;
    }
    return VGS_OK;
}

// grow-only reservation of a handle-owned device buffer of n elements
template <typename T>
int vgs_reserve(vgs_context *h, T **p, int64_t n) {
    if (n <= 0) n = 1;
    const int64_t bytes = n * (int64_t)sizeof(T);
    auto it = h->caps.find((void *)p);
    if (*p && it != h->caps.end() && it->second >= bytes) return VGS_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    h->caps.erase((void *)p);
    cudaError_t e = cudaMalloc((void **)p, (size_t)bytes);
    if (e != cudaSuccess) {
        vgs_set_error("cudaMalloc of %lld bytes failed: %s", (long long)bytes, cudaGetErrorString(e));
        *p = nullptr;
        return VGS_ERR_MEMORY;
    }
    h->caps[(void *)p] = bytes;
    return VGS_OK;
}

// tile bookkeeping shared by host and device
struct TilePlan {
    int64_t n, tile, nt;   // matrix size, tile edge, tiles per side
    int world, symmetric;
    int64_t n_tiles, slots;  // total tiles, slots per rank
};
__host__ __device__ inline int64_t vgs_tri_index(int64_t I, int64_t J, int64_t nt) {  // I <= J, row-major upper triangle
    return I * nt - I * (I - 1) / 2 + (J - I);
}
void vgs_make_plan(TilePlan *p, int64_t n, int64_t tile, int world, int symmetric);
// tile k -> (I, J)
void vgs_tile_coords(const TilePlan &p, int64_t k, int64_t *I, int64_t *J);

// launchers implemented in the kernel translation units
int vgs_nll_score_lists(vgs_context *h, const std::vector<std::vector<int32_t>> &seq_lists_per_block, int64_t block,
                        int64_t m_begin, int64_t m_end, bool add_diagonal, int skip_mode, int mode, const VgsScoreDst &dst,
                        cudaStream_t s, uint64_t cache_key = 0);
WorkCache *vgs_wcache_find(vgs_context *h, uint64_t key);
WorkCache *vgs_wcache_add(vgs_context *h, uint64_t key, int64_t bytes);   // nullptr on allocation failure
void vgs_wcache_invalidate(vgs_context *h);   // keys dropped, allocations kept
void vgs_wcache_clear(vgs_context *h);
inline uint64_t vgs_mix(uint64_t a, uint64_t b) {
    a ^= b + 0x9E3779B97F4A7C15ull + (a << 6) + (a >> 2);
    return a;
}
int vgs_ensure_ph(vgs_context *h, cudaStream_t s);  // builds the perfect hashes on first use
bool vgs_kmer_applicable(vgs_context *h);
// returns VGS_OK, an error, or -1 when the set needs the scan kernel (a model without a root context)
int vgs_kmer_scores(vgs_context *h, const std::vector<char> &need, int64_t nb_s, int64_t nb_m, bool diag_separately,
                    const VgsScoreDst &dst, cudaStream_t s, uint64_t cache_key = 0);
int vgs_kmer_i8_prepare_tables(vgs_context *h, cudaStream_t s);
bool vgs_kmer_applicable_orders(vgs_context *h);                          // every order <= 6 (known after the host's planning pass)
int vgs_kmer_i8_digits_begin(vgs_context *h, bool wide, cudaStream_t s);   // choose the digit count, allocate + clear the operand of the current set
double vgs_kmer_i8_narrow_limit(vgs_context *h);                           // |ln p| below this fits the narrow digit count
int vgs_kmer_i8_digits_range(vgs_context *h, int64_t m_begin, int64_t m_end, cudaStream_t s);   // digits of models [m_begin, m_end)
bool vgs_kmer_i8_settle(vgs_context *h);   // after a synchronised pass with i8_defer_check: true = void, redo
int vgs_kmer_i8_contexts_ahead(vgs_context *h, cudaStream_t s);   // the digit builder's context lookups, before ln p exists (hashes only)
int vgs_kmer_i8_digits_end(vgs_context *h, cudaStream_t s);                // after a synchronisation: record whether the operand is usable
int vgs_kmer_i8_scores(vgs_context *h, const I8Work &w, bool diag_separately, const VgsScoreDst &dst, cudaStream_t s);
int vgs_ensure_exact_logp(vgs_context *h, cudaStream_t s);   // libm's ln p in h->d_logp (no-op unless VGS_LOAD_DEVICE_LOG)
int vgs_kmer_i8_gran(vgs_context *h);   // unit granule (B rows) of the operand as built
bool vgs_pair_tc_wanted(vgs_context *h, int kind);
int vgs_pair_tc_matrix(vgs_context *h, int kind, float *D32, double *D64, cudaStream_t s);   // -1: not applicable
int vgs_peer_barrier_impl(vgs_context *h, uint32_t *const *flags_h, int rank, int world, uint32_t epoch, cudaStream_t stream);
int vgs_ensure_scan_tables(vgs_context *h, cudaStream_t s);   // tier 1-3 NLL tables of the scan kernels
int vgs_ensure_umap(vgs_context *h, cudaStream_t s);          // universe maps (shallow sets) of the pair kernels
int vgs_assemble_impl(vgs_context *h, const TilePlan &p, const float *gathered32, const double *gathered64,
                      float *D32, double *D64, cudaStream_t s);

#ifdef __CUDACC__
// ---- device helpers ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t vgs_key(int len, uint32_t code) { return (1u << (2 * len)) | code; }
__host__ __device__ __forceinline__ uint32_t vgs_lowmask(int len) {  // 2*len low bits set, len <= 16
    return len >= 16 ? 0xFFFFFFFFu : ((1u << (2 * len)) - 1u);
}
__host__ __device__ __forceinline__ uint32_t vgs_hash_slot(uint32_t key, int shift) { return (key * 0x9E3779B1u) >> shift; }

// exact-context probe: ctx_id or -1
__device__ __forceinline__ int vgs_hash_find(const uint2 *__restrict__ slots, uint32_t mask, int shift, uint32_t key) {
    uint32_t s = vgs_hash_slot(key, shift);
    while (true) {
        uint2 e = slots[s];
        if (e.x == key) return (int)e.y;
        if (e.x == 0u) return -1;
        s = (s + 1) & mask;
    }
}

// longest suffix of (hist_code, hist_len) of length <= order that is a context (vlmc.pyx:131-141)
__device__ __forceinline__ int vgs_longest_suffix(const uint2 *__restrict__ slots, uint32_t mask, int shift, int order,
                                                  uint32_t hist_code, int hist_len) {
    int top = hist_len < order ? hist_len : order;
    for (int ln = top; ln >= 0; --ln) {
        int id = vgs_hash_find(slots, mask, shift, vgs_key(ln, hist_code & vgs_lowmask(ln)));
        if (id >= 0) return id;
    }
    return -1;
}

// ---- hash-and-displace perfect hash -------------------------------------------------------------------------
// multiplier pairs by seed (a switch, not an indexed local array: an array would be materialised in local memory and
// re-written on every call)
__host__ __device__ __forceinline__ uint32_t vgs_ph_a2(int seed) {
    switch (seed & 7) {
        case 0: return 0x85EBCA6Bu; case 1: return 0xC2B2AE35u; case 2: return 0x27D4EB2Fu; case 3: return 0x165667B1u;
        case 4: return 0xD3A2646Du; case 5: return 0xFD7046C5u; case 6: return 0xB55A4F09u; default: return 0x2545F491u;
    }
}
__host__ __device__ __forceinline__ uint32_t vgs_ph_a3(int seed) {
    switch (seed & 7) {
        case 0: return 0x9E3779B9u; case 1: return 0x7F4A7C15u; case 2: return 0x94D049BBu; case 3: return 0xBF58476Du;
        case 4: return 0x1B873593u; case 5: return 0xCC9E2D51u; case 6: return 0x38495AB5u; default: return 0x6C8E9CF5u;
    }
}
__device__ __forceinline__ uint32_t vgs_ph_slot(uint32_t key, uint32_t disp, int lgcap, int seed) {
    const uint32_t h2 = (key * vgs_ph_a2(seed)) >> (32 - lgcap);
    const uint32_t h3 = ((key * vgs_ph_a3(seed)) >> (32 - lgcap)) | 1u;
    return (h2 + disp * h3) & ((1u << lgcap) - 1u);
}
// blob = [u16 disp[1 << lgr] padded to 16 bytes | uint2 table[1 << lgcap]]; returns ctx_id or -1
__device__ __forceinline__ int vgs_ph_find(const uint8_t *__restrict__ blob, int lgr, int lgcap, int seed, uint32_t key) {
    const uint16_t *disp = (const uint16_t *)blob;
    const uint2 *table = (const uint2 *)(blob + (((2u << lgr) + 15u) & ~15u));
    const uint32_t d = disp[(key * 0x9E3779B1u) >> (32 - lgr)];
    const uint2 e = table[vgs_ph_slot(key, d, lgcap, seed)];
    return e.x == key ? (int)e.y : -1;
}

__device__ __forceinline__ double vgs_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

// ---- mbarrier + 1-D bulk async copy (TMA without a tensor map; SASS: UBLKCP) --------------------
__device__ __forceinline__ uint32_t vgs_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void vgs_mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vgs_smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void vgs_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(vgs_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void vgs_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(vgs_smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void vgs_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     vgs_smem_addr(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(vgs_smem_addr(bar))
                 : "memory");
}
// one thread issues the copies for `bytes` (multiple of 16) in <= 32 KB pieces; the caller has armed
// the barrier with vgs_mbar_expect_tx for the total of all pieces it is about to issue
__device__ __forceinline__ void vgs_copy_pieces(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    const uint32_t piece = 32768u;
    for (uint32_t o = 0; o < bytes; o += piece) {
        uint32_t n = bytes - o < piece ? bytes - o : piece;
        vgs_bulk_g2s((char *)smem_dst + o, (const char *)gmem_src + o, n, bar);
    }
}
__device__ __forceinline__ void vgs_stage(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    vgs_mbar_expect_tx(bar, bytes);  // one arrival (barrier count is 1) + the byte count
    vgs_copy_pieces(smem_dst, gmem_src, bytes, bar);
}
#endif  // __CUDACC__
