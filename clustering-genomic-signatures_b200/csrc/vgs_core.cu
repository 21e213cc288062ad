// vgs_core.cu -- handle lifetime, model-table construction (K1), longest-suffix lookup, sequences.
//
// Reference behaviour restated here (paths relative to the reference repository):
//   src/vlmc/vlmc.pyx:8-22,179-180   VLMC data model, order = max context length
//   src/vlmc/vlmc.pyx:131-154        get_context / get_all_contexts (longest / all suffix contexts)
//   src/vlmc/vlmc.pyx:94-99          ln p on the host (libm log == math.log), -1000 where p == 0
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <type_traits>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "vgs_common.cuh"

#if defined(__x86_64__)
#include <emmintrin.h>
#define VGS_STREAM_STORE_F64(ptr, v) _mm_stream_si64((long long *)(ptr), __builtin_bit_cast(long long, (double)(v)))
#define VGS_STREAM_FENCE() _mm_sfence()
#else
#define VGS_STREAM_STORE_F64(ptr, v) (*(ptr) = (v))
#define VGS_STREAM_FENCE() std::atomic_thread_fence(std::memory_order_release)
#endif

// ---------------------------------------------------------------------------------------------
// small persistent host thread pool (per handle): the host side of vgs_load_models (validation, ln p) runs on it
class VgsPool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::function<void(int)> fn;
    int gen = 0, pending = 0, n;
    bool stop = false;

  public:
    explicit VgsPool(int n_) : n(n_) {
        for (int i = 0; i < n; ++i)
            th.emplace_back([this, i]() {
                int seen = 0;
                for (;;) {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return stop || gen != seen; });
                    if (stop) return;
                    seen = gen;
                    std::function<void(int)> f = fn;
                    lk.unlock();
                    f(i);
                    lk.lock();
                    if (--pending == 0) cv_done.notify_all();
                }
            });
    }
    ~VgsPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto &t : th) t.join();
    }
    int size() const { return n; }
    void start(std::function<void(int)> f) {
        std::lock_guard<std::mutex> lk(mu);
        fn = std::move(f);
        pending = n;
        ++gen;
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};

static VgsPool *get_pool(vgs_context *h) {
    if (!h->pool) {
        // one process per GPU (torchrun) means several handles share the host's cores: split them, or N x 15 threads on 16
        // cores make every rank's ln p pass N times slower (e2e at 8 GPUs).  VGS_HOST_THREADS overrides.
        unsigned hc = std::thread::hardware_concurrency();
        unsigned share = 1;
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) share = (unsigned)std::max(1, atoi(e));
        // alone: all cores but one (this thread feeds the copy engine); shared: an equal slice of all cores (the feeding
        // thread sleeps between polls, see the upload loop)
        int n = (int)std::max(1u, std::min(share > 1 ? hc / share : (hc > 2 ? hc - 1 : 1u), 24u));
        if (const char *e = getenv("VGS_HOST_THREADS")) n = std::max(1, std::min(64, atoi(e)));
        h->pool = new VgsPool(n);
    }
    return (VgsPool *)h->pool;
}

bool vgs_pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("VGS_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
    return on != 0;
}

static thread_local std::string g_last_error;

void vgs_set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

int vgs_cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    vgs_set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return VGS_ERR_CUDA;
}

extern "C" const char *vgs_last_error(void) { return g_last_error.c_str(); }
extern "C" int vgs_version(void) { return 100; }

static int ensure_buf(void **p, int64_t *have, int64_t need) {
    if (need <= *have) return VGS_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    int64_t want = need + need / 8 + 256;
    cudaError_t e = cudaMalloc(p, (size_t)want);
    if (e != cudaSuccess) {
        vgs_set_error("cudaMalloc of %lld scratch bytes failed: %s", (long long)want, cudaGetErrorString(e));
        return VGS_ERR_MEMORY;
    }
    *have = want;
    return VGS_OK;
}
int vgs_ensure_scratch(vgs_context *h, int64_t bytes) { return ensure_buf(&h->d_scratch, &h->scratch_bytes, bytes); }
int vgs_ensure_scratch2(vgs_context *h, int64_t bytes) { return ensure_buf(&h->d_scratch2, &h->scratch2_bytes, bytes); }
int vgs_ensure_M(vgs_context *h) {
    int64_t need = h->n_seq * h->n_models;
    int rc = vgs_reserve(h, &h->d_M, need);
    if (rc) return rc;
    h->M_elems = need;
    return VGS_OK;
}

WorkCache *vgs_wcache_find(vgs_context *h, uint64_t key) {
    if (!key) return nullptr;
    for (auto &w : h->wcache)
        if (w.key == key) return &w;
    return nullptr;
}
WorkCache *vgs_wcache_add(vgs_context *h, uint64_t key, int64_t bytes) {
    // an invalidated entry (key 0) keeps its allocation: reuse it when it is large enough
    for (auto &w : h->wcache)
        if (w.key == 0 && w.bytes >= bytes) {
            void *dev = w.dev;
            const int64_t cap = w.bytes;
            w = WorkCache();
            w.key = key; w.dev = dev; w.bytes = cap;
            return &w;
        }
    if (h->wcache.size() >= 32) {  // oldest out (cudaFree waits for the device)
        cudaFree(h->wcache.front().dev);
        h->wcache.erase(h->wcache.begin());
    }
    WorkCache w;
    w.key = key;
    w.bytes = std::max<int64_t>(bytes, 16);
    if (cudaMalloc(&w.dev, (size_t)w.bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    h->wcache.push_back(w);
    return &h->wcache.back();
}
// a new model set invalidates the lists; the allocations stay for the next set (stream order keeps in-flight readers
// safe: the next writer is a copy enqueued behind them on the caller's stream)
void vgs_wcache_invalidate(vgs_context *h) {
    for (auto &w : h->wcache) w.key = 0;
}
void vgs_wcache_clear(vgs_context *h) {
    for (auto &w : h->wcache) cudaFree(w.dev);
    h->wcache.clear();
}

int vgs_read_error_flag(vgs_context *h, cudaStream_t s) {
    int flag = 0;
    VGS_CUDA(cudaMemcpyAsync(&flag, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    VGS_CUDA(cudaStreamSynchronize(s));
    if (flag) {
        VGS_CUDA(cudaMemsetAsync(h->d_err, 0, sizeof(int), s));
        VGS_CUDA(cudaStreamSynchronize(s));
    }
    if (flag & 2) {
        vgs_set_error("sequence holds a symbol outside A,C,G,T");
        return VGS_ERR_SYMBOL;
    }
    if (flag & 8) {
        vgs_set_error("Total of weights must be greater than zero");  // random.choices' ValueError (src/vlmc/vlmc.pyx:177)
        return VGS_ERR_INVALID;
    }
    if (flag & 1) {
        vgs_set_error("get_context vlmc.pyx");  // the reference's RuntimeError text (src/vlmc/vlmc.pyx:141)
        return VGS_ERR_NO_CONTEXT;
    }
    return VGS_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" int vgs_create(int device, vgs_handle *out) {
    VGS_CHECK_ARG(out != nullptr, "vgs_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        vgs_set_error("no CUDA device available (%s); libvgs has no CPU fallback",
                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return VGS_ERR_CUDA;
    }
    VGS_CHECK_ARG(device >= 0 && device < n, "vgs_create: bad device index");
    VGS_CUDA(cudaSetDevice(device));
    vgs_context *h = new vgs_context();
    h->device = device;
    cudaDeviceProp prop;
    VGS_CUDA(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    h->cc_major = prop.major;
    h->cc_minor = prop.minor;
    h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (prop.major < 10) {
        vgs_set_error("libvgs is built for sm_100a (Blackwell); device %d is sm_%d%d", device, prop.major, prop.minor);
        delete h;
        return VGS_ERR_UNSUPPORTED;
    }
    VGS_CUDA(cudaMalloc((void **)&h->d_err, sizeof(int)));
    VGS_CUDA(cudaMemset(h->d_err, 0, sizeof(int)));
    VGS_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    if (const char *e = getenv("VGS_I8_FRAC_BITS")) h->i8_frac_bits = std::max(35, std::min(51, atoi(e)));
    *out = h;
    return VGS_OK;
}

// the device buffers stay allocated (grow-only, vgs_reserve); only the logical contents are dropped
static void free_models(vgs_context *h) {
    h->n_models = 0; h->total_ctx = 0; h->kmer_tables_ready = false;
    h->i8_counts_state = 0;   // the k-mer length follows the deepest model: the frequent-k-mer check is redone
    h->model_gen++;
    vgs_wcache_invalidate(h);
}
static void free_sequences(vgs_context *h) { h->n_seq = 0; h->total_symbols = 0; h->total_words = 0; h->kmer_counts_ready = false; h->i8_counts_state = 0; }
static void free_all_buffers(vgs_context *h) {
    for (auto &kv : h->caps) {
        void **p = (void **)kv.first;
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
    h->caps.clear();
}

extern "C" int vgs_destroy(vgs_handle h) {
    if (!h) return VGS_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    free_all_buffers(h);
    vgs_wcache_clear(h);
    cudaFree(h->d_scratch);
    cudaFree(h->d_scratch2);
    cudaFree(h->d_err);
    if (h->h_logp) cudaFreeHost(h->h_logp);
    if (h->h_meta_pin) cudaFreeHost(h->h_meta_pin);
    if (h->h_i8_flag) cudaFreeHost(h->h_i8_flag);
    delete (VgsPool *)h->pool;
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (auto &e : h->ev_copy) if (e) cudaEventDestroy(e);
    delete h;
    return VGS_OK;
}

extern "C" int vgs_device_info(vgs_handle h, int *sm_count, int *cc_major, int *cc_minor, int64_t *free_bytes) {
    VGS_CHECK_ARG(h, "null handle");
    VGS_CUDA(cudaSetDevice(h->device));
    if (sm_count) *sm_count = h->sm_count;
    if (cc_major) *cc_major = h->cc_major;
    if (cc_minor) *cc_minor = h->cc_minor;
    if (free_bytes) {
        size_t f = 0, t = 0;
        VGS_CUDA(cudaMemGetInfo(&f, &t));
        *free_bytes = (int64_t)f;
    }
    return VGS_OK;
}

extern "C" int64_t vgs_launch_count(vgs_handle h) { return h ? h->launches : 0; }

extern "C" int vgs_profile_enable(vgs_handle h, int on) {
    VGS_CHECK_ARG(h, "null handle");
    h->profile = on != 0;
    return VGS_OK;
}

extern "C" int vgs_profile_read(vgs_handle h, int kernel_id, double *total_ms, int64_t *launches) {
    VGS_CHECK_ARG(h && kernel_id >= 0 && kernel_id < 4, "vgs_profile_read: bad kernel id");
    VGS_CUDA(cudaSetDevice(h->device));
    double sum = 0.0;
    int64_t n = 0;
    for (auto &pr : h->prof[kernel_id]) {
        VGS_CUDA(cudaEventSynchronize(pr.second));
        float ms = 0.f;
        VGS_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        sum += ms;
        ++n;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    h->prof[kernel_id].clear();
    if (total_ms) *total_ms = sum;
    if (launches) *launches = n;
    return VGS_OK;
}

extern "C" int vgs_poll_error(vgs_handle h, void *stream) {
    VGS_CHECK_ARG(h, "null handle");
    VGS_CUDA(cudaSetDevice(h->device));
    return vgs_read_error_flag(h, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// K1: device-side table construction
// ---------------------------------------------------------------------------------------------
__global__ void k_build_keys(int64_t total, const uint8_t *__restrict__ len, const uint32_t *__restrict__ code,
                             const double *__restrict__ prob, uint32_t *__restrict__ key, float4 *__restrict__ prob32) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    key[i] = vgs_key(len[i], code[i]);
    if (!prob) return;   // NLL-only load: no transition rows on the device
    const double *p = prob + 4 * i;
    prob32[i] = make_float4((float)p[0], (float)p[1], (float)p[2], (float)p[3]);  // round-to-nearest like numpy's float32 store
}

// one block per model: insert the model's contexts into its open-addressing hash
__global__ void k_build_hash(const ModelMeta *__restrict__ meta, const uint32_t *__restrict__ key, uint2 *__restrict__ hash) {
    const ModelMeta mm = meta[blockIdx.x];
    unsigned long long *slots = (unsigned long long *)(hash + mm.hash_off);
    for (int c = threadIdx.x; c < mm.n_ctx; c += blockDim.x) {
        uint32_t k = key[mm.ctx_off + c];
        uint32_t s = vgs_hash_slot(k, mm.hash_shift);
        // uint2 {x = key, y = id} viewed as u64: x is the low word
        unsigned long long want = ((unsigned long long)(uint32_t)c << 32) | k;
        while (true) {
            unsigned long long old = atomicCAS(&slots[s], 0ull, want);
            if (old == 0ull) break;
            if ((uint32_t)old == k) {  // duplicate key cannot come from a dict; keep the smaller id
                atomicMin(&slots[s], want);
                break;
            }
            s = (s + 1) & mm.hash_mask;
        }
    }
}

// tier 1: dense[(h << 2) | x] = logp[longest_suffix(h)][x] for every order-mer h; NaN if none
__global__ void k_build_dense(const ModelMeta *__restrict__ meta, const uint2 *__restrict__ hash,
                              const double *__restrict__ logp, uint8_t *__restrict__ arena) {
    const ModelMeta mm = meta[blockIdx.y];
    if (mm.tier != 1) return;
    const uint32_t n_hist = 1u << (2 * mm.order);
    double *dense = (double *)(arena + mm.nll_off);
    const uint2 *slots = hash + mm.hash_off;
    for (uint32_t hcode = blockIdx.x * blockDim.x + threadIdx.x; hcode < n_hist; hcode += gridDim.x * blockDim.x) {
        int id = vgs_longest_suffix(slots, mm.hash_mask, mm.hash_shift, mm.order, hcode, mm.order);
        double4 v;
        if (id >= 0) {
            const double *r = logp + 4 * (mm.ctx_off + id);
            v = make_double4(r[0], r[1], r[2], r[3]);
        } else {
            v = make_double4(vgs_nan(), vgs_nan(), vgs_nan(), vgs_nan());
        }
        double *o = dense + 4 * (size_t)hcode;
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
}

// tier 2 / 3: first-level map over D1-mers + extension trie + logp rows (layout: vgs_load_models_ex)
//   map entry   : ctx_id of the longest context that is a suffix of the D1-mer (n_ctx = the NaN row if none), or
//                 FLAG | block when some longer context ends in this D1-mer
//   trie entry  : {value = ctx_id of the longest real context that is a suffix of the node string, child block or NONE};
//                 entry x of block b belongs to the string "x + string(b)"; entries without a node of their own keep
//                 the value of the block's own string, so a scan needs no "missing child" test beyond child == NONE
template <bool WIDE>
__device__ void build_extension_trie(const ModelMeta &mm, void *l1v, void *extv, const uint8_t *__restrict__ len,
                                     const uint32_t *__restrict__ code, int *counter) {
    typedef typename std::conditional<WIDE, unsigned long long, uint32_t>::type Entry;
    typedef typename std::conditional<WIDE, uint32_t, uint16_t>::type MapT;
    const int HALF = WIDE ? 32 : 16;
    const Entry NONE = WIDE ? 0xFFFFFFFFull : 0xFFFFull;
    const uint32_t FLAG = WIDE ? 0x80000000u : 0x8000u;
    MapT *l1 = (MapT *)l1v;
    Entry *ext = (Entry *)extv;
    const int d1 = mm.l1_depth & 0xFF;
    const uint32_t n1 = 1u << (2 * d1);
    const Entry CLAIM = (Entry)FLAG;   // child field = CLAIM | context id while a node's block is being allocated
    auto mk = [&](Entry child, Entry value) { return (child << HALF) | value; };
    auto child_of = [&](Entry e) { return (uint32_t)(e >> HALF); };
    // (a) flag the D1-mers in which a longer context ends
    for (int c = threadIdx.x; c < mm.n_ctx; c += blockDim.x) {
        if (len[mm.ctx_off + c] <= d1) continue;
        const uint32_t li = code[mm.ctx_off + c] & (n1 - 1);
        if (WIDE) atomicOr((unsigned int *)l1 + li, FLAG);
        else atomicOr((unsigned int *)l1 + (li >> 1), (li & 1u) ? 0x80000000u : 0x00008000u);  // the u16 flag inside its word
    }
    __syncthreads();
    // (b) one block per flagged D1-mer, pre-filled with the map's own answer
    for (uint32_t li = threadIdx.x; li < n1; li += blockDim.x) {
        const uint32_t e = l1[li];
        if (!(e & FLAG)) continue;
        const uint32_t blk = (uint32_t)atomicAdd(counter, 1);
        const Entry fill = mk(NONE, (Entry)(e & (FLAG - 1u)));
        for (int x = 0; x < 4; ++x) ext[4 * (size_t)blk + x] = fill;
        l1[li] = (MapT)(FLAG | blk);
    }
    __syncthreads();
    // (c) level by level: the node of length l2 of every context at least that long
    for (int l2 = d1 + 1; l2 <= mm.order; ++l2) {
        // every thread revisits its contexts in the three sub-steps; `slot` is recomputed (cheap walk from the map)
        auto slot_of = [&](uint32_t cd) -> size_t {
            uint32_t blk = (uint32_t)l1[cd & (n1 - 1)] & (FLAG - 1u);
            for (int l = d1; l < l2 - 1; ++l) blk = child_of(ext[4 * (size_t)blk + ((cd >> (2 * l)) & 3u)]);
            return 4 * (size_t)blk + ((cd >> (2 * (l2 - 1))) & 3u);
        };
        // (i) a context that IS this node stamps its id
        for (int c = threadIdx.x; c < mm.n_ctx; c += blockDim.x) {
            if (len[mm.ctx_off + c] != l2) continue;
            const size_t sl = slot_of(code[mm.ctx_off + c]);
            if (WIDE) ((uint32_t *)ext)[2 * sl] = (uint32_t)c;       // low half = value (little endian)
            else ((uint16_t *)ext)[2 * sl] = (uint16_t)c;
        }
        __syncthreads();
        // (ii) contexts that continue below this node elect one of them (smallest id) ...
        for (int c = threadIdx.x; c < mm.n_ctx; c += blockDim.x) {
            if (len[mm.ctx_off + c] <= l2) continue;
            const size_t sl = slot_of(code[mm.ctx_off + c]);
            const Entry cur = ext[sl];
            const Entry value = cur & ((((Entry)1) << HALF) - 1);
            if (WIDE) atomicMin((unsigned long long *)ext + sl, (unsigned long long)mk(CLAIM | (Entry)c, value));
            else atomicMin((unsigned int *)ext + sl, (unsigned int)mk(CLAIM | (Entry)c, value));
        }
        __syncthreads();
        // (iii) ... which allocates the node's child block, pre-filled with the node's value
        for (int c = threadIdx.x; c < mm.n_ctx; c += blockDim.x) {
            if (len[mm.ctx_off + c] <= l2) continue;
            const size_t sl = slot_of(code[mm.ctx_off + c]);
            const Entry cur = ext[sl];
            if (child_of(cur) != (FLAG | (uint32_t)c)) continue;   // block numbers have the FLAG bit clear: no confusion
            const Entry value = cur & ((((Entry)1) << HALF) - 1);
            const uint32_t blk = (uint32_t)atomicAdd(counter, 1);
            const Entry fill = mk(NONE, value);
            for (int x = 0; x < 4; ++x) ext[4 * (size_t)blk + x] = fill;
            ext[sl] = mk((Entry)blk, value);
        }
        __syncthreads();
    }
}

__global__ void k_build_two_level(const ModelMeta *__restrict__ meta, const uint2 *__restrict__ hash,
                                  const uint32_t *__restrict__ key, const uint8_t *__restrict__ len,
                                  const uint32_t *__restrict__ code, const double *__restrict__ logp,
                                  uint8_t *__restrict__ arena) {
    const ModelMeta mm = meta[blockIdx.x];
    if (mm.tier < 2) return;
    __shared__ int n_blocks_used;
    if (threadIdx.x == 0) n_blocks_used = 0;
    uint32_t *l1 = (uint32_t *)(arena + mm.nll_off);
    void *ext = (void *)(arena + mm.nll_off + mm.deep_off);
    double *rows = (double *)(arena + mm.nll_off + mm.logp_off);
    const uint2 *slots = hash + mm.hash_off;
    const int d1 = mm.l1_depth & 0xFF;
    const bool wide = (mm.l1_depth & 0x100) != 0;
    const uint32_t n1 = 1u << (2 * d1);
    for (int i = threadIdx.x; i < 4 * mm.n_ctx; i += blockDim.x) rows[i] = logp[4 * mm.ctx_off + i];
    if (threadIdx.x < 4) rows[4 * mm.n_ctx + threadIdx.x] = vgs_nan();
    // "no context" points at the extra all-NaN row behind the model's rows: no special case in the scan
    if (wide) {
        for (uint32_t hc = threadIdx.x; hc < n1; hc += blockDim.x) {
            int id = vgs_longest_suffix(slots, mm.hash_mask, mm.hash_shift, d1 < mm.order ? d1 : mm.order, hc, d1);
            l1[hc] = id >= 0 ? (uint32_t)id : (uint32_t)mm.n_ctx;
        }
        __syncthreads();
        build_extension_trie<true>(mm, l1, ext, len, code, &n_blocks_used);
    } else {
        // u16 map, built in shared memory level by level without any hash probe: every context of length l <= D1
        // stamps its id on the 4^(D1-l) histories it is a suffix of; longer contexts overwrite shorter ones
        extern __shared__ uint16_t smap[];
        for (uint32_t hc = threadIdx.x; hc < n1; hc += blockDim.x) smap[hc] = (uint16_t)mm.n_ctx;
        __syncthreads();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        for (int l = 0; l <= d1; ++l) {
            const uint32_t reps = 1u << (2 * (d1 - l));
            if (reps >= 32u) {
                for (int c = warp; c < mm.n_ctx; c += nwarps) {
                    if (len[mm.ctx_off + c] != l) continue;
                    const uint32_t cd = code[mm.ctx_off + c];
                    for (uint32_t pfx = lane; pfx < reps; pfx += 32) smap[(pfx << (2 * l)) | cd] = (uint16_t)c;
                }
            } else {
                for (int c = threadIdx.x; c < mm.n_ctx; c += blockDim.x) {
                    if (len[mm.ctx_off + c] != l) continue;
                    const uint32_t cd = code[mm.ctx_off + c];
                    for (uint32_t pfx = 0; pfx < reps; ++pfx) smap[(pfx << (2 * l)) | cd] = (uint16_t)c;
                }
            }
            __syncthreads();
        }
        build_extension_trie<false>(mm, smap, ext, len, code, &n_blocks_used);
        uint16_t *l1n = (uint16_t *)l1;
        for (uint32_t hc = threadIdx.x; hc < n1; hc += blockDim.x) l1n[hc] = smap[hc];
    }
}

// dense universe map for the pair kernels (shallow sets)
__global__ void k_build_umap(const ModelMeta *__restrict__ meta, const uint2 *__restrict__ hash, int depth, int64_t stride,
                             uint16_t *__restrict__ umap) {
    const ModelMeta mm = meta[blockIdx.y];
    const uint2 *slots = hash + mm.hash_off;
    uint16_t *out = umap + (int64_t)blockIdx.y * stride;
    const int64_t total = ((((int64_t)1) << (2 * (depth + 1))) - 1) / 3;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < stride; u += (int64_t)gridDim.x * blockDim.x) {
        uint16_t v = VGS_UMAP_NONE;
        if (u < total) {
            int len = 0;
            int64_t base = 0;
            while (base + (((int64_t)1) << (2 * len)) <= u) { base += ((int64_t)1) << (2 * len); ++len; }
            const uint32_t code = (uint32_t)(u - base);
            const int top = len < mm.order ? len : mm.order;
            for (int ln = top; ln >= 0; --ln) {
                int id = vgs_hash_find(slots, mm.hash_mask, mm.hash_shift, vgs_key(ln, code & vgs_lowmask(ln)));
                if (id >= 0) { v = (uint16_t)(id | (ln == len ? 0x8000 : 0)); break; }
            }
        }
        out[u] = v;
    }
}

// hash-and-displace build, one thread per model (serial per model, all models in parallel).  Buckets of the first-level
// hash are placed largest first; a bucket's displacement is the first d for which all its keys fall into free,
// pairwise distinct slots.  scratch per model: cnt[r] | start[r + 1] | items[n] | order[r]  (ints)
__global__ void k_build_ph(int64_t n_models, ModelMeta *__restrict__ meta, const uint32_t *__restrict__ key,
                           uint8_t *__restrict__ arena, int *__restrict__ scratch, const int64_t *__restrict__ scratch_off) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_models) return;
    ModelMeta mm = meta[m];
    const int lgcap = 32 - mm.hash_shift, lgr = mm.ph_lgr;
    const int r = 1 << lgr, n = mm.n_ctx;
    const uint32_t cap_mask = (1u << lgcap) - 1u;
    uint8_t *blob = arena + mm.ph_off;
    uint16_t *disp = (uint16_t *)blob;
    unsigned long long *table = (unsigned long long *)(blob + (((2u << lgr) + 15u) & ~15u));
    int *cnt = scratch + scratch_off[m], *start = cnt + r, *items = start + r + 1, *order = items + n;
    const uint32_t *keys = key + mm.ctx_off;
    for (int b = 0; b < r; ++b) cnt[b] = 0;
    for (int c = 0; c < n; ++c) cnt[(keys[c] * 0x9E3779B1u) >> (32 - lgr)]++;
    start[0] = 0;
    int max_size = 0;
    for (int b = 0; b < r; ++b) { start[b + 1] = start[b] + cnt[b]; if (cnt[b] > max_size) max_size = cnt[b]; }
    for (int b = 0; b < r; ++b) cnt[b] = 0;
    for (int c = 0; c < n; ++c) {
        const int b = (int)((keys[c] * 0x9E3779B1u) >> (32 - lgr));
        items[start[b] + cnt[b]++] = c;
    }
    // buckets by decreasing size (selection by size classes; sizes are small)
    int no = 0;
    for (int sz = max_size; sz >= 1; --sz)
        for (int b = 0; b < r; ++b)
            if (cnt[b] == sz) order[no++] = b;
    int seed_ok = -1;
    for (int seed = 0; seed < 8 && seed_ok < 0; ++seed) {
        for (uint32_t s = 0; s <= cap_mask; ++s) table[s] = 0ull;
        for (int b = 0; b < r; ++b) disp[b] = 0;
        bool ok = true;
        for (int oi = 0; oi < no && ok; ++oi) {
            const int b = order[oi], sz = cnt[b];
            const int *it = items + start[b];
            bool placed = false;
            for (uint32_t d = 0; d < 65536u && !placed; ++d) {
                bool fits = true;
                for (int a = 0; a < sz && fits; ++a) {
                    const uint32_t sa = vgs_ph_slot(keys[it[a]], d, lgcap, seed);
                    if (table[sa] != 0ull) { fits = false; break; }
                    for (int a2 = 0; a2 < a; ++a2)
                        if (vgs_ph_slot(keys[it[a2]], d, lgcap, seed) == sa) { fits = false; break; }
                }
                if (fits) {
                    for (int a = 0; a < sz; ++a) {
                        const uint32_t sa = vgs_ph_slot(keys[it[a]], d, lgcap, seed);
                        table[sa] = ((unsigned long long)(uint32_t)it[a] << 32) | keys[it[a]];
                    }
                    disp[b] = (uint16_t)d;
                    placed = true;
                }
            }
            if (!placed) ok = false;
        }
        if (ok) seed_ok = seed;
    }
    meta[m].ph_seed = seed_ok;
}

static int log2_ceil_pow2(int64_t v, int64_t *cap_out) {
    int64_t cap = 8;
    int lg = 3;
    while (cap < v) { cap <<= 1; ++lg; }
    *cap_out = cap;
    return lg;
}

// VGS_LOAD_DEVICE_LOG: ln p from the transition rows on the device (src/vlmc/vlmc.pyx:94-99: -1000 where p == 0; a negative
// probability is the reference's "math domain error")
// flag[0] |= 1: a negative probability; flag[1] |= 1: some |ln p| >= narrow_limit (the contraction's digit count follows
// from it: what k_i8_range finds for sets whose ln p came from the host)
__global__ void k_device_logp(int64_t n, const double *__restrict__ prob, double *__restrict__ logp, int *__restrict__ flag,
                              double narrow_limit) {
    bool wide = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double p = prob[i];
        if (p < 0) atomicOr(flag, 1);
        const double l = (p == 0) ? -1000.0 : log(p);
        wide |= !(fabs(l) < narrow_limit);
        logp[i] = l;
    }
    if (__any_sync(0xffffffffu, wide) && (threadIdx.x & 31) == 0) atomicOr(flag + 1, 1);
}

// libm's ln p on the device (h->d_logp) for the bit-exact consumers of a set that was loaded with VGS_LOAD_DEVICE_LOG: the
// transition rows come back from the device, the host pool takes the logarithms, the result goes up.  Once per model set;
// the k-mer contraction keeps using the device's own ln p (h->d_logp_dev), so its results do not depend on whether a
// bit-exact mode ran in between.
int vgs_ensure_exact_logp(vgs_context *h, cudaStream_t s) {
    if (h->logp_exact) return VGS_OK;
    const int64_t n_elem = h->total_ctx * 4;
    int rc;
    if (h->h_logp_cap < n_elem) {
        if (h->h_logp) cudaFreeHost(h->h_logp);
        h->h_logp = nullptr;
        h->h_logp_cap = 0;
        VGS_CUDA(cudaHostAlloc((void **)&h->h_logp, sizeof(double) * (size_t)n_elem, cudaHostAllocDefault));
        h->h_logp_cap = n_elem;
    }
    if ((rc = vgs_reserve(h, &h->d_logp, n_elem))) return rc;
    VGS_CUDA(cudaStreamSynchronize(s));
    VGS_CUDA(cudaMemcpy(h->h_logp, h->d_prob64, sizeof(double) * (size_t)n_elem, cudaMemcpyDeviceToHost));
    VgsPool *pool = get_pool(h);
    const int nt = pool->size();
    double *v = h->h_logp;
    pool->start([&, nt](int tid) {
        const int64_t b0 = n_elem * tid / nt, e0 = n_elem * (tid + 1) / nt;
        for (int64_t i = b0; i < e0; ++i) v[i] = (v[i] == 0) ? -1000.0 : log(v[i]);
    });
    pool->wait();
    VGS_CUDA(cudaMemcpy(h->d_logp, h->h_logp, sizeof(double) * (size_t)n_elem, cudaMemcpyHostToDevice));
    h->logp_exact = true;
    return VGS_OK;
}

extern "C" int vgs_load_models(vgs_handle h, int64_t n_models, const int64_t *ctx_offsets_h, const uint8_t *ctx_len_h,
                               const uint32_t *ctx_code_h, const double *prob_h, const double *logp_h) {
    return vgs_load_models_ex(h, n_models, ctx_offsets_h, ctx_len_h, ctx_code_h, prob_h, logp_h, 0);
}

extern "C" int vgs_load_models_ex(vgs_handle h, int64_t n_models, const int64_t *ctx_offsets_h, const uint8_t *ctx_len_h,
                                  const uint32_t *ctx_code_h, const double *prob_h, const double *logp_h, int flags) {
    VGS_CHECK_ARG(h, "null handle");
    VGS_CHECK_ARG(n_models > 0 && ctx_offsets_h && ctx_len_h && ctx_code_h && prob_h, "vgs_load_models: null/empty input");
    VGS_CUDA(cudaSetDevice(h->device));
    static int load_timing = -1;
    if (load_timing < 0) { const char *e = getenv("VGS_LOAD_TIMING"); load_timing = (e && e[0] == '1') ? 1 : 0; }
    const auto t_begin = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (load_timing) fprintf(stderr, "[vgs_load_models] %-28s %8.3f ms\n", what,
                                 std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    free_models(h);
    // developer aid (VGS_LOAD_TIMING=1): device-side timestamps of the load's stream, relative to the first enqueued operation
    cudaEvent_t ev_t[4] = {nullptr, nullptr, nullptr, nullptr};
    if (load_timing) for (auto &e : ev_t) cudaEventCreate(&e);
    const int64_t total = ctx_offsets_h[n_models];
    VGS_CHECK_ARG(ctx_offsets_h[0] == 0 && total > 0, "vgs_load_models: ctx_offsets must start at 0");
    VGS_CHECK_ARG(total < (int64_t)1 << 31, "vgs_load_models: more than 2^31 contexts in one set");
    const bool want_nll = !(flags & VGS_LOAD_SKIP_NLL);
    const bool want_prob = !(flags & VGS_LOAD_NLL_ONLY);
    VGS_CHECK_ARG(want_nll || want_prob, "vgs_load_models: VGS_LOAD_SKIP_NLL and VGS_LOAD_NLL_ONLY exclude each other");
    // VGS_LOAD_DEVICE_LOG: ln p is taken on the device (CUDA's log(), <= 1 ulp from libm's) for the k-mer contraction -- its
    // fixed-point quantisation is 64 times coarser than that -- so the load is one pass over PCIe with no host arithmetic.
    // The bit-exact modes still get libm's ln p: vgs_ensure_exact_logp computes it on first need.
    bool device_log = (flags & VGS_LOAD_DEVICE_LOG) && want_nll && !logp_h;
    int rc;
    cudaStream_t s = h->own_stream;
    // the caller's arrays start crossing PCIe before anything else happens; every return below drains the stream first
    // A device-log load that keeps no fp32 rows reads the transition rows for ONE thing, ln p: their upload (7/8 of the
    // bytes) goes to a second stream, and everything that needs only the context lengths / codes -- keys, hashes, the
    // context lookups of the digit builder -- runs beside it instead of behind it.
    const bool split_copy = device_log && !want_prob;
    const int64_t prob_first = (total * 4) / 2;   // elements of the transition rows uploaded ahead of the metadata
    if (split_copy && !h->copy_stream) {
        VGS_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        VGS_CUDA(cudaEventCreateWithFlags(&h->ev_copy[0], cudaEventDisableTiming));
        VGS_CUDA(cudaEventCreateWithFlags(&h->ev_copy[1], cudaEventDisableTiming));
    }
    struct Drain { cudaStream_t st, st2; ~Drain() { if (st2) cudaStreamSynchronize(st2); cudaStreamSynchronize(st); } } drain{s, split_copy ? h->copy_stream : nullptr};
    if ((rc = vgs_reserve(h, &h->d_len, total))) return rc;
    if ((rc = vgs_reserve(h, &h->d_code, total))) return rc;
    if ((want_prob || device_log) && (rc = vgs_reserve(h, &h->d_prob64, total * 4))) return rc;
    if (load_timing) cudaEventRecord(ev_t[0], s);
    VGS_CUDA(cudaMemcpyAsync(h->d_len, ctx_len_h, (size_t)total, cudaMemcpyHostToDevice, s));
    VGS_CUDA(cudaMemcpyAsync(h->d_code, ctx_code_h, sizeof(uint32_t) * (size_t)total, cudaMemcpyHostToDevice, s));
    if (split_copy) {
        // (after the two small arrays on the bus: the builders wait for those)
        VGS_CUDA(cudaEventRecord(h->ev_copy[0], s));
        VGS_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_copy[0], 0));
        // first half now; the second follows the per-model metadata, which the host is about to plan (a copy engine serves
        // its queue in order: enqueued behind ALL the rows the metadata would arrive with the last of them)
        VGS_CUDA(cudaMemcpyAsync(h->d_prob64, prob_h, sizeof(double) * (size_t)prob_first, cudaMemcpyHostToDevice, h->copy_stream));
    } else if (want_prob || device_log) {
        VGS_CUDA(cudaMemcpyAsync(h->d_prob64, prob_h, sizeof(double) * (size_t)total * 4, cudaMemcpyHostToDevice, s));
    }

    h->h_ctx_off.assign(ctx_offsets_h, ctx_offsets_h + n_models + 1);
    h->h_meta.assign((size_t)n_models, ModelMeta());
    h->max_order = 0;
    h->min_order = VGS_MAX_ORDER + 1;
    h->max_nll_bytes_t1 = h->max_nll_bytes_t2 = h->max_nll_bytes_t3 = h->max_pair_bytes = 0;
    h->max_n_ctx = 0;
    h->umap_stride = 0;
    int64_t hash_slots = 0, nll_bytes = 0, ph_bytes = 0;
    h->max_ph_bytes = 0;
    h->ph_ready = false;
    // per-model scan of the context lengths (order, length histogram, validity) on the host pool
    VgsPool *pool = get_pool(h);
    std::vector<int32_t> len_hists((size_t)n_models * 16, 0);
    std::vector<int64_t> bad_model((size_t)pool->size(), -1);
    std::vector<int> bad_kind((size_t)pool->size(), 0);
    {
        const int nt = pool->size();
        pool->start([&, nt](int tid) {
            for (int64_t m = tid; m < n_models; m += nt) {
                const int64_t a = ctx_offsets_h[m], e = ctx_offsets_h[m + 1];
                int32_t *hist = &len_hists[(size_t)m * 16];
                for (int64_t c = a; c < e; ++c) {
                    const int ln = ctx_len_h[c];
                    if (ln > VGS_MAX_ORDER) { if (bad_model[(size_t)tid] < 0) { bad_model[(size_t)tid] = m; bad_kind[(size_t)tid] = 1; } continue; }
                    if ((ctx_code_h[c] >> (2 * ln)) != 0 && bad_model[(size_t)tid] < 0) { bad_model[(size_t)tid] = m; bad_kind[(size_t)tid] = 2; }
                    hist[ln]++;
                }
            }
        });
        pool->wait();
        for (int t = 0; t < nt; ++t)
            if (bad_model[(size_t)t] >= 0) {
                if (bad_kind[(size_t)t] == 1) {
                    vgs_set_error("context longer than %d symbols is not supported (model %lld)", VGS_MAX_ORDER, (long long)bad_model[(size_t)t]);
                    return VGS_ERR_UNSUPPORTED;
                }
                vgs_set_error("context code has bits above its length (model %lld)", (long long)bad_model[(size_t)t]);
                return VGS_ERR_INVALID;
            }
    }
    for (int64_t m = 0; m < n_models; ++m) {
        ModelMeta &mm = h->h_meta[(size_t)m];
        int64_t a = ctx_offsets_h[m], n = ctx_offsets_h[m + 1] - a;
        VGS_CHECK_ARG(n > 0, "vgs_load_models: a VLMC needs at least one context");
        VGS_CHECK_ARG(n < (int64_t)1 << 30, "vgs_load_models: too many contexts in one model");
        int order = 0;
        int64_t len_hist[VGS_MAX_ORDER + 1] = {0};
        for (int ln = 0; ln <= VGS_MAX_ORDER; ++ln) {
            len_hist[ln] = len_hists[(size_t)m * 16 + ln];
            if (len_hist[ln] > 0) order = ln;
        }
        mm.ctx_off = a;
        mm.n_ctx = (int32_t)n;
        mm.order = order;
        int64_t cap;
        int lg = log2_ceil_pow2(2 * n, &cap);
        mm.hash_off = hash_slots;
        mm.hash_mask = (uint32_t)(cap - 1);
        mm.hash_shift = 32 - lg;
        hash_slots += cap;
        mm.ph_lgr = std::max(1, lg - 2);
        mm.ph_seed = -1;
        mm.ph_bytes = (int32_t)((((int64_t)2 << mm.ph_lgr) + 15) / 16 * 16 + cap * 8);
        mm.ph_off = ph_bytes;
        mm.l1_depth = 0;
        ph_bytes += ((int64_t)mm.ph_bytes + 255) / 256 * 256;
        h->max_ph_bytes = std::max<int64_t>(h->max_ph_bytes, mm.ph_bytes);
        h->max_pair_bytes = std::max(h->max_pair_bytes, cap * 8 + n * 16);
        h->max_n_ctx = std::max<int>(h->max_n_ctx, (int)n);
        h->max_order = std::max(h->max_order, order);
        h->min_order = std::min(h->min_order, order);
        mm.nll_off = nll_bytes;
        if (order <= VGS_DENSE_MAX_ORDER) {
            mm.tier = 1;
            mm.nll_bytes = (int32_t)(((int64_t)1 << (2 * (order + 1))) * 8);
            if (mm.nll_bytes < 16) mm.nll_bytes = 32;  // order 0: 4 doubles
            mm.deep_mask = 0; mm.deep_shift = 0; mm.deep_off = 0; mm.logp_off = 0;
            h->max_nll_bytes_t1 = std::max<int64_t>(h->max_nll_bytes_t1, mm.nll_bytes);
        } else {
            // first-level map over the last D1 symbols: as deep as shared memory allows (D1 = order means the map
            // alone answers every history: two gathers per base).  Histories whose D1-mer is the suffix of a longer
            // context continue in an extension trie: blocks of four entries {value, child block}, one per further
            // history symbol; a missing child carries its parent's value, so the walk is one load per level and ends
            // at "no child".  u16 map entries / u32 trie entries (tier 2) unless the model has >= 32767 contexts or
            // may need >= 32767 blocks (tier 3: u32 / u64).
            int64_t l1_bytes = 0, n_blocks = 0;
            bool wide = n >= 0x7FFF;
            int d1 = 0;
            for (int attempt = 0; attempt < 2; ++attempt) {
                d1 = wide ? VGS_L1_DEPTH : std::min(order, VGS_L1_MAX_DEPTH);
                for (;; --d1) {
                    // upper bound on the blocks: one per context longer than D1 (its D1-mer) plus one per proper
                    // suffix of length > D1 of such a context
                    n_blocks = 0;
                    for (int ln = d1 + 1; ln <= order; ++ln) n_blocks += len_hist[ln] * (ln - d1);
                    l1_bytes = (((int64_t)1 << (2 * d1)) * (wide ? 4 : 2) + 15) / 16 * 16;
                    if (d1 <= VGS_L1_DEPTH || l1_bytes + n_blocks * (wide ? 32 : 16) + (n + 1) * 32 <= VGS_T2_SMEM_BUDGET) break;
                }
                if (wide || n_blocks < 0x7FFF) break;
                wide = true;  // too many blocks for 15-bit block numbers
            }
            mm.tier = wide ? 3 : 2;
            mm.l1_depth = d1 | (wide ? 0x100 : 0);
            mm.deep_off = (int32_t)l1_bytes;
            mm.deep_mask = (uint32_t)n_blocks;   // capacity of the trie in blocks
            mm.deep_shift = 0;
            mm.logp_off = (int32_t)(l1_bytes + std::max<int64_t>(n_blocks, 1) * (wide ? 32 : 16));
            int64_t bytes = mm.logp_off + (n + 1) * 32;  // + one all-NaN row for histories without any context
            if (bytes >= ((int64_t)1 << 31) || n_blocks >= ((int64_t)1 << 31)) {
                vgs_set_error("model %lld is too large for the two-level NLL table", (long long)m);
                return VGS_ERR_UNSUPPORTED;
            }
            mm.nll_bytes = (int32_t)bytes;
            if (mm.tier == 2) h->max_nll_bytes_t2 = std::max<int64_t>(h->max_nll_bytes_t2, bytes);
            else h->max_nll_bytes_t3 = std::max<int64_t>(h->max_nll_bytes_t3, bytes);
        }
        nll_bytes += ((int64_t)mm.nll_bytes + 255) / 256 * 256;
    }

    lap("host: validate + plan");
    // sets the contraction does not cover (deeper than 6) are scored by the scan, which wants libm's ln p: the host pass below
    if (device_log && !vgs_kmer_applicable_orders(h)) device_log = false;
    if ((rc = vgs_reserve(h, &h->d_meta, n_models))) return rc;
    if ((rc = vgs_reserve(h, &h->d_key, total))) return rc;
    if (want_prob && (rc = vgs_reserve(h, &h->d_prob32, total))) return rc;
    h->prob_ready = want_prob;
    if ((rc = vgs_reserve(h, &h->d_hash, hash_slots))) return rc;
    h->hash_slots = hash_slots;
    h->ph_bytes = ph_bytes;
    h->nll_bytes = nll_bytes;
    h->n_models = n_models;
    h->total_ctx = total;
    h->nll_ready = false;

    // 1. behind the uploads of the caller's arrays (enqueued at the top; asynchronous when its buffers are pinned): everything
    //    that does not need ln p -- keys, fp32 rows, hashes
    if (split_copy) {
        // copy stream: [first half of the rows] -> metadata -> [second half]; the builders on `s` wait for the metadata only
        const int64_t mb = (int64_t)sizeof(ModelMeta) * n_models;
        if (h->h_meta_pin_bytes < mb) {
            if (h->h_meta_pin) cudaFreeHost(h->h_meta_pin);
    if (h->h_i8_flag) cudaFreeHost(h->h_i8_flag);
            h->h_meta_pin = nullptr; h->h_meta_pin_bytes = 0;
            VGS_CUDA(cudaHostAlloc(&h->h_meta_pin, (size_t)mb, cudaHostAllocDefault));
            h->h_meta_pin_bytes = mb;
        }
        memcpy(h->h_meta_pin, h->h_meta.data(), (size_t)mb);
        VGS_CUDA(cudaMemcpyAsync(h->d_meta, h->h_meta_pin, (size_t)mb, cudaMemcpyHostToDevice, h->copy_stream));
        VGS_CUDA(cudaEventRecord(h->ev_copy[0], h->copy_stream));
        VGS_CUDA(cudaMemcpyAsync(h->d_prob64 + prob_first, prob_h + prob_first, sizeof(double) * (size_t)(total * 4 - prob_first), cudaMemcpyHostToDevice,
                                 h->copy_stream));
        VGS_CUDA(cudaEventRecord(h->ev_copy[1], h->copy_stream));
        VGS_CUDA(cudaStreamWaitEvent(s, h->ev_copy[0], 0));
    } else {
        VGS_CUDA(cudaMemcpyAsync(h->d_meta, h->h_meta.data(), sizeof(ModelMeta) * (size_t)n_models, cudaMemcpyHostToDevice, s));
    }
    VGS_CUDA(cudaMemsetAsync(h->d_hash, 0, sizeof(uint2) * (size_t)hash_slots, s));
    k_build_keys<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(total, h->d_len, h->d_code, want_prob ? h->d_prob64 : nullptr, h->d_key,
                                                              want_prob ? h->d_prob32 : nullptr);
    VGS_LAUNCH_CHECK(h);
    k_build_hash<<<(unsigned)n_models, 256, 0, s>>>(h->d_meta, h->d_key, h->d_hash);
    VGS_LAUNCH_CHECK(h);

    if (load_timing) cudaEventRecord(ev_t[1], s);
    lap("enqueued: uploads + keys/hash");
    h->dense_state = 0;
    h->umap_tried = false;
    h->logp_exact = !device_log;
    h->logp_device = device_log;
    if (!want_nll) {
        // negative probabilities pass here as they do in the reference, which only fails on them when it takes their
        // logarithm at scoring time (src/vlmc/vlmc.pyx:99): the NLL load below reports that "math domain error"
        VGS_CUDA(cudaStreamSynchronize(s));
        return VGS_OK;
    }

    if (device_log) {
        // 2'. ln p on the device, then (sets the contraction covers) its digit operand; nothing for the host to do
        if ((rc = vgs_reserve(h, &h->d_logp_dev, total * 4))) return rc;
        if ((rc = vgs_reserve(h, &h->d_nll, nll_bytes))) return rc;
        if ((rc = vgs_reserve(h, &h->d_counters, 16))) return rc;
        VGS_CUDA(cudaMemsetAsync(h->d_counters + 12, 0, 2 * sizeof(int), s));
        const bool eager = !(flags & VGS_LOAD_LAZY_TABLES) && vgs_kmer_applicable_orders(h);
        // still beside the upload of the rows: the digit builder's context lookups (hashes only)
        if (eager && split_copy && (rc = vgs_kmer_i8_contexts_ahead(h, s))) return rc;
        if (load_timing) cudaEventRecord(ev_t[1], s);
        if (split_copy) VGS_CUDA(cudaStreamWaitEvent(s, h->ev_copy[1], 0));
        k_device_logp<<<std::min<int64_t>(8 * h->sm_count, (total * 4 + 255) / 256), 256, 0, s>>>(total * 4, h->d_prob64, h->d_logp_dev, h->d_counters + 12,
                                                                                                  vgs_kmer_i8_narrow_limit(h));
        VGS_LAUNCH_CHECK(h);
        if (load_timing) cudaEventRecord(ev_t[2], s);
        // The digit count follows from the range of ln p, which only the device knows at this point.  No round trip to the
        // host for it: the operand is built at once for the count the handle's previous set needed (a first load: the
        // narrow one), and the flags come back with the builder's own -- a wrong guess costs one rebuild.
        h->nll_ready = true;
        if (eager) {
            if ((rc = vgs_kmer_i8_digits_begin(h, h->i8_guess_wide, s))) return rc;
            if ((rc = vgs_kmer_i8_digits_range(h, 0, h->n_models, s))) return rc;
        }
        int fl[2] = {0, 0};   // negative probability | a value outside the narrow digit range
        VGS_CUDA(cudaMemcpyAsync(fl, h->d_counters + 12, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
        if (eager) {
            if ((rc = vgs_kmer_i8_digits_end(h, s))) return rc;   // synchronises; rebuilds wide if a value did not fit
        } else {
            VGS_CUDA(cudaStreamSynchronize(s));
        }
        const int neg = fl[0];
        if (eager && !neg && h->i8_guess_wide && !fl[1]) {
            // guessed wide, narrow would have done (the GEMM's work follows the digit count): once more
            if ((rc = vgs_kmer_i8_digits_begin(h, false, s))) return rc;
            if ((rc = vgs_kmer_i8_digits_range(h, 0, h->n_models, s))) return rc;
            if ((rc = vgs_kmer_i8_digits_end(h, s))) return rc;
        }
        if (eager && !neg) h->i8_guess_wide = fl[1] != 0;
        lap("device ln p + digits: stream drained");
        if (load_timing) {
            cudaEventRecord(ev_t[3], s);
            cudaEventSynchronize(ev_t[3]);
            float a = 0.f, b = 0.f, c = 0.f;
            cudaEventElapsedTime(&a, ev_t[0], ev_t[1]);
            cudaEventElapsedTime(&b, ev_t[0], ev_t[2]);
            cudaEventElapsedTime(&c, ev_t[0], ev_t[3]);
            fprintf(stderr, "[vgs_load_models] device: keys / hashes / contexts done %.3f ms, rows uploaded + ln p done %.3f, digits done %.3f ms after the first copy began\n", a, b, c);
            for (auto &e : ev_t) cudaEventDestroy(e);
        }
        if (neg) {
            vgs_set_error("math domain error: negative transition probability");
            return VGS_ERR_INVALID;
        }
        return VGS_OK;
    }
    // 2. meanwhile on the host: ln p with libm (bit-identical to CPython's math.log on this host), threaded,
    //    into a pinned staging buffer
    if (h->h_logp_cap < total * 4) {
        if (h->h_logp) cudaFreeHost(h->h_logp);
        h->h_logp = nullptr;
        h->h_logp_cap = 0;
        VGS_CUDA(cudaHostAlloc((void **)&h->h_logp, sizeof(double) * (size_t)total * 4, cudaHostAllocDefault));
        h->h_logp_cap = total * 4;
    }
    if (logp_h) memcpy(h->h_logp, logp_h, sizeof(double) * (size_t)total * 4);
    if ((rc = vgs_reserve(h, &h->d_logp, total * 4))) return rc;
    if ((rc = vgs_reserve(h, &h->d_nll, nll_bytes))) return rc;
    // sets the k-mer contraction covers (every order <= 6): its digit operand is built here, model block by model block
    // as soon as a block's ln p has been uploaded, so the build hides under the host's ln p pass instead of following it
    const bool eager_digits = !(flags & VGS_LOAD_LAZY_TABLES) && vgs_kmer_applicable_orders(h);
    int64_t digits_built = 0;
    // The operand's digit count (= the GEMM's work) follows from the range of ln p, and the operand is laid out for it.  The
    // workers of the ln p pass note a value outside the narrow range (p == 0 or p < e^-16); the count is fixed when the first
    // block is built, from what has been seen by then -- a set with such values shows them early -- and a late surprise
    // costs one rebuild (vgs_kmer_i8_digits_end).
    std::atomic<int> wide_seen(0);
    bool digits_begun = false;
    const double narrow_limit = vgs_kmer_i8_narrow_limit(h);
    std::vector<cudaEvent_t> ev_batch;
    std::vector<double> ev_host;
    auto build_digits_upto = [&](int64_t elems_uploaded, bool final) -> int {
        if (load_timing) {
            cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); ev_batch.push_back(e);
            ev_host.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
        }
        if (!eager_digits) return VGS_OK;
        int64_t m_done = digits_built;
        while (m_done < n_models && ctx_offsets_h[m_done + 1] * 4 <= elems_uploaded) ++m_done;
        if (m_done - digits_built >= 64 || (final && m_done > digits_built)) {
            if (!digits_begun) {
                const int rcb = vgs_kmer_i8_digits_begin(h, wide_seen.load() != 0, s);
                if (rcb) return rcb;
                digits_begun = true;
            }
            const int rcd = vgs_kmer_i8_digits_range(h, digits_built, m_done, s);
            if (rcd) return rcd;
            digits_built = m_done;
        }
        if (load_timing) {
            cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); ev_batch.push_back(e);
            ev_host.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
        }
        return VGS_OK;
    };
    if (logp_h) {
        for (int64_t i = 0; i < total * 4 && !wide_seen.load(std::memory_order_relaxed); ++i)
            if (!(fabs(logp_h[i]) < narrow_limit)) wide_seen.store(1);
        VGS_CUDA(cudaMemcpyAsync(h->d_logp, h->h_logp, sizeof(double) * (size_t)total * 4, cudaMemcpyHostToDevice, s));
        if ((rc = build_digits_upto(total * 4, true))) return rc;
    } else {
        // chunks are claimed with an atomic counter; this thread uploads each chunk as soon as it is complete, so the
        // PCIe transfer of ln p overlaps its computation
        const int64_t n_elem = total * 4;
        // small chunks: 31 chunks of 64 K logarithms on 15 threads left ONE chunk for a third round (0.7 ms of a 2.4 ms
        // load spent with one thread working); 8 K-element chunks balance to within a few microseconds
        const int64_t chunk = 1 << 13;
        const int64_t n_chunks = (n_elem + chunk - 1) / chunk;
        std::atomic<int64_t> next(0), bad(-1);
        // The logarithms go to the pinned staging buffer with streaming (non-temporal) stores: written through the caches,
        // the last 2-3 MB were still dirty in the cores' caches when the copy engine came for them and that copy ran at
        // 8-10 GB/s instead of 37 (the load's stream drained 0.45 ms after the host had finished; now 0.08).
        // VGS_LOAD_NT=0 restores plain stores.
        static int nt_stores = -1;
        if (nt_stores < 0) { const char *e = getenv("VGS_LOAD_NT"); nt_stores = (e && e[0] == '0') ? 0 : 1; }
        const int batch_chunks = 32;
        std::vector<std::atomic<int>> done((size_t)n_chunks);
        for (auto &d : done) d.store(0);
        double *out = h->h_logp;
        pool->start([&](int) {
            for (;;) {
                const int64_t c = next.fetch_add(1);
                if (c >= n_chunks) break;
                const int64_t b0 = c * chunk, e0 = std::min(n_elem, b0 + chunk);
                bool wide = false;
                for (int64_t i = b0; i < e0; ++i) {
                    const double p = prob_h[i];
                    if (p < 0) { int64_t expect = -1; bad.compare_exchange_strong(expect, i); }
                    const double lg = (p == 0) ? -1000.0 : log(p);
                    wide |= !(fabs(lg) < narrow_limit);
                    if (nt_stores) VGS_STREAM_STORE_F64(out + i, lg);
                    else out[i] = lg;
                }
                if (nt_stores) VGS_STREAM_FENCE();
                if (wide) wide_seen.store(1, std::memory_order_relaxed);
                done[(size_t)c].store(1, std::memory_order_release);
            }
        });
        cudaError_t ce = cudaSuccess;
        int64_t sent = 0;  // chunks [0, sent) are uploaded; batches of completed consecutive chunks go out together
        while (sent < n_chunks && ce == cudaSuccess) {
            int64_t upto = sent;
            while (upto < n_chunks && done[(size_t)upto].load(std::memory_order_acquire)) ++upto;
            if (upto - sent >= batch_chunks || (upto == n_chunks && upto > sent)) {   // >= 2 MB per copy
                const int64_t b0 = sent * chunk, e0 = std::min(n_elem, upto * chunk);
                ce = cudaMemcpyAsync(h->d_logp + b0, out + b0, sizeof(double) * (size_t)(e0 - b0), cudaMemcpyHostToDevice, s);
                sent = upto;
                if (ce == cudaSuccess && build_digits_upto(e0, upto == n_chunks) != VGS_OK) ce = cudaErrorUnknown;
            } else {
                if (getenv("LOCAL_WORLD_SIZE")) std::this_thread::sleep_for(std::chrono::microseconds(5));   // other ranks' workers may need this core
                else std::this_thread::yield();
            }
        }
        pool->wait();
        if (ce != cudaSuccess) return vgs_cuda_fail(ce, "ln p upload", __FILE__, __LINE__);
        if (bad.load() >= 0) {
            cudaStreamSynchronize(s);
            vgs_set_error("math domain error: negative transition probability (entry %lld)", (long long)bad.load());
            return VGS_ERR_INVALID;
        }
    }

    if (load_timing) cudaEventRecord(ev_t[2], s);
    lap("host: ln p computed + streamed");
    // 3. the NLL tables of the scan kernels, the digit operand of the k-mer contraction and the universe maps of the pair
    //    kernels are derived on first use (vgs_ensure_scan_tables, vgs_kmer_i8_prepare_tables, vgs_ensure_umap): a run
    //    builds only what its kind / mode reads
    h->nll_ready = true;
    VGS_CUDA(cudaStreamSynchronize(s));
    lap("stream drained");
    if (load_timing) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, ev_t[0], ev_t[1]);
        cudaEventElapsedTime(&b, ev_t[0], ev_t[2]);
        for (size_t i = 0; i + 1 < ev_batch.size(); i += 2) {
            float u = 0.f, d = 0.f;
            cudaEventElapsedTime(&u, ev_t[0], ev_batch[i]);
            cudaEventElapsedTime(&d, ev_t[0], ev_batch[i + 1]);
            fprintf(stderr, "[vgs_load_models]   batch %2zu: host enqueued at %.3f ms; device: upload done %.3f, digits done %.3f ms after the first copy began\n",
                    i / 2, ev_host[i], u, d);
        }
        for (auto &e : ev_batch) cudaEventDestroy(e);
        fprintf(stderr, "[vgs_load_models] device: keys/hash done %.3f ms, last ln p upload + digits done %.3f ms after the first copy began\n", a, b);
        for (auto &e : ev_t) cudaEventDestroy(e);
    }
    if (eager_digits && (rc = vgs_kmer_i8_digits_end(h, s))) return rc;
    return VGS_OK;
}

int vgs_ensure_scan_tables(vgs_context *h, cudaStream_t s) {
    if (h->dense_state) return VGS_OK;
    { const int rce = vgs_ensure_exact_logp(h, s); if (rce) return rce; }
    const int64_t n_models = h->n_models;
    if (h->max_nll_bytes_t1 > 0) {
        // grid.x covers up to 4^6 histories with 256-thread blocks
        dim3 grid(16, 1);
        for (int64_t m0 = 0; m0 < n_models; m0 += 65535) {
            grid.y = (unsigned)std::min<int64_t>(n_models - m0, 65535);
            k_build_dense<<<grid, 256, 0, s>>>(h->d_meta + m0, h->d_hash, h->d_logp, h->d_nll);
            VGS_LAUNCH_CHECK(h);
        }
    }
    if (h->max_nll_bytes_t2 > 0 || h->max_nll_bytes_t3 > 0) {
        const int l1_smem = 2 << (2 * VGS_L1_MAX_DEPTH);  // u16 map of the deepest first level
        VGS_CUDA(cudaFuncSetAttribute(k_build_two_level, cudaFuncAttributeMaxDynamicSharedMemorySize, l1_smem));
        k_build_two_level<<<(unsigned)n_models, 512, l1_smem, s>>>(h->d_meta, h->d_hash, h->d_key, h->d_len, h->d_code, h->d_logp, h->d_nll);
        VGS_LAUNCH_CHECK(h);
    }
    h->dense_state = 1;
    return VGS_OK;
}

int vgs_ensure_umap(vgs_context *h, cudaStream_t s) {
    if (h->umap_tried) return VGS_OK;
    h->umap_tried = true;
    h->umap_stride = 0;
    const int64_t n_models = h->n_models;
    if (h->max_order <= VGS_UMAP_MAX_ORDER && h->max_n_ctx < 0x7FFF) {
        const int64_t total_u = ((((int64_t)1) << (2 * (h->max_order + 1))) - 1) / 3;
        const int64_t stride = (total_u + 7) / 8 * 8;
        if (stride * n_models * 2 <= ((int64_t)8 << 30)) {  // at most 8 GB of maps
            int rc;
            if ((rc = vgs_reserve(h, &h->d_umap, stride * n_models))) return rc;
            for (int64_t m0 = 0; m0 < n_models; m0 += 65535) {
                dim3 grid((unsigned)std::min<int64_t>((stride + 255) / 256, 64), (unsigned)std::min<int64_t>(n_models - m0, 65535));
                k_build_umap<<<grid, 256, 0, s>>>(h->d_meta + m0, h->d_hash, h->max_order, stride, h->d_umap + m0 * stride);
                VGS_LAUNCH_CHECK(h);
            }
            h->umap_stride = stride;
        }
    }
    return VGS_OK;
}

int vgs_ensure_ph(vgs_context *h, cudaStream_t s) {
    if (h->ph_ready) return VGS_OK;
    int rc;
    if ((rc = vgs_reserve(h, &h->d_ph, h->ph_bytes))) return rc;
    // scratch: per model cnt[r] | start[r + 1] | items[n] | order[r]
    std::vector<int64_t> off((size_t)h->n_models + 1, 0);
    for (int64_t m = 0; m < h->n_models; ++m) {
        const ModelMeta &mm = h->h_meta[(size_t)m];
        const int64_t r = (int64_t)1 << mm.ph_lgr;
        off[(size_t)m + 1] = off[(size_t)m] + 3 * r + 1 + mm.n_ctx;
    }
    // one-off build scratch of its own (callers may be holding pointers into the shared scratch buffers)
    const int64_t off_bytes = ((int64_t)off.size() * 8 + 255) / 256 * 256;
    char *tmp = nullptr;
    cudaError_t e = cudaMalloc((void **)&tmp, (size_t)(off_bytes + off.back() * 4 + 256));
    if (e != cudaSuccess) {
        vgs_set_error("cudaMalloc of the perfect-hash build scratch failed: %s", cudaGetErrorString(e));
        return VGS_ERR_MEMORY;
    }
    int64_t *d_off = (int64_t *)tmp;
    int *d_scr = (int *)(tmp + off_bytes);
    e = cudaMemcpyAsync(d_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        k_build_ph<<<(unsigned)((h->n_models + 31) / 32), 32, 0, s>>>(h->n_models, h->d_meta, h->d_key, h->d_ph, d_scr, d_off);
        h->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(tmp);
    if (e != cudaSuccess) return vgs_cuda_fail(e, "perfect-hash build", __FILE__, __LINE__);
    h->ph_ready = true;
    return VGS_OK;
}

extern "C" int vgs_model_count(vgs_handle h, int64_t *n_models, int64_t *total_contexts) {
    VGS_CHECK_ARG(h, "null handle");
    if (n_models) *n_models = h->n_models;
    if (total_contexts) *total_contexts = h->total_ctx;
    return VGS_OK;
}

extern "C" int vgs_model_info(vgs_handle h, int64_t model, int *order, int64_t *n_ctx, int *tier) {
    VGS_CHECK_ARG(h && model >= 0 && model < h->n_models, "vgs_model_info: bad model index");
    const ModelMeta &mm = h->h_meta[(size_t)model];
    if (order) *order = mm.order;
    if (n_ctx) *n_ctx = mm.n_ctx;
    if (tier) *tier = mm.tier;
    return VGS_OK;
}

extern "C" int vgs_get_logp(vgs_handle h, double *logp_out_h) {
    VGS_CHECK_ARG(h && logp_out_h && h->n_models > 0, "vgs_get_logp: no models loaded");
    VGS_CHECK_ARG(h->nll_ready, "vgs_get_logp: models were loaded with VGS_LOAD_SKIP_NLL");
    VGS_CUDA(cudaSetDevice(h->device));
    { const int rce = vgs_ensure_exact_logp(h, h->own_stream); if (rce) return rce; }
    // read back from the device so the test sees what the kernels see
    VGS_CUDA(cudaMemcpy(logp_out_h, h->d_logp, sizeof(double) * (size_t)h->total_ctx * 4, cudaMemcpyDeviceToHost));
    return VGS_OK;
}

// ---------------------------------------------------------------------------------------------
// lookup (a3 / a4)
// ---------------------------------------------------------------------------------------------
__global__ void k_lookup(int64_t n, const ModelMeta *__restrict__ meta, const uint2 *__restrict__ hash,
                         const int32_t *__restrict__ model, const uint32_t *__restrict__ hcode,
                         const uint8_t *__restrict__ hlen, int32_t *__restrict__ out, uint32_t *__restrict__ all_mask) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const ModelMeta mm = meta[model[q]];
    const uint2 *slots = hash + mm.hash_off;
    int hl = hlen[q];
    uint32_t hc = hcode[q];
    int top = hl < mm.order ? hl : mm.order;
    int best = -1;
    uint32_t mask = 0;
    for (int ln = top; ln >= 0; --ln) {
        int id = vgs_hash_find(slots, mm.hash_mask, mm.hash_shift, vgs_key(ln, hc & vgs_lowmask(ln)));
        if (id >= 0) {
            mask |= 1u << ln;
            if (best < 0) best = id;
        }
    }
    out[q] = best;
    if (all_mask) all_mask[q] = mask;
}

extern "C" int vgs_lookup(vgs_handle h, int64_t n, const int32_t *model_h, const uint32_t *hist_code_h,
                          const uint8_t *hist_len_h, int32_t *ctx_id_out_h, uint32_t *all_mask_out_h) {
    VGS_CHECK_ARG(h && h->n_models > 0, "vgs_lookup: no models loaded");
    VGS_CHECK_ARG(n >= 0 && (n == 0 || (model_h && hist_code_h && hist_len_h && ctx_id_out_h)), "vgs_lookup: null input");
    if (n == 0) return VGS_OK;
    for (int64_t q = 0; q < n; ++q) {
        VGS_CHECK_ARG(model_h[q] >= 0 && model_h[q] < h->n_models, "vgs_lookup: model index out of range");
        VGS_CHECK_ARG(hist_len_h[q] <= 16, "vgs_lookup: history longer than 16 symbols (pass its last 16)");
    }
    VGS_CUDA(cudaSetDevice(h->device));
    int64_t bytes = n * (4 + 4 + 1 + 4 + 4) + 64;
    int rc = vgs_ensure_scratch(h, bytes);
    if (rc) return rc;
    char *base = (char *)h->d_scratch;
    int32_t *d_model = (int32_t *)base;
    uint32_t *d_code = (uint32_t *)(base + 4 * n);
    int32_t *d_out = (int32_t *)(base + 8 * n);
    uint32_t *d_mask = (uint32_t *)(base + 12 * n);
    uint8_t *d_len = (uint8_t *)(base + 16 * n);
    cudaStream_t s = h->own_stream;
    VGS_CUDA(cudaMemcpyAsync(d_model, model_h, 4 * (size_t)n, cudaMemcpyHostToDevice, s));
    VGS_CUDA(cudaMemcpyAsync(d_code, hist_code_h, 4 * (size_t)n, cudaMemcpyHostToDevice, s));
    VGS_CUDA(cudaMemcpyAsync(d_len, hist_len_h, (size_t)n, cudaMemcpyHostToDevice, s));
    k_lookup<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(n, h->d_meta, h->d_hash, d_model, d_code, d_len, d_out, d_mask);
    VGS_LAUNCH_CHECK(h);
    VGS_CUDA(cudaMemcpyAsync(ctx_id_out_h, d_out, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (all_mask_out_h) VGS_CUDA(cudaMemcpyAsync(all_mask_out_h, d_mask, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
    VGS_CUDA(cudaStreamSynchronize(s));
    return VGS_OK;
}

// ---------------------------------------------------------------------------------------------
// sequences
// ---------------------------------------------------------------------------------------------
static int set_sequence_layout(vgs_context *h, int64_t n_seq, const int64_t *lengths) {
    h->h_seq_len.assign(lengths, lengths + n_seq);
    h->h_seq_off.assign((size_t)n_seq + 1, 0);
    int64_t total = 0;
    for (int64_t i = 0; i < n_seq; ++i) {
        if (lengths[i] < 0) { vgs_set_error("negative sequence length"); return VGS_ERR_INVALID; }
        int64_t words = ((lengths[i] + 15) / 16 + 3) / 4 * 4;
        h->h_seq_off[(size_t)i + 1] = h->h_seq_off[(size_t)i] + words;
        total += lengths[i];
    }
    h->n_seq = n_seq;
    h->total_symbols = total;
    h->total_words = h->h_seq_off[(size_t)n_seq];
    int rc;
    if ((rc = vgs_reserve(h, &h->d_seq, h->total_words + 4))) return rc;
    if ((rc = vgs_reserve(h, &h->d_seq_off, n_seq + 1))) return rc;
    if ((rc = vgs_reserve(h, &h->d_seq_len, n_seq))) return rc;
    return VGS_OK;
}

// one thread packs one output word from 16 ASCII symbols
__global__ void k_pack_ascii(int64_t n_seq, const uint8_t *__restrict__ ascii, const int64_t *__restrict__ in_off,
                             const int64_t *__restrict__ word_off, uint32_t *__restrict__ words, int *__restrict__ err) {
    int64_t s = blockIdx.y;
    int64_t len = in_off[s + 1] - in_off[s];
    int64_t nw = word_off[s + 1] - word_off[s];
    const uint8_t *src = ascii + in_off[s];
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nw; w += (int64_t)gridDim.x * blockDim.x) {
        uint32_t v = 0;
        for (int j = 0; j < 16; ++j) {
            int64_t t = w * 16 + j;
            uint32_t c = 0;
            if (t < len) {
                uint8_t ch = src[t];
                if (ch == 'A') c = 0; else if (ch == 'C') c = 1; else if (ch == 'G') c = 2; else if (ch == 'T') c = 3;
                else atomicOr(err, 2);
            }
            v |= c << (30 - 2 * j);
        }
        words[word_off[s] + w] = v;
    }
}

extern "C" int vgs_load_sequences_ascii(vgs_handle h, int64_t n_seq, const uint8_t *ascii_h, const int64_t *offsets_h) {
    VGS_CHECK_ARG(h && n_seq > 0 && ascii_h && offsets_h, "vgs_load_sequences_ascii: null/empty input");
    VGS_CUDA(cudaSetDevice(h->device));
    free_sequences(h);
    std::vector<int64_t> lens((size_t)n_seq);
    for (int64_t i = 0; i < n_seq; ++i) lens[(size_t)i] = offsets_h[i + 1] - offsets_h[i];
    int rc = set_sequence_layout(h, n_seq, lens.data());
    if (rc) return rc;
    int64_t total = offsets_h[n_seq] - offsets_h[0];
    rc = vgs_ensure_scratch(h, total + (n_seq + 1) * 8 + 64);
    if (rc) return rc;
    cudaStream_t s = h->own_stream;
    int64_t *d_in_off = (int64_t *)h->d_scratch;
    uint8_t *d_ascii = (uint8_t *)h->d_scratch + (n_seq + 1) * 8;
    std::vector<int64_t> rel((size_t)n_seq + 1);
    for (int64_t i = 0; i <= n_seq; ++i) rel[(size_t)i] = offsets_h[i] - offsets_h[0];
    VGS_CUDA(cudaMemcpyAsync(d_in_off, rel.data(), 8 * (size_t)(n_seq + 1), cudaMemcpyHostToDevice, s));
    VGS_CUDA(cudaMemcpyAsync(d_ascii, ascii_h + offsets_h[0], (size_t)total, cudaMemcpyHostToDevice, s));
    VGS_CUDA(cudaMemcpyAsync(h->d_seq_off, h->h_seq_off.data(), 8 * (size_t)(n_seq + 1), cudaMemcpyHostToDevice, s));
    VGS_CUDA(cudaMemcpyAsync(h->d_seq_len, h->h_seq_len.data(), 8 * (size_t)n_seq, cudaMemcpyHostToDevice, s));
    int64_t max_words = 1;
    for (int64_t i = 0; i < n_seq; ++i) max_words = std::max(max_words, h->h_seq_off[(size_t)i + 1] - h->h_seq_off[(size_t)i]);
    for (int64_t s0 = 0; s0 < n_seq; s0 += 65535) {
        dim3 grid((unsigned)std::min<int64_t>((max_words + 127) / 128, 1024), (unsigned)std::min<int64_t>(n_seq - s0, 65535));
        k_pack_ascii<<<grid, 128, 0, s>>>(n_seq - s0, d_ascii, d_in_off + s0, h->d_seq_off + s0, h->d_seq, h->d_err);
        VGS_LAUNCH_CHECK(h);
    }
    return vgs_read_error_flag(h, s);
}

extern "C" int vgs_load_sequences_packed(vgs_handle h, int64_t n_seq, const uint32_t *words_h, const int64_t *word_offsets_h,
                                         const int64_t *lengths_h) {
    VGS_CHECK_ARG(h && n_seq > 0 && words_h && word_offsets_h && lengths_h, "vgs_load_sequences_packed: null/empty input");
    VGS_CUDA(cudaSetDevice(h->device));
    free_sequences(h);
    int rc = set_sequence_layout(h, n_seq, lengths_h);
    if (rc) return rc;
    for (int64_t i = 0; i <= n_seq; ++i)
        if (word_offsets_h[i] != h->h_seq_off[(size_t)i]) {
            vgs_set_error("vgs_load_sequences_packed: word_offsets must be 16-byte aligned starts, ceil(len/16) words rounded up to 4");
            return VGS_ERR_INVALID;
        }
    cudaStream_t s = h->own_stream;
    VGS_CUDA(cudaMemcpyAsync(h->d_seq, words_h, 4 * (size_t)h->total_words, cudaMemcpyHostToDevice, s));
    VGS_CUDA(cudaMemcpyAsync(h->d_seq_off, h->h_seq_off.data(), 8 * (size_t)(n_seq + 1), cudaMemcpyHostToDevice, s));
    VGS_CUDA(cudaMemcpyAsync(h->d_seq_len, h->h_seq_len.data(), 8 * (size_t)n_seq, cudaMemcpyHostToDevice, s));
    VGS_CUDA(cudaStreamSynchronize(s));
    return VGS_OK;
}

extern "C" int vgs_sequence_count(vgs_handle h, int64_t *n_seq, int64_t *total_symbols) {
    VGS_CHECK_ARG(h, "null handle");
    if (n_seq) *n_seq = h->n_seq;
    if (total_symbols) *total_symbols = h->total_symbols;
    return VGS_OK;
}

extern "C" int vgs_get_sequence(vgs_handle h, int64_t sidx, uint8_t *codes_out_h, int64_t capacity) {
    VGS_CHECK_ARG(h && sidx >= 0 && sidx < h->n_seq && codes_out_h, "vgs_get_sequence: bad index");
    int64_t len = h->h_seq_len[(size_t)sidx];
    VGS_CHECK_ARG(capacity >= len, "vgs_get_sequence: output buffer too small");
    VGS_CUDA(cudaSetDevice(h->device));
    int64_t nw = h->h_seq_off[(size_t)sidx + 1] - h->h_seq_off[(size_t)sidx];
    std::vector<uint32_t> w((size_t)nw);
    VGS_CUDA(cudaMemcpy(w.data(), h->d_seq + h->h_seq_off[(size_t)sidx], 4 * (size_t)nw, cudaMemcpyDeviceToHost));
    for (int64_t t = 0; t < len; ++t) codes_out_h[t] = (uint8_t)((w[(size_t)(t >> 4)] >> (30 - 2 * (t & 15))) & 3u);
    return VGS_OK;
}

// ---------------------------------------------------------------------------------------------
// sampler (a12 / f2): src/vlmc/vlmc.pyx:156-177 with random.choices spelled out:
//   cum = accumulate(p_A, p_C, p_G, p_T); total = cum[3] + 0.0; x = bisect_right(cum, random()*total, 0, 3)
// The MT19937 stream is advanced on the host exactly as CPython does (genrand_res53); the sequential
// chain (symbol t depends on symbols < t) runs one thread per model on the device.
// ---------------------------------------------------------------------------------------------
struct HostMT {
    uint32_t mt[624];
    int idx;
    uint32_t next() {
        if (idx >= 624) {
            int kk;
            for (kk = 0; kk < 624 - 397; kk++) {
                uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            for (; kk < 623; kk++) {
                uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    double random() {
        uint32_t a = next() >> 5, b = next() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
};

__global__ void k_sample(int64_t n_models, const ModelMeta *__restrict__ meta, const uint2 *__restrict__ hash,
                         const double *__restrict__ prob64, const double *__restrict__ uniforms, int64_t length,
                         int64_t pre, const int64_t *__restrict__ word_off, uint32_t *__restrict__ words, int *__restrict__ err,
                         uint32_t init_code, int init_len) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_models) return;
    const ModelMeta mm = meta[m];
    const uint2 *slots = hash + mm.hash_off;
    const double *u = uniforms + m * (length + pre);
    uint32_t *out = words + word_off[m];
    // generate_sequence_from(length, context): the chain starts from the caller's context; only its last `order`
    // symbols are ever read (vlmc.pyx:168: generated_sequence[-self.order:])
    const uint32_t omask = vgs_lowmask(mm.order);
    uint32_t hist = init_code & omask;
    uint32_t cur = 0;
    for (int64_t t = 0; t < length + pre; ++t) {
        int hl = t + init_len < mm.order ? (int)t + init_len : mm.order;
        int id = vgs_longest_suffix(slots, mm.hash_mask, mm.hash_shift, mm.order, hist, hl);
        if (id < 0) { atomicOr(err, 1); return; }
        const double *p = prob64 + 4 * (mm.ctx_off + id);
        double c0 = p[0];
        double c1 = __dadd_rn(c0, p[1]);
        double c2 = __dadd_rn(c1, p[2]);
        double c3 = __dadd_rn(c2, p[3]);
        if (!(c3 > 0.0)) { atomicOr(err, 8); return; }   // random.choices: "Total of weights must be greater than zero"
        double x = __dmul_rn(u[t], __dadd_rn(c3, 0.0));
        uint32_t sym = (x < c0) ? 0u : (x < c1) ? 1u : (x < c2) ? 2u : 3u;
        if (t >= pre) {
            int64_t k = t - pre;
            cur |= sym << (30 - 2 * (int)(k & 15));
            if ((k & 15) == 15 || k == length - 1) { out[k >> 4] = cur; cur = 0; }
        }
        hist = ((hist << 2) | sym) & omask;
    }
}

extern "C" int vgs_sample_sequences(vgs_handle h, int64_t length, int64_t pre_sample, uint32_t *mt_state_h) {
    return vgs_sample_sequences_from(h, length, pre_sample, mt_state_h, nullptr, 0);
}

extern "C" int vgs_sample_sequences_from(vgs_handle h, int64_t length, int64_t pre_sample, uint32_t *mt_state_h,
                                         const uint8_t *context_codes_h, int64_t context_len) {
    VGS_CHECK_ARG(h && h->n_models > 0, "vgs_sample_sequences: no models loaded");
    VGS_CHECK_ARG(context_len >= 0 && (context_len == 0 || context_codes_h), "vgs_sample_sequences_from: bad context");
    uint32_t init_code = 0;
    const int init_len = (int)std::min<int64_t>(context_len, VGS_MAX_ORDER);
    for (int64_t t = context_len - init_len; t < context_len; ++t) {
        VGS_CHECK_ARG(context_codes_h[t] < 4, "vgs_sample_sequences_from: context symbol outside 0..3");
        init_code = (init_code << 2) | context_codes_h[t];
    }
    VGS_CHECK_ARG(h->prob_ready, "vgs_sample_sequences: models were loaded with VGS_LOAD_NLL_ONLY: no transition rows on the device");
    VGS_CHECK_ARG(length > 0 && pre_sample >= 0 && mt_state_h, "vgs_sample_sequences: bad arguments");
    VGS_CHECK_ARG(mt_state_h[624] <= 624, "vgs_sample_sequences: bad MT19937 position");
    VGS_CUDA(cudaSetDevice(h->device));
    free_sequences(h);
    const int64_t n = h->n_models, per = length + pre_sample;
    std::vector<int64_t> lens((size_t)n, length);
    int rc = set_sequence_layout(h, n, lens.data());
    if (rc) return rc;
    std::vector<double> u((size_t)(n * per));
    HostMT mt;
    memcpy(mt.mt, mt_state_h, sizeof(mt.mt));
    mt.idx = (int)mt_state_h[624];
    for (int64_t i = 0; i < n * per; ++i) u[(size_t)i] = mt.random();
    rc = vgs_ensure_scratch(h, 8 * n * per);
    if (rc) return rc;
    cudaStream_t s = h->own_stream;
    VGS_CUDA(cudaMemcpyAsync(h->d_scratch, u.data(), 8 * (size_t)(n * per), cudaMemcpyHostToDevice, s));
    VGS_CUDA(cudaMemcpyAsync(h->d_seq_off, h->h_seq_off.data(), 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, s));
    VGS_CUDA(cudaMemcpyAsync(h->d_seq_len, h->h_seq_len.data(), 8 * (size_t)n, cudaMemcpyHostToDevice, s));
    VGS_CUDA(cudaMemsetAsync(h->d_seq, 0, 4 * (size_t)(h->total_words + 4), s));
    k_sample<<<(unsigned)((n + 31) / 32), 32, 0, s>>>(n, h->d_meta, h->d_hash, h->d_prob64, (const double *)h->d_scratch,
                                                     length, pre_sample, h->d_seq_off, h->d_seq, h->d_err, init_code, init_len);
    VGS_LAUNCH_CHECK(h);
    rc = vgs_read_error_flag(h, s);
    if (rc) return rc;
    memcpy(mt_state_h, mt.mt, sizeof(mt.mt));
    mt_state_h[624] = (uint32_t)mt.idx;
    return VGS_OK;
}
