// vgs_peer.cu -- peer memory over NVLink for the multi-GPU paths (one process per GPU): buffers allocated by the library
// are exported as CUDA IPC handles, the other ranks map them, and kernels then store straight into every rank's copy
// (the int8 GEMM's epilogue writes its score band into all of them) -- no staging payload, no separate all-gather.
// A tiny flag protocol in the same kind of memory replaces the collective's implicit barrier.
#include <string.h>

#include "vgs_common.cuh"

extern "C" int vgs_peer_alloc(vgs_handle h, int64_t bytes, void **dev_ptr_out, void *ipc_handle_out) {
    VGS_CHECK_ARG(h && bytes > 0 && dev_ptr_out && ipc_handle_out, "vgs_peer_alloc: bad arguments");
    VGS_CUDA(cudaSetDevice(h->device));
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) {
        vgs_set_error("cudaMalloc of %lld peer-visible bytes failed: %s", (long long)bytes, cudaGetErrorString(e));
        return VGS_ERR_MEMORY;
    }
    VGS_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t ipc;
    e = cudaIpcGetMemHandle(&ipc, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return vgs_cuda_fail(e, "cudaIpcGetMemHandle", __FILE__, __LINE__);
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(ipc_handle_out, &ipc, sizeof(ipc));
    *dev_ptr_out = p;
    return VGS_OK;
}

extern "C" int vgs_peer_open(vgs_handle h, const void *ipc_handle, void **dev_ptr_out) {
    VGS_CHECK_ARG(h && ipc_handle && dev_ptr_out, "vgs_peer_open: bad arguments");
    VGS_CUDA(cudaSetDevice(h->device));
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, ipc_handle, sizeof(ipc));
    void *p = nullptr;
    VGS_CUDA(cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr_out = p;
    return VGS_OK;
}

extern "C" int vgs_peer_close(vgs_handle h, void *dev_ptr) {
    VGS_CHECK_ARG(h && dev_ptr, "vgs_peer_close: bad arguments");
    VGS_CUDA(cudaSetDevice(h->device));
    VGS_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return VGS_OK;
}

extern "C" int vgs_peer_free(vgs_handle h, void *dev_ptr) {
    VGS_CHECK_ARG(h && dev_ptr, "vgs_peer_free: bad arguments");
    VGS_CUDA(cudaSetDevice(h->device));
    VGS_CUDA(cudaDeviceSynchronize());
    VGS_CUDA(cudaFree(dev_ptr));
    return VGS_OK;
}

// every rank: "everything I enqueued before this point has reached your memory" -> flags[peer][rank] = epoch, then wait
// until flags[rank][peer] >= epoch for every peer.  One warp; the stores are system-scope releases, the polls
// system-scope acquires; a peer that never arrives sets the error word after ~2 s instead of hanging the GPU.
struct PeerFlags {
    uint32_t *flags[VGS_MAX_PEERS];   // flags[r] = rank r's array of `world` words (mapped here)
};
__global__ void k_peer_barrier(PeerFlags pf, int rank, int world, uint32_t epoch, int *err) {
    const int t = threadIdx.x;
    vgs_dep_trigger();
    vgs_dep_wait();   // "everything I enqueued before": the kernels before this one have completed, their stores are visible
    if (t < world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.flags[t] + rank), "r"(epoch) : "memory");
        const uint32_t *mine = pf.flags[rank] + t;
        const long long t0 = clock64();
        for (;;) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((int32_t)(v - epoch) >= 0) break;
            if (clock64() - t0 > (1ll << 32)) {
                atomicOr(err, 4);
                break;
            }
        }
    }
}

int vgs_peer_barrier_impl(vgs_context *h, uint32_t *const *flags_h, int rank, int world, uint32_t epoch, cudaStream_t stream);
extern "C" int vgs_peer_barrier(vgs_handle h, uint32_t *const *flags_h, int rank, int world, uint32_t epoch, void *stream) {
    VGS_CHECK_ARG(h, "null handle");
    VGS_CUDA(cudaSetDevice(h->device));
    return vgs_peer_barrier_impl(h, flags_h, rank, world, epoch, (cudaStream_t)stream);
}

int vgs_peer_barrier_impl(vgs_context *h, uint32_t *const *flags_h, int rank, int world, uint32_t epoch, cudaStream_t stream) {
    VGS_CHECK_ARG(h && flags_h && world >= 1 && world <= VGS_MAX_PEERS && rank >= 0 && rank < world, "vgs_peer_barrier: bad arguments");
    PeerFlags pf;
    for (int r = 0; r < VGS_MAX_PEERS; ++r) pf.flags[r] = r < world ? flags_h[r] : nullptr;
    for (int r = 0; r < world; ++r) VGS_CHECK_ARG(pf.flags[r] != nullptr, "vgs_peer_barrier: null flag array");
    vgs_launch_dep(k_peer_barrier, dim3(1), dim3(32), 0, stream, pf, rank, world, epoch, h->d_err);
    VGS_LAUNCH_CHECK(h);
    return VGS_OK;
}
