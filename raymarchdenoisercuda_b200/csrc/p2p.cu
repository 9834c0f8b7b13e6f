// p2p.cu — NVLink peer-to-peer plumbing between ranks (one process per GPU): CUDA-IPC mappable
// buffers and stream-ordered flags used by the row-banded mode (no reference counterpart: the
// reference is single-GPU).  The data itself moves with ordinary device-to-device
// copies whose destination is a peer-mapped pointer; these helpers only provide the mapping and the
// cross-GPU "data has landed" ordering without any host synchronisation.
#include "common.cuh"

namespace {

// number of flag waits that gave up: a pinned, device-mapped host word, so that the host can poll it after every
// exchange without synchronising the device (round 1 kept it in device memory: reading it meant cudaDeviceSynchronize,
// so nobody did, and a timed-out wait went on to unpack stale history rows)
unsigned int* g_timeouts_host = nullptr;
unsigned int* g_timeouts_dev = nullptr;

int timeouts_init() {
    if (g_timeouts_host) return 0;
    unsigned int* h = nullptr;
    RMD_CUDA_TRY(cudaHostAlloc((void**)&h, sizeof(unsigned int), cudaHostAllocMapped | cudaHostAllocPortable));
    *h = 0u;
    RMD_CUDA_TRY(cudaHostGetDevicePointer((void**)&g_timeouts_dev, h, 0));
    g_timeouts_host = h;
    return 0;
}

__global__ void signal_kernel(unsigned long long* flag, unsigned long long value) {
    __threadfence_system();  // everything this stream wrote before (incl. peer copies) is visible system-wide
    *reinterpret_cast<volatile unsigned long long*>(flag) = value;
    __threadfence_system();
}

// Bounded spin on a word another GPU writes.  Safe to run concurrently with the producer because the
// two run on different devices; the bound turns a protocol bug into a visible error instead of a hang.
__global__ void wait_kernel(const unsigned long long* flag, unsigned long long value, unsigned int* timeouts) {
    const long long t0 = clock64();
    const volatile unsigned long long* f = reinterpret_cast<const volatile unsigned long long*>(flag);
    while (*f < value) {
        __nanosleep(200);
        if (clock64() - t0 > 4000000000LL) {  // ~2 s at 1.9 GHz
            atomicAdd_system(timeouts, 1u);
            break;
        }
    }
    __threadfence_system();
}

// SM clock measured from the device: cycles of %clock64 per nanosecond of %globaltimer over a short spin of one warp.
// bench.py uses it under torchrun, where an NVML query inside the timed region stalls the band path's cross-GPU
// hand-offs by ~2 ms per query (profiles/r2_scaling.md).
__global__ void clock_probe_kernel(unsigned long long* out, unsigned long long spin_ns) {
    if (threadIdx.x != 0) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const long long c0 = clock64();
    do {
        __nanosleep(500);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < spin_ns);
    const long long c1 = clock64();
    out[0] = (unsigned long long)(c1 - c0);
    out[1] = t1 - t0;
}

}  // namespace

extern "C" int rmd_debug_clock_probe(unsigned long long* dev_out2, unsigned int spin_us, void* stream) {
    if (!dev_out2) return RMD_E_NULL;
    clock_probe_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(dev_out2, (unsigned long long)spin_us * 1000ull);
    return (int)cudaGetLastError();
}

static_assert(sizeof(cudaIpcMemHandle_t) == RMD_IPC_HANDLE_BYTES, "IPC handle size");

extern "C" int rmd_p2p_alloc(void** dev_ptr, size_t bytes) {
    if (!dev_ptr || bytes == 0) return RMD_E_NULL;
    RMD_CUDA_TRY(cudaMalloc(dev_ptr, bytes));
    RMD_CUDA_TRY(cudaMemset(*dev_ptr, 0, bytes));
    RMD_CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}
extern "C" int rmd_p2p_free(void* dev_ptr) {
    if (!dev_ptr) return RMD_E_NULL;
    RMD_CUDA_TRY(cudaFree(dev_ptr));
    return 0;
}
extern "C" int rmd_p2p_export(void* dev_ptr, void* handle_out) {
    if (!dev_ptr || !handle_out) return RMD_E_NULL;
    RMD_CUDA_TRY(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_out), dev_ptr));
    return 0;
}
extern "C" int rmd_p2p_open(const void* handle, void** peer_ptr) {
    if (!handle || !peer_ptr) return RMD_E_NULL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    RMD_CUDA_TRY(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int rmd_p2p_close(void* peer_ptr) {
    if (!peer_ptr) return RMD_E_NULL;
    RMD_CUDA_TRY(cudaIpcCloseMemHandle(peer_ptr));
    return 0;
}
extern "C" int rmd_p2p_signal(void* flag, unsigned long long value, void* stream) {
    if (!flag) return RMD_E_NULL;
    signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(flag), value);
    return (int)cudaGetLastError();
}
extern "C" int rmd_p2p_wait(const void* flag, unsigned long long value, void* stream) {
    if (!flag) return RMD_E_NULL;
    const int rc = timeouts_init();
    if (rc) return rc;
    if (*(volatile unsigned int*)g_timeouts_host) return RMD_E_TIMEOUT;  // sticky: an earlier wait gave up
    wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(flag), value, g_timeouts_dev);
    return (int)cudaGetLastError();
}
extern "C" int rmd_p2p_timeouts(void) {
    return g_timeouts_host ? (int)*(volatile unsigned int*)g_timeouts_host : 0;
}
