#include "common.h"

namespace umgap {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

void use_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device available (%s); libumgap_gpu has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        throw StatusError{UMGAP_ERR_CUDA};
    }
    if (device < 0 || device >= n) {
        set_error("device %d out of range (have %d)", device, n);
        throw StatusError{UMGAP_ERR_INVALID};
    }
    UMGAP_CUDA(cudaSetDevice(device));
}

void* Workspace::get(int slot, size_t bytes) {
    if (bytes <= cap[slot] && ptr[slot]) return ptr[slot];
    if (ptr[slot]) {
        UMGAP_CUDA(cudaDeviceSynchronize());
        cudaFree(ptr[slot]);
        ptr[slot] = nullptr;
        cap[slot] = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    UMGAP_CUDA(cudaMalloc(&ptr[slot], want));
    cap[slot] = want;
    return ptr[slot];
}

void Workspace::release() {
    for (int i = 0; i < kSlots; ++i) {
        if (ptr[i]) cudaFree(ptr[i]);
        ptr[i] = nullptr;
        cap[i] = 0;
    }
}

}  // namespace umgap

extern "C" {
const char* umgap_last_error(void) { return umgap::get_error(); }
int umgap_abi_version(void) { return UMGAP_ABI_VERSION; }
void* umgap_host_alloc(size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        umgap::set_error("cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
void umgap_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
int umgap_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        umgap::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return UMGAP_ERR_CUDA;
    }
    return n;
}
}
