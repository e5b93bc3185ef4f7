// Shared host-side plumbing for libumgap_gpu.so: thread-local error text, CUDA checks and the
// grow-only device workspace the handles cache between calls.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/umgap_gpu.h"

namespace umgap {

void set_error(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
const char* get_error();

struct StatusError {
    int code;
};

// Throwing check used inside entry points; every entry point catches StatusError.
#define UMGAP_CUDA(expr)                                                                  \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            ::umgap::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__),       \
                               __FILE__, __LINE__, cudaGetErrorString(e__));              \
            throw ::umgap::StatusError{UMGAP_ERR_CUDA};                                   \
        }                                                                                 \
    } while (0)

#define UMGAP_FAIL(code, ...)                  \
    do {                                       \
        ::umgap::set_error(__VA_ARGS__);       \
        throw ::umgap::StatusError{(code)};    \
    } while (0)

// Wraps an entry-point body: converts exceptions to status codes.
template <class F>
int guarded(F&& f) {
    try {
        f();
        return UMGAP_OK;
    } catch (const StatusError& s) {
        return s.code;
    } catch (const std::bad_alloc&) {
        set_error("out of host memory");
        return UMGAP_ERR_NOMEM;
    } catch (const std::exception& e) {
        set_error("internal error: %s", e.what());
        return UMGAP_ERR_INVALID;
    }
}

void use_device(int device);  // cudaSetDevice with a clear error when no GPU is present

// Grow-only device scratch, one buffer per slot; freed with the owning handle.
struct Workspace {
    static const int kSlots = 56;
    void* ptr[kSlots] = {};
    size_t cap[kSlots] = {};
    void* get(int slot, size_t bytes);
    void release();
};

template <class T>
struct DevBuf {  // RAII device allocation for one call
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    void alloc(size_t count) {
        free();
        n = count;
        if (count) UMGAP_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
    }
    void free() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { free(); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

static inline uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

}  // namespace umgap
