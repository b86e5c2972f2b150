// Shared helpers for the aqml_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include <vector>
#include <stdexcept>
#include <atomic>

#include "../../include/aqml_b200.h"

namespace aqml {

// ---------------------------------------------------------------------------------------------
// error plumbing: C++ exceptions inside, integer codes + thread-local message at the C boundary
// ---------------------------------------------------------------------------------------------
struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

std::string &last_error_ref();
extern std::atomic<long long> g_launches;

[[noreturn]] inline void fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw Error(code, buf);
}

#define AQML_CUDA(call)                                                                            \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            ::aqml::fail(AQML_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define AQML_REQUIRE(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) ::aqml::fail(AQML_EINVAL, __VA_ARGS__);      \
    } while (0)

#define AQML_LAUNCH_CHECK()                   \
    do {                                      \
        ::aqml::g_launches.fetch_add(1);      \
        AQML_CUDA(cudaGetLastError());        \
    } while (0)

template <class F>
int guarded(F &&f) {
    try {
        f();
        return AQML_OK;
    } catch (const Error &e) {
        last_error_ref() = e.what();
        return e.code;
    } catch (const std::exception &e) {
        last_error_ref() = e.what();
        return AQML_EINVAL;
    }
}

void require_device();

// ---------------------------------------------------------------------------------------------
// stream-ordered workspace: every temporary comes from cudaMallocAsync on the call's stream and
// is returned with cudaFreeAsync when the Workspace dies, so `_dev` entry points never block.
// ---------------------------------------------------------------------------------------------
class Workspace {
  public:
    explicit Workspace(cudaStream_t s) : stream_(s) {}
    ~Workspace() {
        for (void *p : ptrs_) cudaFreeAsync(p, stream_);
    }
    Workspace(const Workspace &) = delete;
    Workspace &operator=(const Workspace &) = delete;

    template <class T>
    T *alloc(size_t n) {
        void *p = nullptr;
        size_t bytes = (n ? n : 1) * sizeof(T);
        if (slab_) {  // carve from the slab reserved up front (256-byte aligned pieces, same order on every rank)
            const size_t o = slab_aligned(slab_used_);
            if (o + bytes > slab_bytes_) fail(AQML_EINVAL, "internal: slab of %zu bytes too small (%zu + %zu)", slab_bytes_, o, bytes);
            slab_used_ = o + bytes;
            return reinterpret_cast<T *>(slab_ + o);
        }
        AQML_CUDA(cudaMallocAsync(&p, bytes, stream_));
        ptrs_.push_back(p);
        return reinterpret_cast<T *>(p);
    }
    // one allocation that every following alloc() is carved from: the owner's storage becomes a single contiguous
    // buffer that can be sent to another GPU as it is (aqml_rep_slab)
    void reserve(size_t bytes) {
        void *p = nullptr;
        AQML_CUDA(cudaMallocAsync(&p, bytes ? bytes : 1, stream_));
        ptrs_.push_back(p);
        slab_ = reinterpret_cast<unsigned char *>(p);
        slab_bytes_ = bytes;
        slab_used_ = 0;
    }
    static size_t slab_aligned(size_t v) { return (v + 255) / 256 * 256; }
    static size_t slab_piece(size_t bytes) { return slab_aligned(bytes ? bytes : 1); }
    unsigned char *slab() const { return slab_; }
    size_t slab_bytes() const { return slab_bytes_; }
    template <class T>
    T *upload(const T *host, size_t n) {
        T *d = alloc<T>(n);
        if (n) AQML_CUDA(cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, stream_));
        return d;
    }
    template <class T>
    T *upload(const std::vector<T> &v) {
        return upload(v.data(), v.size());
    }
    template <class T>
    T *zeros(size_t n) {
        T *d = alloc<T>(n);
        AQML_CUDA(cudaMemsetAsync(d, 0, (n ? n : 1) * sizeof(T), stream_));
        return d;
    }
    cudaStream_t stream() const { return stream_; }
    // hand a pointer over to a longer-lived owner (aqml_rep)
    void release_all() { ptrs_.clear(); }
    const std::vector<void *> &pointers() const { return ptrs_; }

  private:
    cudaStream_t stream_;
    std::vector<void *> ptrs_;
    unsigned char *slab_ = nullptr;
    size_t slab_bytes_ = 0, slab_used_ = 0;
};

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// FP64 tensor-core tile: D(8x8) += A(8x4, row) * B(4x8, col); SASS: DMMA.8x8x4 (the only native
// FP64 MMA shape on sm_100a -- m16n8k{4,8,16} lower to sequences of it).
// lane l holds A[l/4][l%4], B[l%4][l/4], C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace aqml
