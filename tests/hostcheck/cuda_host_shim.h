// Serial stand-ins for the few CUDA facilities the fluid driver uses (csrc/tm_fluid_cuda.cuh,
// csrc/tm_trimg_cuda.cuh), so that the WHOLE driver -- allocation sizes, pointer offsets, level
// loops, the order of the launches -- runs on the CPU under g++ (-DTM_HOST_SHIM).  A "launch" is a
// plain call with a 1x1 grid: every kernel of the driver is a grid-stride loop, so one "thread"
// sweeps the whole range; grid_reduce degenerates to a store.  Stream capture is modelled too: between
// cudaStreamBeginCapture and cudaStreamEndCapture launches and async memory operations are RECORDED
// on a tape instead of executed (as CUDA does), cudaGraphLaunch replays the tape, and the calls that
// are illegal inside a capture (allocation, synchronous copies, synchronisation) throw -- so the
// contents and the legality of a captured region are checked on the CPU.  Test infrastructure only.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(x)

struct ShimIdx {
    unsigned x = 0, y = 0, z = 0;
};
struct ShimDim {
    unsigned x = 1, y = 1, z = 1;
};
static const ShimIdx threadIdx, blockIdx;
static const ShimDim blockDim, gridDim;

typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void* cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaDevAttrMultiProcessorCount = 16 };

namespace shim {
struct Tape {
    std::vector<std::function<void()>> ops;
};
inline Tape*& capturing() {
    static Tape* tape = nullptr;
    return tape;
}
inline void not_in_capture(const char* what) {
    if (capturing()) throw std::runtime_error(std::string("shim: ") + what + " is illegal during stream capture");
}
template <class F>
inline void submit(F&& f) {
    if (capturing()) capturing()->ops.emplace_back(std::forward<F>(f));
    else f();
}
template <class K>
struct Launcher {
    K kernel;
    template <class... A>
    void operator()(A... args) const {
        K k = kernel;
        submit([=] { k(args...); });
    }
};
template <class K>
inline Launcher<K> make_launcher(K k) {
    return Launcher<K>{k};
}
}  // namespace shim

typedef shim::Tape* cudaGraph_t;
typedef shim::Tape* cudaGraphExec_t;
enum { cudaStreamCaptureModeThreadLocal = 1 };

template <typename T>
inline cudaError_t cudaMalloc(T** p, size_t bytes) {
    shim::not_in_capture("cudaMalloc");
    *p = static_cast<T*>(std::malloc(bytes ? bytes : 1));
    return *p ? cudaSuccess : 2;
}
template <typename T>
inline cudaError_t cudaMallocHost(T** p, size_t bytes) { return cudaMalloc(p, bytes); }
inline cudaError_t cudaFree(void* p) { shim::not_in_capture("cudaFree"); std::free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemset(void* p, int v, size_t n) { shim::not_in_capture("cudaMemset"); std::memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
    shim::submit([=] { std::memset(p, v, n); });
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
    shim::not_in_capture("cudaMemcpy");
    std::memmove(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
    shim::submit([=] { std::memmove(d, s, n); });
    return cudaSuccess;
}
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { shim::not_in_capture("cudaStreamSynchronize"); return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 4; return cudaSuccess; }
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) {
    shim::not_in_capture("cudaStreamBeginCapture");
    shim::capturing() = new shim::Tape;
    return cudaSuccess;
}
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) {
    *g = shim::capturing();
    shim::capturing() = nullptr;
    return *g ? cudaSuccess : 3;
}
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long) {
    *e = new shim::Tape(*g);
    return cudaSuccess;
}
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t) {
    shim::not_in_capture("cudaGraphLaunch");
    for (auto& op : e->ops) op();
    return cudaSuccess;
}

#define TM_CUDA(expr)                                                        \
    do {                                                                     \
        if ((expr) != cudaSuccess) throw std::runtime_error("shim: " #expr); \
    } while (0)
#define TM_CHECK_LAUNCH() \
    do {                  \
    } while (0)
#define TM_LAUNCH(kernel, grid, block, stream) shim::make_launcher([](auto... a) { kernel(a...); })
#define TM_GRID_STRIDE(i, n) for (size_t i = 0; i < (size_t)(n); ++i)

namespace tmx {

struct ReduceScratch {
    double* partials;
    unsigned int* counter;
    int capacity;
};
template <int NV>
inline void grid_reduce(double (&val)[NV], ReduceScratch, double* const (&outs)[NV]) {
    for (int v = 0; v < NV; ++v) *outs[v] = val[v];
}
constexpr int kVecThreads = 256;
template <typename T>
inline void dot_kernel(size_t n, const T* a, const T* b, ReduceScratch rs, double* out) {
    double val[1] = {0.0};
    for (size_t i = 0; i < n; ++i) val[0] += (double)a[i] * (double)b[i];
    double* const outs[1] = {out};
    grid_reduce<1>(val, rs, outs);
}

}  // namespace tmx
