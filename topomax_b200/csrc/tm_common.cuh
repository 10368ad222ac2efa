// Shared declarations of libtopomax_b200: error handling, lattice/level descriptors and the
// deterministic two-stage reduction used by every dot product / integral on the path.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <string>

#include "tm_element.cuh"

namespace tmx {

// ---------------------------------------------------------------------------------------
// errors: every C-ABI entry returns 0 or a negative code and leaves the message here
// ---------------------------------------------------------------------------------------
void set_error(const std::string& msg);
const char* last_error();
// every kernel launch of the library passes through TM_CHECK_LAUNCH (tm_launch_count)
extern std::atomic<long long> g_launches;

struct CudaFailure {
    std::string what;
};

#define TM_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            throw ::tmx::CudaFailure{std::string(#expr) + " -> " + cudaGetErrorString(_e) +      \
                                    " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"};   \
        }                                                                                      \
    } while (0)

// kernel launch through one macro: the serial host build of the fluid driver (tests/hostcheck)
// replaces it by a plain call
#define TM_LAUNCH(kernel, grid, block, stream) kernel<<<(grid), (block), 0, (stream)>>>

#define TM_CHECK_LAUNCH()            \
    do {                             \
        ++::tmx::g_launches;         \
        TM_CUDA(cudaGetLastError()); \
    } while (0)

// Programmatic dependent launch (the V-cycle's chain of small dependent kernels): a kernel first
// lets its programmatic dependents be scheduled, then waits for the grids it depends on to
// complete and flush.  Every global access of the kernel comes after the wait, so dependents
// only overlap their launch latency and prologue with the tail of their predecessor.  Both
// instructions are no-ops in launches without a programmatic edge.
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---------------------------------------------------------------------------------------
// lattice description of one multigrid level
// ---------------------------------------------------------------------------------------
// nx x ny cells; P2 nodes live on the (2ny+1) x (2nx+1) half-step lattice, 2 dofs per node,
// stored row-major as interleaved (u_x, u_y) pairs: dof = 2*(j*Lx + i) + comp.
// A node is Dirichlet (u = 0) when i <= dl || i >= dr || j <= db || j >= dt; free sides use
// the sentinels -1 / INT_MAX.  On level 0 the thresholds are the domain sides
// (reference: FEM_src/elasisity_problem.py:183-192, FEM_src/domains.py:18-38); coarse levels
// may pull dr/dt inwards (see tm_mg.cu).
template <typename T>
struct LevelGeom {
    int nx, ny, Lx, Ly;
    int dl, dr, db, dt;
    // coefficient source: level 0 derives the six moments per triangle from the P1 density
    // xi (vertex grid (ny+1) x (nx+1)); coarse levels read them from W (SoA, 12 planes of
    // nx*ny: 6 for T_A then 6 for T_B).
    const T* xi;
    const T* W;
    T simp_min;
    Material<T> mat;
    // Row-strip sharding (one GPU: j_off = 0 and every row is owned).  A rank holds its own
    // lattice rows plus halo rows as ONE local lattice that kernels treat as a standalone
    // mesh; they only WRITE rows in [own_j0, own_j1), and the Dirichlet thresholds db/dt
    // are compared against the global row index j + j_off.
    int j_off, own_j0, own_j1;
    int nyg;  // global cell rows of this level (= ny on one GPU)

    __host__ __device__ bool fixed(int i, int j) const {
        const int jg = j + j_off;
        return i <= dl || i >= dr || jg <= db || jg >= dt;
    }
    __host__ __device__ bool row_fixed(int j) const {
        const int jg = j + j_off;
        return jg <= db || jg >= dt;
    }
    __host__ __device__ bool owns_row(int j) const { return j >= own_j0 && j < own_j1; }
};

// ---------------------------------------------------------------------------------------
// deterministic reduction: block partials, then the last block to finish folds them in a
// fixed order.  NV values per thread, accumulated in double regardless of T.
// ---------------------------------------------------------------------------------------
struct ReduceScratch {
    double* partials;       // [NV][capacity]
    unsigned int* counter;  // zero before first use; reset by the finishing block
    int capacity;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Must be called by every thread of every block of the grid (1-D or 2-D grids, any block
// shape with blockDim.x*blockDim.y*blockDim.z <= 1024).  *outs[v] receives the grid total.
template <int NV>
__device__ __forceinline__ void grid_reduce(double (&val)[NV], ReduceScratch rs,
                                            double* const (&outs)[NV]) {
    __shared__ double s_part[NV][32];
    __shared__ bool s_last;
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    const int lane = tid & 31, wid = tid >> 5, nwarps = (nthreads + 31) >> 5;
    const int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    const int nblocks = gridDim.x * gridDim.y * gridDim.z;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double s = warp_sum(val[v]);
        if (lane == 0) s_part[v][wid] = s;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double s = (lane < nwarps) ? s_part[v][lane] : 0.0;
            s = warp_sum(s);
            if (lane == 0) rs.partials[(size_t)v * rs.capacity + bid] = s;
        }
    }
    if (tid == 0) {
        __threadfence();
        unsigned int ticket = atomicAdd(rs.counter, 1u);
        s_last = (ticket == (unsigned int)(nblocks - 1));
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        // fixed-order fold: thread t sums partials t, t+nthreads, ... then tree over threads
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double s = 0.0;
            for (int b = tid; b < nblocks; b += nthreads)
                s += __ldcg(&rs.partials[(size_t)v * rs.capacity + b]);
            s = warp_sum(s);
            __syncthreads();
            if (lane == 0) s_part[v][wid] = s;
            __syncthreads();
            if (wid == 0) {
                double t = (lane < nwarps) ? s_part[v][lane] : 0.0;
                t = warp_sum(t);
                if (lane == 0) *outs[v] = t;
            }
        }
        if (tid == 0) *rs.counter = 0u;
    }
}

inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace tmx
