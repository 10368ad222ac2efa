// Peer-memory halo exchange and scalar all-reduce for the row-strip sharded solver: our own kernels
// over NVLink / NVSwitch instead of ncclSend/ncclRecv/ncclAllReduce launches.
//
// Every rank owns one WINDOW of device memory (cudaMalloc), exported with cudaIpcGetMemHandle and
// mapped by every other rank of the node (cudaIpcOpenMemHandle, lazy peer access).  A window holds
//   * P2PHeader: arrival flags written REMOTELY by the peers, reduction slots, local bookkeeping;
//   * four halo mailboxes: {from the rank below, from the rank above} x {epoch parity}.
//
// Halo exchange of an array = ONE kernel per rank (p2p_halo_kernel):
//   push    my boundary rows go straight into the neighbours' mailboxes (posted remote stores),
//           system-scope fence, the last block to finish raises the neighbours' arrival flags with
//           st.release.sys to the exchange's epoch number;
//   unpack  every block polls ITS OWN window's flags (local memory, no NVLink round trip) until the
//           neighbours' data of this epoch has landed, then copies mailbox -> halo rows.
// No acknowledgement is needed: exchanges are bidirectional between neighbours, so a neighbour can
// reach epoch e+2 (the next use of the same-parity mailbox) only after it unpacked my push of epoch
// e+1, which I issue after my unpack of epoch e (stream order).  Two mailboxes per direction suffice.
// The push phase never waits and the unpack phase waits only on another GPU.  A neighbour's flag
// rises once ALL its blocks have pushed, so the grid is kept <= the SM count (every block resident:
// a block polling in its unpack phase never keeps a block that still has to push off the machine).
//
// All-reduce of <= P2P_RED_MAX doubles = one single-block kernel: every rank writes its values into
// slot [parity][rank] of EVERY window, raises the flags, waits for all peers, and sums the slots in
// rank order -- the same order on every rank, so all ranks hold bit-identical sums (the PCG takes
// identical branches everywhere) and the result is reproducible from run to run.
//
// Epoch counters live in the window, are advanced by the kernels themselves, and never by the host:
// the kernels replay unchanged from a captured CUDA graph.  A poll that does not complete within
// ~2^33 clocks (seconds) raises header.error instead of hanging the GPU; the host turns it into an
// exception at the next scalar read-back.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace tmx {

constexpr int P2P_MAX_RANKS = 16;
constexpr int P2P_RED_MAX = 8;
constexpr size_t P2P_HEADER_BYTES = 8192;
constexpr long long P2P_SPIN_LIMIT = 1LL << 33;

struct P2PHeader {
    // written by the peers
    unsigned long long halo_flag[2][2];                // [0: from below, 1: from above][parity]
    unsigned long long red_flag[P2P_MAX_RANKS][2];     // [source rank][parity]
    double red_slot[2][P2P_MAX_RANKS][P2P_RED_MAX];    // [parity][source rank][value]
    // local
    unsigned long long halo_epoch, red_epoch;
    unsigned int push_count, done_count;
    unsigned int error;
};
static_assert(sizeof(P2PHeader) <= P2P_HEADER_BYTES, "P2P header overflows its reserved bytes");

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// polls a flag of my own window until a peer has raised it to `epoch`; once a poll has timed out
// (`*error` set) later polls give up at once, so a broken run drains in seconds, not hours
__device__ __forceinline__ bool p2p_wait(const unsigned long long* flag, unsigned long long epoch,
                                         const unsigned int* error) {
    if (*reinterpret_cast<const volatile unsigned int*>(error) != 0u) return false;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < epoch) {
        if (clock64() - t0 > P2P_SPIN_LIMIT) return false;
        __nanosleep(20);
    }
    return true;
}

struct P2PHaloArgs {
    char* self;         // my window
    char* below;        // window of rank-1 as mapped here, or nullptr
    char* above;        // window of rank+1 as mapped here, or nullptr
    size_t slot_bytes;  // capacity of one mailbox
    char* v;            // the local array
    size_t up_src, up_bytes;      // rows I send up:   v + up_src,   up_bytes
    size_t down_src, down_bytes;  // rows I send down: v + down_src, down_bytes
    size_t from_below_dst;        // rows arriving from below (up_bytes of them) land at v + from_below_dst
    size_t from_above_dst;        // rows arriving from above (down_bytes) land at v + from_above_dst
};

__host__ __device__ __forceinline__ size_t p2p_mailbox_offset(int from, int parity, size_t slot_bytes) {
    return P2P_HEADER_BYTES + (size_t)(from * 2 + parity) * slot_bytes;
}

// V = copy granule (uint4 / uint2 / unsigned int), chosen by the host from the alignment of the rows
template <typename V>
__global__ void __launch_bounds__(256) p2p_halo_kernel(P2PHaloArgs a) {
    P2PHeader* me = reinterpret_cast<P2PHeader*>(a.self);
    const unsigned long long epoch = me->halo_epoch + 1;  // advanced by the last block to leave
    const int par = (int)(epoch & 1ULL);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    __shared__ int s_flag;

    // ---- push: my boundary rows -> the neighbours' mailboxes
    if (a.above) {
        const V* src = reinterpret_cast<const V*>(a.v + a.up_src);
        V* dst = reinterpret_cast<V*>(a.above + p2p_mailbox_offset(0, par, a.slot_bytes));
        for (size_t i = tid; i < a.up_bytes / sizeof(V); i += nth) dst[i] = src[i];
    }
    if (a.below) {
        const V* src = reinterpret_cast<const V*>(a.v + a.down_src);
        V* dst = reinterpret_cast<V*>(a.below + p2p_mailbox_offset(1, par, a.slot_bytes));
        for (size_t i = tid; i < a.down_bytes / sizeof(V); i += nth) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&me->push_count, 1u);
        s_flag = (t == gridDim.x - 1);
        if (s_flag) {  // every block's stores are fenced: publish
            __threadfence_system();
            if (a.above) st_release_sys(&reinterpret_cast<P2PHeader*>(a.above)->halo_flag[0][par], epoch);
            if (a.below) st_release_sys(&reinterpret_cast<P2PHeader*>(a.below)->halo_flag[1][par], epoch);
            me->push_count = 0;
        }
    }

    // ---- unpack: wait for the neighbours' rows of this epoch in MY window
    __syncthreads();
    if (threadIdx.x == 0) {
        bool ok = true;
        if (a.below) ok = p2p_wait(&me->halo_flag[0][par], epoch, &me->error) && ok;
        if (a.above) ok = p2p_wait(&me->halo_flag[1][par], epoch, &me->error) && ok;
        if (!ok) me->error = 1u;
        s_flag = ok;
    }
    __syncthreads();
    if (s_flag) {
        if (a.below) {
            const V* src = reinterpret_cast<const V*>(a.self + p2p_mailbox_offset(0, par, a.slot_bytes));
            V* dst = reinterpret_cast<V*>(a.v + a.from_below_dst);
            for (size_t i = tid; i < a.up_bytes / sizeof(V); i += nth) dst[i] = __ldcg(src + i);
        }
        if (a.above) {
            const V* src = reinterpret_cast<const V*>(a.self + p2p_mailbox_offset(1, par, a.slot_bytes));
            V* dst = reinterpret_cast<V*>(a.v + a.from_above_dst);
            for (size_t i = tid; i < a.down_bytes / sizeof(V); i += nth) dst[i] = __ldcg(src + i);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&me->done_count, 1u);
        if (t == gridDim.x - 1) {
            me->done_count = 0;
            me->halo_epoch = epoch;
        }
    }
}

struct P2PReduceArgs {
    char* win[P2P_MAX_RANKS];  // every rank's window as mapped here (win[rank] = my own)
    int rank, nranks;
    double* p;  // n values, summed over ranks in place
    int n;
};

__global__ void __launch_bounds__(64) p2p_allreduce_kernel(P2PReduceArgs a) {
    P2PHeader* me = reinterpret_cast<P2PHeader*>(a.win[a.rank]);
    const unsigned long long epoch = me->red_epoch + 1;
    const int par = (int)(epoch & 1ULL);
    const int t = threadIdx.x;
    __shared__ int s_ok;
    if (t == 0) s_ok = 1;
    if (t < a.n) {
        const double val = a.p[t];
        for (int q = 0; q < a.nranks; ++q)
            reinterpret_cast<P2PHeader*>(a.win[q])->red_slot[par][a.rank][t] = val;
    }
    __threadfence_system();
    __syncthreads();
    if (t < a.nranks && t != a.rank) {
        st_release_sys(&reinterpret_cast<P2PHeader*>(a.win[t])->red_flag[a.rank][par], epoch);
        if (!p2p_wait(&me->red_flag[t][par], epoch, &me->error)) {
            me->error = 2u;
            s_ok = 0;
        }
    }
    __syncthreads();
    if (t < a.n && s_ok) {
        double s = 0.0;
        for (int q = 0; q < a.nranks; ++q) s += __ldcg(&me->red_slot[par][q][t]);
        a.p[t] = s;
    }
    if (t == 0) me->red_epoch = epoch;
}

}  // namespace tmx
