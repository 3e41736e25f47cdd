// One-shot all-reduce of a few KB over NVLink peer memory, for the latency-bound reductions of the sharded step
// (batch-norm forward sums [2 d_d doubles], backward column sums + loss [2 d_d + 1 doubles]). Every
// rank owns an inbox with one slot per peer; a reduction is: push my vector into my slot of every peer's inbox
// (plain NVLink stores), publish an epoch flag with release semantics at system scope, wait for the flags of all
// peers in my own inbox, sum the slots in rank order (so every rank computes bit-identical results). Slots and
// flags are double-buffered by epoch parity: a rank can only be one exchange ahead of any peer, because finishing
// exchange e needs every peer's flag e. One 256-thread block, ~3 us on NVSwitch vs ~17 us per small ncclAllReduce.
// The inboxes are cudaMalloc'ed per process and mapped into the peers with CUDA IPC (nvsm_comm_peer_export / _import).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace nvsm {

constexpr int kPeerMaxRanks = 16;
constexpr int kPeerKinds = 4;   // independent reduction sites per step, each with its own epoch counter

struct PeerXchg {
    int nranks, rank;
    int slot_doubles;                       // capacity of one slot
    double* inbox[kPeerMaxRanks];           // inbox[p] = base of rank p's inbox as mapped in THIS process
    unsigned long long* flags[kPeerMaxRanks];
};

// inbox layout: [kind][parity][src_rank][slot_doubles]; flags: [kind][parity][src_rank]
__device__ __forceinline__ size_t peer_slot_index(const PeerXchg& x, int kind, int parity, int src) {
    return ((size_t)(kind * 2 + parity) * x.nranks + src) * x.slot_doubles;
}
__device__ __forceinline__ size_t peer_flag_index(const PeerXchg& x, int kind, int parity, int src) {
    return (size_t)(kind * 2 + parity) * x.nranks + src;
}

__global__ void __launch_bounds__(256) peer_allreduce_kernel(const PeerXchg x, double* __restrict__ buf, int n, int kind,
                                                             unsigned long long epoch, int* __restrict__ error_flag) {
    const int parity = (int)(epoch & 1ull);
    // push
    for (int p = 0; p < x.nranks; ++p) {
        double* dst = x.inbox[p] + peer_slot_index(x, kind, parity, x.rank);
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = buf[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < x.nranks) {
        unsigned long long* f = x.flags[threadIdx.x] + peer_flag_index(x, kind, parity, x.rank);
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
    }
    // wait for every peer's contribution to my inbox (bounded: a lost peer sets the error flag instead of hanging)
    if (threadIdx.x < x.nranks) {
        const unsigned long long* f = x.flags[x.rank] + peer_flag_index(x, kind, parity, threadIdx.x);
        unsigned long long v = 0;
        long spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (++spins > (1L << 28)) { atomicExch(error_flag, 1); break; }
        } while (v < epoch);
    }
    __syncthreads();
    const double* mine = x.inbox[x.rank];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int p = 0; p < x.nranks; ++p) s += mine[peer_slot_index(x, kind, parity, p) + i];
        buf[i] = s;
    }
}

}  // namespace nvsm
