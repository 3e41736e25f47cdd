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
constexpr int kPeerKinds = 5;   // independent reduction sites per step, each with its own epoch counter
constexpr int kPeerKindGt = 4;  // grad_transform (floats, its own inbox region: PeerXchg::gt_inbox)
constexpr int kPeerFlagStride = 128;   // flags per (kind, parity, source rank): one per block of a fused exchange (block 0 =
                                       // the stand-alone one-block kernel and the score kernels' last-block tail)

struct PeerXchg {
    int nranks, rank;
    int slot_doubles;                       // capacity of one slot
    double* inbox[kPeerMaxRanks];           // inbox[p] = base of rank p's inbox as mapped in THIS process
    unsigned long long* flags[kPeerMaxRanks];
    // grad_transform exchange: [parity][src_rank][gt_elems] floats behind the double slots of the same allocation
    float* gt_inbox[kPeerMaxRanks];
    long gt_elems;
};

// inbox layout: [kind][parity][src_rank][slot_doubles]; flags: [kind][parity][src_rank][kPeerFlagStride]
__device__ __forceinline__ size_t peer_slot_index(const PeerXchg& x, int kind, int parity, int src) {
    return ((size_t)(kind * 2 + parity) * x.nranks + src) * x.slot_doubles;
}
__device__ __forceinline__ size_t peer_flag_index(const PeerXchg& x, int kind, int parity, int src, int blk = 0) {
    return ((size_t)(kind * 2 + parity) * x.nranks + src) * kPeerFlagStride + blk;
}

// Building blocks of the exchanges that are FUSED into compute kernels (the producer's last stage pushes its partial
// result over NVLink, the same kernel -- or block -- waits for the peers' and goes on with the reduced values):
//   col_stats_reduce_finalize_kernel<true>  batch-norm forward sums: partial-row reduction -> exchange -> mean / invstd
//   score_sums_tail (score kernels)         backward column sums + loss: the last block to finish exchanges them
// Call with threads `t` = 0 .. nranks-1 of one warp / block, after every pushing thread's __threadfence_system() and a
// barrier among them.
__device__ __forceinline__ void peer_publish(const PeerXchg& x, int kind, int parity, int blk, unsigned long long epoch, int t) {
    if (t < x.nranks) {
        unsigned long long* f = x.flags[t] + peer_flag_index(x, kind, parity, x.rank, blk);
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
    }
}
// (bounded: a lost peer sets the error flag instead of hanging the GPU)
__device__ __forceinline__ void peer_wait(const PeerXchg& x, int kind, int parity, int blk, unsigned long long epoch, int t,
                                          int* error_flag) {
    if (t < x.nranks) {
        const unsigned long long* f = x.flags[x.rank] + peer_flag_index(x, kind, parity, t, blk);
        unsigned long long v = 0;
        long spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (++spins > (1L << 28)) { atomicExch(error_flag, 1); break; }
        } while (v < epoch);
    }
}
// value i of source rank p in MY inbox (written by the peer over NVLink: read past L1)
__device__ __forceinline__ double peer_inbox_value(const PeerXchg& x, int kind, int parity, int p, int i) {
    return __ldcg(x.inbox[x.rank] + peer_slot_index(x, kind, parity, p) + i);
}

// Tail of the score kernels: every block has added its column sums / loss to `buf` [n] with double atomics (loss = last
// element). The last block to arrive (done_counter)
//   * at N > 1 with the peer exchange (xp != null): pushes the GPU's totals to every peer, waits for theirs and leaves
//     the global sums in `buf` -- in rank order, so bit-identical on every rank. Replaces a separate one-block all-reduce
//     launch between the score kernel and batch-norm backward;
//   * writes the loss to `loss_host` (mapped pinned memory, nullable): no device-to-host copy operation sits in the
//     stream between the score kernel and the backward pass.
// All threads of the block must call it. No static shared memory (score_ring_kernel opts into all 227 KB as dynamic).
__device__ __forceinline__ void score_sums_tail(const PeerXchg* __restrict__ xp, double* __restrict__ buf, int n,
                                                unsigned long long epoch, int kind, unsigned int* __restrict__ done_counter,
                                                int* __restrict__ error_flag, double* __restrict__ loss_host) {
    if (!done_counter) return;
    __threadfence();      // this thread's atomics on buf are performed before the counter moves
    __syncthreads();
    int mine = 0;
    if (threadIdx.x == 0) mine = atomicAdd(done_counter, 1u) == gridDim.x - 1 ? 1 : 0;
    if (!__syncthreads_or(mine)) return;
    __threadfence();
    if (xp) {
        const PeerXchg& x = *xp;
        const int parity = (int)(epoch & 1ull);
        for (int p = 0; p < x.nranks; ++p) {
            double* dst = x.inbox[p] + peer_slot_index(x, kind, parity, x.rank);
            for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldcg(buf + i);
        }
        __threadfence_system();
        __syncthreads();
        peer_publish(x, kind, parity, 0, epoch, threadIdx.x);
        peer_wait(x, kind, parity, 0, epoch, threadIdx.x, error_flag);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            double s = 0.0;
            for (int p = 0; p < x.nranks; ++p) s += peer_inbox_value(x, kind, parity, p, i);
            buf[i] = s;
            if (i == n - 1 && loss_host) { *loss_host = s; __threadfence_system(); }
        }
    } else if (threadIdx.x == 0 && loss_host) {
        *loss_host = __ldcg(buf + n - 1);
        __threadfence_system();
    }
    if (threadIdx.x == 0) *done_counter = 0u;   // next launch
}

// grad_transform at N > 1 in fused steps, without NCCL: the split-K partial reduction of the GEMM is the producer --
// block b sums the partials of its chunk of gT (fixed order), stores the result into ITS slot of every rank's gT inbox
// (own included) with plain NVLink stores and publishes flag b -- and the projection update is the consumer:
// transform_update_kernel waits for the flags of the chunk an element lives in and sums the nranks slots in rank order
// (bit-identical T, b on every rank). chunk = elements per block (gridDim.x <= kPeerFlagStride).
__global__ void __launch_bounds__(256) gt_reduce_push_kernel(const PeerXchg* __restrict__ xp, const float* __restrict__ part,
                                                             int nparts, long n, int chunk, unsigned long long epoch) {
    const PeerXchg& x = *xp;
    const int parity = (int)(epoch & 1ull);
    const long e0 = (long)blockIdx.x * chunk, e1 = min(n, e0 + chunk);
    const size_t slot = ((size_t)parity * x.nranks + x.rank) * (size_t)x.gt_elems;
    for (long i = e0 + threadIdx.x; i < e1; i += blockDim.x) {
        float s = 0.f;
        int z = 0;
        for (; z + 8 <= nparts; z += 8) {   // same summation order as reduce_partials_kernel / transform_update_kernel
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(part + (long)(z + u) * n + i);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; z < nparts; ++z) s += __ldg(part + (long)z * n + i);
        for (int p = 0; p < x.nranks; ++p) x.gt_inbox[p][slot + i] = s;
    }
    __threadfence_system();
    __syncthreads();
    peer_publish(x, kPeerKindGt, parity, blockIdx.x, epoch, threadIdx.x);
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
