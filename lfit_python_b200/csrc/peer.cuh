// peer.cuh -- the all-gather of positions and log-probabilities over NVLink peer memory.
//
// north_star: "an NCCL all-gather over NVLink of positions and log-probs per stretch-move half-step"
// (replaces the multiprocessing.Pool result hand-back of /root/reference/mcmcfit.py:273-288).  The rows a
// rank has to hand out are a few hundred KB and follow a kernel that has just produced them, so the cost of a
// library collective is its launch and handshake latency, not bandwidth.  Here every rank owns a WINDOW in
// its HBM that its peers map (CUDA IPC); one kernel packs the rank's rows, stores them straight into its slot
// of every peer's window over NVLink / NVSwitch, and the last CTA to finish raises the rank's flag in every
// window and waits for the other ranks' flags in its own.  Two buffers alternate by step: a rank can be at
// most one exchange ahead of the slowest reader (it needs that reader's next flag to go further), so a buffer
// is never rewritten while a peer still reads it -- provided the reader's consumers run on the stream the
// exchange was launched on, before its next exchange.
#pragma once
#include <cuda_runtime.h>

namespace lfb {

constexpr int kMaxPeers = 16;
constexpr long long kPeerHeader = 256;  // bytes: flags[kMaxPeers] (u64), then the window's status word

struct PeerArgs {
    const double* a;  // rows x ca
    const double* b;  // rows x cb (may be NULL when cb == 0)
    int ca, cb;
    long long rows;
    unsigned char* win[kMaxPeers];  // every rank's window as mapped in this process (win[rank]: the local one)
    int rank, world, buf;
    long long slot_bytes;
    unsigned long long step;
    unsigned int* ticket;  // local: CTAs that have finished their stores
    long long spin_limit;  // clock64 ticks to wait for the peers before giving up
};

// out row = [a row | b row]; element i of the packed rows goes to the same place in every window
__global__ void __launch_bounds__(256) peer_allgather_kernel(const __grid_constant__ PeerArgs P)
{
    const int cols = P.ca + P.cb;
    const long long total = P.rows * cols;
    const long long off = kPeerHeader + ((long long)P.buf * P.world + P.rank) * P.slot_bytes;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / cols;
        const int j = (int)(i - row * cols);
        const double v = j < P.ca ? P.a[row * P.ca + j] : P.b[row * P.cb + (j - P.ca)];
        for (int p = 0; p < P.world; ++p) ((double*)(P.win[p] + off))[i] = v;
    }
    // this CTA's stores are visible system-wide before its ticket is
    __threadfence_system();
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(P.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    // the last CTA: all rows of this rank are out -- raise its flag in every window, wait for everybody's here
    if (threadIdx.x == 0) *P.ticket = 0u;
    __threadfence_system();
    if (threadIdx.x < P.world) {
        const int p = threadIdx.x;
        ((volatile unsigned long long*)P.win[p])[P.rank] = P.step;
        volatile unsigned long long* mine = (volatile unsigned long long*)P.win[P.rank];
        const long long t0 = clock64();
        while (mine[p] < P.step) {
            __nanosleep(64);
            if (clock64() - t0 > P.spin_limit) {
                atomicExch((unsigned int*)(P.win[P.rank] + kMaxPeers * 8), 1u);  // status: a peer never arrived
                break;
            }
        }
    }
    __threadfence_system();
}

}  // namespace lfb
