// kernels_p2p.cu -- the sweep's exchange step over NVLink peer memory (one process per GPU).
//
// The reference merges the per-thread +-1 deltas of all z-threads through one shared
// AtomicInteger[K][V] (topics/UncollapsedParallelLDA.java:102,363-368,1107-1221) and every Phi
// thread reads the merged counts; across GPUs that shared matrix becomes: every rank maps the other
// ranks' n_wk / Phi^T / segment sums with CUDA IPC (engine.cu, ldagpu_comm_init) and
//   * the Phi draw loads and sums the peers' partial counts itself (kernels_phi.cu, REDUCE),
//   * the segment sums and the normalised Phi rows are stored into every rank's buffers by the
//     kernels that produce them,
// ordered by epoch flags with release/acquire at system scope.  This file holds the small pieces
// around those kernels: pushing the per-rank topic totals, the stand-alone count reduce for the
// step-wise API (setZIndicators / ldagpu_rebuild_counts), and signal / wait.
#include "common.cuh"

namespace ldagpu {

// n_k of this rank's partial counts -> slot [rank] of every rank's nk_parts; then "my counts are complete".
// Runs after the z-step and topic_totals in stream order, so everything this rank contributed is in L2.
__global__ void __launch_bounds__(256) p2p_push_topic_totals_kernel(PeerTable pt, const int32_t *__restrict__ n_k_local,
                                                                     int32_t Ks, uint32_t epoch)
{
    for (int k = threadIdx.x; k < Ks; k += blockDim.x) {
        const int32_t v = n_k_local[k];
#pragma unroll
        for (int q = 0; q < P2P_MAX; ++q)
            if (q < pt.world) pt.nk_parts[q][(size_t)pt.rank * Ks + k] = v;
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < pt.world) p2p_signal_one(pt, P2P_FLAG_COUNTS, epoch, (int)threadIdx.x);
}

cudaError_t launch_p2p_push_topic_totals(const PeerTable &pt, const int32_t *n_k_local, int32_t Ks, uint32_t epoch,
                                         cudaStream_t st)
{
    p2p_push_topic_totals_kernel<<<1, 256, 0, st>>>(pt, n_k_local, Ks, epoch);
    return cudaGetLastError();
}

// Stand-alone reduce-scatter: global counts of the rank's rows [row0, row1) = sum of every rank's partial
// counts (int4 loads from the peers), written over the rank's own rows; n_k from the pushed parts.
__global__ void __launch_bounds__(256) p2p_reduce_counts_kernel(PeerTable pt, Dims dm, int32_t *n_k, int32_t row0,
                                                                 int32_t row1, uint32_t epoch)
{
    if (threadIdx.x == 0) p2p_wait_all(pt, P2P_FLAG_COUNTS, epoch);
    __syncthreads();
    const size_t first = (size_t)row0 * dm.Ks / 4, last = (size_t)row1 * dm.Ks / 4;   // Ks % 32 == 0
    int32_t *own = pt.n_wk[pt.rank];
    for (size_t i = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < last; i += (size_t)gridDim.x * blockDim.x) {
        int4 acc = make_int4(0, 0, 0, 0);
#pragma unroll
        for (int q = 0; q < P2P_MAX; ++q) {
            if (q < pt.world) {
                const int4 v = __ldcg(reinterpret_cast<const int4 *>(pt.n_wk[q]) + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        reinterpret_cast<int4 *>(own)[i] = acc;
    }
    if (blockIdx.x == 0) {
        const int32_t *parts = pt.nk_parts[pt.rank];
        for (int k = threadIdx.x; k < dm.Ks; k += blockDim.x) {
            int32_t t = 0;
            for (int q = 0; q < pt.world; ++q) t += __ldcg(parts + (size_t)q * dm.Ks + k);
            n_k[k] = t;
        }
    }
}

cudaError_t launch_p2p_reduce_counts(const PeerTable &pt, const Dims &dm, int32_t *n_k, int32_t row0, int32_t row1,
                                     uint32_t epoch, int sm_count, cudaStream_t st)
{
    p2p_reduce_counts_kernel<<<sm_count * 4, 256, 0, st>>>(pt, dm, n_k, row0, row1, epoch);
    return cudaGetLastError();
}

__global__ void p2p_signal_kernel(PeerTable pt, int kind, uint32_t epoch)
{
    __threadfence_system();
    if ((int)threadIdx.x < pt.world) p2p_signal_one(pt, kind, epoch, (int)threadIdx.x);
}
__global__ void p2p_wait_kernel(PeerTable pt, int kind, uint32_t epoch)
{
    if (threadIdx.x == 0) p2p_wait_all(pt, kind, epoch);
}

cudaError_t launch_p2p_signal(const PeerTable &pt, int kind, uint32_t epoch, cudaStream_t st)
{
    p2p_signal_kernel<<<1, 32, 0, st>>>(pt, kind, epoch);
    return cudaGetLastError();
}
cudaError_t launch_p2p_wait(const PeerTable &pt, int kind, uint32_t epoch, cudaStream_t st)
{
    p2p_wait_kernel<<<1, 32, 0, st>>>(pt, kind, epoch);
    return cudaGetLastError();
}

}  // namespace ldagpu
