// kernels_phi.cu -- K2: the per-topic Dirichlet draw of Phi over the vocabulary.
//
// Replaces (reference, src/main/java/cc/mallet/):
//   topics/LDAGroupedGibbsSampler.java:139-209, topics/LDAPartiallyCollapsedGibbsSampler.java:48-118
//     samplePhi/loopOverTopics: phi[k] ~ Dir(beta + n_.k)
//   topics/UncollapsedParallelLDA.java:1287-1294 + types/MarsagliaSparseDirichlet.java:31-55
//     the initial Phi of every scheme (same distribution)
//   types/ParallelDirichlet.java:46-70  Dirichlet = independent Gammas, normalise, floor
//   util/ParallelRandoms.java:60-70,148-159  Marsaglia-Tsang Gamma (contract_math.cuh)
//
// Layout: Phi^T [Vp][Ks] fp32 word-major (a token's K-vector is contiguous for the z-step), so a
// topic's Dirichlet normaliser is a column sum.  Three kernels:
//   phi_draw      g = Gamma(beta + n_wk) in fp64, stored rounded to fp32; fp64 partial column sums
//                 over blocks of 8 words (sequential within the block)
//   phi_segments  the blocks of each of the 8 vocabulary segments summed sequentially
//   phi_normalise S_k = ((s0+s1)+(s2+s3))+((s4+s5)+(s6+s7)); phi = (float)(g / S_k), floored
// The fixed summation tree makes the result independent of the grid and of the number of GPUs
// (rank r owns segments [8r/G, 8(r+1)/G)).
#include "common.cuh"
#include "contract_math.cuh"

namespace ldagpu {

constexpr int PHI_THREADS = 128;

// One CTA = 8 words x 128 topics.  Phase 1: every thread runs attempt 0 of its 8 cells in lock step
// for the ZERO-COUNT cells (shape = beta: the Marsaglia-Tsang constants are shared, as the reference's
// MarsagliaSparseDirichlet.java:20-29,37-38 precomputes them) and keeps what the squeeze accepts.
// Phase 2: the remaining cells (non-zero counts, squeeze failures) are compacted into a CTA-wide list
// and drained by all threads, each taking the next entry when its cell is accepted.  Cells are keyed
// by (w, k, attempt), so the evaluation order does not change any value; the column sums are taken
// afterwards in row order.
__global__ void __launch_bounds__(PHI_THREADS)
phi_draw_kernel(Dims dm, const int32_t *__restrict__ n_wk, double beta, float *__restrict__ phiT,
                double *__restrict__ partial, int32_t row0, uint32_t seed_lo, uint32_t seed_hi, uint32_t sweep)
{
    __shared__ int32_t s_n[PHI_ROW_BLOCK][PHI_THREADS];
    __shared__ float s_g[PHI_ROW_BLOCK][PHI_THREADS];
    __shared__ unsigned short s_list[PHI_ROW_BLOCK * PHI_THREADS];
    __shared__ int s_count, s_next;
    const int tid = threadIdx.x;
    const int k = blockIdx.y * PHI_THREADS + tid;
    const int32_t wb = row0 + blockIdx.x * PHI_ROW_BLOCK;
    const bool col_ok = k < dm.Ks;
    if (tid == 0) { s_count = 0; s_next = PHI_THREADS; }
    // stage the 8 counts of this column with coalesced loads
#pragma unroll
    for (int r = 0; r < PHI_ROW_BLOCK; ++r) {
        int32_t w = wb + r;
        s_n[r][tid] = (col_ok && w < dm.V) ? n_wk[(size_t)w * dm.Ks + k] : 0;
        s_g[r][tid] = 0.0f;
    }
    __syncthreads();
    bool boost0;
    double d0, c0, inva0;
    gamma_setup<double>(__dadd_rn(beta, 0.0), boost0, d0, c0, inva0);
    // ---- phase 1
    unsigned pend = 0;
    if (k < dm.K) {
#pragma unroll 1
        for (int r = 0; r < PHI_ROW_BLOCK; ++r) {
            const int32_t w = wb + r;
            if (w >= dm.V) break;   // padding rows stay zero
            bool done = false;
            if (s_n[r][tid] == 0) {
                const unsigned long long cell = (unsigned long long)w * (unsigned long long)dm.K + (unsigned long long)k;
                uint4 rnd = philox4x32_10((uint32_t)cell, (uint32_t)(cell >> 32), sweep, STREAM_PHI << 24, seed_lo, seed_hi);
                double g;
                done = gamma_attempt_squeeze<double>(boost0, d0, c0, inva0, rnd, g);
                if (done) s_g[r][tid] = __double2float_rn(g);
            }
            pend |= (done ? 0u : 1u) << r;
        }
    }
    // ---- phase 2
    {
        int pos = pend ? atomicAdd(&s_count, __popc(pend)) : 0;
        while (pend) {
            const int r = __ffs(pend) - 1;
            pend &= pend - 1;
            s_list[pos++] = (unsigned short)(r * PHI_THREADS + tid);
        }
    }
    __syncthreads();
    const int total = s_count;
    int idx = tid;
    while (idx < total) {
        const int e = s_list[idx];
        const int r = e / PHI_THREADS, col = e % PHI_THREADS;
        const unsigned long long cell = (unsigned long long)(wb + r) * (unsigned long long)dm.K +
                                        (unsigned long long)(blockIdx.y * PHI_THREADS + col);
        const double g = c_gamma<double>(__dadd_rn(beta, __int2double_rn(s_n[r][col])), seed_lo, seed_hi, cell,
                                         sweep, STREAM_PHI);
        s_g[r][col] = __double2float_rn(g);
        idx = atomicAdd(&s_next, 1);
    }
    __syncthreads();
    if (col_ok) {
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < PHI_ROW_BLOCK; ++r) {
            const float g32 = s_g[r][tid];
            acc = __dadd_rn(acc, (double)g32);
            phiT[(size_t)(wb + r) * dm.Ks + k] = g32;
        }
        partial[(size_t)(wb / PHI_ROW_BLOCK) * dm.Ks + k] = acc;
    }
}

cudaError_t launch_phi_draw(const Dims &dm, const int32_t *n_wk, double beta, float *phiT,
                            double *partial, int32_t row0, int32_t row1, uint32_t seed_lo,
                            uint32_t seed_hi, uint32_t sweep, cudaStream_t st)
{
    if (row1 <= row0) return cudaSuccess;
    dim3 grid((row1 - row0) / PHI_ROW_BLOCK, (dm.Ks + PHI_THREADS - 1) / PHI_THREADS);
    phi_draw_kernel<<<grid, PHI_THREADS, 0, st>>>(dm, n_wk, beta, phiT, partial, row0, seed_lo, seed_hi, sweep);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(PHI_THREADS)
phi_segment_kernel(Dims dm, const double *__restrict__ partial, double *__restrict__ seg, int seg0)
{
    const int k = blockIdx.x * PHI_THREADS + threadIdx.x;
    const int s = seg0 + blockIdx.y;
    if (k >= dm.Ks) return;
    const int blocks_per_seg = dm.Vp / PHI_SEGMENTS / PHI_ROW_BLOCK;
    const double *p = partial + (size_t)s * blocks_per_seg * dm.Ks + k;
    double acc = 0.0;
    for (int b = 0; b < blocks_per_seg; ++b) acc = __dadd_rn(acc, p[(size_t)b * dm.Ks]);
    seg[(size_t)s * dm.Ks + k] = acc;
}

cudaError_t launch_phi_segment_sums(const Dims &dm, const double *partial, double *seg, int seg0,
                                    int seg1, cudaStream_t st)
{
    if (seg1 <= seg0) return cudaSuccess;
    dim3 grid((dm.Ks + PHI_THREADS - 1) / PHI_THREADS, seg1 - seg0);
    phi_segment_kernel<<<grid, PHI_THREADS, 0, st>>>(dm, partial, seg, seg0);
    return cudaGetLastError();
}

constexpr int NORM_ROWS = 16;

__global__ void __launch_bounds__(PHI_THREADS)
phi_normalise_kernel(Dims dm, const double *__restrict__ seg, double *__restrict__ topic_sum,
                     float *__restrict__ phiT, double *__restrict__ mean_sum, int32_t row0, int32_t row1)
{
    const int k = blockIdx.y * PHI_THREADS + threadIdx.x;
    if (k >= dm.K) return;
    double s[PHI_SEGMENTS];
#pragma unroll
    for (int i = 0; i < PHI_SEGMENTS; ++i) s[i] = seg[(size_t)i * dm.Ks + k];
    const double S = __dadd_rn(__dadd_rn(__dadd_rn(s[0], s[1]), __dadd_rn(s[2], s[3])),
                               __dadd_rn(__dadd_rn(s[4], s[5]), __dadd_rn(s[6], s[7])));
    if (blockIdx.x == 0 && topic_sum) topic_sum[k] = S;
    const int32_t wa = row0 + blockIdx.x * NORM_ROWS;
    for (int32_t w = wa; w < wa + NORM_ROWS && w < row1 && w < dm.V; ++w) {
        const size_t idx = (size_t)w * dm.Ks + k;
        float v = phiT[idx];
        if (S != 0.0) {
            v = __double2float_rn(__ddiv_rn((double)v, S));
            if (v <= 0.0f) v = 0x1p-149f;   // ParallelDirichlet.java:63-65 floors at Double.MIN_VALUE
            phiT[idx] = v;
        }
        if (mean_sum) mean_sum[idx] += (double)v;   // LDAGroupedGibbsSampler.java:193-197
    }
}

cudaError_t launch_phi_normalise(const Dims &dm, const double *seg, double *topic_sum, float *phiT,
                                 double *phi_mean_sum, int32_t row0, int32_t row1, cudaStream_t st)
{
    if (row1 <= row0) return cudaSuccess;
    dim3 grid((row1 - row0 + NORM_ROWS - 1) / NORM_ROWS, (dm.K + PHI_THREADS - 1) / PHI_THREADS);
    phi_normalise_kernel<<<grid, PHI_THREADS, 0, st>>>(dm, seg, topic_sum, phiT, phi_mean_sum, row0, row1);
    return cudaGetLastError();
}

}  // namespace ldagpu
