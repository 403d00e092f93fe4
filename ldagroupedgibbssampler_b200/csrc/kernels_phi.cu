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
//   phi_normalise S_k = ((s0+s1)+(s2+s3))+((s4+s5)+(s6+s7)); phi = (float)(g * (1 / S_k)), floored
// The fixed summation tree makes the result independent of the grid and of the number of GPUs
// (rank r owns segments [8r/G, 8(r+1)/G)).
#include "common.cuh"
#include "contract_math.cuh"

namespace ldagpu {

constexpr int PHI_THREADS = 128;
constexpr int MAX_GRID_Y = 65535;

// One CTA = 8 words x 128 topics.  Phase 1: every thread runs attempt 0 of its 8 cells in lock step
// for the ZERO-COUNT cells (shape = beta: the Marsaglia-Tsang constants are shared, as the reference's
// MarsagliaSparseDirichlet.java:20-29,37-38 precomputes them) and keeps what the squeeze accepts.
// Phase 2: the remaining cells (non-zero counts, squeeze failures) are compacted into a CTA-wide list
// and drained by all threads, each taking the next entry when its cell is accepted.  Cells are keyed
// by (w, k, attempt), so the evaluation order does not change any value; the column sums are taken
// afterwards in row order.
//
// REDUCE (multi-GPU, peer memory): the counts of a cell are the sum of every rank's partial counts,
// loaded straight from the peers' n_wk over NVLink while the CTA's previous/next neighbours are in
// their Gamma loops -- the reduce-scatter of the sweep happens inside this kernel.  The global counts
// are written back to the rank's own rows (accessors, log-likelihood), n_k comes from the per-rank
// totals the peers pushed before they signalled.
// POLYA (poisson_L > 0): the Poisson Polya-urn draw instead of the Gammas -- X = Poisson(beta + n_wk) per cell
// (types/PolyaUrnDirichletFixedCoeffPoisson.java:17-44, contract_math.cuh c_poisson), an integer, so most cells
// of a zero count are exactly 0 and the rows of Phi are sparse.  Same two phases: a zero-count cell is settled by
// one comparison (u <= exp(-beta)); the rest go through the list.
template <bool REDUCE, bool POLYA>
__global__ void __launch_bounds__(PHI_THREADS)
phi_draw_kernel(PeerTable pt, uint32_t epoch_counts, Dims dm, int32_t *n_wk, int32_t *n_k,
                double beta, float *__restrict__ phiT, double *__restrict__ partial, int32_t row0, uint32_t seed_lo,
                uint32_t seed_hi, uint32_t sweep, int32_t poisson_L)
{
    __shared__ int32_t s_n[PHI_ROW_BLOCK][PHI_THREADS];
    __shared__ float s_g[PHI_ROW_BLOCK][PHI_THREADS];
    __shared__ unsigned short s_list[PHI_ROW_BLOCK * PHI_THREADS];
    __shared__ int s_count, s_next;
    const int tid = threadIdx.x;
    // topic chunks vary fastest over the grid: the CTAs in flight together cover whole rows, so local and
    // peer accesses walk Phi^T / n_wk contiguously
    const int k = blockIdx.x * PHI_THREADS + tid;          // column of the rows (common.cuh: tpos / ttopic)
    const int kt = ttopic(dm, k);                          // its topic: Philox counters and n_k are keyed by topic
    const int32_t wb = row0 + blockIdx.y * PHI_ROW_BLOCK;
    const bool col_ok = k < dm.Ks;
    if (tid == 0) { s_count = 0; s_next = PHI_THREADS; }
    if (REDUCE) {
        if (tid == 0) p2p_wait_all(pt, P2P_FLAG_COUNTS, epoch_counts);   // every rank's z-step has finished
        __syncthreads();
        int32_t acc[PHI_ROW_BLOCK];
#pragma unroll
        for (int r = 0; r < PHI_ROW_BLOCK; ++r) acc[r] = 0;
#pragma unroll
        for (int q = 0; q < P2P_MAX; ++q) {
            if (q < pt.world) {
                const int32_t *src = pt.n_wk[q];
#pragma unroll
                for (int r = 0; r < PHI_ROW_BLOCK; ++r) {
                    const int32_t w = wb + r;
                    if (col_ok && w < dm.V) acc[r] += __ldcg(src + (size_t)w * dm.Ks + k);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < PHI_ROW_BLOCK; ++r) {
            const int32_t w = wb + r;
            s_n[r][tid] = acc[r];
            s_g[r][tid] = 0.0f;
            if (col_ok && w < dm.V) n_wk[(size_t)w * dm.Ks + k] = acc[r];
        }
        if (blockIdx.y == 0 && k < dm.K) {
            const int32_t *parts = pt.nk_parts[pt.rank];
            int32_t t = 0;
            for (int q = 0; q < pt.world; ++q) t += __ldcg(parts + (size_t)q * dm.Ks + k);
            n_k[k] = t;
        }
    } else {
        // stage the 8 counts of this column with coalesced loads
#pragma unroll
        for (int r = 0; r < PHI_ROW_BLOCK; ++r) {
            int32_t w = wb + r;
            s_n[r][tid] = (col_ok && w < dm.V) ? n_wk[(size_t)w * dm.Ks + k] : 0;
            s_g[r][tid] = 0.0f;
        }
    }
    __syncthreads();
    bool boost0 = false;
    double d0 = 0.0, c0 = 0.0, inva0 = 0.0, p0 = 0.0;
    if (POLYA) p0 = c_exp_neg<double>(-beta);
    else gamma_setup<double>(__dadd_rn(beta, 0.0), boost0, d0, c0, inva0);
    // ---- phase 1
    unsigned pend = 0;
    if (col_ok && kt < dm.K) {
#pragma unroll 1
        for (int r = 0; r < PHI_ROW_BLOCK; ++r) {
            const int32_t w = wb + r;
            if (w >= dm.V) break;   // padding rows stay zero
            bool done = false;
            if (s_n[r][tid] == 0) {
                const unsigned long long cell = (unsigned long long)w * (unsigned long long)dm.K + (unsigned long long)kt;
                if (POLYA) {
                    uint4 rnd = philox4x32_10((uint32_t)cell, (uint32_t)(cell >> 32), sweep, STREAM_POISSON << 24, seed_lo, seed_hi);
                    done = !(uniform52(rnd.x, rnd.y) > p0);   // k = 0: the cell stays exactly zero
                } else {
                    uint4 rnd = philox4x32_10((uint32_t)cell, (uint32_t)(cell >> 32), sweep, STREAM_PHI << 24, seed_lo, seed_hi);
                    double g;
                    done = gamma_attempt_squeeze<double>(boost0, d0, c0, inva0, rnd, g);
                    if (done) s_g[r][tid] = __double2float_rn(g);
                }
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
                                        (unsigned long long)ttopic(dm, blockIdx.x * PHI_THREADS + col);
        if (POLYA) {
            uint4 rnd = philox4x32_10((uint32_t)cell, (uint32_t)(cell >> 32), sweep, STREAM_POISSON << 24, seed_lo, seed_hi);
            s_g[r][col] = __int2float_rn(c_poisson(beta, s_n[r][col], poisson_L, p0, rnd));
        } else {
            const double g = c_gamma<double>(__dadd_rn(beta, __int2double_rn(s_n[r][col])), seed_lo, seed_hi, cell,
                                             sweep, STREAM_PHI);
            s_g[r][col] = __double2float_rn(g);
        }
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
                            uint32_t seed_hi, uint32_t sweep, int32_t poisson_L, cudaStream_t st)
{
    // grid.y is limited to 65535: very large vocabularies go in slabs of rows
    for (int32_t r0 = row0; r0 < row1; r0 += MAX_GRID_Y * PHI_ROW_BLOCK) {
        const int32_t r1 = r0 + MAX_GRID_Y * PHI_ROW_BLOCK < row1 ? r0 + MAX_GRID_Y * PHI_ROW_BLOCK : row1;
        dim3 grid((dm.Ks + PHI_THREADS - 1) / PHI_THREADS, (r1 - r0) / PHI_ROW_BLOCK);
        if (poisson_L > 0)
            phi_draw_kernel<false, true><<<grid, PHI_THREADS, 0, st>>>(PeerTable{}, 0u, dm, const_cast<int32_t *>(n_wk), nullptr,
                                                                       beta, phiT, partial, r0, seed_lo, seed_hi, sweep, poisson_L);
        else
            phi_draw_kernel<false, false><<<grid, PHI_THREADS, 0, st>>>(PeerTable{}, 0u, dm, const_cast<int32_t *>(n_wk), nullptr,
                                                                        beta, phiT, partial, r0, seed_lo, seed_hi, sweep, 0);
    }
    return cudaGetLastError();
}

cudaError_t launch_phi_draw_p2p(const PeerTable &pt, bool reduce_counts, uint32_t epoch_counts, const Dims &dm,
                                int32_t *n_wk, int32_t *n_k, double beta, float *phiT, double *partial, int32_t row0,
                                int32_t row1, uint32_t seed_lo, uint32_t seed_hi, uint32_t sweep, int32_t poisson_L,
                                cudaStream_t st)
{
    if (!reduce_counts) return launch_phi_draw(dm, n_wk, beta, phiT, partial, row0, row1, seed_lo, seed_hi, sweep, poisson_L, st);
    for (int32_t r0 = row0; r0 < row1; r0 += MAX_GRID_Y * PHI_ROW_BLOCK) {
        const int32_t r1 = r0 + MAX_GRID_Y * PHI_ROW_BLOCK < row1 ? r0 + MAX_GRID_Y * PHI_ROW_BLOCK : row1;
        dim3 grid((dm.Ks + PHI_THREADS - 1) / PHI_THREADS, (r1 - r0) / PHI_ROW_BLOCK);
        if (poisson_L > 0)
            phi_draw_kernel<true, true><<<grid, PHI_THREADS, 0, st>>>(pt, epoch_counts, dm, n_wk, n_k, beta, phiT, partial, r0,
                                                                      seed_lo, seed_hi, sweep, poisson_L);
        else
            phi_draw_kernel<true, false><<<grid, PHI_THREADS, 0, st>>>(pt, epoch_counts, dm, n_wk, n_k, beta, phiT, partial, r0,
                                                                       seed_lo, seed_hi, sweep, 0);
    }
    return cudaGetLastError();
}

// P2P: the rank's segment sums are stored into every rank's seg buffer (the all-gather of the 8 x Ks
// doubles), and the last CTA publishes "segments of rank r are in place".
template <bool P2P>
__global__ void __launch_bounds__(PHI_THREADS)
phi_segment_kernel(PeerTable pt, uint32_t epoch_seg, Dims dm, const double *__restrict__ partial,
                   double *__restrict__ seg, int seg0)
{
    const int k = blockIdx.x * PHI_THREADS + threadIdx.x;
    const int s = seg0 + blockIdx.y;
    if (k < dm.Ks) {
        const int blocks_per_seg = dm.Vp / PHI_SEGMENTS / PHI_ROW_BLOCK;
        const double *p = partial + (size_t)s * blocks_per_seg * dm.Ks + k;
        // sequential sum (contract 4.4); eight loads in flight per step so the chain is not latency bound
        double acc = 0.0;
        int b = 0;
        for (; b + 8 <= blocks_per_seg; b += 8) {
            double v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = p[(size_t)(b + i) * dm.Ks];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc = __dadd_rn(acc, v[i]);
        }
        for (; b < blocks_per_seg; ++b) acc = __dadd_rn(acc, p[(size_t)b * dm.Ks]);
        if (P2P) {
#pragma unroll
            for (int q = 0; q < P2P_MAX; ++q)
                if (q < pt.world) pt.seg[q][(size_t)s * dm.Ks + k] = acc;
        } else {
            seg[(size_t)s * dm.Ks + k] = acc;
        }
    }
    if (P2P) p2p_cta_done_signal(pt, P2P_FLAG_SEG, epoch_seg, gridDim.x * gridDim.y);
}

cudaError_t launch_phi_segment_sums(const Dims &dm, const double *partial, double *seg, int seg0,
                                    int seg1, cudaStream_t st)
{
    if (seg1 <= seg0) return cudaSuccess;
    dim3 grid((dm.Ks + PHI_THREADS - 1) / PHI_THREADS, seg1 - seg0);
    phi_segment_kernel<false><<<grid, PHI_THREADS, 0, st>>>(PeerTable{}, 0u, dm, partial, seg, seg0);
    return cudaGetLastError();
}

cudaError_t launch_phi_segment_sums_p2p(const PeerTable &pt, uint32_t epoch_seg, const Dims &dm, const double *partial,
                                        int seg0, int seg1, cudaStream_t st)
{
    if (seg1 <= seg0) return cudaErrorInvalidValue;   // every rank owns at least one segment
    dim3 grid((dm.Ks + PHI_THREADS - 1) / PHI_THREADS, seg1 - seg0);
    phi_segment_kernel<true><<<grid, PHI_THREADS, 0, st>>>(pt, epoch_seg, dm, partial, nullptr, seg0);
    return cudaGetLastError();
}

constexpr int NORM_ROWS = 16;

// P2P: waits for every rank's segment sums, and stores the normalised rows of the rank's vocabulary
// slice into EVERY rank's Phi^T (the all-gather of the sweep as NVLink stores from the producing
// kernel); the last CTA publishes "Phi rows of rank r are in place".
template <bool P2P>
__global__ void __launch_bounds__(PHI_THREADS)
phi_normalise_kernel(PeerTable pt, uint32_t epoch_seg, uint32_t epoch_phi, Dims dm, const double *__restrict__ seg,
                     double *__restrict__ topic_sum, float *__restrict__ phiT, double *__restrict__ mean_sum,
                     int32_t row0, int32_t row1, int keep_zeros)
{
    if (P2P) {
        if (threadIdx.x == 0) p2p_wait_all(pt, P2P_FLAG_SEG, epoch_seg);
        __syncthreads();
        seg = pt.seg[pt.rank];
        phiT = pt.phiT[pt.rank];
    }
    const int k = blockIdx.x * PHI_THREADS + threadIdx.x;   // column; column chunks fastest (see phi_draw_kernel)
    if (k < dm.Ks && ttopic(dm, k) < dm.K) {
        double s[PHI_SEGMENTS];
#pragma unroll
        for (int i = 0; i < PHI_SEGMENTS; ++i) s[i] = P2P ? __ldcg(seg + (size_t)i * dm.Ks + k) : seg[(size_t)i * dm.Ks + k];
        const double S = __dadd_rn(__dadd_rn(__dadd_rn(s[0], s[1]), __dadd_rn(s[2], s[3])),
                                   __dadd_rn(__dadd_rn(s[4], s[5]), __dadd_rn(s[6], s[7])));
        if (blockIdx.y == 0 && topic_sum) topic_sum[k] = S;
        const double invS = S != 0.0 ? __ddiv_rn(1.0, S) : 0.0;   // one reciprocal per topic, one product per cell
        const int32_t wa = row0 + blockIdx.y * NORM_ROWS;
        for (int32_t w = wa; w < wa + NORM_ROWS && w < row1 && w < dm.V; ++w) {
            const size_t idx = (size_t)w * dm.Ks + k;
            float v = phiT[idx];
            if (S != 0.0) {
                v = __double2float_rn(__dmul_rn((double)v, invS));
                // ParallelDirichlet.java:63-65 floors at Double.MIN_VALUE; "with the Poisson it is allowed to have 0's"
                // (PolyaUrnDirichletFixedCoeffPoisson.java:36-39)
                if (v <= 0.0f && !keep_zeros) v = 0x1p-149f;
                if (!P2P) phiT[idx] = v;
            }
            if (P2P) {
#pragma unroll
                for (int q = 0; q < P2P_MAX; ++q)
                    if (q < pt.world) pt.phiT[q][idx] = v;
            }
            if (mean_sum) mean_sum[idx] += (double)v;   // LDAGroupedGibbsSampler.java:193-197
        }
    }
    if (P2P) p2p_cta_done_signal(pt, P2P_FLAG_PHI, epoch_phi, gridDim.x * gridDim.y);
}

cudaError_t launch_phi_normalise(const Dims &dm, const double *seg, double *topic_sum, float *phiT,
                                 double *phi_mean_sum, int32_t row0, int32_t row1, int keep_zeros, cudaStream_t st)
{
    for (int32_t r0 = row0; r0 < row1; r0 += MAX_GRID_Y * NORM_ROWS) {
        const int32_t r1 = r0 + MAX_GRID_Y * NORM_ROWS < row1 ? r0 + MAX_GRID_Y * NORM_ROWS : row1;
        dim3 grid((dm.Ks + PHI_THREADS - 1) / PHI_THREADS, (r1 - r0 + NORM_ROWS - 1) / NORM_ROWS);
        phi_normalise_kernel<false><<<grid, PHI_THREADS, 0, st>>>(PeerTable{}, 0u, 0u, dm, seg, topic_sum, phiT,
                                                                  phi_mean_sum, r0, r1, keep_zeros);
    }
    return cudaGetLastError();
}

cudaError_t launch_phi_normalise_p2p(const PeerTable &pt, uint32_t epoch_seg, uint32_t epoch_phi, const Dims &dm,
                                     double *topic_sum, double *phi_mean_sum, int32_t row0, int32_t row1, int keep_zeros,
                                     cudaStream_t st)
{
    if (row1 <= row0) return cudaErrorInvalidValue;   // every rank owns rows
    if ((row1 - row0 + NORM_ROWS - 1) / NORM_ROWS > MAX_GRID_Y) return cudaErrorInvalidConfiguration;   // > 1M rows per rank
    dim3 grid((dm.Ks + PHI_THREADS - 1) / PHI_THREADS, (row1 - row0 + NORM_ROWS - 1) / NORM_ROWS);
    phi_normalise_kernel<true><<<grid, PHI_THREADS, 0, st>>>(pt, epoch_seg, epoch_phi, dm, nullptr, topic_sum, nullptr,
                                                             phi_mean_sum, row0, row1, keep_zeros);
    return cudaGetLastError();
}

}  // namespace ldagpu
