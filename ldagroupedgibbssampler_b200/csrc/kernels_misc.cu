// kernels_misc.cu -- K3 (count rebuild), K4 (log-likelihood, log-posterior) and accessor kernels.
//
// Replaces (reference, src/main/java/cc/mallet/topics/):
//   K3  UncollapsedParallelLDA.java:1547-1557 (+-1 deltas), :1107-1221 (updateCounts merge),
//       :1797-1830 (setZIndicators rebuild), :471-482 (updateTypeTopicCount)
//   K4  UncollapsedParallelLDA.java:1644-1758 (modelLogLikelihood), :1573-1634 (computeLogPosterior)
//   accessors  UncollapsedParallelLDA.java:226-234 (getTypeTopicCounts), ModifiedSimpleLDA.java:536-547
//       (getDocumentTopicMatrix), UncollapsedParallelLDA.java:1946-1966 (getPhi / getPhiMeans)
#include "common.cuh"
#include "contract_math.cuh"

namespace ldagpu {

constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------
// K3: n_wk[w][k] = #{i : w_i = w, z_i = k},  n_k = column sums.  Integer, exact.
// int4 loads of (w, z); equal neighbours are merged inside the thread (documents are bags of words:
// equal types are adjacent), one fire-and-forget reduction per distinct cell of the thread's four
// tokens.  Merging equal cells across the warp as well (match.any + redux) was measured slower on
// B200 (2.7 ms against 2.0 ms on the PubMed-shaped shard, profiles/README.md): the cells of 32 lanes
// rarely coincide and the scatter over a 578 MB matrix is bound by L2/DRAM sector traffic, not by the
// number of reductions.
// ---------------------------------------------------------------------------------------
constexpr int CNT_THREADS = 256;

// CHECK: z comes straight from the host (setZIndicators): an indicator outside [0, K) raises the error
// word and is skipped instead of being counted (UncollapsedParallelLDA.java:475-481 throws)
template <bool CHECK>
__global__ void __launch_bounds__(CNT_THREADS)
counts_kernel(Dims dm, const int32_t *__restrict__ tokens, const int32_t *__restrict__ z,
              int32_t *__restrict__ n_wk, int *__restrict__ bad)
{
    const int64_t n4 = dm.N / 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int4 w4 = __ldg(reinterpret_cast<const int4 *>(tokens) + i);
        const int4 z4 = __ldg(reinterpret_cast<const int4 *>(z) + i);
        size_t key[4];
        int c[4] = {1, 1, 1, 1};
        const int ww[4] = {w4.x, w4.y, w4.z, w4.w};
        const int zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
        for (int s = 0; s < 4; ++s) key[s] = (size_t)ww[s] * (size_t)dm.Ks + (size_t)tpos(dm, zz[s]);
        if (CHECK) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if ((unsigned)zz[s] >= (unsigned)dm.K) { c[s] = 0; key[s] = ~(size_t)0 - (size_t)s; atomicOr(bad, 2); }
        }
#pragma unroll
        for (int s = 3; s > 0; --s)
            if (key[s] == key[s - 1]) { c[s - 1] += c[s]; c[s] = 0; }
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (c[s] > 0) atomicAdd(&n_wk[key[s]], c[s]);
    }
    // scalar tail (N % 4 tokens), done by block 0
    if (blockIdx.x == 0)
        for (int64_t i = n4 * 4 + threadIdx.x; i < dm.N; i += blockDim.x) {
            if (CHECK && (unsigned)z[i] >= (unsigned)dm.K) { atomicOr(bad, 2); continue; }
            atomicAdd(&n_wk[(size_t)tokens[i] * dm.Ks + tpos(dm, z[i])], 1);
        }
}

// counts of tokens [0, n) of the given (offset) arrays added into n_wk, with the range check; no reset, no
// totals: ldagpu_set_z pipelines the host-to-device copy of z with this kernel chunk by chunk
cudaError_t launch_counts_chunk(const Dims &dm, const int32_t *tokens, const int32_t *z, int64_t n, int32_t *n_wk,
                                int *bad, int sm_count, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    Dims d = dm;
    d.N = n;
    int64_t need = (n / 4 + CNT_THREADS - 1) / CNT_THREADS;
    int64_t grid = (int64_t)sm_count * 8;
    if (need < grid) grid = need;
    if (grid < 1) grid = 1;
    counts_kernel<true><<<(unsigned)grid, CNT_THREADS, 0, st>>>(d, tokens, z, n_wk, bad);
    return cudaGetLastError();
}

// n_k = column sums of n_wk.  (A shared-memory histogram of z inside counts_kernel serialised on the
// popular topics: 21 short-scoreboard stall cycles per issue in the round-1 profile.)
__global__ void __launch_bounds__(128)
topic_totals_kernel(Dims dm, const int32_t *__restrict__ n_wk, int32_t *__restrict__ n_k)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= dm.K) return;
    int acc = 0;
    const int rows = (dm.V + gridDim.y - 1) / gridDim.y;
    const int w0 = blockIdx.y * rows, w1 = min(w0 + rows, dm.V);
    const int col = tpos(dm, k);   // n_k is in natural topic order, the rows are not (common.cuh)
#pragma unroll 8
    for (int w = w0; w < w1; ++w) acc += n_wk[(size_t)w * dm.Ks + col];
    if (acc) atomicAdd(&n_k[k], acc);
}

cudaError_t launch_topic_totals(const Dims &dm, const int32_t *n_wk, int32_t *n_k, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(n_k, 0, sizeof(int32_t) * (size_t)dm.Ks, st);
    if (e != cudaSuccess) return e;
    int gy = (dm.V + 63) / 64;   // ~64 rows per CTA, eight loads in flight per thread
    if (gy > 1024) gy = 1024;
    if (gy < 1) gy = 1;
    dim3 grid((dm.K + 127) / 128, gy);
    topic_totals_kernel<<<grid, 128, 0, st>>>(dm, n_wk, n_k);
    return cudaGetLastError();
}

cudaError_t launch_counts(const Dims &dm, const int32_t *tokens, const int32_t *z, int32_t *n_wk,
                          int32_t *n_k, int sm_count, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(n_wk, 0, sizeof(int32_t) * (size_t)dm.Vp * dm.Ks, st);
    if (e != cudaSuccess) return e;
    if (dm.N > 0) {
        int64_t need = (dm.N / 4 + CNT_THREADS - 1) / CNT_THREADS;
        int64_t grid = (int64_t)sm_count * 8;
        if (need < grid) grid = need;
        if (grid < 1) grid = 1;
        counts_kernel<false><<<(unsigned)grid, CNT_THREADS, 0, st>>>(dm, tokens, z, n_wk, nullptr);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return launch_topic_totals(dm, n_wk, n_k, st);
}

// getDocumentTopicMatrix: dense [D][K]
__global__ void doc_topic_kernel(Dims dm, const int64_t *__restrict__ doc_off, const int32_t *__restrict__ z,
                                 int32_t *__restrict__ n_dk)
{
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t d = wid; d < dm.D; d += nw)
        for (int64_t t = doc_off[d] + lane; t < doc_off[d + 1]; t += 32)
            atomicAdd(&n_dk[(size_t)d * dm.K + z[t]], 1);
}
cudaError_t launch_doc_topic_counts(const Dims &dm, const int64_t *doc_off, const int32_t *z,
                                    int32_t *n_dk, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(n_dk, 0, sizeof(int32_t) * (size_t)dm.D * dm.K, st);
    if (e != cudaSuccess || dm.D == 0) return e;
    int64_t grid = (dm.D + 7) / 8;
    if (grid > 65535 * 16) grid = 65535 * 16;
    doc_topic_kernel<<<(unsigned)grid, 256, 0, st>>>(dm, doc_off, z, n_dk);
    return cudaGetLastError();
}

// 16-bit transport of the topic indicators over PCIe (K <= 65 536): ldagpu_set_z16 / ldagpu_get_z16 /
// ldagpu_sweep_get_z16 move uint16 and convert on the device; the device copy stays int32
// `head` elements in front of the first 16-byte boundary of the int32 side (the two arrays are offset by the same number
// of elements, so their vector alignment coincides) and the tail go element by element.
__global__ void unpack16_kernel(const uint16_t *__restrict__ in, int32_t *__restrict__ out, int64_t n, int head)
{
    const int64_t n4 = (n - head) / 4, stride = (int64_t)gridDim.x * blockDim.x;
    const ushort4 *in4 = reinterpret_cast<const ushort4 *>(in + head);
    int4 *out4 = reinterpret_cast<int4 *>(out + head);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const ushort4 v = __ldg(in4 + i);
        out4[i] = make_int4(v.x, v.y, v.z, v.w);
    }
    if (blockIdx.x == 0) {
        if ((int)threadIdx.x < head) out[threadIdx.x] = in[threadIdx.x];
        for (int64_t i = head + n4 * 4 + threadIdx.x; i < n; i += blockDim.x) out[i] = in[i];
    }
}
__global__ void pack16_kernel(const int32_t *__restrict__ in, uint16_t *__restrict__ out, int64_t n, int head)
{
    const int64_t n4 = (n - head) / 4, stride = (int64_t)gridDim.x * blockDim.x;
    const int4 *in4 = reinterpret_cast<const int4 *>(in + head);
    ushort4 *out4 = reinterpret_cast<ushort4 *>(out + head);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int4 v = __ldg(in4 + i);
        out4[i] = make_ushort4((unsigned short)v.x, (unsigned short)v.y, (unsigned short)v.z, (unsigned short)v.w);
    }
    if (blockIdx.x == 0) {
        if ((int)threadIdx.x < head) out[threadIdx.x] = (uint16_t)in[threadIdx.x];
        for (int64_t i = head + n4 * 4 + threadIdx.x; i < n; i += blockDim.x) out[i] = (uint16_t)in[i];
    }
}
static unsigned copy_grid(int64_t n, int sm_count)
{
    int64_t need = (n / 4 + 255) / 256, grid = (int64_t)sm_count * 8;
    if (need < grid) grid = need;
    return (unsigned)(grid < 1 ? 1 : grid);
}
static int vec_head(const int32_t *p32, int64_t n)
{
    const int h = (int)((4 - (reinterpret_cast<uintptr_t>(p32) / sizeof(int32_t)) % 4) % 4);
    return (int64_t)h < n ? h : (int)n;
}
cudaError_t launch_unpack16(const uint16_t *in, int32_t *out, int64_t n, int sm_count, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    unpack16_kernel<<<copy_grid(n, sm_count), 256, 0, st>>>(in, out, n, vec_head(out, n));
    return cudaGetLastError();
}
cudaError_t launch_pack16(const int32_t *in, uint16_t *out, int64_t n, int sm_count, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    pack16_kernel<<<copy_grid(n, sm_count), 256, 0, st>>>(in, out, n, vec_head(in, n));
    return cudaGetLastError();
}

__global__ void validate_kernel(Dims dm, const int32_t *__restrict__ tokens, const int32_t *__restrict__ z, int *bad)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dm.N; i += (int64_t)gridDim.x * blockDim.x) {
        if (tokens && (tokens[i] < 0 || tokens[i] >= dm.V)) atomicOr(bad, 1);
        if (z && (z[i] < 0 || z[i] >= dm.K)) atomicOr(bad, 2);
    }
}
cudaError_t launch_validate(const Dims &dm, const int32_t *tokens, const int32_t *z, int *bad, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(bad, 0, sizeof(int), st);
    if (e != cudaSuccess || dm.N == 0) return e;
    int64_t grid = (dm.N + 255) / 256;
    if (grid > 4096) grid = 4096;
    validate_kernel<<<(unsigned)grid, 256, 0, st>>>(dm, tokens, z, bad);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// K4: log-likelihood (UncollapsedParallelLDA.java:1644-1758).
// logGammaStirling is MALLET 2.0.8's Dirichlet.logGammaStirling restated (the jar is not in the
// reference tree, pom.xml:130-141): shift up to >= 2, Stirling series, subtract the shifted logs.
// fp64 throughout, contract ln; per-block partial sums, combined by sum_partials_kernel.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double lgamma_stirling(double z)
{
    int shift = 0;
    while (z < 2.0) { z += 1.0; ++shift; }
    const double zi = 1.0 / z;
    const double z3 = zi * zi * zi;
    double r = 0x1.d67f1c864beb5p-1 /* ln(2 pi)/2 */ + (z - 0.5) * c_ln<double>(z) - z + zi / 12.0 - z3 / 360.0 +
               z3 * zi * zi / 1260.0;
    while (shift > 0) { --shift; z -= 1.0; r -= c_ln<double>(z); }
    return r;
}

constexpr int RED_THREADS = 256;

// deterministic block reduction; result valid in thread 0
__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double s_red[RED_THREADS / 32];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
    return t;
}

// lgS(alpha_k) for every topic, once per handle (alpha is fixed after create)
__global__ void lgs_table_kernel(int K, const double *__restrict__ alpha, double *__restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < K) out[k] = lgamma_stirling(alpha[k]);
}
cudaError_t launch_lgs_table(int K, const double *alpha, double *out, cudaStream_t st)
{
    lgs_table_kernel<<<(K + 127) / 128, 128, 0, st>>>(K, alpha, out);
    return cudaGetLastError();
}

// document part: sum_d [ sum_{k: n_dk>0} (lgS(alpha_k + n_dk) - lgS(alpha_k)) - lgS(alphaSum + N_d) ]
// One warp per document, work proportional to the document's length (not to K): pass 1 builds the
// shared-memory histogram of z; pass 2 walks the tokens again and the FIRST token of every topic
// (lowest lane of a match group, earliest 32-token block) takes the topic's count, clears it and adds
// the term, so the lgamma evaluations run with most lanes active.  Documents are assigned to warps by
// a fixed stride and the winner of a topic is fixed, so the sum is reproducible.
__global__ void __launch_bounds__(RED_THREADS)
ll_doc_kernel(Dims dm, const int64_t *__restrict__ doc_off, const int32_t *__restrict__ z,
              const double *__restrict__ alpha, const double *__restrict__ lgs_alpha, double alpha_sum,
              double *__restrict__ partials)
{
    extern __shared__ int32_t s_cnt[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int32_t *cnt = s_cnt + (size_t)warp * dm.Ks;
    for (int i = lane; i < dm.Ks; i += 32) cnt[i] = 0;
    __syncwarp();
    double acc = 0.0;
    for (int64_t d = (int64_t)blockIdx.x * nwarp + warp; d < dm.D; d += (int64_t)gridDim.x * nwarp) {
        const int64_t t0 = doc_off[d], t1 = doc_off[d + 1];
        for (int64_t t = t0 + lane; t < t1; t += 32) atomicAdd(&cnt[z[t]], 1);
        __syncwarp();
        for (int64_t tb = t0; tb < t1; tb += 32) {
            const bool valid = tb + lane < t1;
            const int k = valid ? z[tb + lane] : -1 - lane;
            const unsigned same = __match_any_sync(FULL, k);
            if (valid && lane == __ffs(same) - 1) {
                const int c = cnt[k];
                if (c > 0) {
                    cnt[k] = 0;
                    acc += lgamma_stirling(alpha[k] + (double)c) - lgs_alpha[k];
                }
            }
            __syncwarp();
        }
        if (lane == 0) acc -= lgamma_stirling(alpha_sum + (double)(t1 - t0));
    }
    double t = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

cudaError_t launch_ll_doc(const Dims &dm, const int64_t *doc_off, const int32_t *z,
                          const double *alpha, const double *lgs_alpha, double alpha_sum, double *partials,
                          int n_partials, int sm_count, cudaStream_t st)
{
    (void)sm_count;
    int warps = RED_THREADS / 32;
    while (warps > 1 && sizeof(int32_t) * (size_t)dm.Ks * warps > 200 * 1024) warps /= 2;
    size_t smem = sizeof(int32_t) * (size_t)dm.Ks * warps;
    cudaError_t e = kernel_config(reinterpret_cast<const void *>(ll_doc_kernel), warps * 32, smem, nullptr);
    if (e != cudaSuccess) return e;
    ll_doc_kernel<<<n_partials, warps * 32, smem, st>>>(dm, doc_off, z, alpha, lgs_alpha, alpha_sum, partials);
    return cudaGetLastError();
}

// type part over rows [row0,row1): sum_{n_wk>0} lgS(beta + n_wk), and the count of such cells
__global__ void __launch_bounds__(RED_THREADS)
ll_type_kernel(Dims dm, const int32_t *__restrict__ n_wk, double beta, int32_t row0, int32_t row1,
               double *__restrict__ partials)
{
    double acc = 0.0, nnz = 0.0;
    const int64_t cells = (int64_t)(row1 - row0) * dm.Ks;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t w = row0 + (int32_t)(i / dm.Ks);
        const int k = (int)(i % dm.Ks);   // column; padding columns belong to no topic
        if (w >= dm.V || ttopic(dm, k) >= dm.K) continue;
        int c = n_wk[(size_t)w * dm.Ks + k];
        if (c > 0) { acc += lgamma_stirling(beta + (double)c); nnz += 1.0; }
    }
    double t = block_sum(acc);
    double u = block_sum(nnz);
    if (threadIdx.x == 0) { partials[2 * blockIdx.x] = t; partials[2 * blockIdx.x + 1] = u; }
}
cudaError_t launch_ll_type(const Dims &dm, const int32_t *n_wk, double beta, int32_t row0,
                           int32_t row1, double *partials, int n_partials, cudaStream_t st)
{
    ll_type_kernel<<<n_partials, RED_THREADS, 0, st>>>(dm, n_wk, beta, row0, row1, partials);
    return cudaGetLastError();
}

// out[j] = sum_i partials[i*stride + j] for j < stride, sequential over i (deterministic)
__global__ void sum_partials_kernel(const double *__restrict__ partials, int n, int stride, double *out)
{
    int j = threadIdx.x;
    if (j >= stride) return;
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc += partials[(size_t)i * stride + j];
    out[j] = acc;
}
cudaError_t launch_sum_partials(const double *partials, int n, int stride, double *out, cudaStream_t st)
{
    sum_partials_kernel<<<1, 32, 0, st>>>(partials, n, stride, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// Sufficient statistics of the hyper-parameter optimisation (ModifiedSimpleLDA.java:812-905): how many
// (document, topic) pairs have n_dk = c, and how many (type, topic) cells have n_wk = c, c >= 1 (bin 0 is
// filled in by the host: all pairs minus the others).  Counts above the last bin land in the last bin.
// Equal bins of a warp are merged before the atomic (bins 1 and 2 would otherwise serialise).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void hist_add(unsigned long long *hist, int bin, bool active)
{
    const unsigned same = __match_any_sync(FULL, active ? bin : -1 - (int)(threadIdx.x & 31));
    if (active && (int)(threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&hist[bin], (unsigned long long)__popc(same));
}

__global__ void __launch_bounds__(RED_THREADS)
hist_doc_topic_kernel(Dims dm, const int64_t *__restrict__ doc_off, const int32_t *__restrict__ z,
                      unsigned long long *__restrict__ hist, int nbins)
{
    extern __shared__ int32_t s_cnt[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int32_t *cnt = s_cnt + (size_t)warp * dm.Ks;
    for (int i = lane; i < dm.Ks; i += 32) cnt[i] = 0;
    __syncwarp();
    for (int64_t d = (int64_t)blockIdx.x * nwarp + warp; d < dm.D; d += (int64_t)gridDim.x * nwarp) {
        const int64_t t0 = doc_off[d], t1 = doc_off[d + 1];
        for (int64_t t = t0 + lane; t < t1; t += 32) atomicAdd(&cnt[z[t]], 1);
        __syncwarp();
        for (int64_t tb = t0; tb < t1; tb += 32) {
            const bool valid = tb + lane < t1;
            const int k = valid ? z[tb + lane] : -1 - lane;
            const unsigned same = __match_any_sync(FULL, k);
            int c = 0;
            if (valid && lane == __ffs(same) - 1) { c = cnt[k]; cnt[k] = 0; }   // the first token of a topic takes its count
            __syncwarp();
            hist_add(hist, c < nbins ? c : nbins - 1, c > 0);
        }
    }
}

__global__ void __launch_bounds__(RED_THREADS)
hist_type_topic_kernel(Dims dm, const int32_t *__restrict__ n_wk, int32_t row0, int32_t row1,
                       unsigned long long *__restrict__ hist, int nbins)
{
    const int64_t cells = (int64_t)(row1 - row0) * dm.Ks;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (cells + stride - 1) / stride;
    for (int64_t r = 0; r < rounds; ++r) {
        const int64_t i = r * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        int c = 0;
        if (i < cells) {
            const int32_t w = row0 + (int32_t)(i / dm.Ks);
            const int k = (int)(i % dm.Ks);
            if (w < dm.V && ttopic(dm, k) < dm.K) c = n_wk[(size_t)w * dm.Ks + k];
        }
        hist_add(hist, c < nbins ? c : nbins - 1, c > 0);
    }
}

cudaError_t launch_count_histograms(const Dims &dm, const int64_t *doc_off, const int32_t *z, const int32_t *n_wk,
                                    int32_t row0, int32_t row1, unsigned long long *doc_hist, int n_doc_bins,
                                    unsigned long long *type_hist, int n_type_bins, cudaStream_t st)
{
    cudaError_t e;
    if (doc_hist && n_doc_bins > 0) {
        e = cudaMemsetAsync(doc_hist, 0, sizeof(unsigned long long) * (size_t)n_doc_bins, st);
        if (e != cudaSuccess) return e;
        int warps = RED_THREADS / 32;
        while (warps > 1 && sizeof(int32_t) * (size_t)dm.Ks * warps > 200 * 1024) warps /= 2;
        const size_t smem = sizeof(int32_t) * (size_t)dm.Ks * warps;
        e = kernel_config(reinterpret_cast<const void *>(hist_doc_topic_kernel), warps * 32, smem, nullptr);
        if (e != cudaSuccess) return e;
        if (dm.D > 0) hist_doc_topic_kernel<<<592, warps * 32, smem, st>>>(dm, doc_off, z, doc_hist, n_doc_bins);
    }
    if (type_hist && n_type_bins > 0) {
        e = cudaMemsetAsync(type_hist, 0, sizeof(unsigned long long) * (size_t)n_type_bins, st);
        if (e != cudaSuccess) return e;
        if (row1 > row0) hist_type_topic_kernel<<<592, RED_THREADS, 0, st>>>(dm, n_wk, row0, row1, type_hist, n_type_bins);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// log-posterior (UncollapsedParallelLDA.java:1573-1634):
//   sum_i ln(phi[z_i][w_i] + 1e-12) + sum_d sum_k (n_dk + alpha_k - 1) ln(theta_dk + 1e-12)
//   + (beta - 1) sum_kv ln(phi_kv + 1e-12)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS)
lp_tokens_kernel(Dims dm, const int32_t *__restrict__ tokens, const int32_t *__restrict__ z,
                 const float *__restrict__ phiT, double *__restrict__ partials)
{
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dm.N; i += (int64_t)gridDim.x * blockDim.x)
        acc += c_ln<double>((double)phiT[(size_t)tokens[i] * dm.Ks + tpos(dm, z[i])] + 1e-12);
    double t = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}
cudaError_t launch_lp_tokens(const Dims &dm, const int32_t *tokens, const int32_t *z,
                             const float *phiT, double *partials, int n_partials, cudaStream_t st)
{
    lp_tokens_kernel<<<n_partials, RED_THREADS, 0, st>>>(dm, tokens, z, phiT, partials);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(RED_THREADS)
lp_theta_kernel(Dims dm, const int64_t *__restrict__ doc_off, const int32_t *__restrict__ z,
                const float *__restrict__ theta, const double *__restrict__ alpha, double *__restrict__ partials)
{
    extern __shared__ int32_t s_cnt[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int32_t *cnt = s_cnt + (size_t)warp * dm.Ks;
    for (int i = lane; i < dm.Ks; i += 32) cnt[i] = 0;
    __syncwarp();
    double acc = 0.0;
    for (int64_t d = (int64_t)blockIdx.x * nwarp + warp; d < dm.D; d += (int64_t)gridDim.x * nwarp) {
        for (int64_t t = doc_off[d] + lane; t < doc_off[d + 1]; t += 32) atomicAdd(&cnt[z[t]], 1);
        __syncwarp();
        const float *trow = theta + (size_t)d * dm.Ks;
        for (int k = lane; k < dm.K; k += 32) {
            acc += ((double)cnt[k] + alpha[k] - 1.0) * c_ln<double>((double)trow[tpos(dm, k)] + 1e-12);
            cnt[k] = 0;
        }
        __syncwarp();
    }
    double t = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}
cudaError_t launch_lp_theta(const Dims &dm, const int64_t *doc_off, const int32_t *z,
                            const float *theta, const double *alpha, double *partials,
                            int n_partials, int sm_count, cudaStream_t st)
{
    (void)sm_count;
    int warps = RED_THREADS / 32;
    while (warps > 1 && sizeof(int32_t) * (size_t)dm.Ks * warps > 200 * 1024) warps /= 2;
    size_t smem = sizeof(int32_t) * (size_t)dm.Ks * warps;
    cudaError_t e = kernel_config(reinterpret_cast<const void *>(lp_theta_kernel), warps * 32, smem, nullptr);
    if (e != cudaSuccess) return e;
    lp_theta_kernel<<<n_partials, warps * 32, smem, st>>>(dm, doc_off, z, theta, alpha, partials);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(RED_THREADS)
lp_phi_kernel(Dims dm, const float *__restrict__ phiT, double beta, int32_t row0, int32_t row1,
              double *__restrict__ partials)
{
    double acc = 0.0;
    const int64_t cells = (int64_t)(row1 - row0) * dm.Ks;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t w = row0 + (int32_t)(i / dm.Ks);
        const int k = (int)(i % dm.Ks);
        if (w >= dm.V || ttopic(dm, k) >= dm.K) continue;
        acc += c_ln<double>((double)phiT[(size_t)w * dm.Ks + k] + 1e-12);
    }
    double t = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = (beta - 1.0) * t;
}
cudaError_t launch_lp_phi(const Dims &dm, const float *phiT, double beta, int32_t row0, int32_t row1,
                          double *partials, int n_partials, cudaStream_t st)
{
    lp_phi_kernel<<<n_partials, RED_THREADS, 0, st>>>(dm, phiT, beta, row0, row1, partials);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// accessors: layout conversion between the device layout and the reference's Java arrays
// ---------------------------------------------------------------------------------------
__global__ void export_phi_kernel(Dims dm, const float *__restrict__ phiT, double *__restrict__ out)
{
    // tile transpose [V][Ks] -> [K][V]
    __shared__ float tile[32][33];
    const int w0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int w = w0 + r, k = k0 + threadIdx.x;
        tile[r][threadIdx.x] = (w < dm.V && k < dm.K) ? phiT[(size_t)w * dm.Ks + tpos(dm, k)] : 0.0f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int k = k0 + r, w = w0 + threadIdx.x;
        if (k < dm.K && w < dm.V) out[(size_t)k * dm.V + w] = (double)tile[threadIdx.x][r];
    }
}
cudaError_t launch_export_phi(const Dims &dm, const float *phiT, double *phi_kv, cudaStream_t st)
{
    dim3 grid((dm.V + 31) / 32, (dm.K + 31) / 32), block(32, 8);
    export_phi_kernel<<<grid, block, 0, st>>>(dm, phiT, phi_kv);
    return cudaGetLastError();
}

__global__ void import_phi_kernel(Dims dm, const double *__restrict__ in, float *__restrict__ phiT)
{
    __shared__ float tile[32][33];
    const int w0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int k = k0 + r, w = w0 + threadIdx.x;
        tile[r][threadIdx.x] = (w < dm.V && k < dm.K) ? (float)in[(size_t)k * dm.V + w] : 0.0f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int w = w0 + r, k = k0 + threadIdx.x;
        if (w < dm.V && k < dm.K) phiT[(size_t)w * dm.Ks + tpos(dm, k)] = tile[threadIdx.x][r];
    }
}
cudaError_t launch_import_phi(const Dims &dm, const double *phi_kv, float *phiT, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(phiT, 0, sizeof(float) * (size_t)dm.Vp * dm.Ks, st);
    if (e != cudaSuccess) return e;
    dim3 grid((dm.V + 31) / 32, (dm.K + 31) / 32), block(32, 8);
    import_phi_kernel<<<grid, block, 0, st>>>(dm, phi_kv, phiT);
    return cudaGetLastError();
}

__global__ void export_mean_kernel(Dims dm, const double *__restrict__ sum_vk, double scale, double *__restrict__ out)
{
    __shared__ double tile[32][33];
    const int w0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int w = w0 + r, k = k0 + threadIdx.x;
        tile[r][threadIdx.x] = (w < dm.V && k < dm.K) ? sum_vk[(size_t)w * dm.Ks + tpos(dm, k)] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int k = k0 + r, w = w0 + threadIdx.x;
        if (k < dm.K && w < dm.V) out[(size_t)k * dm.V + w] = tile[threadIdx.x][r] * scale;
    }
}
cudaError_t launch_export_mean(const Dims &dm, const double *sum_vk, double scale, double *out_kv, cudaStream_t st)
{
    dim3 grid((dm.V + 31) / 32, (dm.K + 31) / 32), block(32, 8);
    export_mean_kernel<<<grid, block, 0, st>>>(dm, sum_vk, scale, out_kv);
    return cudaGetLastError();
}

__global__ void export_counts_kernel(Dims dm, const int32_t *__restrict__ n_wk, int32_t *__restrict__ out)
{
    const int64_t cells = (int64_t)dm.V * dm.K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = n_wk[(size_t)(i / dm.K) * dm.Ks + tpos(dm, (int)(i % dm.K))];
}
cudaError_t launch_export_counts(const Dims &dm, const int32_t *n_wk, int32_t *out_vk, cudaStream_t st)
{
    int64_t cells = (int64_t)dm.V * dm.K;
    int64_t grid = (cells + 255) / 256;
    if (grid > 65535) grid = 65535;
    export_counts_kernel<<<(unsigned)grid, 256, 0, st>>>(dm, n_wk, out_vk);
    return cudaGetLastError();
}

__global__ void export_theta_kernel(Dims dm, const float *__restrict__ theta, double *__restrict__ out)
{
    const int64_t cells = dm.D * dm.K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (double)theta[(size_t)(i / dm.K) * dm.Ks + tpos(dm, (int)(i % dm.K))];
}
cudaError_t launch_export_theta(const Dims &dm, const float *theta, double *out, cudaStream_t st)
{
    int64_t cells = dm.D * dm.K;
    if (cells == 0) return cudaSuccess;
    int64_t grid = (cells + 255) / 256;
    if (grid > 65535) grid = 65535;
    export_theta_kernel<<<(unsigned)grid, 256, 0, st>>>(dm, theta, out);
    return cudaGetLastError();
}
__global__ void import_theta_kernel(Dims dm, const double *__restrict__ in, float *__restrict__ theta)
{
    const int64_t cells = dm.D * dm.Ks;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t d = i / dm.Ks;
        const int k = ttopic(dm, (int)(i % dm.Ks));
        theta[i] = k < dm.K ? (float)in[(size_t)d * dm.K + k] : 0.0f;
    }
}
cudaError_t launch_import_theta(const Dims &dm, const double *in, float *theta, cudaStream_t st)
{
    int64_t cells = dm.D * dm.Ks;
    if (cells == 0) return cudaSuccess;
    int64_t grid = (cells + 255) / 256;
    if (grid > 65535) grid = 65535;
    import_theta_kernel<<<(unsigned)grid, 256, 0, st>>>(dm, in, theta);
    return cudaGetLastError();
}

}  // namespace ldagpu
