// kernels_sparse.cu -- the sparse PCGS z-step ("spalias") for large K, and its alias tables.
//
// Replaces (reference, src/main/java/cc/mallet/):
//   topics/SpaliasUncollapsedParallelLDA.java:39-60    per-type alias table over alpha_k * phi[k][w]
//   topics/SpaliasUncollapsedParallelLDA.java:124-245  token loop: p(k) = alpha_k phi_kw (alias draw)
//                                                      + n_dk phi_kw (sparse cumulative sum over n_dk > 0)
//   topics/SpaliasUncollapsedParallelLDA.java:262-312  sampleNewTopic, insert / remove of the non-zero list
//   util/OptimizedGentleAliasMethod.java:52-79,100-107 table construction and generateSample(u)
//
// Cost per token is O(nnz_d) gathered Phi entries instead of the dense step's K: 12 + ~32 nnz_d bytes
// (a 4-byte gather costs a 32-byte sector) against 4 K.  Arithmetic contract: DESIGN.md 4.6; the CPU
// oracle (oracle/lda_oracle_sparse.c) reproduces tables and z bit for bit.
#include "common.cuh"
#include "contract_math.cuh"

namespace ldagpu {

constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------
// Alias tables.  One thread per word type runs the reference's sequential stack algorithm
// (it is order dependent, so it stays sequential per type; types are independent).  Scratch per
// thread: b[K] fp64 and one int stack[K] holding the "low" stack from the front and the "high"
// stack from the back; scratch is interleaved across threads (element i of thread t at i*T + t)
// so the classification pass is coalesced.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
alias_build_kernel(Dims dm, const float *__restrict__ alpha, const float *__restrict__ phiT,
                   float *__restrict__ ps, int32_t *__restrict__ al, float *__restrict__ type_norm,
                   double *__restrict__ bs, int32_t *__restrict__ stack)
{
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int K = dm.K;
    const double k1 = 1.0 / (double)K;
    for (int64_t w = tid; w < dm.V; w += T) {
        const float *ph = phiT + (size_t)w * dm.Ks;
        float *pw = ps + (size_t)w * dm.Ks;
        int32_t *aw = al + (size_t)w * dm.Ks;
        double norm = 0.0;
        for (int k = 0; k < K; ++k) norm = __dadd_rn(norm, (double)__fmul_rn(alpha[k], ph[k]));
        type_norm[w] = __double2float_rn(norm);
        int low = 0, high = 0;   // low stack: stack[0..low), high stack: stack[K-high..K) (top = K-high)
        for (int i = 0; i < K; ++i) {
            aw[i] = i;
            pw[i] = 0.0f;
            const double b = __dsub_rn(__ddiv_rn((double)__fmul_rn(alpha[i], ph[i]), norm), k1);
            bs[(size_t)i * T + tid] = b;
            if (b < 0.0) stack[(size_t)(low++) * T + tid] = i;
            else stack[(size_t)(K - 1 - (high++)) * T + tid] = i;
        }
        while (low > 0 && high > 0) {
            const int l = stack[(size_t)(--low) * T + tid];
            const int h = stack[(size_t)(K - high) * T + tid];
            const double c = bs[(size_t)l * T + tid], d = bs[(size_t)h * T + tid];
            const double nb = __dadd_rn(c, d);
            bs[(size_t)l * T + tid] = 0.0;
            bs[(size_t)h * T + tid] = nb;
            if (nb <= 0.0) high--;
            if (nb < 0.0) stack[(size_t)(low++) * T + tid] = h;
            aw[l] = h;
            pw[l] = __double2float_rn(__dadd_rn(1.0, __dmul_rn((double)K, c)));
        }
    }
}

int64_t alias_scratch_threads(const Dims &dm, int sm_count)
{
    int64_t t = (int64_t)sm_count * 256;
    int64_t need = ((int64_t)dm.V + 127) / 128 * 128;
    return need < t ? need : t;
}

cudaError_t launch_alias_build(const Dims &dm, const float *alpha, const float *phiT, float *ps, int32_t *al,
                               float *type_norm, double *bs_scratch, int32_t *stack_scratch, int sm_count,
                               cudaStream_t st)
{
    const int64_t T = alias_scratch_threads(dm, sm_count);
    alias_build_kernel<<<(unsigned)(T / 128), 128, 0, st>>>(dm, alpha, phiT, ps, al, type_norm, bs_scratch, stack_scratch);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// Sparse z-step.  One warp per document (PCGS is sequential inside a document).  The document's
// non-zero topics live in a per-warp shared-memory list in the reference's order (append on first
// use, swap-remove when a count reaches 0), with their counts beside them.
// ---------------------------------------------------------------------------------------
struct SparseArgs {
    ZArgs z;
    const float *ps;
    const int32_t *al;
    const float *type_norm;
    int cap;   // list capacity per warp (multiple of 32)
};

__device__ __forceinline__ int list_find(const int *nz, int nnz, int k, int lane)
{
    for (int c0 = 0; c0 < nnz; c0 += 32) {
        const int i = c0 + lane;
        const unsigned m = __ballot_sync(FULL, i < nnz && nz[i] == k);
        if (m) return c0 + __ffs(m) - 1;
    }
    return -1;
}

__global__ void __launch_bounds__(256) z_spalias_kernel(SparseArgs sa)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ZArgs &a = sa.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = a.dm.K, Ks = a.dm.Ks, cap = sa.cap;
    int *nz = reinterpret_cast<int *>(smem_raw) + (size_t)warp * cap * 3;
    int *cnt = nz + cap;
    float *cum = reinterpret_cast<float *>(cnt + cap);

    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(a.work_counter, 1ull);
        item = __shfl_sync(FULL, item, 0);
        if ((int64_t)item >= a.n_items) break;
        const int64_t d = a.item_doc[item], t0 = a.doc_off[d], t1 = a.doc_off[d + 1];
        if (t0 == t1) continue;
        // non-zero topic list in first-occurrence order (SpaliasUncollapsedParallelLDA.java:147-153)
        int nnz = 0;
        for (int64_t tb = t0; tb < t1; tb += 32) {
            const int zz = (tb + lane < t1) ? a.z[tb + lane] : -1;
            const int nv = (int)((t1 - tb) < 32 ? (t1 - tb) : 32);
            for (int tt = 0; tt < nv; ++tt) {
                const int k = __shfl_sync(FULL, zz, tt);
                const int i = list_find(nz, nnz, k, lane);
                if (lane == 0) {
                    if (i < 0) { nz[nnz] = k; cnt[nnz] = 1; }
                    else cnt[i] += 1;
                }
                if (i < 0) ++nnz;
                __syncwarp();
            }
        }

        for (int64_t tb = t0; tb < t1; tb += 32) {
            const int64_t t = tb + lane;
            const bool valid = t < t1;
            const int nv = (int)((t1 - tb) < 32 ? (t1 - tb) : 32);
            const int w = valid ? a.tokens[t] : 0;
            const int zold = valid ? a.z[t] : 0;
            float U = 0.0f;
            if (valid) {
                unsigned long long gt = (unsigned long long)(a.dm.token_base + t);
                uint4 r = philox4x32_10((uint32_t)gt, (uint32_t)(gt >> 32), a.sweep, STREAM_Z << 24, a.seed_lo, a.seed_hi);
                U = uniform23(r.x);
            }
            int znew = 0;
            for (int tt = 0; tt < nv; ++tt) {
                const int wt = __shfl_sync(FULL, w, tt);
                const int old = __shfl_sync(FULL, zold, tt);
                const float u = __shfl_sync(FULL, U, tt);
                const float *ph = a.phiT + (size_t)wt * Ks;
                // remove the token from its topic (:159-166, :295-304)
                {
                    const int i = list_find(nz, nnz, old, lane);
                    bool emptied = false;
                    if (lane == 0) {
                        const int c = cnt[i] - 1;
                        cnt[i] = c;
                        if (c == 0) { nz[i] = nz[nnz - 1]; cnt[i] = cnt[nnz - 1]; emptied = true; }
                    }
                    emptied = __shfl_sync(FULL, emptied, 0);
                    if (emptied) --nnz;
                    __syncwarp();
                }
                // sparse cumulative sum over the list (:178-191), chunks of 32 with sequential carries
                float carry = 0.0f;
                for (int c0 = 0; c0 < nnz; c0 += 32) {
                    const int i = c0 + lane;
                    float x = i < nnz ? __fmul_rn(__int2float_rn(cnt[i]), __ldg(ph + nz[i])) : 0.0f;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const float y = __shfl_up_sync(FULL, x, off);
                        if (lane >= off) x = __fadd_rn(x, y);
                    }
                    const float cv = c0 == 0 ? x : __fadd_rn(carry, x);
                    if (i < nnz) cum[i] = cv;
                    carry = __shfl_sync(FULL, cv, 31);
                }
                __syncwarp();
                const float sum = carry;
                const float tn = __ldg(sa.type_norm + wt);
                const float tot = __fadd_rn(tn, sum);
                int nw;
                if (u < __fdiv_rn(tn, tot) || nnz == 0) {
                    // prior part: alias draw (:265-267, OptimizedGentleAliasMethod.java:100-107)
                    const float up = __fadd_rn(u, __fdiv_rn(__fmul_rn(sum, u), tn));
                    const float ups = __fmul_rn(up, __int2float_rn(K));
                    int i = __float2int_rz(ups);
                    if (i > K - 1) i = K - 1;
                    const size_t cell = (size_t)wt * Ks + i;
                    if (__fsub_rn(ups, __int2float_rn(i)) > __ldg(sa.ps + cell)) i = __ldg(sa.al + cell);
                    nw = i;
                } else {
                    // likelihood part: first slot with u*tot - tn <= cum (:269-275, findIdx :347-375)
                    const float ul = __fsub_rn(__fmul_rn(u, tot), tn);
                    int slot = nnz - 1;
                    for (int c0 = 0; c0 < nnz; c0 += 32) {
                        const int i = c0 + lane;
                        const unsigned m = __ballot_sync(FULL, i < nnz && ul <= cum[i]);
                        if (m) { slot = c0 + __ffs(m) - 1; break; }
                    }
                    nw = nz[slot];
                }
                if (lane == tt) znew = nw;
                // add the token under its new topic (:223-230, :306-312)
                {
                    const int i = list_find(nz, nnz, nw, lane);
                    if (lane == 0) {
                        if (i < 0) { nz[nnz] = nw; cnt[nnz] = 1; }
                        else cnt[i] += 1;
                    }
                    if (i < 0) ++nnz;
                    __syncwarp();
                }
            }
            if (valid) {
                a.z[t] = znew;
                if (a.n_wk_out) atomicAdd(&a.n_wk_out[(size_t)w * Ks + znew], 1);
            }
        }
    }
}

cudaError_t launch_z_spalias(const ZArgs &z, const float *ps, const int32_t *al, const float *type_norm,
                             int max_doc_len, int sm_count, cudaStream_t st)
{
    if (z.n_items == 0) return cudaSuccess;
    SparseArgs sa;
    sa.z = z; sa.ps = ps; sa.al = al; sa.type_norm = type_norm;
    int cap = max_doc_len < z.dm.K ? max_doc_len : z.dm.K;
    cap = (cap + 32 + 31) / 32 * 32;
    sa.cap = cap;
    const size_t per_warp = (size_t)cap * 12;
    int warps = (int)((size_t)(200 * 1024) / per_warp);
    if (warps < 1) return cudaErrorInvalidValue;   // a document with > ~17 000 distinct topics
    if (warps > 8) warps = 8;
    const size_t smem = warps * per_warp;
    cudaError_t e = cudaFuncSetAttribute(z_spalias_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, z_spalias_kernel, warps * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)sm_count * per_sm;
    const int64_t need = (z.n_items + warps - 1) / warps;
    if (need < grid) grid = need;
    z_spalias_kernel<<<(unsigned)grid, warps * 32, smem, st>>>(sa);
    return cudaGetLastError();
}

}  // namespace ldagpu
