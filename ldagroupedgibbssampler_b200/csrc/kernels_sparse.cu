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
// Alias tables.  The reference's stack algorithm is order dependent, so the pairing loop stays
// sequential per word type; types are independent.  A warp takes 32 types at a time:
//   phase A (cooperative, coalesced): for each of the 32 types the lanes stream the Phi^T row, form the
//            normaliser (lane-strided fp64 partial sums + xor butterfly), classify every topic as "low"
//            (b < 0) or "high" and compact the indices, in topic order, into the type's two stacks
//            (ballot + popc prefix) -- the same stacks the sequential classification would build;
//   phase B (one lane per type): the sequential pairing loop on the type's private scratch
//            (b[K] fp64, one int stack[K]: "low" from the front, "high" from the back; both are popped
//            in decreasing order, so the loop walks its own 12 K bytes with sector reuse).
// Only the types that occur in this rank's tokens get a table (no token ever reads the others).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
alias_build_kernel(Dims dm, const float *__restrict__ alpha, const float *__restrict__ phiT,
                   AliasSlot *__restrict__ table, float *__restrict__ type_norm,
                   double *__restrict__ bs_all, int32_t *__restrict__ stack_all,
                   const int32_t *__restrict__ active, int32_t n_active)
{
    const int lane = threadIdx.x & 31;
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t warp_first = tid - lane;          // scratch slot of lane 0 of this warp
    const int K = dm.K;
    const double k1 = 1.0 / (double)K;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int64_t base = warp_first; base < n_active; base += T) {
        int mylow = 0, myhigh = 0;
        // ---- phase A
        for (int j = 0; j < 32; ++j) {
            if (base + j >= n_active) break;
            const int64_t w = active[base + j];
            const float *ph = phiT + (size_t)w * dm.Ks;
            AliasSlot *tw = table + (size_t)w * dm.Ks;
            double *bs = bs_all + (size_t)(warp_first + j) * K;
            int32_t *stack = stack_all + (size_t)(warp_first + j) * K;
            double acc = 0.0;
            for (int k = lane; k < K; k += 32) acc = __dadd_rn(acc, (double)__fmul_rn(alpha[k], ph[k]));
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(FULL, acc, off));
            const double norm = acc;
            if (lane == 0) type_norm[w] = __double2float_rn(norm);
            int low = 0, high = 0;
            for (int c0 = 0; c0 < K; c0 += 32) {
                const int i = c0 + lane;
                const bool valid = i < K;
                double b = 0.0;
                if (valid) {
                    b = __dsub_rn(__ddiv_rn((double)__fmul_rn(alpha[i], ph[i]), norm), k1);
                    bs[i] = b;
                    tw[i] = AliasSlot{0.0f, i};
                }
                const bool is_low = valid && b < 0.0;
                const unsigned lm = __ballot_sync(FULL, is_low), hm = __ballot_sync(FULL, valid && !is_low);
                if (is_low) stack[low + __popc(lm & lt_mask)] = i;
                else if (valid) stack[K - 1 - (high + __popc(hm & lt_mask))] = i;
                low += __popc(lm);
                high += __popc(hm);
            }
            if (lane == j) { mylow = low; myhigh = high; }
        }
        __syncwarp();
        // ---- phase B
        if (base + lane < n_active) {
            const int64_t w = active[base + lane];
            AliasSlot *tw = table + (size_t)w * dm.Ks;
            double *bs = bs_all + (size_t)tid * K;
            int32_t *stack = stack_all + (size_t)tid * K;
            // low stack: stack[0..low), high stack: stack[K-high..K) (top = K-high).  The reference loop
            //   while (low>0 && high>0) { l = pop low; h = top high; b[h] += b[l]; b[l] = 0;
            //                             if (b[h] <= 0) pop high; if (b[h] < 0) push h on low; a[l] = h; ps[l] = 1 + K c }
            // with the two values it keeps re-reading held in registers: the residual of the current top
            // high, and a high that just turned low (it is pushed on top of the low stack, so it is the
            // next one popped).  b[] is then read once per topic and never written back.
            int low = mylow, high = myhigh, h = 0, pl = 0;
            double d = 0.0, pc = 0.0;
            bool have_h = false, pending = false;
            while ((pending || low > 0) && high > 0) {
                int l;
                double c;
                if (pending) { l = pl; c = pc; pending = false; }
                else { l = stack[--low]; c = bs[l]; }
                if (!have_h) { h = stack[K - high]; d = bs[h]; have_h = true; }
                const double nb = __dadd_rn(c, d);
                d = nb;
                if (nb <= 0.0) { high--; have_h = false; }
                if (nb < 0.0) { pending = true; pl = h; pc = nb; }
                tw[l] = AliasSlot{__double2float_rn(__dadd_rn(1.0, __dmul_rn((double)K, c))), h};
            }
        }
        __syncwarp();
    }
}

int64_t alias_scratch_threads(const Dims &dm, int sm_count)
{
#ifndef ALIAS_TPS
#define ALIAS_TPS 512
#endif
    int64_t t = (int64_t)sm_count * ALIAS_TPS;
    int64_t need = ((int64_t)dm.V + 127) / 128 * 128;
    return need < t ? need : t;
}

cudaError_t launch_alias_build(const Dims &dm, const float *alpha, const float *phiT, AliasSlot *table,
                               float *type_norm, double *bs_scratch, int32_t *stack_scratch,
                               const int32_t *active, int32_t n_active, int sm_count, cudaStream_t st)
{
    if (n_active == 0) return cudaSuccess;
    const int64_t T = alias_scratch_threads(dm, sm_count);
    // balanced rounds: when the types do not fit one wave of T threads, split them evenly over the rounds
    const int64_t rounds = ((int64_t)n_active + T - 1) / T;
    int64_t blocks = (((int64_t)n_active + rounds - 1) / rounds + 127) / 128;
    if (blocks > T / 128) blocks = T / 128;
    alias_build_kernel<<<(unsigned)blocks, 128, 0, st>>>(dm, alpha, phiT, table, type_norm, bs_scratch,
                                                       stack_scratch, active, n_active);
    return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------
// Sparse z-step.  One warp per document (PCGS is sequential inside a document).  The document's
// non-zero topics live in a per-warp shared-memory list in the reference's order (append on first
// use, swap-remove when a count reaches 0), with their counts beside them.
// ---------------------------------------------------------------------------------------
struct SparseArgs {
    ZArgs z;
    const AliasSlot *table;
    const float *type_norm;
    int *lists;   // per resident warp: nz[cap], cnt[cap], cum[cap] in global memory (L1/L2 resident)
    int cap;      // list capacity per warp (multiple of 32)
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
    const ZArgs &a = sa.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = a.dm.K, Ks = a.dm.Ks, cap = sa.cap;
    // the lists are private to the warp; __syncwarp() orders lane 0's updates before the other lanes' reads
    int *nz = sa.lists + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * (size_t)cap * 3;
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
                for (int g0 = 0; g0 < nnz; g0 += 128) {
                    // four chunks at a time: all gathers are issued before the first scan needs one
                    float xs[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int i = g0 + 32 * c + lane;
                        xs[c] = i < nnz ? __fmul_rn(__int2float_rn(cnt[i]), __ldg(ph + nz[i])) : 0.0f;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int c0 = g0 + 32 * c;
                        if (c0 < nnz) {
                            const int i = c0 + lane;
                            float x = xs[c];
#pragma unroll
                            for (int off = 1; off < 32; off <<= 1) {
                                const float y = __shfl_up_sync(FULL, x, off);
                                if (lane >= off) x = __fadd_rn(x, y);
                            }
                            const float cv = c0 == 0 ? x : __fadd_rn(carry, x);
                            if (i < nnz) cum[i] = cv;
                            carry = __shfl_sync(FULL, cv, 31);
                        }
                    }
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
                    const int2 slot = __ldg(reinterpret_cast<const int2 *>(sa.table + cell));   // {ps, alias}
                    if (__fsub_rn(ups, __int2float_rn(i)) > __int_as_float(slot.x)) i = slot.y;
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

size_t spalias_list_bytes(const Dims &dm, int max_doc_len, int sm_count)
{
    int cap = max_doc_len < dm.K ? max_doc_len : dm.K;
    cap = (cap + 32 + 31) / 32 * 32;
    return (size_t)sm_count * 8 * 8 * (size_t)cap * 12;   // up to 8 CTAs of 8 warps per SM
}

cudaError_t launch_z_spalias(const ZArgs &z, const AliasSlot *table, const float *type_norm,
                             int *lists, int max_doc_len, int sm_count, cudaStream_t st)
{
    if (z.n_items == 0) return cudaSuccess;
    SparseArgs sa;
    sa.z = z; sa.table = table; sa.type_norm = type_norm; sa.lists = lists;
    int cap = max_doc_len < z.dm.K ? max_doc_len : z.dm.K;
    sa.cap = (cap + 32 + 31) / 32 * 32;
    int per_sm = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, z_spalias_kernel, 256, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int64_t grid = (int64_t)sm_count * per_sm;
    const int64_t need = (z.n_items + 7) / 8;
    if (need < grid) grid = need;
    z_spalias_kernel<<<(unsigned)grid, 256, 0, st>>>(sa);
    return cudaGetLastError();
}

}  // namespace ldagpu
