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
                    // a type whose Phi column is all zero (possible with the Polya-urn Phi draw) has no prior part:
                    // its table is never consulted (tn = 0), keep the arithmetic finite
                    b = norm > 0.0 ? __dsub_rn(__ddiv_rn((double)__fmul_rn(alpha[i], ph[i]), norm), k1) : 0.0;
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
// non-zero topics and their counts form a list in the reference's order (append on first use,
// swap-remove when a count reaches 0).  The list is cut into blocks of 256 entries and lane l owns
// entries 8l..8l+7 of a block (two 16-byte loads per array): finding a topic, gathering Phi at the
// listed topics, the cumulative sum (lane-local prefix + one warp scan per block) and the search all
// work on those eight registers.
//   fast path  lists of up to 256 topics live in shared memory (2 KB per warp, 64 warps per SM);
//   slow path  a document whose list outgrows that continues on per-warp lists in global memory with
//              the same arithmetic block by block (early sweeps from a random start, very long documents).
// ---------------------------------------------------------------------------------------
constexpr int SP_BLOCK = 256;   // list entries per block (8 per lane)
constexpr int SP_WARPS = 8;     // warps per CTA

struct SparseArgs {
    ZArgs z;
    const AliasSlot *table;
    const float *type_norm;
    int *lists;   // per resident warp: nz[cap], cnt[cap], cum[cap] in global memory (slow path)
    int cap;      // list capacity per warp (multiple of SP_BLOCK)
};

__device__ __forceinline__ void load8(const int *a, int (&v)[8])
{
    const int4 x = reinterpret_cast<const int4 *>(a)[0], y = reinterpret_cast<const int4 *>(a)[1];
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
}

// position of topic k in the list, -1 if absent
template <bool MULTI>
__device__ __forceinline__ int list_find(const int *nz, int nnz, int k, int lane)
{
    for (int b0 = 0; b0 < (MULTI ? nnz : 1); b0 += SP_BLOCK) {
        int v[8];
        load8(nz + b0 + 8 * lane, v);
        const int nvalid = nnz - (b0 + 8 * lane);
        int pe = -1;
#pragma unroll
        for (int e = 7; e >= 0; --e)
            if (e < nvalid && v[e] == k) pe = e;
        const unsigned m = __ballot_sync(FULL, pe >= 0);
        if (m) {
            const int l = __ffs(m) - 1;
            return b0 + 8 * l + __shfl_sync(FULL, pe, l);
        }
    }
    return -1;
}

// add one token of topic k (SpaliasUncollapsedParallelLDA.java:223-230,306-312); pos = its list position when known
__device__ __forceinline__ void list_add_at(int *nz, int *cnt, int &nnz, int k, int pos, int lane)
{
    const int c = pos >= 0 ? cnt[pos] : 0;
    __syncwarp();
    if (lane == 0) {
        if (pos < 0) { nz[nnz] = k; cnt[nnz] = 1; }
        else cnt[pos] = c + 1;
    }
    if (pos < 0) ++nnz;
    __syncwarp();
}

// remove one token of topic `old` (:159-166, :295-304)
template <bool MULTI>
__device__ __forceinline__ void list_remove(int *nz, int *cnt, int &nnz, int old, int lane)
{
    const int pos = list_find<MULTI>(nz, nnz, old, lane);
    const int c = cnt[pos] - 1;
    const int lk = nz[nnz - 1], lc = cnt[nnz - 1];
    __syncwarp();
    if (lane == 0) {
        if (c == 0) { nz[pos] = lk; cnt[pos] = lc; }
        else cnt[pos] = c;
    }
    if (c == 0) --nnz;
    __syncwarp();
}

// One token: scores over the list, cumulative sum, prior/likelihood branch.  Returns the new topic and, in
// `slot`, its list position when the likelihood branch found it (-1 otherwise).
template <bool MULTI>
__device__ __forceinline__ int sparse_draw(const SparseArgs &sa, const int *nz, const int *cnt, float *cum, int nnz,
                                           int wt, float u, int lane, int &slot)
{
    const int K = sa.z.dm.K, Ks = sa.z.dm.Ks;
    const float *ph = sa.z.phiT + (size_t)wt * Ks;
    const float tn = __ldg(sa.type_norm + wt);
    float carry = 0.0f, E = 0.0f;
    float q[8];
    int nvalid0 = 0;
    for (int b0 = 0; b0 < (MULTI ? nnz : 1); b0 += SP_BLOCK) {
        int kk[8], cc[8];
        load8(nz + b0 + 8 * lane, kk);
        load8(cnt + b0 + 8 * lane, cc);
        const int nvalid = nnz - (b0 + 8 * lane);
        float s[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = e < nvalid ? __ldg(ph + kk[e]) : 0.0f;
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = e < nvalid ? __fmul_rn(__int2float_rn(cc[e]), s[e]) : 0.0f;
        q[0] = s[0];
#pragma unroll
        for (int e = 1; e < 8; ++e) q[e] = __fadd_rn(q[e - 1], s[e]);
        float inc = q[7];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const float y = __shfl_up_sync(FULL, inc, off);
            if (lane >= off) inc = __fadd_rn(inc, y);
        }
        E = __shfl_up_sync(FULL, inc, 1);
        if (lane == 0) E = 0.0f;
        if (MULTI) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (e < nvalid) cum[b0 + 8 * lane + e] = __fadd_rn(carry, __fadd_rn(E, q[e]));
        }
        carry = __fadd_rn(carry, __shfl_sync(FULL, inc, 31));
        nvalid0 = nvalid;
    }
    if (MULTI) __syncwarp();
    const float sum = carry;
    const float tot = __fadd_rn(tn, sum);
    slot = -1;
    if (!(tot > 0.0f)) {
        // neither the prior nor the document gives this type any mass (its Phi column is all zero over the topics
        // that matter): the reference draws the topic uniformly (topics/PolyaUrnSpaliasLDA.java:275-277)
        int i = __float2int_rz(__fmul_rn(u, __int2float_rn(K)));
        return i > K - 1 ? K - 1 : i;
    }
    if (u < __fdiv_rn(tn, tot) || nnz == 0) {
        // prior part: alias draw (:265-267, OptimizedGentleAliasMethod.java:100-107)
        const float up = __fadd_rn(u, __fdiv_rn(__fmul_rn(sum, u), tn));
        const float ups = __fmul_rn(up, __int2float_rn(K));
        int i = __float2int_rz(ups);
        if (i > K - 1) i = K - 1;
        const int2 entry = __ldg(reinterpret_cast<const int2 *>(sa.table + (size_t)wt * Ks + i));   // {ps, alias}
        if (__fsub_rn(ups, __int2float_rn(i)) > __int_as_float(entry.x)) i = entry.y;
        return i;
    }
    // likelihood part: first slot with u*tot - tn <= cum (:269-275, findIdx :347-375)
    const float ul = __fsub_rn(__fmul_rn(u, tot), tn);
    slot = nnz - 1;
    if (!MULTI) {
        int pe = -1;
#pragma unroll
        for (int e = 7; e >= 0; --e)
            if (e < nvalid0 && ul <= __fadd_rn(E, q[e])) pe = e;
        const unsigned m = __ballot_sync(FULL, pe >= 0);
        if (m) {
            const int l = __ffs(m) - 1;
            slot = 8 * l + __shfl_sync(FULL, pe, l);
        }
    } else {
        for (int b0 = 0; b0 < nnz; b0 += SP_BLOCK) {
            const int nvalid = nnz - (b0 + 8 * lane);
            int pe = -1;
#pragma unroll
            for (int e = 7; e >= 0; --e)
                if (e < nvalid && ul <= cum[b0 + 8 * lane + e]) pe = e;
            const unsigned m = __ballot_sync(FULL, pe >= 0);
            if (m) {
                const int l = __ffs(m) - 1;
                slot = b0 + 8 * l + __shfl_sync(FULL, pe, l);
                break;
            }
        }
    }
    return nz[slot];
}

// tokens [ta, t1) of one document on the given lists; MULTI = lists of any length (global memory)
template <bool MULTI>
__device__ __forceinline__ int64_t sparse_tokens(const SparseArgs &sa, int *nz, int *cnt, float *cum, int &nnz,
                                                 int64_t ta, int64_t t1, int lane)
{
    const ZArgs &a = sa.z;
    const int Ks = a.dm.Ks;
    for (int64_t tb = ta; tb < t1; tb += 32) {
        const int64_t t = tb + lane;
        const bool valid = t < t1;
        const int nv = (int)((t1 - tb) < 32 ? (t1 - tb) : 32);
        const int w = valid ? a.tokens[t] : 0;
        const int zold = valid ? a.z[t] : 0;
        float U = 0.0f;
        if (valid) {
            const unsigned long long gt = (unsigned long long)(a.dm.token_base + t);
            const uint4 r = philox4x32_10((uint32_t)gt, (uint32_t)(gt >> 32), a.sweep, STREAM_Z << 24, a.seed_lo, a.seed_hi);
            U = uniform23(r.x);
        }
        int znew = 0, done = nv;
        for (int tt = 0; tt < nv; ++tt) {
            if (!MULTI && nnz >= SP_BLOCK) { done = tt; break; }   // the list may outgrow shared memory: hand over
            const int wt = __shfl_sync(FULL, w, tt);
            const int old = __shfl_sync(FULL, zold, tt);
            const float u = __shfl_sync(FULL, U, tt);
            list_remove<MULTI>(nz, cnt, nnz, old, lane);
            int slot;
            const int nw = sparse_draw<MULTI>(sa, nz, cnt, cum, nnz, wt, u, lane, slot);
            if (lane == tt) znew = nw;
            if (slot < 0) slot = list_find<MULTI>(nz, nnz, nw, lane);
            list_add_at(nz, cnt, nnz, nw, slot, lane);
        }
        if (lane < done) {
            a.z[t] = znew;
            if (a.n_wk_out) atomicAdd(&a.n_wk_out[(size_t)w * Ks + znew], 1);
        }
        if (done < nv) return tb + done;   // first token not processed
    }
    return t1;
}

// 4 CTAs (32 warps, 64 registers) per SM: 5, 6 and 8 CTAs were measured slower on B200 (spills; the kernel is
// bound by the DRAM sectors of the Phi gathers, not by latency)
__global__ void __launch_bounds__(SP_WARPS * 32, 4) z_spalias_kernel(SparseArgs sa)
{
    __shared__ __align__(16) int s_nz[SP_WARPS][SP_BLOCK];
    __shared__ __align__(16) int s_cnt[SP_WARPS][SP_BLOCK];
    const ZArgs &a = sa.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cap = sa.cap;
    int *gnz = sa.lists + ((size_t)blockIdx.x * SP_WARPS + warp) * (size_t)cap * 3;
    int *gcnt = gnz + cap;
    float *gcum = reinterpret_cast<float *>(gcnt + cap);
    int *nz = s_nz[warp], *cnt = s_cnt[warp];

    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(a.work_counter, 1ull);
        item = __shfl_sync(FULL, item, 0);
        if ((int64_t)item >= a.n_items) break;
        const int64_t d = a.item_doc[item], t0 = a.doc_off[d], t1 = a.doc_off[d + 1];
        if (t0 == t1) continue;
        // non-zero topic list in first-occurrence order (SpaliasUncollapsedParallelLDA.java:147-153)
        int nnz = 0;
        bool in_smem = true;
        for (int64_t tb = t0; tb < t1 && in_smem; tb += 32) {
            const int zz = (tb + lane < t1) ? a.z[tb + lane] : -1;
            const int nv = (int)((t1 - tb) < 32 ? (t1 - tb) : 32);
            for (int tt = 0; tt < nv; ++tt) {
                const int k = __shfl_sync(FULL, zz, tt);
                const int pos = list_find<false>(nz, nnz, k, lane);
                if (pos < 0 && nnz >= SP_BLOCK) { in_smem = false; break; }
                list_add_at(nz, cnt, nnz, k, pos, lane);
            }
        }
        int64_t next = t0;
        if (in_smem) next = sparse_tokens<false>(sa, nz, cnt, nullptr, nnz, t0, t1, lane);
        if (next < t1) {
            // slow path: continue (or, when the initial list did not fit, start over) on the global lists
            if (in_smem) {
                for (int i = lane; i < nnz; i += 32) { gnz[i] = nz[i]; gcnt[i] = cnt[i]; }
                __syncwarp();
            } else {
                nnz = 0;
                for (int64_t tb = t0; tb < t1; tb += 32) {
                    const int zz = (tb + lane < t1) ? a.z[tb + lane] : -1;
                    const int nv = (int)((t1 - tb) < 32 ? (t1 - tb) : 32);
                    for (int tt = 0; tt < nv; ++tt) {
                        const int k = __shfl_sync(FULL, zz, tt);
                        const int pos = list_find<true>(gnz, nnz, k, lane);
                        list_add_at(gnz, gcnt, nnz, k, pos, lane);
                    }
                }
            }
            sparse_tokens<true>(sa, gnz, gcnt, gcum, nnz, next, t1, lane);
        }
    }
}

size_t spalias_list_bytes(const Dims &dm, int max_doc_len, int sm_count)
{
    int cap = max_doc_len < dm.K ? max_doc_len : dm.K;
    cap = (cap + 1 + SP_BLOCK - 1) / SP_BLOCK * SP_BLOCK + SP_BLOCK;   // whole blocks are loaded
    return (size_t)sm_count * 8 * SP_WARPS * (size_t)cap * 12;   // up to 8 CTAs of 8 warps per SM
}

cudaError_t launch_z_spalias(const ZArgs &z, const AliasSlot *table, const float *type_norm,
                             int *lists, int max_doc_len, int sm_count, cudaStream_t st)
{
    if (z.n_items == 0) return cudaSuccess;
    SparseArgs sa;
    sa.z = z; sa.table = table; sa.type_norm = type_norm; sa.lists = lists;
    int cap = max_doc_len < z.dm.K ? max_doc_len : z.dm.K;
    sa.cap = (cap + 1 + SP_BLOCK - 1) / SP_BLOCK * SP_BLOCK + SP_BLOCK;
    int per_sm = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, z_spalias_kernel, SP_WARPS * 32, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int64_t grid = (int64_t)sm_count * per_sm;
    const int64_t need = (z.n_items + SP_WARPS - 1) / SP_WARPS;
    if (need < grid) grid = need;
    z_spalias_kernel<<<(unsigned)grid, SP_WARPS * 32, 0, st>>>(sa);
    return cudaGetLastError();
}

}  // namespace ldagpu
