// kernels_z.cu -- K1: the z-step of one Gibbs sweep, and the GGS theta draw that feeds it.
//
// Replaces (reference, src/main/java/cc/mallet/topics/):
//   LDAGroupedGibbsSampler.java:47-132      GGS  z-step (theta draw :60-72, token loop :79-130)
//   UncollapsedParallelLDA.java:1466-1545   PCGS z-step
//   UncollapsedParallelLDA.java:1354-1437   RecursiveDocumentSampler / loopOverBatches (scheduling)
//
// Design: one warp owns one work item (GGS: a chunk of 32..256 tokens of one
// document; PCGS: one whole document, because n_dk changes token by token).  The K topics of a
// token are spread over the lanes, lane l owning topics 4l..4l+3 of every 128-topic tile, so one
// Phi^T row is read as coalesced float4.  Rows are fetched by the TMA engine (cp.async.bulk,
// 1-D) into a per-warp shared-memory slot; the row is copied to registers as soon as it lands and
// the slot is refilled with the next run's row while the warp computes.  A run of equal
// word types shares one row fetch (and, for GGS, one set of tile totals).  The categorical draw is a
// fixed three-level fp32 prefix tree (lane-local fma prefix, distributed-butterfly tile totals,
// Kogge-Stone scan inside the chosen tile) so the CPU oracle can reproduce the sampled topic bit for bit; uniforms are
// Philox4x32-10 keyed by the global token index.
#include "common.cuh"
#include "contract_math.cuh"

namespace ldagpu {

constexpr unsigned FULL = 0xffffffffu;
// Tuning (measured on B200, PubMed-shaped K=1000 / Enron-shaped K=400, profiles/README.md): the
// kernel is bound by the SM's shared-memory data pipe (row copy + shuffles + spill reloads), not by
// row fetches (80 % of them hit L2), so resident warps matter more than prefetch depth: one slot
// and 4 CTAs/SM beat three slots and 2 CTAs/SM by 20 %.
#ifndef Z_MINB_DEF
#define Z_MINB_DEF 4
#endif
// warps per CTA.  GGS at 1024 topics keeps theta (32 registers) and the prefixes (32) live: 7 warps x 4 CTAs
// leave 72 registers per thread instead of 64 and halve the spill reloads, which ride on the same
// shared-memory data pipe (128 B per clock per SM) that bounds this kernel (profiles/README.md)
template <int NT, bool PCGS> __host__ __device__ constexpr int z_warps() { return (NT == 8 && !PCGS) ? 7 : 8; }
constexpr int Z_STAGES = 1;               // one Phi^T row slot per warp (the row also lives in registers)

template <int NT> struct RowScan {
    float p[NT][4];   // lane-local inclusive prefix of the 4 owned scores, per tile (p[j][3] = lane total)
    float B;          // cumulative tile total through tile (lane >> 2) & 7
    float S;          // total over all tiles
};

// Scores, lane-local prefixes, and the cumulative tile totals (contract: DESIGN.md 4.2).
// Tile totals come from a distributed butterfly: after the xor-16/8/4 exchanges lane l works for
// tile (l>>2)&7 only, so 8 tiles cost 4+2+1+1+1 shuffles instead of 8 x 5.
template <int NT>
__device__ __forceinline__ void scan_scores(const float4 (&a)[NT], const float4 (&ph)[NT],
                                            RowScan<NT> &rs, int lane)
{
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j < NT) {
            // lane-local prefix as one product and three fused multiply-adds (contract 4.2)
            float p0 = __fmul_rn(a[j < NT ? j : 0].x, ph[j < NT ? j : 0].x);
            float p1 = __fmaf_rn(a[j < NT ? j : 0].y, ph[j < NT ? j : 0].y, p0);
            float p2 = __fmaf_rn(a[j < NT ? j : 0].z, ph[j < NT ? j : 0].z, p1);
            float p3 = __fmaf_rn(a[j < NT ? j : 0].w, ph[j < NT ? j : 0].w, p2);
            rs.p[j < NT ? j : 0][0] = p0; rs.p[j < NT ? j : 0][1] = p1;
            rs.p[j < NT ? j : 0][2] = p2; rs.p[j < NT ? j : 0][3] = p3;
            t[j] = p3;
        } else {
            t[j] = 0.0f;
        }
    }
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0;
    float u[4], v[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float keep = h16 ? t[4 + i] : t[i], send = h16 ? t[i] : t[4 + i];
        u[i] = __fadd_rn(keep, __shfl_xor_sync(FULL, send, 16));
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float keep = h8 ? u[2 + i] : u[i], send = h8 ? u[i] : u[2 + i];
        v[i] = __fadd_rn(keep, __shfl_xor_sync(FULL, send, 8));
    }
    float w;
    {
        float keep = h4 ? v[1] : v[0], send = h4 ? v[0] : v[1];
        w = __fadd_rn(keep, __shfl_xor_sync(FULL, send, 4));
    }
    w = __fadd_rn(w, __shfl_xor_sync(FULL, w, 2));
    w = __fadd_rn(w, __shfl_xor_sync(FULL, w, 1));
    // inclusive scan over the tile index: quads hold tiles, so the offsets are 4, 8, 16 lanes
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
        float y = __shfl_up_sync(FULL, w, off);
        if (lane >= off) w = __fadd_rn(w, y);
    }
    rs.B = w;
    rs.S = __shfl_sync(FULL, w, 31);
}

// Inclusive scan of one tile's 32 lane totals, as the lanes hold it.  GGS keeps it across the tokens of
// a run: they share the scores, so a token that lands in the tile scanned last (always, when K <= 128)
// skips the scan -- the same values, computed once.
struct TileScan {
    int js;                  // tile the fields belong to, -1 = none
    float q0, q1, q2;        // the lane's own prefixes inside the tile
    float inc, prev;         // inclusive scan at this lane and at the lane before (0 for lane 0)
};

// first k with cumsum_k >= U * sum, searched tile -> lane -> element
template <int NT, bool REUSE>
__device__ __forceinline__ int draw_topic(const RowScan<NT> &rs, TileScan &ts, float U, int lane, int K)
{
    const float u = __fmul_rn(U, rs.S);
    int js = 0;
    float base = 0.0f;
    if (NT > 1) {
        const unsigned mt = __ballot_sync(FULL, rs.B >= u);
        js = (mt ? __ffs(mt) - 1 : 31) >> 2;
        if (js > NT - 1) js = NT - 1;
        base = __shfl_sync(FULL, rs.B, js > 0 ? 4 * js - 1 : 0);
        if (js == 0) base = 0.0f;
    }
    const float r = __fsub_rn(u, base);
    if (!REUSE || js != ts.js) {
        // js is warp-uniform: a real branch picks the tile's registers
        float q0 = rs.p[0][0], q1 = rs.p[0][1], q2 = rs.p[0][2], inc = rs.p[0][3];
#define LDAGPU_PICK(J)                                                                         \
    case J:                                                                                    \
        if (J < NT) {                                                                          \
            q0 = rs.p[J < NT ? J : 0][0]; q1 = rs.p[J < NT ? J : 0][1];                         \
            q2 = rs.p[J < NT ? J : 0][2]; inc = rs.p[J < NT ? J : 0][3];                        \
        }                                                                                      \
        break;
        switch (js) {
            LDAGPU_PICK(1) LDAGPU_PICK(2) LDAGPU_PICK(3) LDAGPU_PICK(4) LDAGPU_PICK(5) LDAGPU_PICK(6) LDAGPU_PICK(7)
            default: break;
        }
#undef LDAGPU_PICK
        // inclusive scan of the chosen tile's lane totals
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            float y = __shfl_up_sync(FULL, inc, off);
            if (lane >= off) inc = __fadd_rn(inc, y);
        }
        float prev = __shfl_up_sync(FULL, inc, 1);
        if (lane == 0) prev = 0.0f;
        ts.js = js; ts.q0 = q0; ts.q1 = q1; ts.q2 = q2; ts.inc = inc; ts.prev = prev;
    }
    const unsigned m = __ballot_sync(FULL, ts.inc >= r);
    const int ls = m ? __ffs(m) - 1 : 31;
    const float r2 = __fsub_rn(r, ts.prev);
    int i = 3;
    if (ts.q2 >= r2) i = 2;
    if (ts.q1 >= r2) i = 1;
    if (ts.q0 >= r2) i = 0;
    const int k = __shfl_sync(FULL, TILE * js + 4 * lane + i, ls);
    return k < K ? k : K - 1;
}

// shared-memory footprint of one warp, in bytes
template <int NT, bool PCGS> __host__ __device__ constexpr size_t z_warp_smem()
{
    return (size_t)Z_STAGES * NT * TILE * 4 + (PCGS ? (size_t)NT * TILE * 8 : 0);
}
template <int NT, bool PCGS> __host__ __device__ constexpr size_t z_cta_smem()
{
    return z_warps<NT, PCGS>() * z_warp_smem<NT, PCGS>() + (PCGS ? (size_t)NT * TILE * 4 : 0) +
           (size_t)z_warps<NT, PCGS>() * Z_STAGES * 8;
}

template <int NT, bool PCGS>
__global__ void __launch_bounds__(z_warps<NT, PCGS>() * 32, (PCGS && NT == 8) ? 2 : Z_MINB_DEF) z_kernel(ZArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int ROWF = NT * TILE;
    constexpr int Z_WARPS = z_warps<NT, PCGS>();
    unsigned char *wbase = smem_raw + (size_t)warp * z_warp_smem<NT, PCGS>();
    float *ring = reinterpret_cast<float *>(wbase);
    int *cnt = reinterpret_cast<int *>(wbase + (size_t)Z_STAGES * ROWF * 4);       // PCGS only
    float *av = reinterpret_cast<float *>(wbase + (size_t)Z_STAGES * ROWF * 4 + (size_t)ROWF * 4);
    float *alpha_s = reinterpret_cast<float *>(smem_raw + Z_WARPS * z_warp_smem<NT, PCGS>());
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + Z_WARPS * z_warp_smem<NT, PCGS>() +
                                                  (PCGS ? (size_t)ROWF * 4 : 0)) + warp * Z_STAGES;
    const int K = a.dm.K, Ks = a.dm.Ks;
    const uint32_t row_bytes = (uint32_t)Ks * 4u;

    // zero the ring once: the bulk copies only ever write the first Ks floats of a slot, so the
    // padding topics read as +0 for the whole kernel
    for (int i = lane; i < Z_STAGES * ROWF / 4; i += 32)
        reinterpret_cast<float4 *>(ring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (PCGS) {
        for (int i = lane; i < ROWF / 4; i += 32) reinterpret_cast<int4 *>(cnt)[i] = make_int4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < ROWF; i += blockDim.x) alpha_s[i] = i < Ks ? a.alpha[i] : 0.0f;
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < Z_STAGES; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (PCGS) __syncthreads(); else __syncwarp();

    uint32_t phase = 0;            // parity of the slot's mbarrier (one completed fetch flips it)
    float *const slot = ring;
    uint64_t *const bar = bars;

    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(a.work_counter, 1ull);
        item = __shfl_sync(FULL, item, 0);
        if ((int64_t)item >= a.n_items) break;

        int64_t d, t0, t1;
        float4 th[NT];
        if (PCGS) {
            d = a.item_doc[item];
            t0 = a.doc_off[d];
            t1 = a.doc_off[d + 1];
            if (t0 == t1) continue;   // UncollapsedParallelLDA.java:1474
            for (int64_t t = t0 + lane; t < t1; t += 32) atomicAdd(&cnt[a.z[t]], 1);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                int4 c = reinterpret_cast<const int4 *>(cnt)[j * 32 + lane];
                float4 al = reinterpret_cast<const float4 *>(alpha_s)[j * 32 + lane];
                float4 v;
                v.x = __fadd_rn(__int2float_rn(c.x), al.x);
                v.y = __fadd_rn(__int2float_rn(c.y), al.y);
                v.z = __fadd_rn(__int2float_rn(c.z), al.z);
                v.w = __fadd_rn(__int2float_rn(c.w), al.w);
                reinterpret_cast<float4 *>(av)[j * 32 + lane] = v;
            }
            __syncwarp();
        } else {
            d = a.item_doc[item];
            t0 = a.item_begin[item];
            int64_t de = a.doc_off[d + 1];
            t1 = t0 + a.chunk < de ? t0 + a.chunk : de;
            const float *trow = a.theta + (size_t)d * Ks;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                int k0 = j * TILE + lane * 4;
                th[j] = k0 < Ks ? __ldg(reinterpret_cast<const float4 *>(trow + k0))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }

        for (int64_t tb = t0; tb < t1; tb += 32) {
            const int64_t t = tb + lane;
            const bool valid = t < t1;
            const int nvalid = (int)((t1 - tb) < 32 ? (t1 - tb) : 32);
            int w = valid ? a.tokens[t] : -1;
            int zold = (PCGS && valid) ? a.z[t] : 0;
            int wprev = __shfl_up_sync(FULL, w, 1);
            unsigned heads = __ballot_sync(FULL, valid && (lane == 0 || w != wprev));
            float U = 0.0f;
            if (valid) {
                unsigned long long gt = (unsigned long long)(a.dm.token_base + t);
                uint4 r = philox4x32_10((uint32_t)gt, (uint32_t)(gt >> 32), a.sweep, STREAM_Z << 24,
                                        a.seed_lo, a.seed_hi);
                U = uniform23(r.x);
            }
            int znew = 0;
            // One row slot per warp: the row of a run is copied to registers as soon as it lands, and the
            // slot is refilled with the next run's row while the warp computes (lane 0 drives the TMA).
            auto request = [&](int head_lane) {
                const int wrow = __shfl_sync(FULL, w, head_lane);
                if (lane == 0) {
                    mbar_expect_tx(bar, row_bytes);
                    bulk_g2s(slot, a.phiT + (size_t)wrow * Ks, row_bytes, bar);
                }
            };
            unsigned rem = heads;          // lane 0 of a non-empty block is always a head
            int b = 0;
            request(0);
            for (;;) {
                rem &= rem - 1;
                const int e = rem ? __ffs(rem) - 1 : nvalid;   // run = tokens [b, e) of this block
                mbar_wait(bar, phase);
                phase ^= 1u;
                float4 ph[NT];
#pragma unroll
                for (int j = 0; j < NT; ++j) ph[j] = reinterpret_cast<const float4 *>(slot)[j * 32 + lane];
                __syncwarp();
                if (rem) request(e);

                RowScan<NT> rs;
                TileScan ts;
                ts.js = -1;
                if (!PCGS) scan_scores<NT>(th, ph, rs, lane);
                for (int tt = b; tt < e; ++tt) {
                    if (PCGS) {
                        // remove the token from the document counts (UncollapsedParallelLDA.java:1494);
                        // every lane computes the new entry (broadcast reads), lane 0 stores it
                        const int old = __shfl_sync(FULL, zold, tt);
                        const int c = cnt[old] - 1;
                        const float v = __fadd_rn(__int2float_rn(c), alpha_s[old]);
                        __syncwarp();
                        if (lane == 0) { cnt[old] = c; av[old] = v; }
                        __syncwarp();
                        float4 aa[NT];
#pragma unroll
                        for (int j = 0; j < NT; ++j) aa[j] = reinterpret_cast<const float4 *>(av)[j * 32 + lane];
                        scan_scores<NT>(aa, ph, rs, lane);
                    }
                    const float Ut = __shfl_sync(FULL, U, tt);
                    // (at 8 tiles the six extra live registers cost more in spill reloads than the reuse saves)
                    const int k = draw_topic<NT, !PCGS && NT < 8>(rs, ts, Ut, lane, K);
                    if (lane == tt) znew = k;
                    if (PCGS) {
                        // add it back under its new topic (UncollapsedParallelLDA.java:1535)
                        const int c = cnt[k] + 1;
                        const float v = __fadd_rn(__int2float_rn(c), alpha_s[k]);
                        __syncwarp();
                        if (lane == 0) { cnt[k] = c; av[k] = v; }
                        __syncwarp();
                    }
                }
                if (!rem) break;
                b = e;
            }
            if (valid) {
                a.z[t] = znew;
                // fused count rebuild (the reference adds its +-1 deltas inside the token loop too,
                // UncollapsedParallelLDA.java:1505,1542): fire-and-forget reductions ride on the
                // memory system this issue-bound kernel leaves idle
                if (a.n_wk_out) atomicAdd(&a.n_wk_out[(size_t)w * Ks + znew], 1);
            }
        }
        if (PCGS) {
            // leave the histogram clean for the next document
            for (int i = lane; i < ROWF / 4; i += 32) reinterpret_cast<int4 *>(cnt)[i] = make_int4(0, 0, 0, 0);
            __syncwarp();
        }
    }
}

template <int NT, bool PCGS>
static cudaError_t launch_z_t(const ZArgs &a, int sm_count, cudaStream_t st)
{
    constexpr size_t smem = z_cta_smem<NT, PCGS>();
    constexpr int Z_WARPS = z_warps<NT, PCGS>();
    static bool configured = false;
    static int ctas_per_sm = 1;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(z_kernel<NT, PCGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, z_kernel<NT, PCGS>, Z_WARPS * 32, smem);
        if (e != cudaSuccess) return e;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        configured = true;
    }
    int64_t warps_needed = a.n_items;
    int64_t grid = (int64_t)sm_count * ctas_per_sm;   // persistent: every resident warp pulls items
    int64_t need = (warps_needed + Z_WARPS - 1) / Z_WARPS;
    if (need < grid) grid = need;
    if (grid < 1) grid = 1;
    z_kernel<NT, PCGS><<<(unsigned)grid, Z_WARPS * 32, smem, st>>>(a);
    return cudaGetLastError();
}

template <bool PCGS> static cudaError_t launch_z_any(const ZArgs &a, int sm_count, cudaStream_t st)
{
    if (a.n_items == 0) return cudaSuccess;
    int nt = a.dm.NT;
    if (nt <= 1) return launch_z_t<1, PCGS>(a, sm_count, st);
    if (nt <= 2) return launch_z_t<2, PCGS>(a, sm_count, st);
    if (nt <= 4) return launch_z_t<4, PCGS>(a, sm_count, st);
    if (nt <= 8) return launch_z_t<8, PCGS>(a, sm_count, st);
    return PCGS ? launch_z_pcgs_big(a, sm_count, st) : launch_z_ggs_big(a, sm_count, st);
}

cudaError_t launch_z_ggs(const ZArgs &a, int sm_count, cudaStream_t st) { return launch_z_any<false>(a, sm_count, st); }
cudaError_t launch_z_pcgs(const ZArgs &a, int sm_count, cudaStream_t st) { return launch_z_any<true>(a, sm_count, st); }

// ---------------------------------------------------------------------------------------
// GGS theta draw: theta_d ~ Dir(n_d + alpha) from the counts before the document is resampled
// (LDAGroupedGibbsSampler.java:60-72; ParallelDirichlet.java:46-70 = K Gammas, normalise, floor).
// One warp per document, lane l owning topics 4l..4l+3 of every tile (the z-step's layout).
//   phase 1  branch-free lock step over the lane's 4*NT cells: attempt 0 of every ZERO-COUNT cell
//            (shape = alpha_k, the vast majority; Marsaglia-Tsang constants from a per-CTA table,
//            cf. the reference's MarsagliaSparseDirichlet.java:9-29) is settled when the squeeze
//            accepts it (~92 %); all other cells are appended to a per-warp list in shared memory
//            in the same step (ballot + popc)
//   phase 2  the list is drained by all 32 lanes with a flattened attempt loop: a lane whose cell
//            is accepted takes the next list entry, so rejections do not idle the warp
//   phase 3  normalising sum in the contract's order (lane-sequential, then xor butterfly), divide,
//            floor, coalesced float4 store
// Cells are independent and keyed by (document, topic, attempt): the order of evaluation does not
// change any value.  One shared-memory word per topic holds the count until the cell is drawn and
// the Gamma value afterwards.
// ---------------------------------------------------------------------------------------
constexpr int TH_WARPS = 8;
constexpr int TH_PLIST = 512;   // pending-list capacity per warp (drained early when nearly full)

__global__ void __launch_bounds__(TH_WARPS * 32) theta_kernel(ThetaArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NT = a.dm.NT, K = a.dm.K, Ks = a.dm.Ks;
    const int ROWF = NT * TILE;
    const unsigned lt_mask = (1u << lane) - 1u;
    float *d0 = reinterpret_cast<float *>(smem_raw);
    float *c0 = d0 + ROWF;
    float *i0 = c0 + ROWF;
    unsigned char *wb = smem_raw + (size_t)3 * ROWF * 4 + (size_t)warp * ((size_t)ROWF * 4 + TH_PLIST * 2);
    int *cg = reinterpret_cast<int *>(wb);                     // count, then Gamma value bits
    unsigned short *plist = reinterpret_cast<unsigned short *>(wb + (size_t)ROWF * 4);
    for (int k = threadIdx.x; k < ROWF; k += blockDim.x) {
        bool b; float dd, cc, ii;
        gamma_setup<float>(k < K ? a.alpha[k] : 1.0f, b, dd, cc, ii);
        d0[k] = dd; c0[k] = cc; i0[k] = ii;
    }
    for (int i = lane; i < ROWF; i += 32) cg[i] = 0;
    __syncthreads();

    for (;;) {
        unsigned long long dd = 0;
        if (lane == 0) dd = atomicAdd(a.work_counter, 1ull);
        dd = __shfl_sync(FULL, dd, 0);
        if ((int64_t)dd >= a.dm.D) break;
        const int64_t d = (int64_t)dd;
        const int64_t t0 = a.doc_off[d], t1 = a.doc_off[d + 1];
        float *trow = a.theta + (size_t)d * Ks;
        if (t0 == t1) {   // empty document: the reference skips it, its theta row stays zero
            for (int i = lane; i < Ks; i += 32) trow[i] = 0.0f;
            continue;
        }
        for (int64_t t = t0 + lane; t < t1; t += 32) atomicAdd(&cg[a.z[t]], 1);
        __syncwarp();

        const unsigned long long cell0 = (unsigned long long)(a.dm.doc_base + d) * (unsigned long long)K;
        int npend = 0;

        // phase 2 body: drain plist[0, npend) with a flattened attempt loop
        auto drain = [&]() {
            __syncwarp();
            int next = 32, idx = lane, k = 0;
            bool have = idx < npend, boost = false;
            float dd_ = 0.f, cc_ = 0.f, ii_ = 0.f;
            uint32_t attempt = 0;
            if (have) {
                k = plist[idx];
                gamma_setup<float>(__fadd_rn(__int2float_rn(cg[k]), a.alpha[k]), boost, dd_, cc_, ii_);
            }
            while (__any_sync(FULL, have)) {
                bool finished = false;
                if (have) {
                    const unsigned long long cell = cell0 + (unsigned long long)k;
                    uint4 w = philox4x32_10((uint32_t)cell, (uint32_t)(cell >> 32), a.sweep,
                                            (STREAM_THETA << 24) | attempt, a.seed_lo, a.seed_hi);
                    float gv;
                    if (gamma_attempt<float>(boost, dd_, cc_, ii_, w, gv)) {
                        cg[k] = __float_as_int(gv);
                        finished = true;
                    } else {
                        ++attempt;
                    }
                }
                const unsigned fm = __ballot_sync(FULL, finished);
                if (finished) {
                    idx = next + __popc(fm & lt_mask);
                    have = idx < npend;
                    attempt = 0;
                    if (have) {
                        k = plist[idx];
                        gamma_setup<float>(__fadd_rn(__int2float_rn(cg[k]), a.alpha[k]), boost, dd_, cc_, ii_);
                    }
                }
                next += __popc(fm);
            }
            npend = 0;
            __syncwarp();
        };

        // ---- phase 1; the cells it leaves open are appended to the pending list as they are found
        //      (ballot + popc, all lanes) and the list is drained whenever another 32 might not fit
        for (int q = 0; q < NT * 4; ++q) {
            const int k = (q >> 2) * TILE + lane * 4 + (q & 3);
            const bool valid = k < K;
            bool done = !valid;
            if (valid && cg[k] == 0) {
                const float ii_ = i0[k];
                const unsigned long long cell = cell0 + (unsigned long long)k;
                uint4 w = philox4x32_10((uint32_t)cell, (uint32_t)(cell >> 32), a.sweep, STREAM_THETA << 24,
                                        a.seed_lo, a.seed_hi);
                float gv;
                done = gamma_attempt_squeeze<float>(ii_ > 0.0f, d0[k], c0[k], ii_, w, gv);
                if (done) cg[k] = __float_as_int(gv);
            }
            const unsigned open_mask = __ballot_sync(FULL, !done);
            if (!done) plist[npend + __popc(open_mask & lt_mask)] = (unsigned short)k;
            npend += __popc(open_mask);
            if (npend > TH_PLIST - 32) drain();
        }
        // ---- phase 2
        if (npend > 0) drain();
        __syncwarp();
        // ---- phase 3: sum in contract order, normalise, store
        const float4 *g4 = reinterpret_cast<const float4 *>(cg);
        float acc = 0.0f;
        for (int j = 0; j < NT; ++j) {
            const float4 v = g4[j * 32 + lane];
            acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y);
            acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, off));
        const float sum = acc;
        // contract 4.3: theta = g * (1 / sum) -- one division per document; the products handle the
        // many denormal g (alpha << 1) at full speed, where a division would take its slow path
        const float inv = sum != 0.0f ? __fdiv_rn(1.0f, sum) : 0.0f;
        for (int j = 0; j < NT; ++j) {
            int k0 = j * TILE + lane * 4;
            float4 v = g4[j * 32 + lane];
            if (sum != 0.0f) {
                v.x = __fmul_rn(v.x, inv); v.y = __fmul_rn(v.y, inv);
                v.z = __fmul_rn(v.z, inv); v.w = __fmul_rn(v.w, inv);
                if (v.x <= 0.0f) v.x = 0x1p-149f;
                if (v.y <= 0.0f) v.y = 0x1p-149f;
                if (v.z <= 0.0f) v.z = 0x1p-149f;
                if (v.w <= 0.0f) v.w = 0x1p-149f;
            }
            // padding topics (k >= K) carry no probability
            if (k0 + 0 >= K) v.x = 0.0f;
            if (k0 + 1 >= K) v.y = 0.0f;
            if (k0 + 2 >= K) v.z = 0.0f;
            if (k0 + 3 >= K) v.w = 0.0f;
            if (k0 < Ks) reinterpret_cast<float4 *>(trow)[k0 >> 2] = v;
            reinterpret_cast<int4 *>(cg)[j * 32 + lane] = make_int4(0, 0, 0, 0);
        }
        __syncwarp();
    }
}

cudaError_t launch_theta(const ThetaArgs &a, int sm_count, cudaStream_t st)
{
    if (a.dm.D == 0) return cudaSuccess;
    if (a.dm.NT > MAX_REG_TILES) return launch_theta_big(a, sm_count, st);
    const size_t rowf = (size_t)a.dm.NT * TILE;
    size_t smem = 3 * rowf * 4 + TH_WARPS * (rowf * 4 + TH_PLIST * 2);
    static size_t configured_smem = 0;
    if (smem > configured_smem) {
        cudaError_t e = cudaFuncSetAttribute(theta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured_smem = smem;
    }
    int per_sm = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, theta_kernel, TH_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)sm_count * per_sm;
    int64_t need = (a.dm.D + TH_WARPS - 1) / TH_WARPS;
    if (need < grid) grid = need;
    theta_kernel<<<(unsigned)grid, TH_WARPS * 32, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ldagpu
