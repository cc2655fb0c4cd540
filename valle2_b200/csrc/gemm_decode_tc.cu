// Decode-shape linear layers that FINISH inside one launch: swap-AB tcgen05 GEMM (weight rows on the 128-row MMA-M, the
// batch on MMA-N, split-K over CTAs so that tiles x splits fills the SMs) + in-kernel split-K reduction + the layer's
// arithmetic in the epilogue.  A decoder layer of the batched AR step becomes QKV -> attention -> out-proj -> FFN1 -> FFN2
// (5 dependent launches instead of 8: no LayerNorm kernel, no GELU-reduce kernel):
//
//   split-K reduction  every CTA stores its 128 x BN fp32 accumulator to an L2-resident exchange buffer, arrives on the
//                      tile's counter (release), waits until the tile's n_split CTAs have arrived (acquire), then reduces
//                      ITS share of the tile (all 128 weight rows x M / n_split batch rows) over the n_split partials in
//                      split order -- deterministic, and the reduction is spread over all CTAs instead of one.  All CTAs of
//                      the grid are co-resident (grid <= #SMs, one CTA per SM), so the wait cannot deadlock; it is bounded.
//   LayerNorm          folded algebraically into the consuming GEMM (modules.py:271,278 + :146 / :220):
//                        LN(x) . W^T = rstd * (x . (gamma (.) W)^T  -  mean * c)  +  (beta . W^T),   c[n] = sum_k gamma[k] W[n][k]
//                      so the MMA runs on the raw residual rows (bf16 copy written by the producer) against pre-scaled
//                      weights, and mean / rstd come from per-row (sum, sum of squares) partials that the producing
//                      epilogue (out-proj / FFN2 / embedding) writes next to the rows.
//   epilogues          PLAIN  y = acc (+ bias)                                     logits        valle_ar.py:158
//                      LN     y = rstd (acc - mean c) + b'                         QKV           modules.py:271 + :146
//                      LN_GELU y = gelu_erf(rstd (acc - mean c) + b')  (bf16)       FFN linear_1  modules.py:278 + :220-221
//                      RESIDUAL x += acc + bias; bf16 copy of x; row statistics    out / linear_2  modules.py:171+:274, :221+:278
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;          // weight rows per tile = UMMA M = TMEM lanes
constexpr int BK = 64;           // k-block: 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int THREADS = 192;     // warp 0: TMA producer, warp 1: MMA issuer + TMEM allocator, warps 2-5: epilogue
constexpr int EPI_THREADS = 128;

struct DecGemmParams {
    int N, M, K;
    int tiles, n_split, kb_total, kb_per_split, rows_per_split;
    int mode, late_trigger;
    float* part;                 // exchange buffer [tiles][n_split][BN batch rows][128] fp32
    unsigned* counters;          // [tiles] monotonically increasing arrival counters (zeroed once by the caller)
    const float* bias;           // PLAIN: optional; LN / LN_GELU: b' = beta . W^T (+ bias); RESIDUAL: bias
    const float* colsum;         // LN / LN_GELU: c[n]
    const float2* stats_in;      // LN / LN_GELU: [M][n_chunks_in] (sum, sum of squares) partials of the input rows
    int n_chunks_in;
    float inv_d, eps;
    float* y32;                  // PLAIN / LN: fp32 [M][ldy32]
    int64_t ldy32;
    __nv_bfloat16* y16;          // LN_GELU: bf16 out [M][ldy16]; RESIDUAL: bf16 copy of the updated residual rows
    int64_t ldy16;
    float* xres;                 // RESIDUAL: fp32 residual stream [M][ldx], updated in place
    int64_t ldx;
    float2* stats_out;           // RESIDUAL: [M][tiles]
    unsigned long long* dbg;     // optional %globaltimer stamps [cta][16]
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define DG_STAMP(i) do { if (p.dbg) p.dbg[blockIdx.x * 16 + (i)] = gtimer(); } while (0)

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

constexpr int recv_rows(int BN) { return BN + 16; }      // n_split * ceil(M / n_split) <= M + n_split - 1, n_split <= 16

// CLUSTER: the n_split CTAs of a tile form a thread-block cluster (blockIdx.x = tile * n_split + split) and exchange their
// partial accumulators through distributed shared memory (remote 16-byte stores + one cluster barrier) instead of the
// L2-resident buffer + release/acquire counter: ~1 us less per launch (tools/dg_timeline.py).  Same summation order.
template <int BN, bool CLUSTER>
__global__ void __launch_bounds__(THREADS, 1) decode_gemm_kernel(const __grid_constant__ CUtensorMap tm_w,
                                                                 const __grid_constant__ CUtensorMap tm_x, DecGemmParams p) {
    constexpr int A_BYTES = BM * BK * 2;
    constexpr int B_BYTES = BN * BK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int RING_BYTES = STAGES * STAGE_BYTES;
    constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    constexpr uint32_t IDESC = umma_idesc_bf16(BM, BN, 0, 0);
    constexpr int MAXI = 2;                     // epilogue work items per thread whose operands are prefetched
    // ALIAS (BN >= 64): the receive buffer of the exchange lives INSIDE the operand ring, so that two CTAs fit per SM and the
    // next kernel's CTAs (and their weight prefetch) can be resident while this one runs -- at batch 64 a separate buffer made
    // every launch start 6 us late (tools/dg_timeline.py 64).  Price: nobody may push before EVERY CTA of the cluster has finished
    // its MMAs (phase 1 of the cluster barrier moves behind the accumulator), a few 100 ns of skew; not worth it at BN <= 32.
    constexpr bool ALIAS = CLUSTER && BN >= 64;
    constexpr int STG_BYTES = 4 * 32 * 32 * 4;  // the epilogue warps' transpose staging at the start of the ring
    // ONE kernel serves the four epilogue modes (run-time switch): the GEMMs of a layer alternate between them, and a shared
    // function keeps its code in the instruction caches (one kernel per mode measured +0.4 us per launch, tools/dg_timeline.py)
    const bool IS_LN = (p.mode == VB_DG_LN || p.mode == VB_DG_LN_GELU), IS_RES = (p.mode == VB_DG_RESIDUAL);

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t tfull_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ float s_mu[BN], s_rstd[BN];

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    float* s_stage = reinterpret_cast<float*>(smem_al);                      // transpose staging: aliases the operand ring
    float* s_recv = reinterpret_cast<float*>(smem_al + (ALIAS ? STG_BYTES : RING_BYTES));   // CLUSTER: [n_split][rows_per_split][128]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = CLUSTER ? blockIdx.x / p.n_split : blockIdx.x % p.tiles;
    const int split = CLUSTER ? blockIdx.x % p.n_split : blockIdx.x / p.tiles;
    const int kb0 = split * p.kb_per_split;
    const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
    const int nkb = kb1 - kb0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_x);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(&tfull_bar), 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(&tmem_base_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    if (threadIdx.x == 0) DG_STAMP(0);
    if constexpr (CLUSTER && !ALIAS) cluster_arrive();   // phase 1: every CTA of the cluster is running (waited for before the pushes)

    // the epilogue's share of the tile after the exchange: ALL 128 weight rows x batch rows [b_lo, b_hi)
    const int b_lo = min(p.M, split * p.rows_per_split), b_hi = min(p.M, b_lo + p.rows_per_split);
    const int n_base = tile * BM;
    const int et = threadIdx.x - 64;
    const int items = (b_hi - b_lo) * 32;
    float4 pf_b[MAXI], pf_c[MAXI], pf_x[MAXI];   // prefetched bias / colsum / residual values of the first MAXI items

    if (!p.late_trigger) pdl_trigger();
    if (warp == 0) {
        if (elect_one()) {
            // weights are immutable: their tiles are requested before the dependency wait (PDL), so the weight stream of
            // this GEMM overlaps the tail of the predecessor; the activation tile of a slot follows after the wait.
            const int pre = min(nkb, STAGES);
            for (int i = 0; i < pre; ++i) {
                const uint32_t fb = smem_u32(&full_bar[i]);
                mbar_expect_tx(fb, STAGE_BYTES);
                tma_load_2d(smem_base + i * STAGE_BYTES, &tm_w, fb, (kb0 + i) * BK, tile * BM);
            }
            DG_STAMP(1);
            pdl_wait();
            DG_STAMP(2);
            if (p.late_trigger) pdl_trigger();
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < nkb; ++i) {
                const uint32_t fb = smem_u32(&full_bar[stage]);
                const uint32_t sa = smem_base + stage * STAGE_BYTES;
                if (i >= pre) {
                    mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                    mbar_expect_tx(fb, STAGE_BYTES);
                    tma_load_2d(sa, &tm_w, fb, (kb0 + i) * BK, tile * BM);
                }
                tma_load_2d(sa + A_BYTES, &tm_x, fb, (kb0 + i) * BK, 0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        } else if (p.late_trigger) { pdl_wait(); pdl_trigger(); }
        __syncwarp();
        if constexpr (ALIAS) cluster_arrive();
        if constexpr (CLUSTER) cluster_wait();                // phase 1 (every thread pairs each arrive with a wait)
    } else if (warp == 1) {
        if (p.late_trigger) { pdl_wait(); pdl_trigger(); }
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(smem_u32(&full_bar[stage]), phase);
                if (i == 0) DG_STAMP(3);
                tc_fence_after();
                const uint32_t sa = smem_base + stage * STAGE_BYTES;
#pragma unroll
                for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                    const uint64_t da = umma_desc_sw128(sa + kk * UMMA_K * 2, 16, 1024);
                    const uint64_t db = umma_desc_sw128(sa + A_BYTES + kk * UMMA_K * 2, 16, 1024);
                    umma_f16(tmem_base, da, db, IDESC, (i > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(smem_u32(&empty_bar[stage]));
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(smem_u32(&tfull_bar));
            DG_STAMP(4);
        }
        __syncwarp();
        if constexpr (ALIAS) cluster_arrive();
        if constexpr (CLUSTER) cluster_wait();                // phase 1
    } else {
        // ------------------------------------------------------------------ epilogue warps, part 1 -----
        const int ew = warp - 2;
        const int q = warp & 3;                               // TMEM lane quadrant of this warp
        // immutable epilogue operands of this thread's first work items: before the dependency wait
#pragma unroll
        for (int i = 0; i < MAXI; ++i) {
            pf_b[i] = pf_c[i] = pf_x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int gi = et + i * EPI_THREADS;
            const int n = n_base + 4 * (gi & 31);
            if (gi < items && n + 3 < p.N) {
                if (p.bias) pf_b[i] = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                if (IS_LN) pf_c[i] = __ldg(reinterpret_cast<const float4*>(p.colsum + n));
            }
        }
        pdl_wait();
        if (p.late_trigger) pdl_trigger();
        if (IS_RES && (p.ldx & 3) == 0) {
#pragma unroll
            for (int i = 0; i < MAXI; ++i) {
                const int gi = et + i * EPI_THREADS;
                const int n = n_base + 4 * (gi & 31);
                if (gi < items && n + 3 < p.N)
                    pf_x[i] = __ldcg(reinterpret_cast<const float4*>(p.xres + static_cast<int64_t>(b_lo + (gi >> 5)) * p.ldx + n));
            }
        }
        if (IS_LN) {
            // row statistics of this share's input rows from the producer's per-tile partial sums (fixed order)
            for (int b = b_lo + et; b < b_hi; b += EPI_THREADS) {
                const float2* sp = p.stats_in + static_cast<int64_t>(b) * p.n_chunks_in;
                float s = 0.f, ss = 0.f;
                int c = 0;
                if ((p.n_chunks_in & 1) == 0) {               // two chunks per 16-byte load, four loads in flight
                    const float4* sp4 = reinterpret_cast<const float4*>(sp);
#pragma unroll 1
                    for (; c + 8 <= p.n_chunks_in; c += 8) {
                        float4 t[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) t[j] = __ldcg(sp4 + (c >> 1) + j);
#pragma unroll
                        for (int j = 0; j < 4; ++j) { s += t[j].x; ss += t[j].y; s += t[j].z; ss += t[j].w; }
                    }
                }
#pragma unroll 1
                for (; c < p.n_chunks_in; ++c) { const float2 t = __ldcg(sp + c); s += t.x; ss += t.y; }
                const float mu = s * p.inv_d;
                const float var = fmaxf(ss * p.inv_d - mu * mu, 0.f);
                s_mu[b - b_lo] = mu;
                s_rstd[b - b_lo] = rsqrtf(var + p.eps);
            }
        }
        // ---- accumulator -> the owners of its batch rows.  32 n x 32 b per warp and chunk go through a shared-memory
        // transpose so that every store instruction writes 16 bytes per lane along n.
        mbar_wait(smem_u32(&tfull_bar), 0);
        if (et == 0) DG_STAMP(5);
        tc_fence_after();
        if constexpr (ALIAS) cluster_arrive();                // ALIAS: phase 1 = every CTA's accumulator is complete, its ring is free
        if constexpr (CLUSTER) cluster_wait();                // phase 1 complete: remote shared memory may be written
        {
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
            float* stg = s_stage + ew * (32 * 32);
            float* dst = p.part + (static_cast<int64_t>(tile) * p.n_split + split) * BM * BN + q * 32;
            constexpr int CH = BN >= 32 ? 32 : 16;
            const int n4 = (lane & 7) * 4, bsub = lane >> 3;
            const uint32_t recv_u32 = smem_u32(s_recv);
            const uint32_t inv_rps = (65536u + p.rows_per_split - 1) / p.rows_per_split;   // b / rows_per_split == (b * inv_rps) >> 16 for b < 256
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += CH) {
                if (c0 >= p.M) break;                         // batch columns past M are never read back
                uint32_t v[CH];
                if constexpr (CH == 32) tmem_ld_32x32(t_addr + c0, reinterpret_cast<uint32_t(&)[32]>(v));
                else tmem_ld_32x16(t_addr + c0, reinterpret_cast<uint32_t(&)[16]>(v));
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < CH; ++j) stg[j * 32 + lane] = __uint_as_float(v[j]);
                __syncwarp();
#pragma unroll
                for (int it = 0; it < CH / 4; ++it) {
                    const int bb = it * 4 + bsub, b = c0 + bb;
                    if (b < p.M) {
                        const float4 val = *reinterpret_cast<const float4*>(stg + bb * 32 + n4);
                        if constexpr (CLUSTER) {
                            const int owner = static_cast<int>((static_cast<uint32_t>(b) * inv_rps) >> 16), bl = b - owner * p.rows_per_split;
                            const uint32_t local = recv_u32 + static_cast<uint32_t>(((split * p.rows_per_split + bl) * BM + q * 32 + n4) * 4);
                            uint32_t remote;
                            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(owner));
                            asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "f"(val.x), "f"(val.y),
                                         "f"(val.z), "f"(val.w) : "memory");
                        } else {
                            __stcg(reinterpret_cast<float4*>(dst + static_cast<int64_t>(b) * BM + n4), val);
                        }
                    }
                }
                __syncwarp();
            }
        }
        tc_fence_before();
        if constexpr (!CLUSTER) {
            // ---- arrive on the tile's counter, wait for the tile's other splits
            epi_bar();
            if (et == 0) {
                DG_STAMP(8);
                __threadfence();
                const unsigned ticket = atomicAdd(p.counters + tile, 1u);
                const unsigned target = (ticket / static_cast<unsigned>(p.n_split) + 1u) * static_cast<unsigned>(p.n_split);
                DG_STAMP(9);
                unsigned spins = 0;
                while (static_cast<int>(ld_acquire_u32(p.counters + tile) - target) < 0) {
                    if (++spins > (1u << 24)) __trap();       // a lost arrival becomes a CUDA error, not a hung GPU
                }
                __threadfence();
                DG_STAMP(6);
            }
            epi_bar();
        } else if (et == 0) {
            DG_STAMP(8);
        }
    }
    if constexpr (CLUSTER) {                                  // phase 2: every partial of the cluster has been delivered
        cluster_arrive();
        if (warp >= 2) {
            cluster_wait();
            if (et == 0) DG_STAMP(6);
        }
    }
    if (warp >= 2) {
        // ---- part 2: reduce the share over the splits (split order) and finish the layer's arithmetic;
        //      thread = (batch row, 4 consecutive n)
        constexpr bool from_smem = CLUSTER;
        const float* pbase = from_smem ? s_recv : p.part + static_cast<int64_t>(tile) * p.n_split * BM * BN;
        const int64_t sstr4 = from_smem ? static_cast<int64_t>(p.rows_per_split) * BM / 4 : static_cast<int64_t>(BM) * BN / 4;
        int slot = 0;
#pragma unroll 1
        for (int gi = et; gi < ((items + 31) & ~31); gi += EPI_THREADS, ++slot) {   // whole warps stay together (row sums)
            const bool live = gi < items;
            const int bl = live ? gi >> 5 : 0, g = gi & 31;
            const int b = b_lo + bl, n = n_base + 4 * g;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                const float4* src = reinterpret_cast<const float4*>(pbase + static_cast<int64_t>(from_smem ? bl : b) * BM) + g;
                int s = 0;
#pragma unroll 1
                for (; s + 8 <= p.n_split; s += 8) {
                    float4 t[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t[j] = from_smem ? src[(s + j) * sstr4] : __ldcg(src + (s + j) * sstr4);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { a.x += t[j].x; a.y += t[j].y; a.z += t[j].z; a.w += t[j].w; }
                }
#pragma unroll 1
                for (; s < p.n_split; ++s) {
                    const float4 t = from_smem ? src[s * sstr4] : __ldcg(src + s * sstr4);
                    a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                }
            }
            if (et == 0 && slot == 0) DG_STAMP(10);
            const bool full4 = live && (n + 3 < p.N);
            float av[4] = {a.x, a.y, a.z, a.w};
            float bv[4] = {0.f, 0.f, 0.f, 0.f}, cv[4] = {0.f, 0.f, 0.f, 0.f}, xo[4] = {0.f, 0.f, 0.f, 0.f};
            const bool res_vec = IS_RES && full4 && ((p.ldx & 3) == 0);
            if (live) {
                if (full4 && slot < MAXI) {
#pragma unroll
                    for (int i = 0; i < MAXI; ++i) {
                        if (i == slot) {
                            bv[0] = pf_b[i].x; bv[1] = pf_b[i].y; bv[2] = pf_b[i].z; bv[3] = pf_b[i].w;
                            cv[0] = pf_c[i].x; cv[1] = pf_c[i].y; cv[2] = pf_c[i].z; cv[3] = pf_c[i].w;
                            xo[0] = pf_x[i].x; xo[1] = pf_x[i].y; xo[2] = pf_x[i].z; xo[3] = pf_x[i].w;
                        }
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (n + e < p.N) {
                            if (p.bias) bv[e] = __ldg(p.bias + n + e);
                            if (IS_LN) cv[e] = __ldg(p.colsum + n + e);
                        }
                    }
                    if (res_vec) {
                        const float4 t = __ldcg(reinterpret_cast<const float4*>(p.xres + static_cast<int64_t>(b) * p.ldx + n));
                        xo[0] = t.x; xo[1] = t.y; xo[2] = t.z; xo[3] = t.w;
                    }
                }
            }
            if (et == 0 && slot == 0) DG_STAMP(11);
            if (IS_RES) {
                float sum = 0.f, sq = 0.f;
                if (live) {
                    float* xr = p.xres + static_cast<int64_t>(b) * p.ldx + n;
                    if (!res_vec) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) if (n + e < p.N) xo[e] = __ldcg(xr + e);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        av[e] = (n + e < p.N) ? xo[e] + (av[e] + bv[e]) : 0.f;
                        sum += av[e];
                        sq = fmaf(av[e], av[e], sq);
                    }
                    if (res_vec) *reinterpret_cast<float4*>(xr) = make_float4(av[0], av[1], av[2], av[3]);
                    else
#pragma unroll
                        for (int e = 0; e < 4; ++e) if (n + e < p.N) xr[e] = av[e];
                    if (p.y16) {
                        __nv_bfloat16* yb = p.y16 + static_cast<int64_t>(b) * p.ldy16 + n;
                        if (full4 && ((p.ldy16 & 3) == 0)) *reinterpret_cast<uint2*>(yb) = make_uint2(pack_bf16x2(av[0], av[1]), pack_bf16x2(av[2], av[3]));
                        else
#pragma unroll
                            for (int e = 0; e < 4; ++e) if (n + e < p.N) yb[e] = __float2bfloat16_rn(av[e]);
                    }
                }
                // statistics of (row b, this tile's 128 columns): lanes of the warp = the 32 column groups, fixed order
                sum = warp_sum(sum);
                sq = warp_sum(sq);
                if (live && g == 0 && p.stats_out) p.stats_out[static_cast<int64_t>(b) * p.tiles + tile] = make_float2(sum, sq);
            } else if (live) {
                if (IS_LN) {
                    const float mu = s_mu[bl], rstd = s_rstd[bl];
#pragma unroll
                    for (int e = 0; e < 4; ++e) av[e] = fmaf(rstd, fmaf(-mu, cv[e], av[e]), bv[e]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) av[e] += bv[e];
                }
                if (p.mode == VB_DG_LN_GELU) {
                    __nv_bfloat16* yr = p.y16 + static_cast<int64_t>(b) * p.ldy16 + n;
#pragma unroll
                    for (int e = 0; e < 4; ++e) av[e] = gelu_erf_fast(av[e]);
                    if (full4 && ((p.ldy16 & 3) == 0)) *reinterpret_cast<uint2*>(yr) = make_uint2(pack_bf16x2(av[0], av[1]), pack_bf16x2(av[2], av[3]));
                    else
#pragma unroll
                        for (int e = 0; e < 4; ++e) if (n + e < p.N) yr[e] = __float2bfloat16_rn(av[e]);
                } else {
                    float* yr = p.y32 + static_cast<int64_t>(b) * p.ldy32 + n;
                    if (full4 && ((p.ldy32 & 3) == 0)) *reinterpret_cast<float4*>(yr) = make_float4(av[0], av[1], av[2], av[3]);
                    else
#pragma unroll
                        for (int e = 0; e < 4; ++e) if (n + e < p.N) yr[e] = av[e];
                }
            }
        }
        if (et == 0) DG_STAMP(7);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

unsigned long long* g_dg_dbg = nullptr;

int plan(int64_t N, int64_t K, int* tiles, int* n_split, int* kb_per_split, int* rows_per_split) {
    const int t = static_cast<int>(vb_ceil_div(N, BM)), kb_total = static_cast<int>(vb_ceil_div(K, BK));
    if (t > vb_sm_count()) return VB_ERR_UNSUPPORTED;
    const int want = vb_sm_count() / t;                     // tiles x splits <= #SMs: every CTA of the grid is resident
    int ns = max(1, min(min(want, 32), max(1, kb_total / 2)));
    const int kps = static_cast<int>(vb_ceil_div(kb_total, ns));
    ns = static_cast<int>(vb_ceil_div(kb_total, kps));
    *tiles = t; *n_split = ns; *kb_per_split = kps; *rows_per_split = 0;    // batch rows per split: set by the caller (needs M)
    return VB_OK;
}

bool cluster_enabled() {          // VALLE_B200_DG_CLUSTER=0: exchange through the L2 buffer + counters only (A/B)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VALLE_B200_DG_CLUSTER");
        v = (e && e[0] == '0') ? 0 : (e ? atoi(e) : 16);     // value = largest cluster size used (8: portable sizes only)
        if (v != 0 && v < 2) v = 16;
    }
    return v != 0;
}
int cluster_max() { cluster_enabled(); const char* e = getenv("VALLE_B200_DG_CLUSTER"); const int v = e ? atoi(e) : 16; return v >= 2 ? v : 16; }

template <int BN, bool CLUSTER>
int launch(const CUtensorMap& tw, const CUtensorMap& tx, const DecGemmParams& p, cudaStream_t st) {
    constexpr bool ALIAS = CLUSTER && BN >= 64;
    static_assert(!ALIAS || 4 * 32 * 32 * 4 + recv_rows(BN) * BM * 4 <= STAGES * (BM * BK * 2 + BN * BK * 2), "receive buffer must fit in the ring");
    constexpr int SMEM = STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 + ((CLUSTER && !ALIAS) ? recv_rows(BN) * BM * 4 : 0);
    static bool configured = false;
    auto kern = decode_gemm_kernel<BN, CLUSTER>;
    if (!configured) {
        VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        if (CLUSTER) VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        configured = true;
    }
    if constexpr (!CLUSTER) {
        VB_CUDA(vb_launch(true, kern, dim3(p.tiles * p.n_split), dim3(THREADS), SMEM, st, tw, tx, p));
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(p.tiles * p.n_split); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM; cfg.stream = st;
        cudaLaunchAttribute attr[2];
        int na = 0;
        if (vb_pdl_enabled()) {
            attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = p.n_split; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
        cfg.attrs = attr; cfg.numAttrs = na;
        VB_CUDA(cudaLaunchKernelEx(&cfg, kern, tw, tx, p));
    }
    return VB_OK;
}

int bn_of(int M) { return M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : M <= 128 ? 128 : 256; }

}  // namespace

extern "C" int vb_decode_gemm_set_debug(void* buf) {   /* device buffer of grid * 8 uint64 stamps, or NULL */
    g_dg_dbg = static_cast<unsigned long long*>(buf);
    return VB_OK;
}

extern "C" int vb_decode_gemm_plan(int M, int64_t N, int64_t K, int* tiles, int* n_split, int64_t* ws_bytes) {
    VB_REQUIRE(M >= 1 && M <= 256 && N >= 1 && K >= 8, VB_ERR_BAD_ARG, "vb_decode_gemm_plan: bad shape");
    int t, ns, kps, rps;
    VB_REQUIRE(plan(N, K, &t, &ns, &kps, &rps) == VB_OK, VB_ERR_UNSUPPORTED, "vb_decode_gemm: N = %lld needs more tiles than SMs", (long long)N);
    if (tiles) *tiles = t;
    if (n_split) *n_split = ns;
    if (ws_bytes) *ws_bytes = static_cast<int64_t>(t) * ns * BM * bn_of(M) * 4;
    return VB_OK;
}

extern "C" int vb_decode_gemm(const void* x, int64_t ldx, const void* w, int64_t ldw, int M, int64_t N, int64_t K, int mode,
                              const float* bias, const float* colsum, const float* stats_in, int n_chunks_in, float eps,
                              float* y32, int64_t ldy32, void* y16, int64_t ldy16, float* xres, int64_t ldxres,
                              float* stats_out, void* ws_part, void* counters, int flags, void* stream) {
    VB_REQUIRE(x && w && ws_part && counters, VB_ERR_BAD_ARG, "vb_decode_gemm: null pointer");
    VB_REQUIRE(M >= 1 && M <= 256, VB_ERR_UNSUPPORTED, "vb_decode_gemm: M must be in [1,256] (got %d)", M);
    VB_REQUIRE(K % 8 == 0 && N >= 1, VB_ERR_UNSUPPORTED, "vb_decode_gemm: K %% 8 != 0 or N < 1");
    VB_REQUIRE(mode >= VB_DG_PLAIN && mode <= VB_DG_RESIDUAL, VB_ERR_BAD_ARG, "vb_decode_gemm: bad mode %d", mode);
    VB_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(colsum) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(xres) & 15) == 0, VB_ERR_BAD_ARG, "vb_decode_gemm: bias / colsum / xres must be 16-byte aligned");
    if (mode == VB_DG_PLAIN) VB_REQUIRE(y32, VB_ERR_BAD_ARG, "vb_decode_gemm: PLAIN needs y32");
    if (mode == VB_DG_LN) VB_REQUIRE(y32 && colsum && bias && stats_in && n_chunks_in >= 1, VB_ERR_BAD_ARG, "vb_decode_gemm: LN needs y32, colsum, bias, stats_in");
    if (mode == VB_DG_LN_GELU) VB_REQUIRE(y16 && colsum && bias && stats_in && n_chunks_in >= 1, VB_ERR_BAD_ARG, "vb_decode_gemm: LN_GELU needs y16, colsum, bias, stats_in");
    if (mode == VB_DG_RESIDUAL) VB_REQUIRE(xres && bias, VB_ERR_BAD_ARG, "vb_decode_gemm: RESIDUAL needs xres and bias");
    DecGemmParams p{};
    p.N = static_cast<int>(N); p.M = M; p.K = static_cast<int>(K);
    VB_REQUIRE(plan(N, K, &p.tiles, &p.n_split, &p.kb_per_split, &p.rows_per_split) == VB_OK, VB_ERR_UNSUPPORTED,
               "vb_decode_gemm: N = %lld needs more tiles than SMs", (long long)N);
    p.rows_per_split = static_cast<int>(vb_ceil_div(M, p.n_split));       // the reduce is shared out by batch rows
    p.kb_total = static_cast<int>(vb_ceil_div(K, BK));
    p.mode = mode;
    p.late_trigger = (flags & VB_FLAG_LATE_TRIGGER) ? 1 : 0;
    p.part = static_cast<float*>(ws_part);
    p.counters = static_cast<unsigned*>(counters);
    p.bias = bias; p.colsum = colsum;
    p.stats_in = reinterpret_cast<const float2*>(stats_in); p.n_chunks_in = n_chunks_in;
    p.inv_d = 1.0f / static_cast<float>(K); p.eps = eps;
    p.y32 = y32; p.ldy32 = ldy32;
    p.y16 = static_cast<__nv_bfloat16*>(y16); p.ldy16 = ldy16;
    p.xres = xres; p.ldx = ldxres;
    p.stats_out = reinterpret_cast<float2*>(stats_out);
    p.dbg = g_dg_dbg;
    CUtensorMap tw, tx;
    int rc;
    if ((rc = vb_make_tmap_bf16_2d(&tw, w, N, K, ldw, BM, BK)) != VB_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // cluster exchange: the tile's splits as one thread-block cluster (<= 16 CTAs), partials through distributed shared memory
    const bool cl = cluster_enabled() && p.n_split >= 2 && p.n_split <= cluster_max() && !(flags & VB_FLAG_DG_GLOBAL);
#define DG_CASE(BNV, CL)                                                                 \
    {                                                                                    \
        if ((rc = vb_make_tmap_bf16_2d(&tx, x, M, K, ldx, BNV, BK)) != VB_OK) return rc; \
        return launch<BNV, CL>(tw, tx, p, st);                                           \
    }
    switch (bn_of(M)) {
        case 16: if (cl) DG_CASE(16, true) else DG_CASE(16, false)
        case 32: if (cl) DG_CASE(32, true) else DG_CASE(32, false)
        case 64: if (cl) DG_CASE(64, true) else DG_CASE(64, false)
        case 128: DG_CASE(128, false)
        default: DG_CASE(256, false)
    }
#undef DG_CASE
}
