// bf16 GEMM on the 5th-generation tensor cores: tcgen05.mma (cta_group::1, M=128) with fp32 accumulators in TMEM,
// operands staged in shared memory by TMA (SWIZZLE_128B), warp-specialised persistent CTAs:
//   warp 0   TMA producer (one elected lane)         smem ring: full[]/empty[] mbarriers
//   warp 1   MMA issuer (one elected lane) + TMEM allocator
//   warp 2-5 epilogue: tcgen05.ld -> bias / erf-GELU / residual -> global, double-buffered TMEM accumulators so the
//            epilogue of tile i overlaps the main loop of tile i+1.
// Two orientations share the kernel:
//   normal : A = x [M][K], B = w [N][K]  ->  y[m][n]           (prefill / NAR / training shapes)
//   swap-AB: A = w [N][K], B = x [M<=256][K] -> part[s][m][n]  (decode: weight rows fill the 128-row MMA-M, the batch
//            sits on MMA-N, split-K slices are written as deterministic fp32 partials)
// Replaces the nn.Linear call sites modules.py:146,171,220-221 and valle_ar.py:158 / valle_nar.py:157.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;      // UMMA M (rows of the A operand per tile)
constexpr int BK = 64;       // k-block: 64 bf16 = one 128-byte swizzle atom row
constexpr int UMMA_K = 16;
// Epilogue warps: 4 (swap-AB), 8 (large-M), 12 for the CTA-pair kernel when it is compiled for ONE non-residual epilogue: without the
// residual prefetch registers the epilogue fits the 128 registers that 448 threads leave, and three warps per sub-partition instead
// of two lift the ~0.45 IPC that bounds the GELU epilogue.
constexpr int epi_warps(bool swap, bool wide = false) { return swap ? 4 : (wide ? 12 : 8); }
constexpr int num_threads(bool swap, bool wide = false) { return 64 + 32 * epi_warps(swap, wide); }
constexpr int STG_BYTES_PER_WARP = 32 * 128;   // epilogue staging: 32 rows x 32 fp32, 128B-swizzled
constexpr int stg_bufs(bool swap) { return swap ? 2 : 1; }   // the split-K slices leave by TMA store: two staging tiles in flight

struct GemmParams {
    int rows_a, rows_b, K;
    int tiles_a, tiles_b;      // tiles along A rows / B rows
    int kb_total, kb_per_split, n_split;
    int epilogue;
    const float* bias;
    const float* residual;
    int64_t ldr;
    void* y;
    int y_bf16;
    int64_t ldy;
    float* part;
    int64_t part_stride;
    int64_t part_ld;           // N (row pitch of a partial slice)
    int late_trigger;          // release the dependent kernel only after our own pdl_wait (see vb_linear_decode flags)
    int tma_store;             // swap-AB: the slices are written by TMA stores through tm_p (needs 16-byte aligned slice rows)
    unsigned long long* dbg;   // optional %globaltimer stamps [cta][8] (vb_linear_decode_set_debug)
    // VB_EPI_ARGMAX only: rng_on != 0 turns the pick into a Categorical draw (Gumbel-max): score = logit * inv_temp + Gumbel noise
    // from a counter-based hash of (rng_key, row, column)
    float inv_temp;
    int rng_on;
    unsigned long long rng_key;
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// stamps 0..6: %globaltimer (comparable across SMs, ~0.26 us resolution); the same events as SM cycle counts go to
// slots 8..14 (exact intervals inside one CTA)
#define DBG_STAMP(i) do { if (p.dbg) { p.dbg[blockIdx.x * 16 + (i)] = gtimer(); p.dbg[blockIdx.x * 16 + 8 + (i)] = clock64(); } } while (0)

// A_MN / B_MN: the operand is given TRANSPOSED in memory -- [K][rows] row-major, i.e. MN-major for the MMA (the backward
// GEMMs: dgrad reads W (N,K) as the MN-major B operand of dy.W, wgrad reads dy and x as MN-major A and B of dy^T.x, so no
// operand is transposed in HBM first).  Tile = 64-wide MN atoms of [64 k rows][64 columns], 8 KB each, 128B-swizzled by
// TMA; UMMA descriptor: start advances 2 KB per 16 k rows, LBO = atom pitch (8 KB), SBO = 8-row group pitch (1 KB).
// TWO: CTA pair (`cta_group::2`, cluster of two CTAs on one TPC).  The pair owns a 256 x BN output tile: each CTA loads its 128
// rows of A and HALF of the B tile (BN/2 rows) per k-block, the leader CTA issues ONE tcgen05.mma of M = 256 that reads both
// halves of B out of both CTAs' shared memory, and each CTA's tensor core accumulates its own 128 x BN rows in its own TMEM.
// Per SM and k-block that is 32 KB of TMA writes and operand reads instead of 48 KB -- the single-CTA kernel streams
// 94 B/clk/SM of operands at full MMA rate.  TMA loads of both CTAs complete on the LEADER's mbarrier, the MMA commits are
// multicast to both CTAs' barriers, the accumulator-empty barrier collects the epilogue warps of both CTAs.
template <int BN, int STAGES, bool SWAP, bool A_MN = false, bool B_MN = false, bool TWO = false, int EPI = -1>
__global__ void __launch_bounds__(num_threads(SWAP, TWO && EPI >= 0 && EPI != VB_EPI_BIAS_RESIDUAL), 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a,
                                                                 const __grid_constant__ CUtensorMap tm_b,
                                                                 const __grid_constant__ CUtensorMap tm_p, GemmParams p) {
    constexpr bool WIDE = TWO && EPI >= 0 && EPI != VB_EPI_BIAS_RESIDUAL;      // 12 epilogue warps (see epi_warps)
    constexpr int A_BYTES = BM * BK * 2;
    constexpr int B_BYTES = (TWO ? BN / 2 : BN) * BK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // NACC independent accumulators per tile (summed by the epilogue) were tried for the narrow decode tiles, on the theory
    // that four MMAs of a k-block accumulating into ONE TMEM tile run at MMA latency: measured no change (~400 cycles per
    // k-block either way, the issue rate of 128x32x16 MMAs) and +190 cycles of TMEM loads, so one accumulator it is.
    constexpr int NACC = 1;
    constexpr int ACC_COLS = NACC * BN;             // TMEM columns of one (multi-)accumulator; two of them are in flight
    constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
    static_assert(!(SWAP && (A_MN || B_MN)), "MN-major operands are for the large-M orientation");
    static_assert(!TWO || !SWAP, "the CTA-pair form exists for the large-M orientation");
    constexpr uint32_t IDESC = umma_idesc_bf16(TWO ? 2 * BM : BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;      // clears the CTA-rank bit of a shared::cluster address: the leader's copy
    uint32_t cta_rank = 0;
    if constexpr (TWO) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_b);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&tfull_bar[a]), 1);
            mbar_init(smem_u32(&tempty_bar[a]), (TWO ? 2 : 1) * epi_warps(SWAP, WIDE));
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        if constexpr (TWO) {      // the same warp of both CTAs allocates the same columns in both tensor memories
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "n"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            tmem_alloc<TMEM_COLS>(smem_u32(&tmem_base_slot));
        }
    }
    tc_fence_before();
    if constexpr (TWO) {          // barrier inits and the allocation must be visible to the peer CTA before anything remote happens
        asm volatile("barrier.cluster.arrive.release;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    } else {
        __syncthreads();
    }
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    const int epilogue = (EPI >= 0) ? EPI : p.epilogue;      // EPI >= 0: compiled for one epilogue (dead code and its registers go)
    const int total_tiles = p.tiles_a * p.tiles_b * p.n_split;
    // work distribution: a CTA (a CTA pair when TWO) takes every tile_step-th tile starting at tile_first
    const int tile_first = TWO ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int tile_step = TWO ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    if (threadIdx.x == 0) DBG_STAMP(0);                                   // prologue done

    if (!p.late_trigger) pdl_trigger();
    if (warp == 0) {
        if (elect_one()) {
            // The weight operand (A when SWAP, else B) is immutable, so its tiles are requested BEFORE waiting on the
            // predecessor kernel (PDL): the weight stream of this GEMM overlaps the tail of whatever runs before it.
            // The activation operand of the same ring slots is requested after pdl_wait(); each slot's mbarrier
            // expects the bytes of both tiles.
            auto tile_of = [&](int t, int& ta, int& tb, int& kb0, int& kb1) {
                int split;
                if (SWAP) { ta = t % p.tiles_a; const int r = t / p.tiles_a; tb = r % p.tiles_b; split = r / p.tiles_b; }
                else      { tb = t % p.tiles_b; ta = t / p.tiles_b; split = 0; }
                if constexpr (TWO) ta = ta * 2 + static_cast<int>(cta_rank);      // this CTA's 128-row block of the pair's 256 rows
                kb0 = split * p.kb_per_split;
                kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            };
            // CTA pair: both CTAs' loads complete on the leader's barrier, which expects the bytes of both
            auto tma2 = [&](uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
                asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(dst), "l"(map), "r"(bar & PEER_MASK), "r"(c0), "r"(c1) : "memory");
            };
            auto expect = [&](uint32_t bar) {
                if constexpr (TWO) { if (cta_rank == 0) mbar_expect_tx(bar, 2 * STAGE_BYTES); }
                else mbar_expect_tx(bar, STAGE_BYTES);
            };
            auto load_a = [&](uint32_t dst, uint32_t fb, int kb, int ta) {
                if constexpr (A_MN) {
#pragma unroll
                    for (int i = 0; i < BM / 64; ++i) {
                        if constexpr (TWO) tma2(dst + i * 8192, &tm_a, fb, ta * BM + i * 64, kb * BK);
                        else tma_load_2d(dst + i * 8192, &tm_a, fb, ta * BM + i * 64, kb * BK);
                    }
                } else if constexpr (TWO) {
                    tma2(dst, &tm_a, fb, kb * BK, ta * BM);
                } else {
                    tma_load_2d(dst, &tm_a, fb, kb * BK, ta * BM);
                }
            };
            auto load_b = [&](uint32_t dst, uint32_t fb, int kb, int tb) {
                if constexpr (B_MN) {
                    constexpr int COLS = TWO ? BN / 2 : BN;      // CTA pair: this CTA's half of the tile's columns
#pragma unroll
                    for (int i = 0; i < COLS / 64; ++i) {
                        if constexpr (TWO) tma2(dst + i * 8192, &tm_b, fb, tb * BN + static_cast<int>(cta_rank) * COLS + i * 64, kb * BK);
                        else tma_load_2d(dst + i * 8192, &tm_b, fb, tb * BN + i * 64, kb * BK);
                    }
                } else if constexpr (TWO) {
                    tma2(dst, &tm_b, fb, kb * BK, tb * BN + static_cast<int>(cta_rank) * (BN / 2));      // this CTA's half of the B tile
                } else {
                    tma_load_2d(dst, &tm_b, fb, kb * BK, tb * BN);
                }
            };
            int pre = 0;   // ring slots already armed with their weight tile
            for (int t = tile_first; t < total_tiles && pre < STAGES; t += tile_step) {
                int ta, tb, kb0, kb1;
                tile_of(t, ta, tb, kb0, kb1);
                for (int kb = kb0; kb < kb1 && pre < STAGES; ++kb, ++pre) {
                    const uint32_t fb = smem_u32(&full_bar[pre]);
                    expect(fb);
                    const uint32_t sa = smem_base + pre * STAGE_BYTES;
                    if (SWAP) load_a(sa, fb, kb, ta);
                    else      load_b(sa + A_BYTES, fb, kb, tb);
                }
            }
            DBG_STAMP(1);                                                     // weight tiles requested
            pdl_wait();
            DBG_STAMP(2);                                                     // dependency resolved
            if (p.late_trigger) pdl_trigger();
            int stage = 0, n = 0;
            uint32_t phase = 0;
            for (int t = tile_first; t < total_tiles; t += tile_step) {
                int ta, tb, kb0, kb1;
                tile_of(t, ta, tb, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb, ++n) {
                    const uint32_t fb = smem_u32(&full_bar[stage]);
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
                    if (n < pre) {      // weight tile already in flight: only the activation tile is missing
                        if (SWAP) load_b(sa + A_BYTES, fb, kb, tb);
                        else      load_a(sa, fb, kb, ta);
                    } else {
                        if (SWAP) mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                        else mbar_wait_relaxed(smem_u32(&empty_bar[stage]), phase ^ 1);
                        expect(fb);
                        load_a(sa, fb, kb, ta);
                        load_b(sa + A_BYTES, fb, kb, tb);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        else if (p.late_trigger) { pdl_wait(); pdl_trigger(); }
    } else if (warp == 1) {
        if (p.late_trigger) { pdl_wait(); pdl_trigger(); }
        if (elect_one() && (!TWO || cta_rank == 0)) {      // CTA pair: only the leader issues
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            // commits: to this CTA's barrier, or (CTA pair) multicast to the barrier at the same offset in both CTAs
            auto commit = [&](uint32_t bar) {
                if constexpr (TWO) {
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
                } else {
                    umma_commit(bar);
                }
            };
            for (int t = tile_first; t < total_tiles; t += tile_step, ++local) {
                const int split = SWAP ? t / (p.tiles_a * p.tiles_b) : 0;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
                const int acc = local & 1;
                const uint32_t acc_phase = (local >> 1) & 1;
                if (SWAP) mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
                else mbar_wait_relaxed(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (SWAP) mbar_wait(smem_u32(&full_bar[stage]), phase);
                    else mbar_wait_relaxed(smem_u32(&full_bar[stage]), phase);
                    if (kb == kb0) DBG_STAMP(3);                              // first k-block (weights + activations) landed
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                        const uint64_t da = A_MN ? umma_desc_sw128(sa + kk * UMMA_K * 128, 8192, 1024)
                                                 : umma_desc_sw128(sa + kk * UMMA_K * 2, 16, 1024);
                        const uint64_t db = B_MN ? umma_desc_sw128(sa + A_BYTES + kk * UMMA_K * 128, 8192, 1024)
                                                 : umma_desc_sw128(sa + A_BYTES + kk * UMMA_K * 2, 16, 1024);
                        if constexpr (TWO) {
                            asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                                         " tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                                         ::"r"(d_tmem), "l"(da), "l"(db), "r"(IDESC), "r"((kb > kb0 || kk > 0) ? 1u : 0u) : "memory");
                        } else {
                            umma_f16(d_tmem + (kk % NACC) * BN, da, db, IDESC, (kb > kb0 || kk >= NACC) ? 1u : 0u);
                        }
                    }
                    commit(smem_u32(&empty_bar[stage]));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                commit(smem_u32(&tfull_bar[acc]));
                DBG_STAMP(4);                                                 // last MMA of the tile issued
            }
        }
    } else {
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        int local = 0;
        pdl_wait();              // residual reads / output writes must not overtake the predecessor kernel
        if (p.late_trigger) pdl_trigger();
        for (int t = tile_first; t < total_tiles; t += tile_step, ++local) {
            int ta, tb, split;
            if (SWAP) { ta = t % p.tiles_a; const int r = t / p.tiles_a; tb = r % p.tiles_b; split = r / p.tiles_b; }
            else      { tb = t % p.tiles_b; ta = t / p.tiles_b; split = 0; }
            if constexpr (TWO) ta = ta * 2 + static_cast<int>(cta_rank);
            const int kb0 = split * p.kb_per_split;
            const bool empty_slice = kb0 >= p.kb_total;   // a split past the end of K: contributes zeros
            const int acc = local & 1;
            const uint32_t acc_phase = (local >> 1) & 1;
            // Residual epilogue (normal orientation): the fp32 residual rows come from HBM.  They are requested one 32-column
            // chunk AHEAD -- the first chunk of a tile even before its accumulator is complete -- so that their latency
            // overlaps the main loop / the previous chunk's TMEM load, staging and stores instead of sitting between them
            // (out-proj at M = 57 600: 532 TFLOP/s with the loads issued per chunk).
            float4 r_next[8];
            bool res_vec = false;
            // large-M orientation: a quadrant's warps split the tile's 32-column chunks among them (2 warps: 4 + 4, 3 warps: 2 + 3 + 3)
            constexpr int E_PARTS = epi_warps(SWAP, WIDE) / 4, E_CHUNKS = BN / 32;
            int e_c_begin = 0, e_c_end = 0, e_c4 = 0, e_rsub = 0;
            if (!SWAP) {
                const int part = (warp - 2) >> 2;
                e_c_begin = (part * E_CHUNKS / E_PARTS) * 32; e_c_end = ((part + 1) * E_CHUNKS / E_PARTS) * 32;
                e_c4 = lane & 7; e_rsub = lane >> 3;
                res_vec = (epilogue == VB_EPI_BIAS_RESIDUAL) && ((p.rows_b & 3) == 0) && ((p.ldy & 3) == 0) && ((p.ldr & 3) == 0);
            }
            auto load_res = [&](int c0, float4 (&r)[8]) {
                const int gcol = tb * BN + c0 + e_c4 * 4;
                const bool col_ok = gcol < p.rows_b;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int grow = ta * BM + q * 32 + it * 4 + e_rsub;
                    r[it] = (grow < p.rows_a && col_ok)
                                ? *reinterpret_cast<const float4*>(p.residual + static_cast<int64_t>(grow) * p.ldr + gcol)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            if (res_vec) load_res(e_c_begin, r_next);
            mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
            if (threadIdx.x == 64) DBG_STAMP(5);                              // accumulator complete
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * ACC_COLS;
            const int row = ta * BM + q * 32 + lane;  // row of the A operand owned by this thread
            if (SWAP) {
                // D[n][m]: this thread owns weight row n; columns are batch rows m0 + m of batch tile tb.
                const int m0 = tb * BN;
                float* dst = p.part + split * p.part_stride + row;
                constexpr int CH = (BN >= 32) ? 32 : 16;
                const int ew = warp - 2;
                float* stg = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + STAGES * STAGE_BYTES) + ew * (stg_bufs(SWAP) * 32 * 32);
                if (p.tma_store) {
                    // The warp's 32 weight rows x 32 batch rows of a chunk are transposed into a dense [m][n] staging tile and
                    // leave as ONE TMA store (box 32 x 32 of the [split][m][n] slice tensor, rows / columns past the end clipped
                    // by the tensor map): with 16-byte stores from the lanes the epilogue of a 128 x 128 tile took 3 530 cycles,
                    // ~50 cycles per store instruction (tools/gemm_timeline.py 256).  Two staging tiles per warp alternate.
                    int buf = 0, issued = 0;
#pragma unroll 1
                    for (int c0 = 0; c0 < BN; c0 += 32) {
                        if (m0 + c0 >= p.rows_b) break;
                        uint32_t v[32];
                        tmem_ld_32x32(t_addr + c0, v);
                        tmem_ld_wait();
                        if (threadIdx.x == 64 && c0 == 0) DBG_STAMP(6);           // accumulator in registers
                        float* sb = stg + buf * (32 * 32);
                        if (issued >= 2) {      // the store that last read this staging tile must have finished reading it
                            if (lane == 0) bulk_wait_group_read<1>();
                            __syncwarp();
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) sb[j * 32 + lane] = empty_slice ? 0.f : __uint_as_float(v[j]);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_3d(&tm_p, smem_u32(sb), ta * BM + q * 32, m0 + c0, split);
                            bulk_commit_group();
                        }
                        ++issued;
                        buf ^= 1;
                    }
                    if (lane == 0) bulk_wait_group_read<0>();
                    __syncwarp();
                } else {
                // vector path: the warp's 32 weight rows x 32 batch rows go through a shared-memory transpose so that each
                // store instruction writes 16 bytes per lane along n (8 lanes = one 128-byte row segment) -- four times fewer
                // store instructions than one 4-byte store per (n, m) (measured: ~50 cycles per store instruction)
                const bool vec = ((p.part_ld & 3) == 0) && ((p.part_stride & 3) == 0) && (ta * BM + q * 32 + 32 <= p.rows_a);
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += CH) {
                    if (m0 + c0 >= p.rows_b) break;
                    uint32_t v[CH];
                    if constexpr (CH == 32) tmem_ld_32x32(t_addr + c0, reinterpret_cast<uint32_t(&)[32]>(v));
                    else tmem_ld_32x16(t_addr + c0, reinterpret_cast<uint32_t(&)[16]>(v));
                    if constexpr (NACC > 1) {
#pragma unroll
                        for (int a = 1; a < NACC; ++a) {
                            uint32_t w[CH];
                            if constexpr (CH == 32) tmem_ld_32x32(t_addr + a * BN + c0, reinterpret_cast<uint32_t(&)[32]>(w));
                            else tmem_ld_32x16(t_addr + a * BN + c0, reinterpret_cast<uint32_t(&)[16]>(w));
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < CH; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
                        }
                    } else {
                        tmem_ld_wait();
                    }
                    if (threadIdx.x == 64 && c0 == 0) DBG_STAMP(6);           // accumulator in registers
                    if (vec) {
#pragma unroll
                        for (int j = 0; j < CH; ++j) stg[j * 32 + lane] = empty_slice ? 0.f : __uint_as_float(v[j]);
                        __syncwarp();
                        const int n4 = (lane & 7) * 4, msub = lane >> 3;
#pragma unroll
                        for (int it = 0; it < CH / 4; ++it) {
                            const int mm = it * 4 + msub, m = m0 + c0 + mm;
                            if (m < p.rows_b) {
                                const float4 val = *reinterpret_cast<const float4*>(stg + mm * 32 + n4);
                                *reinterpret_cast<float4*>(p.part + split * p.part_stride + static_cast<int64_t>(m) * p.part_ld + ta * BM + q * 32 + n4) = val;
                            }
                        }
                        __syncwarp();
                    } else if (row < p.rows_a) {
#pragma unroll
                        for (int j = 0; j < CH; ++j) {
                            const int m = m0 + c0 + j;
                            if (m < p.rows_b) dst[m * p.part_ld] = empty_slice ? 0.f : __uint_as_float(v[j]);
                        }
                    }
                }
                }
            } else if constexpr (EPI == VB_EPI_ARGMAX) {
                // Fused logits + greedy pick (vb_linear_argmax): thread = accumulator row; its warp's share of the tile's columns is
                // scanned straight out of tensor memory (ascending columns, strict '>' keeps the lowest column of a tie) and the
                // row's (value, column) maximum joins the other warps / tiles of the row through one 64-bit atomicMax.  Nothing
                // is staged and no logit is stored.
                // rng_on (vb_linear_categorical, valle_nar.py:160): the same scan over logit / temperature + Gumbel noise -- the arg-max
                // of that is an exact draw from Categorical(softmax(logits / temperature)) (Gumbel-max), so the sampled stage needs
                // no logits in HBM either.  Noise = -log(-log(u)), u from a counter-based hash of (key, row, column): a 64-bit
                // splitmix round per row, one 32-bit finaliser per logit; 24-bit uniforms strictly inside (0, 1).
                float best = 0.f;
                int best_col = -1;
                const int n_base = tb * BN;
                const bool noisy = p.rng_on != 0;
                uint32_t row_key = 0;
                if (noisy) {
                    unsigned long long z = p.rng_key + 0x9E3779B97F4A7C15ull * (static_cast<unsigned long long>(row) + 1ull);
                    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                    z ^= z >> 31;
                    row_key = static_cast<uint32_t>(z >> 32) ^ static_cast<uint32_t>(z);
                }
#pragma unroll 1
                for (int c0 = e_c_begin; c0 < e_c_end; c0 += 32) {
                    const int n0 = n_base + c0;
                    if (n0 >= p.rows_b) break;
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float val = __uint_as_float(v[j]);
                        if (noisy) {
                            uint32_t h = row_key ^ (static_cast<uint32_t>(n0 + j) * 0x9E3779B9u);
                            h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
                            const float u = (static_cast<float>(h >> 8) + 0.5f) * (1.f / 16777216.f);
                            // E = -log(u) ~ Exp(1).  The largest noise values (the winners) come from u next to 1, where the fast
                            // logarithm's ABSOLUTE error (2^-22) would swamp E: there E = -log1p(-w), w = 1 - u exact, by its series
                            const float w1 = 1.f - u;
                            const float e = (w1 < 0.015625f) ? w1 * (1.f + w1 * (0.5f + w1 * 0.33333334f)) : -__logf(u);
                            val = fmaf(val, p.inv_temp, -__logf(e));
                        }
                        if (n0 + j < p.rows_b && val > -INFINITY && (best_col < 0 || val > best)) { best = val; best_col = n0 + j; }
                    }
                }
                if (row < p.rows_a && best_col >= 0) {
                    uint32_t u = __float_as_uint(best);
                    if ((u << 1) == 0u) u = 0u;                           // -0.0 and +0.0 compare equal: one key for both
                    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // order-preserving map float -> unsigned
                    const unsigned long long key = (static_cast<unsigned long long>(u) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(best_col));
                    atomicMax(static_cast<unsigned long long*>(p.y) + row, key);
                }
            } else {
                // Normal orientation.  Each quadrant is served by two warps (column halves).  A 32-column chunk goes
                // TMEM -> registers (thread = row) -> a warp-private 128B-swizzled smem tile -> registers again with
                // lane = (row, 16-byte column group), so that residual loads and output stores are coalesced
                // (8 lanes cover one 128-byte row segment) instead of 32 scattered rows per instruction.
                const int ew = warp - 2;
                const uint32_t stg = smem_base + STAGES * STAGE_BYTES + ew * STG_BYTES_PER_WARP;
                const int n_base = tb * BN;
                const int row_base = ta * BM + q * 32;
                const bool vec_ok = ((p.rows_b & 3) == 0) && ((p.ldy & 3) == 0) && (epilogue != VB_EPI_BIAS_RESIDUAL || (p.ldr & 3) == 0);
                const int c4 = lane & 7, rsub = lane >> 3;
#pragma unroll 1
                for (int c0 = e_c_begin; c0 < e_c_end; c0 += 32) {
                    const int n0 = n_base + c0;
                    if (n0 >= p.rows_b) break;
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t a = stg + lane * 128 + (static_cast<uint32_t>(j ^ (lane & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                                     "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
                    }
                    __syncwarp();
                    const int gcol = n0 + c4 * 4;
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (epilogue != VB_EPI_NONE) {
                        if (gcol + 3 < p.rows_b) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gcol));
                        else {
                            if (gcol < p.rows_b) b4.x = p.bias[gcol];
                            if (gcol + 1 < p.rows_b) b4.y = p.bias[gcol + 1];
                            if (gcol + 2 < p.rows_b) b4.z = p.bias[gcol + 2];
                        }
                    }
                    // phase A: every shared and global LOAD of the chunk is issued before any store (y may alias the
                    // residual, so interleaving them would serialise the loads behind the stores)
                    float4 f[8], r4[8];
                    const bool col_ok = gcol < p.rows_b;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rr = it * 4 + rsub;
                        const uint32_t a = stg + rr * 128 + (static_cast<uint32_t>(c4 ^ (rr & 7)) << 4);
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[it].x), "=f"(f[it].y), "=f"(f[it].z), "=f"(f[it].w) : "r"(a));
                    }
                    if (res_vec) {
#pragma unroll
                        for (int it = 0; it < 8; ++it) r4[it] = r_next[it];
                        // next chunk of this warp's column half (different columns: no overlap with the stores below)
                        if (c0 + 32 < e_c_end && n0 + 32 < p.rows_b) load_res(c0 + 32, r_next);
                    }
                    // phase B: epilogue math + stores
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int grow = row_base + it * 4 + rsub;
                        if (grow >= p.rows_a || !col_ok) continue;
                        float4 v4 = f[it];
                        v4.x += b4.x; v4.y += b4.y; v4.z += b4.z; v4.w += b4.w;
                        if (epilogue == VB_EPI_BIAS_GELU) {
                            #ifdef VB_GELU_MUFU      // experiment: the Abramowitz-Stegun form with two MUFU ops per value
                            const float2 g01 = gelu_erf_fast2(make_float2(v4.x, v4.y)), g23 = gelu_erf_fast2(make_float2(v4.z, v4.w));
#else
                            const float2 g01 = gelu_erf_poly2(make_float2(v4.x, v4.y)), g23 = gelu_erf_poly2(make_float2(v4.z, v4.w));
#endif
                            v4 = make_float4(g01.x, g01.y, g23.x, g23.y);
                        }
                        if (vec_ok) {
                            if (epilogue == VB_EPI_BIAS_RESIDUAL) { v4.x += r4[it].x; v4.y += r4[it].y; v4.z += r4[it].z; v4.w += r4[it].w; }
                            if (p.y_bf16) {
                                *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.y) + static_cast<int64_t>(grow) * p.ldy + gcol) =
                                    make_uint2(pack_bf16x2(v4.x, v4.y), pack_bf16x2(v4.z, v4.w));
                            } else {
                                *reinterpret_cast<float4*>(static_cast<float*>(p.y) + static_cast<int64_t>(grow) * p.ldy + gcol) = v4;
                            }
                        } else {
                            const float fv[4] = {v4.x, v4.y, v4.z, v4.w};
                            for (int e = 0; e < 4; ++e) {
                                if (gcol + e >= p.rows_b) break;
                                float val = fv[e];
                                if (epilogue == VB_EPI_BIAS_RESIDUAL) val += p.residual[static_cast<int64_t>(grow) * p.ldr + gcol + e];
                                if (p.y_bf16) static_cast<__nv_bfloat16*>(p.y)[static_cast<int64_t>(grow) * p.ldy + gcol + e] = __float2bfloat16_rn(val);
                                else static_cast<float*>(p.y)[static_cast<int64_t>(grow) * p.ldy + gcol + e] = val;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (TWO) {      // the leader's barrier collects the epilogue warps of both CTAs
                    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(&tempty_bar[acc]) & PEER_MASK) : "memory");
                } else {
                    mbar_arrive(smem_u32(&tempty_bar[acc]));
                }
            }
            if (threadIdx.x == 64) DBG_STAMP(7);                              // epilogue stores issued, TMEM released
        }
    }
    tc_fence_before();
    if constexpr (TWO) {          // the peer may still be reading its accumulators / receiving multicast arrivals
        asm volatile("barrier.cluster.arrive.release;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    } else {
        __syncthreads();
    }
    if (warp == 1) {
        tc_fence_after();
        if constexpr (TWO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
        else tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

template <int BN, int STAGES, bool SWAP, bool A_MN = false, bool B_MN = false>
int launch_gemm_tc(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st, const CUtensorMap* tp = nullptr) {
    constexpr int SMEM = STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 + epi_warps(SWAP) * stg_bufs(SWAP) * STG_BYTES_PER_WARP;
    static bool configured = false;
    auto kern = gemm_tc_kernel<BN, STAGES, SWAP, A_MN, B_MN>;
    if (!configured) {
        VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    const int total = p.tiles_a * p.tiles_b * p.n_split;
    const int grid = min(total, vb_sm_count());
    static const CUtensorMap no_map{};
    VB_CUDA(vb_launch(SWAP, kern, dim3(grid), dim3(num_threads(SWAP)), SMEM, st, ta, tb, tp ? *tp : no_map, p));
    return VB_OK;
}

// CTA-pair launch: cluster of two CTAs per 256 x BN tile, one pair per SM pair
template <int BN, int STAGES, bool A_MN = false, bool B_MN = false, int EPI = -1>
int launch_gemm_tc_pair(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
    constexpr bool WIDE = EPI >= 0 && EPI != VB_EPI_BIAS_RESIDUAL;
    constexpr int SMEM = STAGES * (BM * BK * 2 + (BN / 2) * BK * 2) + 1024 + epi_warps(false, WIDE) * STG_BYTES_PER_WARP;
    static_assert(SMEM <= 232448 - 1024, "CTA-pair GEMM: ring + staging exceed the shared memory of an SM");
    static bool configured = false;
    auto kern = gemm_tc_kernel<BN, STAGES, false, A_MN, B_MN, true, EPI>;
    if (!configured) {
        VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    const int total = p.tiles_a * p.tiles_b;
    const int pairs = min(total, vb_sm_count() / 2);
    static const CUtensorMap no_map{};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(num_threads(false, WIDE));
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VB_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, no_map, p));
    return VB_OK;
}

}  // namespace

// y[M,N] = epi(x . w^T), bf16 operands
int vb_linear_tc(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, const float* residual,
                 int64_t ldr, void* y, int y_dtype, int64_t ldy, int64_t M, int64_t N, int64_t K, int epilogue,
                 cudaStream_t st, float inv_temp, int rng_on, unsigned long long rng_key) {
    VB_REQUIRE(K % 8 == 0, VB_ERR_UNSUPPORTED, "vb_linear(bf16): K must be a multiple of 8 (got %lld)", (long long)K);
    CUtensorMap ta, tb;
    int rc;
    GemmParams p{};
    p.rows_a = (int)M; p.rows_b = (int)N; p.K = (int)K;
    p.tiles_a = (int)vb_ceil_div(M, BM);
    p.kb_total = (int)vb_ceil_div(K, BK); p.kb_per_split = p.kb_total; p.n_split = 1;
    p.epilogue = epilogue; p.bias = bias; p.residual = residual; p.ldr = ldr;
    p.y = y; p.y_bf16 = (y_dtype == VB_BF16); p.ldy = ldy;
    p.inv_temp = inv_temp; p.rng_on = rng_on; p.rng_key = rng_key;
    if ((rc = vb_make_tmap_bf16_2d(&ta, x, M, K, ldx, BM, BK)) != VB_OK) return rc;
    if (N > 128) {
        p.tiles_b = (int)vb_ceil_div(N, 256);
        // CTA pairs (cta_group::2) for the large products: 256 x 256 tiles, each CTA of a pair loads half of the weight tile
        static const bool pair_on = !(getenv("VALLE_B200_GEMM_PAIR") != nullptr && getenv("VALLE_B200_GEMM_PAIR")[0] == '0');
        if (pair_on && M >= 1024) {
            p.tiles_a = (int)vb_ceil_div(M, 2 * BM);
            if ((rc = vb_make_tmap_bf16_2d(&tb, w, N, K, ldw, 128, BK)) != VB_OK) return rc;
            // one instantiation per non-residual epilogue with 12 epilogue warps (5-stage ring pays for their staging tiles)
            if (epilogue == VB_EPI_NONE) return launch_gemm_tc_pair<256, 5, false, false, VB_EPI_NONE>(ta, tb, p, st);
            if (epilogue == VB_EPI_BIAS) return launch_gemm_tc_pair<256, 5, false, false, VB_EPI_BIAS>(ta, tb, p, st);
            if (epilogue == VB_EPI_BIAS_GELU) return launch_gemm_tc_pair<256, 5, false, false, VB_EPI_BIAS_GELU>(ta, tb, p, st);
            if (epilogue == VB_EPI_ARGMAX) return launch_gemm_tc_pair<256, 5, false, false, VB_EPI_ARGMAX>(ta, tb, p, st);
            return launch_gemm_tc_pair<256, 6, false, false, VB_EPI_BIAS_RESIDUAL>(ta, tb, p, st);
        }
        VB_REQUIRE(epilogue != VB_EPI_ARGMAX, VB_ERR_UNSUPPORTED, "vb_linear_argmax: needs M >= 1024 (got %lld)", (long long)M);
        if ((rc = vb_make_tmap_bf16_2d(&tb, w, N, K, ldw, 256, BK)) != VB_OK) return rc;
        return launch_gemm_tc<256, 4, false>(ta, tb, p, st);
    }
    VB_REQUIRE(epilogue != VB_EPI_ARGMAX, VB_ERR_UNSUPPORTED, "vb_linear_argmax: needs N > 128 (got %lld)", (long long)N);
    if (N > 64) {
        p.tiles_b = 1;
        if ((rc = vb_make_tmap_bf16_2d(&tb, w, N, K, ldw, 128, BK)) != VB_OK) return rc;
        return launch_gemm_tc<128, 6, false>(ta, tb, p, st);
    }
    p.tiles_b = 1;
    if ((rc = vb_make_tmap_bf16_2d(&tb, w, N, K, ldw, 64, BK)) != VB_OK) return rc;
    return launch_gemm_tc<64, 8, false>(ta, tb, p, st);
}

// y[M,N] = epi(X . W^T) with operands optionally given transposed: x_mn: x is [K][M] (pitch ldx), w_mn: w is [K][N] (pitch ldw)
int vb_linear_tc_t(const void* x, int64_t ldx, int x_mn, const void* w, int64_t ldw, int w_mn, const float* bias,
                   const float* residual, int64_t ldr, void* y, int y_dtype, int64_t ldy, int64_t M, int64_t N, int64_t K,
                   int epilogue, cudaStream_t st) {
    if (!x_mn && !w_mn) return vb_linear_tc(x, ldx, w, ldw, bias, residual, ldr, y, y_dtype, ldy, M, N, K, epilogue, st, 1.f, 0, 0ull);
    VB_REQUIRE(ldx % 8 == 0 && ldw % 8 == 0, VB_ERR_UNSUPPORTED, "vb_linear_t: operand pitches must be multiples of 8");
    CUtensorMap ta, tb;
    int rc;
    GemmParams p{};
    p.rows_a = (int)M; p.rows_b = (int)N; p.K = (int)K;
    p.tiles_a = (int)vb_ceil_div(M, BM);
    p.tiles_b = (int)vb_ceil_div(N, 256);
    p.kb_total = (int)vb_ceil_div(K, BK); p.kb_per_split = p.kb_total; p.n_split = 1;
    p.epilogue = epilogue; p.bias = bias; p.residual = residual; p.ldr = ldr;
    p.y = y; p.y_bf16 = (y_dtype == VB_BF16); p.ldy = ldy;
    // MN-major operand: the tensor map runs over the [K][rows] matrix, boxes of 64 k rows x 64 columns
    if (x_mn) rc = vb_make_tmap_bf16_2d(&ta, x, K, M, ldx, BK, 64);
    else rc = vb_make_tmap_bf16_2d(&ta, x, M, K, ldx, BM, BK);
    if (rc != VB_OK) return rc;
    if (w_mn) rc = vb_make_tmap_bf16_2d(&tb, w, K, N, ldw, BK, 64);
    else rc = vb_make_tmap_bf16_2d(&tb, w, N, K, ldw, 256, BK);
    if (rc != VB_OK) return rc;
    static const bool pair_on = !(getenv("VALLE_B200_GEMM_PAIR") != nullptr && getenv("VALLE_B200_GEMM_PAIR")[0] == '0');
    if (pair_on && M >= 1024) {      // CTA pairs: 256 x 256 tiles, half of the B tile per CTA (K-major B: 128-row boxes)
        p.tiles_a = (int)vb_ceil_div(M, 2 * BM);
        if (!w_mn && (rc = vb_make_tmap_bf16_2d(&tb, w, N, K, ldw, 128, BK)) != VB_OK) return rc;
        if (epilogue == VB_EPI_NONE) {      // the backward GEMMs: compiled for the plain epilogue, 12 epilogue warps
            if (x_mn && w_mn) return launch_gemm_tc_pair<256, 5, true, true, VB_EPI_NONE>(ta, tb, p, st);
            if (w_mn) return launch_gemm_tc_pair<256, 5, false, true, VB_EPI_NONE>(ta, tb, p, st);
            return launch_gemm_tc_pair<256, 5, true, false, VB_EPI_NONE>(ta, tb, p, st);
        }
        if (x_mn && w_mn) return launch_gemm_tc_pair<256, 6, true, true>(ta, tb, p, st);
        if (w_mn) return launch_gemm_tc_pair<256, 6, false, true>(ta, tb, p, st);
        return launch_gemm_tc_pair<256, 6, true, false>(ta, tb, p, st);
    }
    if (x_mn && w_mn) return launch_gemm_tc<256, 4, false, true, true>(ta, tb, p, st);
    if (w_mn) return launch_gemm_tc<256, 4, false, false, true>(ta, tb, p, st);
    return launch_gemm_tc<256, 4, false, true, false>(ta, tb, p, st);
}

// split K so that (N/128 slabs) x splits fills the SMs, with at least 2 k-blocks (128 columns of K) per slice
// Above 128 batch rows the batch is tiled too (128 rows per tile, MMA-N = 128) and the split count is the largest whose
// (weight tile x batch tile x split) grid still fits on the SMs in ONE wave: half the slice bytes to write and re-read, a 64 KB
// instead of a 128 KB epilogue per CTA, no CTA with two tiles.  Decode step at batch 136 / 160 / 192 / 224 / 256: 1.254 / 1.410 /
// 1.667 / 1.865 / 2.058 ms against 1.303 / 1.471 / 1.735 / 1.956 / 2.179 ms with one 256-row tile (profiles/r02c_ab_decode_batch_tiles.jsonl).
static int decode_batch_tiles(int64_t M) { return M > 128 ? (int)vb_ceil_div(M, 128) : 1; }
extern "C" int vb_linear_decode_splits_m(int64_t M, int64_t N, int64_t K, int max_split) {
    const int tb = decode_batch_tiles(M);
    if (tb == 1) return vb_linear_decode_splits(N, K, max_split);
    // tiled batch: every (weight tile, batch tile, split) must get its own CTA -- with 1.3 waves the last epilogue ends 2 us
    // after the median one (tools/gemm_timeline.py 256) -- so the largest split count whose tiles still fit on the SMs
    const int tiles = (int)vb_ceil_div(N, BM) * tb, kb_total = (int)vb_ceil_div(K, BK);
    int n_split = max(1, min(min(vb_sm_count() / tiles, max_split), max(1, kb_total / 2)));
    for (;;) {
        const int kb_per_split = (int)vb_ceil_div(kb_total, n_split);
        const int eff = (int)vb_ceil_div(kb_total, kb_per_split);
        if (eff * tiles <= vb_sm_count() || n_split == 1) return eff;
        --n_split;
    }
}
extern "C" int vb_linear_decode_splits(int64_t N, int64_t K, int max_split) {      // batch <= 128
    const int tiles_a = (int)vb_ceil_div(N, BM), kb_total = (int)vb_ceil_div(K, BK);
    const int want = (int)vb_ceil_div(vb_sm_count(), tiles_a);
    int n_split = max(1, min(min(want, max_split), max(1, kb_total / 2)));
    const int kb_per_split = (int)vb_ceil_div(kb_total, n_split);
    return (int)vb_ceil_div(kb_total, kb_per_split);
}

static unsigned long long* g_gemm_dbg = nullptr;
extern "C" int vb_linear_decode_set_debug(void* buf) {   /* device buffer of #SM * 8 uint64 stamps, or NULL */
    g_gemm_dbg = static_cast<unsigned long long*>(buf);
    return VB_OK;
}

extern "C" int vb_linear_decode(const void* x, int64_t ldx, const void* w, int64_t ldw, float* part, int64_t part_stride,
                                int64_t M, int64_t N, int64_t K, int max_split, int flags, int* n_split_out, void* stream) {
    VB_REQUIRE(x && w && part, VB_ERR_BAD_ARG, "vb_linear_decode: null pointer");
    VB_REQUIRE(M >= 1 && M <= 1024, VB_ERR_UNSUPPORTED, "vb_linear_decode: M must be in [1,1024] (got %lld)", (long long)M);
    VB_REQUIRE(K % 8 == 0 && N >= 1, VB_ERR_UNSUPPORTED, "vb_linear_decode: K %% 8 != 0 or N < 1");
    VB_REQUIRE(max_split >= 1, VB_ERR_BAD_ARG, "vb_linear_decode: max_split must be >= 1");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GemmParams p{};
    p.rows_a = (int)N; p.rows_b = (int)M; p.K = (int)K;
    p.tiles_a = (int)vb_ceil_div(N, BM); p.tiles_b = decode_batch_tiles(M);
    p.kb_total = (int)vb_ceil_div(K, BK);
    const int n_split = vb_linear_decode_splits_m(M, N, K, max_split);
    p.kb_per_split = (int)vb_ceil_div(p.kb_total, n_split);
    p.n_split = n_split;
    p.part = part; p.part_stride = part_stride; p.part_ld = N;
    p.late_trigger = (flags & VB_FLAG_LATE_TRIGGER) ? 1 : 0;
    p.dbg = g_gemm_dbg;

    VB_REQUIRE(n_split == 1 || part_stride >= M * N, VB_ERR_BAD_ARG, "vb_linear_decode: part_stride too small");
    if (n_split_out) *n_split_out = n_split;
    CUtensorMap ta, tb, tp;
    int rc;
    if ((rc = vb_make_tmap_bf16_2d(&ta, w, N, K, ldw, BM, BK)) != VB_OK) return rc;
    // slices [split][m][n] as a TMA store destination (32 x 32 boxes) when their rows are 16-byte aligned
    p.tma_store = (M > 16 && (N & 3) == 0 && (part_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(part) & 15) == 0) ? 1 : 0;
    if (p.tma_store && (rc = vb_make_tmap_f32_3d(&tp, part, N, M, n_split, N, n_split > 1 ? part_stride : M * N, 32, 32)) != VB_OK) return rc;
#define DECODE_CASE(BNV, ST)                                                             \
    {                                                                                    \
        if ((rc = vb_make_tmap_bf16_2d(&tb, x, M, K, ldx, BNV, BK)) != VB_OK) return rc; \
        return launch_gemm_tc<BNV, ST, true>(ta, tb, p, st, p.tma_store ? &tp : nullptr); \
    }
    if (M <= 16) DECODE_CASE(16, 4)
    if (M <= 32) DECODE_CASE(32, 4)
    if (M <= 64) DECODE_CASE(64, 4)
    DECODE_CASE(128, 4)      /* M <= 128, or two 128-row batch tiles */
#undef DECODE_CASE
}
