// Flash-style attention on the 5th-generation tensor cores for the large-M paths (AR prefill, NAR stages, teacher-forced
// forward): S = Q K^T and O += P V are tcgen05.mma with accumulators in TMEM, operands staged by TMA (SWIZZLE_128B)
// straight out of the packed qkv buffer [B*S][3][H][64] written by the QKV GEMM -- no head-split copy, no
// materialised (B,H,S,S) mask: the prefix-LM / padding predicate is evaluated from (x_len, kv_len) per element.
//
// CTA = 128 query rows of one (batch, head); key blocks of 64; 320 threads, two CTAs per SM.
//   warp 0    Q tile + 4-deep ring of K blocks by TMA, and the Q K^T issuer: QK_t (M128 N64 K64) into score buffer t%2 as soon as
//             the row threads have S_{t-2} in registers -- the scores run up to two blocks ahead
//   warp 1    3-deep ring of V blocks by TMA, the P V issuer and the TMEM allocator.  P V_t reads its A operand P_t from TENSOR
//             MEMORY (tcgen05.mma with [a_tmem]) and V_t from shared memory (MN-major B); O accumulates in TMEM.
//             One control thread for both MMA streams needed ~1100 cycles per block (four mbarrier waits at ~90 cycles each
//             plus a serial instruction stream on a busy sub-partition) and S_t arrived late.
//   warp 2-5  softmax of the EVEN key blocks, one thread per query row (tcgen05.ld 32x32b gives a thread its whole row)
//   warp 6-9  softmax of the ODD key blocks, same rows (warps w and w+4 share a TMEM lane quadrant and an SM sub-partition)
// Shared-memory bandwidth is what a tcgen05.mma with small N pays for: M128 N64 K16 reads 4 KB of A + 2 KB of B = 48 cycles at
// 128 B/clk for 32 cycles of math (tools/umma_probe.cu).  With P staged in shared memory a key block moved 80 KB through that
// pipe (48 KB operand reads + 16 KB P stores + 16 KB TMA writes = 640 of the 704 cycles a block took); P through TMEM leaves
// 48 KB.  Beyond that the kernel is bound by the MUFU pipe (64 exp2 per row and block = 525 pipe cycles) AND by the
// per-block latencies of a row thread (mbarrier waits ~90 cycles each, TMEM load, proxy fence: ~800 cycles with the MUFU pipe
// idle).  When all row warps of a sub-partition run their exponentials at the same time they queue on the pipe and then idle
// together (measured: 45 % MUFU utilisation); so the two halves are kept in ANTI-PHASE by a ping-pong of named barriers: a warp
// starts the exponentials of block t only when its partner has finished those of block t-1, and does its waits, loads, row
// maximum and stores under the partner's exponentials.
// The softmax reference m_ref is shared by the two halves and handed on in block order through shared memory (the ping-pong
// barrier orders it); it is raised lazily -- only when a block maximum exceeds it by 2^8 -- and only then O (in TMEM) is rescaled
// by the thread that raised it, so p = exp2(s - m_ref) <= 256 and no per-block traffic on O is needed.
// exp2 offload (compile-time, off): POLY_OF_8 of every 8 score pairs on the FMA pipe (Cody-Waite + degree-3 minimax, 7.5e-5).
// Replaces F.scaled_dot_product_attention + merge_masks on modules.py:160-167 for S > 1.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

#ifndef VB_FWD_POLY
#define VB_FWD_POLY 0          // of every 8 score pairs, how many take exp2 on the FMA pipe
#endif

constexpr int BQ = 128;    // query rows per CTA
constexpr int BKV = 64;    // keys per block
constexpr int DH = 64;
constexpr int K_STAGES = 4;   // K blocks: a stage is free as soon as Q K_t^T has completed
constexpr int V_STAGES = 3;   // V blocks: free after P_t V_t
constexpr int POLY_OF_8 = VB_FWD_POLY;
#ifndef VB_FWD_RELEASE_PART
#define VB_FWD_RELEASE_PART 1
#endif
constexpr int RELEASE_PART = VB_FWD_RELEASE_PART;   // after which 32-column part of its exponentials a row warp releases its partner
constexpr int THREADS = 320;
constexpr int Q_BYTES = BQ * DH * 2;          // 16 KB
constexpr int KV_BYTES = BKV * DH * 2;        // 8 KB per K or V block
constexpr int SMEM_BYTES = Q_BYTES + (K_STAGES + V_STAGES) * KV_BYTES + 1024;
constexpr int O_COL = 2 * BKV;                // score buffers: cols [0, 64) even blocks, [64, 128) odd blocks; O: [128, 192)
constexpr int P_COL = O_COL + DH;             // P tiles (bf16 pairs, 32 columns per 64 keys): [192, 224) even, [224, 256) odd blocks
constexpr int TMEM_COLS = 256;

// three-input maximum (FMNMX3 on sm_100)
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// exp2 of two values on the FMA pipe: x = n + f with n = round(x) (adding 1.5 * 2^23 leaves n in the low mantissa bits),
// 2^f by a degree-3 minimax polynomial on [-0.5, 0.5], n added into the exponent field.  x is clamped to >= -126.
__device__ __forceinline__ float2 poly_exp2(float2 x) {
    const float2 magic = make_float2(12582912.f, 12582912.f);
    x.x = fmaxf(x.x, -126.f);
    x.y = fmaxf(x.y, -126.f);
    const float2 r = __fadd2_rn(x, magic);
    const float2 n = __fadd2_rn(r, make_float2(-12582912.f, -12582912.f));
    const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
    float2 p = __ffma2_rn(make_float2(0.055171460f, 0.055171460f), f, make_float2(0.24261086f, 0.24261086f));
    p = __ffma2_rn(p, f, make_float2(0.69326097f, 0.69326097f));
    p = __ffma2_rn(p, f, make_float2(0.99992812f, 0.99992812f));
    return make_float2(__uint_as_float(__float_as_uint(p.x) + (__float_as_uint(r.x) << 23)),
                       __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(r.y) << 23)));
}

// p = exp2(s * scale - m) for 32 scores of a row, their sum, bf16 pack (pk[i] = columns 2i, 2i+1)
template <bool POLY>
__device__ __forceinline__ float softmax_32(const uint32_t (&sv)[32], uint32_t (&pk)[16], float scale_log2e, float m_use) {
    const float2 sc2 = make_float2(scale_log2e, scale_log2e), nm2 = make_float2(-m_use, -m_use);
    float2 rs_a = make_float2(0.f, 0.f), rs_b = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
        const float2 t = __ffma2_rn(make_float2(__uint_as_float(sv[c]), __uint_as_float(sv[c + 1])), sc2, nm2);
        float2 pp;
        if (POLY && ((c >> 1) & 7) >= 8 - POLY_OF_8) pp = poly_exp2(t);
        else pp = make_float2(fast_exp2(t.x), fast_exp2(t.y));
        if (c & 2) rs_b = __fadd2_rn(rs_b, pp);
        else rs_a = __fadd2_rn(rs_a, pp);
        pk[c >> 1] = pack_bf16x2(pp.x, pp.y);
    }
    const float2 rs = __fadd2_rn(rs_a, rs_b);
    return rs.x + rs.y;
}

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows x 16 bf16 = 8 columns of bf16 pairs) is read from tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor (SWIZZLE_128B, SBO 1024) from its low word: start address >> 4 | LBO >> 4 << 16
__device__ __forceinline__ uint64_t desc_from_lo(uint32_t lo) { return (static_cast<uint64_t>(0x40004040u) << 32) | lo; }
#ifdef VB_FWD_DEBUG_WAIT      // experiment build: a wait that does not complete reports where it is and traps
#include <stdio.h>
__device__ __forceinline__ void dbg_wait(uint32_t bar, uint32_t parity, int code, int t) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 17)) {
            printf("attn_prefill wait %d stuck: cta (%d,%d,%d) thread %d block %d parity %u\n", code, blockIdx.x, blockIdx.y, blockIdx.z,
                   threadIdx.x, t, parity);
            __trap();
        }
    }
}
#define WAIT(bar, parity, code, t) dbg_wait(bar, parity, code, t)
#else
#define WAIT(bar, parity, code, t) mbar_wait(bar, parity)
#endif
// named barriers over the two warps of a pair (64 threads): full sync, or producer arrive / consumer sync
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void pair_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(THREADS, 2) attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap tm_q,
                                                                     const __grid_constant__ CUtensorMap tm_kv,
                                                                     __nv_bfloat16* __restrict__ o, int S, int H, int mask_mode,
                                                                     const int32_t* __restrict__ x_lens,
                                                                     const int32_t* __restrict__ kv_lens, float scale_log2e,
                                                                     float* __restrict__ lse, long long* __restrict__ dbg, int dbg_thread) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_q, bar_done, s_full[2], s_free[2], p_full[2], p_free[2];
    __shared__ __align__(8) uint64_t k_full[K_STAGES], k_empty[K_STAGES], v_full[V_STAGES], v_empty[V_STAGES];
    __shared__ float xch[2][BQ];      // [block parity][row]: softmax reference after that block
    __shared__ float fin_m[2][BQ], fin_l[2][BQ];      // [half][row]: final reference and row sum of each half
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int d = H * DH;
    pdl_trigger();
    const int i0 = qt * BQ;
    long long* dbg_cta = dbg == nullptr ? nullptr
        : dbg + (static_cast<int64_t>(blockIdx.z) * gridDim.y * gridDim.x + blockIdx.y * gridDim.x + blockIdx.x) * 32 * 8;
    if (dbg != nullptr && threadIdx.x == 64) {      // kernel entry of this CTA: %globaltimer and the SM it runs on
        unsigned long long gt; unsigned sm;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        dbg_cta[30 * 8 + 0] = static_cast<long long>(gt);
        dbg_cta[30 * 8 + 1] = sm;
    }

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_smem = base;
    const uint32_t k_smem = base + Q_BYTES;
    const uint32_t v_smem = k_smem + K_STAGES * KV_BYTES;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        mbar_init(smem_u32(&bar_q), 1);
        mbar_init(smem_u32(&bar_done), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&s_full[s]), 1);
            mbar_init(smem_u32(&s_free[s]), 4);
            mbar_init(smem_u32(&p_full[s]), 4);
            mbar_init(smem_u32(&p_free[s]), 1);
        }
        for (int s = 0; s < K_STAGES; ++s) {
            mbar_init(smem_u32(&k_full[s]), 1);
            mbar_init(smem_u32(&k_empty[s]), 1);
        }
        for (int s = 0; s < V_STAGES; ++s) {
            mbar_init(smem_u32(&v_full[s]), 1);
            mbar_init(smem_u32(&v_empty[s]), 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(&tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int row0 = b * S;   // first row of this batch in the packed [B*S][3d] matrix
    pdl_wait();               // qkv / lens are produced by earlier kernels; the set-up above overlapped their tail
    // block 0 is always needed: request Q, K_0 and V_0 before the sequence lengths are even read
    if (threadIdx.x == 0) {
        mbar_expect_tx(smem_u32(&bar_q), Q_BYTES);
        tma_load_2d(q_smem, &tm_q, smem_u32(&bar_q), h * DH, row0 + i0);
        mbar_expect_tx(smem_u32(&k_full[0]), KV_BYTES);
        tma_load_2d(k_smem, &tm_kv, smem_u32(&k_full[0]), d + h * DH, row0);
        mbar_expect_tx(smem_u32(&v_full[0]), KV_BYTES);
        tma_load_2d(v_smem, &tm_kv, smem_u32(&v_full[0]), 2 * d + h * DH, row0);
    }
    const int kv_len = kv_lens ? min(kv_lens[b], S) : S;
    const int x_len = (mask_mode == VB_MASK_PREFIX_LM) ? x_lens[b] : 0;
    // keys any row of this tile may attend: everything (no mask) or the text prefix plus the causal part
    int k_end = kv_len;
    if (mask_mode == VB_MASK_PREFIX_LM) k_end = min(kv_len, max(x_len, i0 + BQ));
    const int nb = max(1, (k_end + BKV - 1) / BKV);

    // descriptor low words; a step of 16 along K is +32 bytes (K-major) or +16 rows of 128 bytes (MN-major V)
    const uint32_t q_lo = ((q_smem & 0x3ffffu) >> 4) | (1u << 16);
    const uint32_t k_lo = ((k_smem & 0x3ffffu) >> 4) | (1u << 16);
    const uint32_t v_lo = ((v_smem & 0x3ffffu) >> 4) | (64u << 16);
    if (warp == 0) {
        // K producer + Q K^T issuer.  MMA-warp stamps (debug slots 4, 5 of block t): {operands + score buffer ready, issued}
        if (elect_one()) {
            constexpr uint32_t IDESC_QK = umma_idesc_bf16(BQ, BKV, 0, 0);   // A = Q (K-major), B = K block (K-major)
            auto load_k = [&](int t) {
                const int st = t % K_STAGES;
                mbar_expect_tx(smem_u32(&k_full[st]), KV_BYTES);
                tma_load_2d(k_smem + st * KV_BYTES, &tm_kv, smem_u32(&k_full[st]), d + h * DH, row0 + t * BKV);
            };
            for (int t = 1; t < min(nb, K_STAGES); ++t) load_k(t);
            WAIT(smem_u32(&bar_q), 0, 1, 0);
            for (int t = 0; t < nb; ++t) {
                const int st = t % K_STAGES;
                WAIT(smem_u32(&k_full[st]), (t / K_STAGES) & 1, 2, t);
                // score buffer t%2 last held S_{t-2}: free once its four row warps have it in registers
                if (t >= 2) WAIT(smem_u32(&s_free[t & 1]), ((t >> 1) - 1) & 1, 3, t);
                tc_fence_after();
                if (dbg_cta && t < 30) dbg_cta[t * 8 + 4] = clock64();
                const uint32_t kl = k_lo + st * (KV_BYTES >> 4);
                const uint32_t s_tmem = tmem_base + (t & 1) * BKV;
#pragma unroll
                for (int kk = 0; kk < DH / 16; ++kk)
                    umma_f16(s_tmem, desc_from_lo(q_lo + kk * 2), desc_from_lo(kl + kk * 2), IDESC_QK, kk > 0 ? 1u : 0u);
                umma_commit(smem_u32(&s_full[t & 1]));
                umma_commit(smem_u32(&k_empty[st]));
                if (dbg_cta && t < 30) dbg_cta[t * 8 + 5] = clock64();
                // refill the stage of the PREVIOUS block (its Q K^T was issued one iteration ago and has completed by now)
                if (t >= 1 && t - 1 + K_STAGES < nb) {
                    WAIT(smem_u32(&k_empty[(t - 1) % K_STAGES]), ((t - 1) / K_STAGES) & 1, 10, t);
                    load_k(t - 1 + K_STAGES);
                }
            }
        }
    } else if (warp == 1) {
        // V producer + P V issuer.  Stamps (debug slots 6, 7 of block t): {P ready, issued}
        if (elect_one()) {
            constexpr uint32_t IDESC_PV = umma_idesc_bf16(BQ, DH, 0, 1);    // A = P (K-major), B = V block (MN-major)
            const uint32_t o_tmem = tmem_base + O_COL;
            auto load_v = [&](int t) {
                const int st = t % V_STAGES;
                mbar_expect_tx(smem_u32(&v_full[st]), KV_BYTES);
                tma_load_2d(v_smem + st * KV_BYTES, &tm_kv, smem_u32(&v_full[st]), 2 * d + h * DH, row0 + t * BKV);
            };
            for (int t = 1; t < min(nb, V_STAGES); ++t) load_v(t);
            for (int t = 0; t < nb; ++t) {
                const int st = t % V_STAGES, pb = t & 1;
                WAIT(smem_u32(&v_full[st]), (t / V_STAGES) & 1, 4, t);
                WAIT(smem_u32(&p_full[pb]), (t >> 1) & 1, 5, t);
                tc_fence_after();
                if (dbg_cta && t < 30) dbg_cta[t * 8 + 6] = clock64();
                const uint32_t vl = v_lo + st * (KV_BYTES >> 4), p_tmem = tmem_base + P_COL + pb * (BKV / 2);
#pragma unroll
                for (int kk = 0; kk < BKV / 16; ++kk)
                    umma_f16_ts(o_tmem, p_tmem + kk * 8, desc_from_lo(vl + kk * 128), IDESC_PV, (t > 0 || kk > 0) ? 1u : 0u);
                umma_commit(smem_u32(&v_empty[st]));
                umma_commit(smem_u32(&p_free[pb]));
                if (t == nb - 1) umma_commit(smem_u32(&bar_done));
                if (dbg_cta && t < 30) dbg_cta[t * 8 + 7] = clock64();
                if (t >= 1 && t - 1 + V_STAGES < nb) {
                    WAIT(smem_u32(&v_empty[(t - 1) % V_STAGES]), ((t - 1) / V_STAGES) & 1, 11, t);
                    load_v(t - 1 + V_STAGES);
                }
            }
        }
    } else {
        const int q = warp & 3;               // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;     // 0: even key blocks, 1: odd key blocks
        const int r = q * 32 + lane;          // row within the tile == TMEM lane
        const int i = i0 + r;                 // query index within the sequence
        // ping-pong barriers of the pair (ids 1..8): `mine` = "my exponentials of this block are done (and my m_ref is
        // published)", arrived by me, awaited by the partner before ITS exponentials; `theirs` the other way round
        const int bar_mine = 1 + half * 4 + q, bar_theirs = 1 + (half ^ 1) * 4 + q;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        float m_ref = -INFINITY, l_run = 0.f;      // l_run: this thread's blocks only, relative to m_ref
        // optional per-block cycle stamps of ONE row thread of every CTA (vb_attention_prefill_set_debug), slots 0..3 of block t:
        // {S ready, row maximum known + partner's exponentials done, exp2 / pack done, P written}
        const long long t_entry = clock64();
        const bool stamp = dbg != nullptr && threadIdx.x == dbg_thread;
#pragma unroll 1
        for (int t = half; t < nb; t += 2) {
            const int u = t >> 1;                 // this half's block counter
            WAIT(smem_u32(&s_full[half]), u & 1, 6, t);
            tc_fence_after();
            if (stamp && t < 30) dbg_cta[t * 8 + 0] = clock64();
            uint32_t sv[64];
            {
                uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[0]);
                uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[32]);
                tmem_ld_32x32(lane_addr + half * BKV, lo);
                tmem_ld_32x32(lane_addr + half * BKV + 32, hi);
                tmem_ld_wait();
            }
            // S_t is in registers: the MMA warp may overwrite the buffer with S_{t+2}
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_free[half]));
            const int kbase = t * BKV;
            bool need_mask = (kbase + BKV > kv_len);
            if (mask_mode == VB_MASK_PREFIX_LM)
                need_mask = need_mask || !((kbase + BKV <= x_len) || (i0 >= x_len && kbase + BKV - 1 <= i0));
            if (need_mask) {
#pragma unroll
                for (int c = 0; c < 64; ++c) {
                    const int kj = kbase + c;
                    bool ok = kj < kv_len;
                    if (mask_mode == VB_MASK_PREFIX_LM) ok = ok && ((kj < x_len) || (i >= x_len && kj <= i));
                    if (!ok) sv[c] = 0xff800000u;   // -inf -> p = 0
                }
            }
            // maximum of the block's raw scores (scale > 0 commutes with max)
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int c = 0; c < 64; c += 8) {
                mx0 = fmax3(mx0, __uint_as_float(sv[c]), __uint_as_float(sv[c + 1]));
                mx1 = fmax3(mx1, __uint_as_float(sv[c + 2]), __uint_as_float(sv[c + 3]));
                mx2 = fmax3(mx2, __uint_as_float(sv[c + 4]), __uint_as_float(sv[c + 5]));
                mx3 = fmax3(mx3, __uint_as_float(sv[c + 6]), __uint_as_float(sv[c + 7]));
            }
            const float m_blk = fmaxf(fmax3(mx0, mx1, mx2), mx3) * scale_log2e;
            // P V_{t-2} complete: this half's P tile is free -- and P V_{t-3} is complete too (the P V issuer works in order),
            // which the rescaling below relies on.  Waited for here, off the hand-over path between the halves.
            if (t >= 2) WAIT(smem_u32(&p_free[half]), (u - 1) & 1, 8, t);
            if (t > 0) {
                // the partner has finished the exponentials of block t-1 and published the reference it used
                pair_sync(bar_theirs);
                const float m_prev = xch[(t - 1) & 1][r];
                if (m_prev != m_ref) {              // references only grow: m_prev > m_ref
                    l_run *= (m_ref == -INFINITY) ? 0.f : fast_exp2(m_ref - m_prev);
                    m_ref = m_prev;
                }
            }
            if (stamp && t < 30) dbg_cta[t * 8 + 1] = clock64();
            const bool grow = m_blk > m_ref + 8.0f;            // also true for the first finite block (m_ref = -inf)
            float alpha = 1.f;
            if (grow) {
                alpha = (m_ref == -INFINITY) ? 0.f : fast_exp2(m_ref - m_blk);
                m_ref = m_blk;
                l_run *= alpha;
            }
            xch[t & 1][r] = m_ref;
            if (t > 0 && __any_sync(0xffffffffu, grow)) {
                // O is about to be corrected: P V_{t-1} (and with it every earlier one) must have completed.  p_free of the
                // partner's buffer is in phase (t-1)/2 (pending) or one later, never further: P V_{t+1} needs P V_t first.
                WAIT(smem_u32(&p_free[half ^ 1]), ((t - 1) >> 1) & 1, 7, t);
                tc_fence_after();
                const float2 a2 = make_float2(alpha, alpha);
#pragma unroll 1
                for (int part = 0; part < 8; ++part) {       // 8 columns at a time: the rare path must not cost registers
                    uint32_t ov[8];
                    tmem_ld_32x8(lane_addr + O_COL + part * 8, ov);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {
                        const float2 tt = __fmul2_rn(make_float2(__uint_as_float(ov[e]), __uint_as_float(ov[e + 1])), a2);
                        ov[e] = __float_as_uint(tt.x);
                        ov[e + 1] = __float_as_uint(tt.y);
                    }
                    tmem_st_32x8(lane_addr + O_COL + part * 8, ov);
                }
                tmem_st_wait();
            }
            const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;     // fully masked so far: p = 0, no NaN
            // p = exp2(s * scale - m_ref), 32 columns at a time
            uint32_t pk[32];
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                uint32_t (&svp)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[part * 32]);
                uint32_t (&pkp)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[part * 16]);
                if (need_mask || POLY_OF_8 == 0) l_run += softmax_32<false>(svp, pkp, scale_log2e, m_use);
                else l_run += softmax_32<true>(svp, pkp, scale_log2e, m_use);
                // the partner may start the exponentials of block t+1 (RELEASE_PART 0: when half of mine are still to come --
                // the overlap fills the hand-over bubble on the MUFU pipe)
                if (part == RELEASE_PART && t + 1 < nb) pair_arrive(bar_mine);
            }
            if (stamp && t < 30) dbg_cta[t * 8 + 2] = clock64();
            // P_t -> this half's TMEM tile: lane = query row, 32 columns of bf16 pairs = the A operand of P V_t, read by the
            // tensor core straight from tensor memory.  (Through shared memory, P cost 16 KB of stores + 16 KB of operand reads
            // per block on a shared-memory pipe that the MMA operand traffic already keeps 60 % busy.)
            tmem_st_32x32(lane_addr + P_COL + half * (BKV / 2), pk);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&p_full[half]));
            if (stamp && t < 30) dbg_cta[t * 8 + 3] = clock64();
        }
        // combine the halves: references only grow, so the final one is the larger of the two; one exchange of (m_ref, l) per row.
        // Own barrier ids: a ping-pong arrival of this warp may still be waiting for its partner.
        fin_m[half][r] = m_ref;
        fin_l[half][r] = l_run;
        pair_sync(9 + q);
        const float m_oth = fin_m[half ^ 1][r], l_oth = fin_l[half ^ 1][r];
        const float m_fin = fmaxf(m_ref, m_oth);
        const float l_tot = ((m_ref == -INFINITY) ? 0.f : l_run * fast_exp2(m_ref - m_fin)) +
                            ((m_oth == -INFINITY) ? 0.f : l_oth * fast_exp2(m_oth - m_fin));
        m_ref = m_fin;
        WAIT(smem_u32(&bar_done), 0, 9, nb);
        tc_fence_after();
        if (stamp) { dbg_cta[31 * 8 + 0] = t_entry; dbg_cta[31 * 8 + 1] = clock64(); }     // row-thread entry, last PV done
        const float inv = (l_tot > 0.f) ? 1.f / l_tot : 0.f;
        // optional: log-sum-exp of the scaled scores (natural log) for the backward pass -- any reference m_ref gives the
        // same value, so the lazy rescaling does not matter; +inf marks a row that attends nothing
        if (lse != nullptr && i < S && half == 0)
            lse[(static_cast<int64_t>(b) * H + h) * S + i] = (l_tot > 0.f) ? (m_ref + log2f(l_tot)) * 0.69314718055994531f : INFINITY;
        __nv_bfloat16* orow = o + (static_cast<int64_t>(row0) + i) * d + h * DH + half * 32;
        uint32_t ov[32];
        tmem_ld_32x32(lane_addr + O_COL + half * 32, ov);
        tmem_ld_wait();
        if (i < S) {
#pragma unroll
            for (int e = 0; e < 32; e += 8) {
                uint4 w;
                w.x = pack_bf16x2(__uint_as_float(ov[e]) * inv, __uint_as_float(ov[e + 1]) * inv);
                w.y = pack_bf16x2(__uint_as_float(ov[e + 2]) * inv, __uint_as_float(ov[e + 3]) * inv);
                w.z = pack_bf16x2(__uint_as_float(ov[e + 4]) * inv, __uint_as_float(ov[e + 5]) * inv);
                w.w = pack_bf16x2(__uint_as_float(ov[e + 6]) * inv, __uint_as_float(ov[e + 7]) * inv);
                *reinterpret_cast<uint4*>(orow + e) = w;
            }
        }
        if (stamp) {      // output stored
            dbg_cta[31 * 8 + 2] = clock64();
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            dbg_cta[30 * 8 + 2] = static_cast<long long>(gt);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

}  // namespace

static long long* g_fwd_dbg = nullptr;
static int g_fwd_dbg_thread = 64;      // which row thread stamps (VALLE_B200_FWD_DBG_THREAD: 64, 96, 128 or 160 = lane 0 of a softmax warp)
extern "C" int vb_attention_prefill_set_debug(void* buf) {   /* device buffer of grid * 32 * 8 int64 cycle stamps, or NULL */
    g_fwd_dbg = static_cast<long long*>(buf);
    const char* t = getenv("VALLE_B200_FWD_DBG_THREAD");
    g_fwd_dbg_thread = t ? atoi(t) : 64;
    return VB_OK;
}

extern "C" int vb_attention_prefill_tc(const void* qkv, void* o, int B, int S, int H, int mask_mode, const int32_t* x_lens,
                                       const int32_t* kv_lens, float* lse, void* stream) {
    VB_REQUIRE(qkv && o, VB_ERR_BAD_ARG, "vb_attention_prefill_tc: null pointer");
    VB_REQUIRE(B >= 1 && S >= 1 && H >= 1 && B <= 65535 && H <= 65535, VB_ERR_BAD_ARG, "vb_attention_prefill_tc: bad shape");
    VB_REQUIRE(mask_mode == VB_MASK_NONE || mask_mode == VB_MASK_PREFIX_LM, VB_ERR_UNSUPPORTED,
               "vb_attention_prefill_tc: mask mode %d (use vb_attention for explicit masks)", mask_mode);
    VB_REQUIRE(mask_mode != VB_MASK_PREFIX_LM || x_lens != nullptr, VB_ERR_BAD_ARG, "vb_attention_prefill_tc: prefix-LM needs x_lens");
    const int64_t d = static_cast<int64_t>(H) * DH;
    const int64_t rows = static_cast<int64_t>(B) * S;
    CUtensorMap tq, tkv;
    int rc;
    if ((rc = vb_make_tmap_bf16_2d(&tq, qkv, rows, 3 * d, 3 * d, BQ, DH)) != VB_OK) return rc;
    if ((rc = vb_make_tmap_bf16_2d(&tkv, qkv, rows, 3 * d, 3 * d, BKV, DH)) != VB_OK) return rc;
    static bool configured = false;
    if (!configured) {
        VB_CUDA(cudaFuncSetAttribute(attn_prefill_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    dim3 grid(static_cast<unsigned>(vb_ceil_div(S, BQ)), H, B);
    const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(DH));
    VB_CUDA(vb_launch(false, attn_prefill_tc_kernel, grid, dim3(THREADS), SMEM_BYTES, static_cast<cudaStream_t>(stream), tq, tkv,
                      static_cast<__nv_bfloat16*>(o), S, H, mask_mode, x_lens, kv_lens, scale_log2e, lse, g_fwd_dbg, g_fwd_dbg_thread));
    return VB_OK;
}
