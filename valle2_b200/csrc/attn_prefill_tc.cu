// Flash-style attention on the 5th-generation tensor cores for the large-M paths (AR prefill, NAR stages, teacher-forced
// forward): S = Q K^T and O += P V are tcgen05.mma with accumulators in TMEM, operands staged by TMA (SWIZZLE_128B)
// straight out of the packed qkv buffer [B*S][3][H][64] written by the QKV GEMM -- no head-split copy, no
// materialised (B,H,S,S) mask: the prefix-LM / padding predicate is evaluated from (x_len, kv_len) per element.
//
// CTA = 128 query rows of one (batch, head); key blocks of 64; 320 threads, two CTAs per SM.
//   warp 0    TMA producer: Q tile once, then a 4-deep ring of K blocks and a 3-deep ring of V blocks (a block's operands must be
//             requested > 1000 cycles ahead: with a 3-deep {K, V} ring the MMA warp waited ~500 cycles per block for them)
//   warp 1    MMA issuer + TMEM allocator: QK_j (M128 N64 K64) into score buffer j%2, PV_j (M128 N64 K64; V is the MN-major
//             B operand) accumulating into O in TMEM.  Issue order QK_0, QK_1, PV_0, QK_2, PV_1, ...
//   warp 2-9  softmax, TWO threads per query row (warps w and w+4 share a TMEM lane quadrant; each takes 32 of the block's
//             64 columns).  The kernel is bound by per-block latencies of the row threads (mbarrier wait ~90 cycles, TMEM
//             load, proxy fence, arrive), not by any pipe -- twice the warps hide twice the latency, and a thread carries 32
//             instead of 64 scores.  The halves exchange their row maxima through shared memory (one 64-thread named barrier).
// Synchronisation per key block: the row threads wait for S_j and arrive once with P_j; everything else is implied by the
// MMA issue order (tcgen05 operations of one thread complete in order):
//   * S_j complete  =>  PV_{j-2} complete  =>  P buffer j%2 may be overwritten;
//   * P_{j-1} seen by the MMA warp  =>  every row thread has read S_{j-1}  =>  QK_{j+1} may overwrite score buffer (j+1)%2.
// O stays in TMEM: the softmax reference is raised lazily (only when a block maximum exceeds it by 2^8) and only then O is
// rescaled.  exp2: in unmasked blocks POLY_OF_8 of every 8 score pairs take exp2 on the FMA pipe instead of MUFU (round-to-
// nearest range reduction by the 1.5*2^23 trick, degree-3 minimax on [-0.5, 0.5]: 7.5e-5 relative error, far below the bf16
// rounding of P).
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
constexpr int K_STAGES = 4;   // K blocks: a stage is free as soon as Q K_j^T has completed
constexpr int V_STAGES = 3;   // V blocks: free after P_j V_j
constexpr int POLY_OF_8 = VB_FWD_POLY;
constexpr int THREADS = 320;
constexpr int Q_BYTES = BQ * DH * 2;          // 16 KB
constexpr int KV_BYTES = BKV * DH * 2;        // 8 KB per K or V block
constexpr int P_BYTES = BQ * BKV * 2;         // 16 KB
constexpr int SMEM_BYTES = Q_BYTES + (K_STAGES + V_STAGES) * KV_BYTES + 2 * P_BYTES + 1024;
constexpr int O_COL = 2 * BKV;                // score buffers: cols [0, 64), [64, 128); O: [128, 192)
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

// p = exp2(s * scale - m) for a thread's 32 scores, their sum, bf16 pack (pk[i] = columns 2i, 2i+1)
template <bool POLY>
__device__ __forceinline__ float softmax_half_row(const uint32_t (&sv)[32], uint32_t (&pk)[16], float scale_log2e, float m_use) {
    const float2 sc2 = make_float2(scale_log2e, scale_log2e), nm2 = make_float2(-m_use, -m_use);
    float2 rs_a = make_float2(0.f, 0.f), rs_b = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
        const float2 t = __ffma2_rn(make_float2(__uint_as_float(sv[c]), __uint_as_float(sv[c + 1])), sc2, nm2);
        float2 pp;
#ifdef VB_FWD_NOEXP      // timing experiment only: no exponential at all (wrong results)
        pp = t;
#else
        if (POLY && ((c >> 1) & 7) >= 8 - POLY_OF_8) pp = poly_exp2(t);
        else pp = make_float2(fast_exp2(t.x), fast_exp2(t.y));
#endif
        if (c & 2) rs_b = __fadd2_rn(rs_b, pp);
        else rs_a = __fadd2_rn(rs_a, pp);
        pk[c >> 1] = pack_bf16x2(pp.x, pp.y);
    }
    const float2 rs = __fadd2_rn(rs_a, rs_b);
    return rs.x + rs.y;
}

// shared-memory matrix descriptor (SWIZZLE_128B, SBO 1024) from its low word: start address >> 4 | LBO >> 4 << 16
__device__ __forceinline__ uint64_t desc_from_lo(uint32_t lo) { return (static_cast<uint64_t>(0x40004040u) << 32) | lo; }
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(THREADS, 2) attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap tm_q,
                                                                     const __grid_constant__ CUtensorMap tm_kv,
                                                                     __nv_bfloat16* __restrict__ o, int S, int H, int mask_mode,
                                                                     const int32_t* __restrict__ x_lens,
                                                                     const int32_t* __restrict__ kv_lens, float scale_log2e,
                                                                     float* __restrict__ lse, long long* __restrict__ dbg, int dbg_thread) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_q, bar_o, s_full[2], p_full[2], k_full[K_STAGES], k_empty[K_STAGES], v_full[V_STAGES], v_empty[V_STAGES];
    __shared__ float xch[2][2][BQ];      // [block parity][column half][row]: the halves' row maxima (and, at the end, sums)
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
    const uint32_t p_smem = v_smem + V_STAGES * KV_BYTES;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        mbar_init(smem_u32(&bar_q), 1);
        mbar_init(smem_u32(&bar_o), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&s_full[s]), 1);
            mbar_init(smem_u32(&p_full[s]), 8);
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
    const int kv_len = kv_lens ? min(kv_lens[b], S) : S;
    const int x_len = (mask_mode == VB_MASK_PREFIX_LM) ? x_lens[b] : 0;
    // keys any row of this tile may attend: everything (no mask) or the text prefix plus the causal part
    int k_end = kv_len;
    if (mask_mode == VB_MASK_PREFIX_LM) k_end = min(kv_len, max(x_len, i0 + BQ));
    const int nb = max(1, (k_end + BKV - 1) / BKV);

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(smem_u32(&bar_q), Q_BYTES);
            tma_load_2d(q_smem, &tm_q, smem_u32(&bar_q), h * DH, row0 + i0);
            // K runs one block ahead of V: K_t and V_{t-1} per iteration (a K stage is released earlier than a V stage)
            int ks = 0, vs = 0;
            uint32_t kph = 0, vph = 0;
            for (int t = 0; t <= nb; ++t) {
                if (t < nb) {
                    mbar_wait_relaxed(smem_u32(&k_empty[ks]), kph ^ 1);
                    mbar_expect_tx(smem_u32(&k_full[ks]), KV_BYTES);
                    tma_load_2d(k_smem + ks * KV_BYTES, &tm_kv, smem_u32(&k_full[ks]), d + h * DH, row0 + t * BKV);
                    if (++ks == K_STAGES) { ks = 0; kph ^= 1; }
                }
                if (t > 0) {
                    mbar_wait_relaxed(smem_u32(&v_empty[vs]), vph ^ 1);
                    mbar_expect_tx(smem_u32(&v_full[vs]), KV_BYTES);
                    tma_load_2d(v_smem + vs * KV_BYTES, &tm_kv, smem_u32(&v_full[vs]), 2 * d + h * DH, row0 + (t - 1) * BKV);
                    if (++vs == V_STAGES) { vs = 0; vph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t IDESC_QK = umma_idesc_bf16(BQ, BKV, 0, 0);   // A = Q (K-major), B = K block (K-major)
            constexpr uint32_t IDESC_PV = umma_idesc_bf16(BQ, DH, 0, 1);    // A = P (K-major), B = V block (MN-major)
            const uint32_t o_tmem = tmem_base + O_COL;
            // descriptor low words; a step of 16 along K is +32 bytes (K-major) or +16 rows of 128 bytes (MN-major V)
            const uint32_t q_lo = ((q_smem & 0x3ffffu) >> 4) | (1u << 16);
            const uint32_t k_lo = ((k_smem & 0x3ffffu) >> 4) | (1u << 16);
            const uint32_t v_lo = ((v_smem & 0x3ffffu) >> 4) | (64u << 16);
            const uint32_t p_lo = ((p_smem & 0x3ffffu) >> 4) | (1u << 16);
            // MMA-warp stamps (debug buffer slots 4..7 of a block): {K/V stage ready, Q K^T issued, P ready, P V issued}
            mbar_wait(smem_u32(&bar_q), 0);
            int pv_stage = 0;
            uint32_t pv_phase = 0;
            auto issue_pv = [&](int jj) {
                const int pb = jj & 1;
                mbar_wait(smem_u32(&v_full[pv_stage]), pv_phase);
                mbar_wait(smem_u32(&p_full[pb]), (jj >> 1) & 1);
                tc_fence_after();
                if (dbg_cta && jj < 30) dbg_cta[jj * 8 + 6] = clock64();
                const uint32_t vl = v_lo + pv_stage * (KV_BYTES >> 4), pl = p_lo + pb * (P_BYTES >> 4);
#pragma unroll
                for (int kk = 0; kk < BKV / 16; ++kk)
                    umma_f16(o_tmem, desc_from_lo(pl + kk * 2), desc_from_lo(vl + kk * 128), IDESC_PV, (jj > 0 || kk > 0) ? 1u : 0u);
                umma_commit(smem_u32(&v_empty[pv_stage]));
                umma_commit(smem_u32(&bar_o));
                if (dbg_cta && jj < 30) dbg_cta[jj * 8 + 7] = clock64();
                if (++pv_stage == V_STAGES) { pv_stage = 0; pv_phase ^= 1; }
            };
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < nb; ++j) {
                mbar_wait(smem_u32(&k_full[stage]), phase);
                tc_fence_after();
                if (dbg_cta && j < 30) dbg_cta[j * 8 + 4] = clock64();
                const uint32_t kl = k_lo + stage * (KV_BYTES >> 4);
                const uint32_t s_tmem = tmem_base + (j & 1) * BKV;
#pragma unroll
                for (int kk = 0; kk < DH / 16; ++kk)
                    umma_f16(s_tmem, desc_from_lo(q_lo + kk * 2), desc_from_lo(kl + kk * 2), IDESC_QK, kk > 0 ? 1u : 0u);
                umma_commit(smem_u32(&s_full[j & 1]));
                umma_commit(smem_u32(&k_empty[stage]));
                if (dbg_cta && j < 30) dbg_cta[j * 8 + 5] = clock64();
                if (++stage == K_STAGES) { stage = 0; phase ^= 1; }
                if (j > 0) issue_pv(j - 1);
            }
            issue_pv(nb - 1);
        }
    } else {
        const int sw = warp - 2;              // 0..7
        const int q = warp & 3;               // TMEM lane quadrant this warp may access
        const int half = sw >> 2;             // which 32 of a block's 64 columns
        const int r = q * 32 + lane;          // row within the tile == TMEM lane
        const int i = i0 + r;                 // query index within the sequence
        const int pair_id = 1 + q;            // named barrier of the two warps that share the quadrant
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        // O accumulates in TMEM across key blocks (PV MMAs with accumulate = 1).  The softmax reference m_ref is only
        // raised when the block maximum exceeds it by more than 2^8 ("lazy rescaling"): then O and l are multiplied by
        // exp2(m_ref - m_new) -- a TMEM load/scale/store done by both warps of a pair (32 columns of O each) if any of
        // their rows needs it.  Otherwise p = exp2(s - m_ref) <= 256, which bf16 / fp32 hold comfortably.
        float m_ref = -INFINITY, l_run = 0.f;      // l_run: this thread's 32 columns only
        // optional per-block cycle stamps of ONE row thread of every CTA (vb_attention_prefill_set_debug): [cta][block][0..3] =
        // {S ready, row maximum known, exp2 / pack done, P written}
        const long long t_entry = clock64();
        const bool stamp = dbg != nullptr && threadIdx.x == dbg_thread;
        for (int j = 0; j < nb; ++j) {
            const int sb = j & 1;
            mbar_wait(smem_u32(&s_full[sb]), (j >> 1) & 1);
            tc_fence_after();
            if (stamp && j < 30) dbg_cta[j * 8 + 0] = clock64();
            uint32_t sv[32];
            tmem_ld_32x32(lane_addr + sb * BKV + half * 32, sv);
            tmem_ld_wait();
            const int kbase = j * BKV;
            bool need_mask = (kbase + BKV > kv_len);
            if (mask_mode == VB_MASK_PREFIX_LM)
                need_mask = need_mask || !((kbase + BKV <= x_len) || (i0 >= x_len && kbase + BKV - 1 <= i0));
            if (need_mask) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int kj = kbase + half * 32 + c;
                    bool ok = kj < kv_len;
                    if (mask_mode == VB_MASK_PREFIX_LM) ok = ok && ((kj < x_len) || (i >= x_len && kj <= i));
                    if (!ok) sv[c] = 0xff800000u;   // -inf -> p = 0
                }
            }
            // maximum of this half's raw scores (scale > 0 commutes with max), then the other half's through shared memory
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                mx0 = fmax3(mx0, __uint_as_float(sv[c]), __uint_as_float(sv[c + 1]));
                mx1 = fmax3(mx1, __uint_as_float(sv[c + 2]), __uint_as_float(sv[c + 3]));
            }
            float mx = fmaxf(mx0, mx1);
            xch[sb][half][r] = mx;
            pair_barrier(pair_id);
            mx = fmaxf(mx, xch[sb][half ^ 1][r]);
            if (stamp && j < 30) dbg_cta[j * 8 + 1] = clock64();
            const float m_blk = mx * scale_log2e;
            const bool grow = m_blk > m_ref + 8.0f;            // also true for the first finite block (m_ref = -inf)
            float alpha = 1.f;
            if (grow) {
                alpha = (m_ref == -INFINITY) ? 0.f : fast_exp2(m_ref - m_blk);
                m_ref = m_blk;
                l_run *= alpha;
            }
            if (j > 0 && __any_sync(0xffffffffu, grow)) {
                // O is about to be corrected: every P V issued so far must have completed.  bar_o completes one phase per
                // P V; it is in phase j-1 (P V_{j-1} pending) or j, so the parity of phase j-1 is unambiguous.
                mbar_wait(smem_u32(&bar_o), (j - 1) & 1);
                tc_fence_after();
                const float2 a2 = make_float2(alpha, alpha);
                uint32_t ov[32];
                tmem_ld_32x32(lane_addr + O_COL + half * 32, ov);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    const float2 t = __fmul2_rn(make_float2(__uint_as_float(ov[e]), __uint_as_float(ov[e + 1])), a2);
                    ov[e] = __float_as_uint(t.x);
                    ov[e + 1] = __float_as_uint(t.y);
                }
                tmem_st_32x32(lane_addr + O_COL + half * 32, ov);
                tmem_st_wait();
            }
            const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;     // fully masked so far: p = 0, no NaN
            uint32_t pk[16];
            if (need_mask || POLY_OF_8 == 0) l_run += softmax_half_row<false>(sv, pk, scale_log2e, m_use);
            else l_run += softmax_half_row<true>(sv, pk, scale_log2e, m_use);
            if (stamp && j < 30) dbg_cta[j * 8 + 2] = clock64();
            // P_j -> smem buffer j%2 as K-major 128B-swizzled rows (16-byte chunk c of row r lives at chunk c ^ (r & 7));
            // the buffer is free: S_j complete implies P V_{j-2} complete (issue order of the MMA warp)
            const uint32_t p_row = p_smem + sb * P_BYTES + r * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t addr = p_row + (static_cast<uint32_t>((half * 4 + c) ^ (r & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[c * 4]), "r"(pk[c * 4 + 1]),
                             "r"(pk[c * 4 + 2]), "r"(pk[c * 4 + 3]) : "memory");
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&p_full[sb]));
            if (stamp && j < 30) dbg_cta[j * 8 + 3] = clock64();
        }
        // the other half's partial row sum (slot parity nb&1 was last used by block nb-2 or not at all)
        xch[nb & 1][half][r] = l_run;
        pair_barrier(pair_id);
        const float l_tot = l_run + xch[nb & 1][half ^ 1][r];
        mbar_wait(smem_u32(&bar_o), (nb - 1) & 1);
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
