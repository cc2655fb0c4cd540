// Attention backward on the 5th-generation tensor cores (teacher-forced training step, bf16, head_dim 64), the dQ half:
// per CTA 128 query rows of one (batch, head), key blocks of 64 -- the forward kernel's pipeline (attn_prefill_tc.cu) with
// one more product and no online maximum, because the forward pass saved the rows' log-sum-exp:
//
//   S_j  = Q K_j^T          tcgen05.mma  M128 N64 K64   (Q, K_j K-major, TMA tiles out of the packed qkv buffer)
//   dP_j = dO V_j^T         tcgen05.mma  M128 N64 K64   (dO tile by TMA out of the [B*S][d] gradient of the attention output)
//   P_j  = exp2(S_j scale log2e - lse log2e)            one thread per query row (tcgen05.ld 32x32b: no shuffles)
//   dS_j = P_j (dP_j - delta) scale  -> bf16, 128B-swizzled smem tile = K-major A operand of the next product
//   dQ  += dS_j K_j         tcgen05.mma  M128 N64 K64   (K_j as the MN-major B operand: the same smem tile, other descriptor)
//
// delta_i = sum_e dO[i][e] O[i][e] is computed by the row threads while the first tiles are in flight and written out for
// the dK / dV kernel.  TMEM: S [0,64), dP [64,128), dQ [128,192) of 256 allocated columns; 81 KB of shared memory -> two CTAs
// per SM.  Replaces the mma.sync kernel attn_bwd_dq_mma_kernel (train.cu) when the forward's lse is available; torch
// autograd of F.scaled_dot_product_attention (modules.py:167) is what it stands in for.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BQ = 128;
constexpr int BKV = 64;
constexpr int DH = 64;
constexpr int KV_STAGES = 2;
constexpr int THREADS = 192;
constexpr int Q_BYTES = BQ * DH * 2;          // 16 KB (Q tile, dO tile)
constexpr int KV_BYTES = BKV * DH * 2;        // 8 KB each for K and V
constexpr int DS_BYTES = BQ * BKV * 2;        // 16 KB
constexpr int DQ_STAGES = 3;                  // {K, V} ring of the dQ kernel: a stage is only free after dQ += dS K of its block, so two
                                              // stages stall the next S / dP behind it; three still leave two CTAs per SM (97 KB each)
constexpr int SMEM_BYTES = 2 * Q_BYTES + DQ_STAGES * 2 * KV_BYTES + DS_BYTES + 1024;
constexpr int TMEM_COLS = 256;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(THREADS, 2) attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_q,
                                                                    const __grid_constant__ CUtensorMap tm_kv,
                                                                    const __grid_constant__ CUtensorMap tm_do,
                                                                    const __nv_bfloat16* __restrict__ o,
                                                                    const __nv_bfloat16* __restrict__ dO,
                                                                    __nv_bfloat16* __restrict__ dqkv, const float* __restrict__ lse,
                                                                    float* __restrict__ delta, int S, int H, int mask_mode,
                                                                    const int32_t* __restrict__ x_lens,
                                                                    const int32_t* __restrict__ kv_lens, float scale) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_q, bar_s, bar_sfree, bar_p, bar_o;
    __shared__ __align__(8) uint64_t kv_full[DQ_STAGES], kv_empty[DQ_STAGES];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int d = H * DH;
    const int i0 = qt * BQ;

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_smem = base;
    const uint32_t do_smem = base + Q_BYTES;
    const uint32_t kv_smem = do_smem + Q_BYTES;
    const uint32_t ds_smem = kv_smem + DQ_STAGES * 2 * KV_BYTES;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        tma_prefetch_desc(&tm_do);
        mbar_init(smem_u32(&bar_q), 1);
        mbar_init(smem_u32(&bar_s), 1);
        mbar_init(smem_u32(&bar_sfree), 4);
        mbar_init(smem_u32(&bar_p), 4);
        mbar_init(smem_u32(&bar_o), 1);
        for (int s = 0; s < DQ_STAGES; ++s) {
            mbar_init(smem_u32(&kv_full[s]), 1);
            mbar_init(smem_u32(&kv_empty[s]), 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(&tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int row0 = b * S;
    const int kv_len = kv_lens ? min(kv_lens[b], S) : S;
    const int x_len = (mask_mode == VB_MASK_PREFIX_LM) ? x_lens[b] : 0;
    int k_end = kv_len;
    if (mask_mode == VB_MASK_PREFIX_LM) k_end = min(kv_len, max(x_len, i0 + BQ));
    const int nb = max(1, (k_end + BKV - 1) / BKV);

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(smem_u32(&bar_q), 2 * Q_BYTES);
            tma_load_2d(q_smem, &tm_q, smem_u32(&bar_q), h * DH, row0 + i0);
            tma_load_2d(do_smem, &tm_do, smem_u32(&bar_q), h * DH, row0 + i0);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < nb; ++j) {
                mbar_wait_relaxed(smem_u32(&kv_empty[stage]), phase ^ 1);
                const uint32_t fb = smem_u32(&kv_full[stage]);
                mbar_expect_tx(fb, 2 * KV_BYTES);
                const uint32_t dst = kv_smem + stage * 2 * KV_BYTES;
                tma_load_2d(dst, &tm_kv, fb, d + h * DH, row0 + j * BKV);
                tma_load_2d(dst + KV_BYTES, &tm_kv, fb, 2 * d + h * DH, row0 + j * BKV);
                if (++stage == DQ_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t IDESC_KK = umma_idesc_bf16(BQ, BKV, 0, 0);   // both operands K-major: Q K^T and dO V^T
            constexpr uint32_t IDESC_MN = umma_idesc_bf16(BQ, DH, 0, 1);    // A = dS (K-major), B = K block (MN-major)
            const uint32_t s_tmem = tmem_base, dp_tmem = tmem_base + BKV, dq_tmem = tmem_base + 2 * BKV;
            mbar_wait_relaxed(smem_u32(&bar_q), 0);
            int stage = 0;
            uint32_t phase = 0;
            auto issue_dq = [&](int jj, int st) {
                mbar_wait_relaxed(smem_u32(&bar_p), jj & 1);
                tc_fence_after();
                const uint32_t k_s = kv_smem + st * 2 * KV_BYTES;
#pragma unroll
                for (int kk = 0; kk < BKV / 16; ++kk) {
                    const uint64_t da = umma_desc_sw128(ds_smem + kk * 32, 16, 1024);
                    const uint64_t db = umma_desc_sw128(k_s + kk * 16 * 128, 1024, 1024);
                    umma_f16(dq_tmem, da, db, IDESC_MN, (jj > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(smem_u32(&kv_empty[st]));
                umma_commit(smem_u32(&bar_o));
            };
            int prev_stage = 0;
            for (int j = 0; j < nb; ++j) {
                mbar_wait_relaxed(smem_u32(&kv_full[stage]), phase);
                mbar_wait_relaxed(smem_u32(&bar_sfree), (j & 1) ^ 1);
                tc_fence_after();
                const uint32_t k_s = kv_smem + stage * 2 * KV_BYTES, v_s = k_s + KV_BYTES;
#pragma unroll
                for (int kk = 0; kk < DH / 16; ++kk) {
                    const uint64_t da = umma_desc_sw128(q_smem + kk * 32, 16, 1024);
                    const uint64_t db = umma_desc_sw128(k_s + kk * 32, 16, 1024);
                    umma_f16(s_tmem, da, db, IDESC_KK, kk > 0 ? 1u : 0u);
                }
#pragma unroll
                for (int kk = 0; kk < DH / 16; ++kk) {
                    const uint64_t da = umma_desc_sw128(do_smem + kk * 32, 16, 1024);
                    const uint64_t db = umma_desc_sw128(v_s + kk * 32, 16, 1024);
                    umma_f16(dp_tmem, da, db, IDESC_KK, kk > 0 ? 1u : 0u);
                }
                umma_commit(smem_u32(&bar_s));
                if (j > 0) issue_dq(j - 1, prev_stage);
                prev_stage = stage;
                if (++stage == DQ_STAGES) { stage = 0; phase ^= 1; }
            }
            issue_dq(nb - 1, prev_stage);
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;          // row within the tile == TMEM lane
        const int i = i0 + r;                 // query index within the sequence
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int64_t bh = static_cast<int64_t>(b) * H + h;
        // delta_i and lse_i while the first tiles are in flight
        float dl = 0.f, lse2 = INFINITY;
        if (i < S) {
            const uint4* po = reinterpret_cast<const uint4*>(o + (static_cast<int64_t>(row0) + i) * d + h * DH);
            const uint4* pd = reinterpret_cast<const uint4*>(dO + (static_cast<int64_t>(row0) + i) * d + h * DH);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 a = po[c], g = pd[c];
                dl += bf16_lo(a.x) * bf16_lo(g.x) + bf16_hi(a.x) * bf16_hi(g.x);
                dl += bf16_lo(a.y) * bf16_lo(g.y) + bf16_hi(a.y) * bf16_hi(g.y);
                dl += bf16_lo(a.z) * bf16_lo(g.z) + bf16_hi(a.z) * bf16_hi(g.z);
                dl += bf16_lo(a.w) * bf16_lo(g.w) + bf16_hi(a.w) * bf16_hi(g.w);
            }
            delta[bh * S + i] = dl;
            const float l = lse[bh * S + i];
            lse2 = (l == INFINITY) ? INFINITY : l * 1.4426950408889634f;
        }
        const float scale_log2e = scale * 1.4426950408889634f;
        // lse2 = +inf (row attends nothing, or i >= S): -inf - inf stays -inf for a masked score, p = 0, dS = 0 * finite = 0
        const float2 sc2 = make_float2(scale_log2e, scale_log2e), nl2 = make_float2(-lse2, -lse2);
        const float2 ss2 = make_float2(scale, scale), nd2 = make_float2(-dl * scale, -dl * scale);
        const uint32_t ds_row = ds_smem + r * 128;
        for (int j = 0; j < nb; ++j) {
            mbar_wait(smem_u32(&bar_s), j & 1);
            tc_fence_after();
            if (j > 0) {      // dQ MMA of block j-1 must have consumed the dS tile before it is overwritten
                mbar_wait(smem_u32(&bar_o), (j - 1) & 1);
                tc_fence_after();
            }
            const int kbase = j * BKV;
            bool need_mask = (kbase + BKV > kv_len);
            if (mask_mode == VB_MASK_PREFIX_LM)
                need_mask = need_mask || !((kbase + BKV <= x_len) || (i0 >= x_len && kbase + BKV - 1 <= i0));
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t sv[32], dv[32];
                tmem_ld_32x32(lane_addr + half * 32, sv);
                tmem_ld_32x32(lane_addr + BKV + half * 32, dv);
                tmem_ld_wait();
                if (half == 1) {      // S_j and dP_j fully consumed: the MMA warp may overwrite them with block j+1
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bar_sfree));
                }
                if (need_mask) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int kj = kbase + half * 32 + c;
                        bool ok = kj < kv_len;
                        if (mask_mode == VB_MASK_PREFIX_LM) ok = ok && ((kj < x_len) || (i >= x_len && kj <= i));
                        if (!ok) sv[c] = 0xff800000u;   // -inf -> p = 0
                    }
                }
                uint32_t pk[16];
#pragma unroll
                for (int c = 0; c < 32; c += 2) {      // packed fp32x2: t = s scale log2e - lse2, g = dp scale - delta scale, dS = p g
                    const float2 t = __ffma2_rn(make_float2(__uint_as_float(sv[c]), __uint_as_float(sv[c + 1])), sc2, nl2);
                    const float2 g2 = __ffma2_rn(make_float2(__uint_as_float(dv[c]), __uint_as_float(dv[c + 1])), ss2, nd2);
                    const float2 ds2 = __fmul2_rn(make_float2(ex2(t.x), ex2(t.y)), g2);
                    pk[c >> 1] = pack_bf16x2(ds2.x, ds2.y);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t addr = ds_row + (static_cast<uint32_t>((half * 4 + c) ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[c * 4]), "r"(pk[c * 4 + 1]),
                                 "r"(pk[c * 4 + 2]), "r"(pk[c * 4 + 3]) : "memory");
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_p));
        }
        mbar_wait(smem_u32(&bar_o), (nb - 1) & 1);
        tc_fence_after();
        __nv_bfloat16* qrow = dqkv + (static_cast<int64_t>(row0) + i) * 3 * d + h * DH;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t ov[32];
            tmem_ld_32x32(lane_addr + 2 * BKV + half * 32, ov);
            tmem_ld_wait();
            if (i < S) {
#pragma unroll
                for (int e = 0; e < 32; e += 8) {
                    uint4 w;
                    w.x = pack_bf16x2(__uint_as_float(ov[e]), __uint_as_float(ov[e + 1]));
                    w.y = pack_bf16x2(__uint_as_float(ov[e + 2]), __uint_as_float(ov[e + 3]));
                    w.z = pack_bf16x2(__uint_as_float(ov[e + 4]), __uint_as_float(ov[e + 5]));
                    w.w = pack_bf16x2(__uint_as_float(ov[e + 6]), __uint_as_float(ov[e + 7]));
                    *reinterpret_cast<uint4*>(qrow + half * 32 + e) = w;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// dK / dV: per CTA 128 keys of one (batch, head), query blocks of 64.  Everything is computed transposed so that the key
// is the accumulator row (TMEM lane) and no reduction across CTAs is needed:
//   S^T_i  = K Q_i^T        M128 N64 K64   (K tile, Q_i block K-major)
//   dP^T_i = V dO_i^T       M128 N64 K64
//   P^T = exp2(S^T scale log2e - lse_q log2e),  dS^T = P^T (dP^T - delta_q) scale      thread = key row, columns = queries
//   dV += P^T dO_i          M128 N64 K64   (P^T tile K-major A, dO_i block as MN-major B)
//   dK += dS^T Q_i          M128 N64 K64   (dS^T tile K-major A, Q_i block as MN-major B)
// lse / delta of the block's 64 queries are staged in shared memory (double-buffered, one named barrier per block).
// TMEM: S^T [0,64) dP^T [64,128) dV [128,192) dK [192,256); 97 KB of shared memory -> two CTAs per SM.
// ------------------------------------------------------------------------------------------------------------------
constexpr int BK2 = 128;                      // keys per CTA
constexpr int BI = 64;                        // queries per block
constexpr int KT_BYTES = BK2 * DH * 2;        // 16 KB (K tile, V tile)
constexpr int QB_BYTES = BI * DH * 2;         // 8 KB (Q block, dO block)
constexpr int PT_BYTES = BK2 * BI * 2;        // 16 KB (P^T, dS^T)
constexpr int SMEM2_BYTES = 2 * KT_BYTES + KV_STAGES * 2 * QB_BYTES + 2 * PT_BYTES + 1024;

__global__ void __launch_bounds__(THREADS, 2) attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tm_kt,   // qkv, box 128 x 64
                                                                     const __grid_constant__ CUtensorMap tm_qb,   // qkv, box 64 x 64
                                                                     const __grid_constant__ CUtensorMap tm_dob,  // dO, box 64 x 64
                                                                     __nv_bfloat16* __restrict__ dqkv, const float* __restrict__ lse,
                                                                     const float* __restrict__ delta, int S, int H, int mask_mode,
                                                                     const int32_t* __restrict__ x_lens,
                                                                     const int32_t* __restrict__ kv_lens, float scale) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_kv, bar_s, bar_sfree, bar_p, bar_o;
    __shared__ __align__(8) uint64_t qd_full[KV_STAGES], qd_empty[KV_STAGES];
    __shared__ __align__(16) float lse_s[2][BI], dl_s[2][BI];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int d = H * DH;
    const int j0 = jt * BK2;

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t k_smem = base;
    const uint32_t v_smem = base + KT_BYTES;
    const uint32_t qd_smem = v_smem + KT_BYTES;
    const uint32_t pt_smem = qd_smem + KV_STAGES * 2 * QB_BYTES;
    const uint32_t dst_smem = pt_smem + PT_BYTES;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_kt);
        tma_prefetch_desc(&tm_qb);
        tma_prefetch_desc(&tm_dob);
        mbar_init(smem_u32(&bar_kv), 1);
        mbar_init(smem_u32(&bar_s), 1);
        mbar_init(smem_u32(&bar_sfree), 4);
        mbar_init(smem_u32(&bar_p), 4);
        mbar_init(smem_u32(&bar_o), 1);
        for (int s = 0; s < KV_STAGES; ++s) {
            mbar_init(smem_u32(&qd_full[s]), 1);
            mbar_init(smem_u32(&qd_empty[s]), 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(&tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int row0 = b * S;
    const int kv_len = kv_lens ? min(kv_lens[b], S) : S;
    const int x_len = (mask_mode == VB_MASK_PREFIX_LM) ? x_lens[b] : 0;
    // queries that may see a key of this tile: everything (text-prefix keys / no mask) or the causal part from the tile on
    int i_begin = 0;
    if (mask_mode == VB_MASK_PREFIX_LM && j0 >= x_len) i_begin = j0;
    const bool live = j0 < kv_len;                                  // a tile past the sequence gets zero gradients
    const int nb = live ? (S - i_begin + BI - 1) / BI : 0;

    if (warp == 0) {
        if (elect_one() && nb > 0) {
            mbar_expect_tx(smem_u32(&bar_kv), 2 * KT_BYTES);
            tma_load_2d(k_smem, &tm_kt, smem_u32(&bar_kv), d + h * DH, row0 + j0);
            tma_load_2d(v_smem, &tm_kt, smem_u32(&bar_kv), 2 * d + h * DH, row0 + j0);
            int stage = 0;
            uint32_t phase = 0;
            for (int ib = 0; ib < nb; ++ib) {
                mbar_wait_relaxed(smem_u32(&qd_empty[stage]), phase ^ 1);
                const uint32_t fb = smem_u32(&qd_full[stage]);
                mbar_expect_tx(fb, 2 * QB_BYTES);
                const uint32_t dst = qd_smem + stage * 2 * QB_BYTES;
                tma_load_2d(dst, &tm_qb, fb, h * DH, row0 + i_begin + ib * BI);
                tma_load_2d(dst + QB_BYTES, &tm_dob, fb, h * DH, row0 + i_begin + ib * BI);
                if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one() && nb > 0) {
            constexpr uint32_t IDESC_KK = umma_idesc_bf16(BK2, BI, 0, 0);
            constexpr uint32_t IDESC_MN = umma_idesc_bf16(BK2, DH, 0, 1);
            const uint32_t st_tmem = tmem_base, dpt_tmem = tmem_base + 64, dv_tmem = tmem_base + 128, dk_tmem = tmem_base + 192;
            mbar_wait_relaxed(smem_u32(&bar_kv), 0);
            int stage = 0;
            uint32_t phase = 0;
            auto issue_acc = [&](int ii, int st) {
                mbar_wait_relaxed(smem_u32(&bar_p), ii & 1);
                tc_fence_after();
                const uint32_t q_s = qd_smem + st * 2 * QB_BYTES, do_s = q_s + QB_BYTES;
#pragma unroll
                for (int kk = 0; kk < BI / 16; ++kk) {
                    const uint64_t da = umma_desc_sw128(pt_smem + kk * 32, 16, 1024);
                    const uint64_t db = umma_desc_sw128(do_s + kk * 16 * 128, 1024, 1024);
                    umma_f16(dv_tmem, da, db, IDESC_MN, (ii > 0 || kk > 0) ? 1u : 0u);
                }
#pragma unroll
                for (int kk = 0; kk < BI / 16; ++kk) {
                    const uint64_t da = umma_desc_sw128(dst_smem + kk * 32, 16, 1024);
                    const uint64_t db = umma_desc_sw128(q_s + kk * 16 * 128, 1024, 1024);
                    umma_f16(dk_tmem, da, db, IDESC_MN, (ii > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(smem_u32(&qd_empty[st]));
                umma_commit(smem_u32(&bar_o));
            };
            int prev_stage = 0;
            for (int ib = 0; ib < nb; ++ib) {
                mbar_wait_relaxed(smem_u32(&qd_full[stage]), phase);
                mbar_wait_relaxed(smem_u32(&bar_sfree), (ib & 1) ^ 1);
                tc_fence_after();
                const uint32_t q_s = qd_smem + stage * 2 * QB_BYTES, do_s = q_s + QB_BYTES;
#pragma unroll
                for (int kk = 0; kk < DH / 16; ++kk) {
                    const uint64_t da = umma_desc_sw128(k_smem + kk * 32, 16, 1024);
                    const uint64_t db = umma_desc_sw128(q_s + kk * 32, 16, 1024);
                    umma_f16(st_tmem, da, db, IDESC_KK, kk > 0 ? 1u : 0u);
                }
#pragma unroll
                for (int kk = 0; kk < DH / 16; ++kk) {
                    const uint64_t da = umma_desc_sw128(v_smem + kk * 32, 16, 1024);
                    const uint64_t db = umma_desc_sw128(do_s + kk * 32, 16, 1024);
                    umma_f16(dpt_tmem, da, db, IDESC_KK, kk > 0 ? 1u : 0u);
                }
                umma_commit(smem_u32(&bar_s));
                if (ib > 0) issue_acc(ib - 1, prev_stage);
                prev_stage = stage;
                if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
            }
            issue_acc(nb - 1, prev_stage);
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;          // key row within the tile == TMEM lane
        const int kj = j0 + r;                // key index within the sequence
        const int sid = threadIdx.x - 64;     // 0..127 among the row threads
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int64_t bh = static_cast<int64_t>(b) * H + h;
        const float scale_log2e = scale * 1.4426950408889634f;
        const bool key_ok = kj < kv_len;
        const float2 sc2 = make_float2(scale_log2e, scale_log2e), ss2 = make_float2(scale, scale);
        const uint32_t pt_row = pt_smem + r * 128, dst_row = dst_smem + r * 128;
        for (int ib = 0; ib < nb; ++ib) {
            const int i0 = i_begin + ib * BI;
            {   // lse (log2 units) and delta of this block's queries: threads 0..63 / 64..127 fetch one value each
                const int qi = i0 + (sid & 63);
                if (sid < 64) {      // staged as the addends of the two fused multiply-adds: -lse log2e and -delta scale
                    const float l = (qi < S) ? lse[bh * S + qi] : INFINITY;
                    lse_s[ib & 1][sid] = (l == INFINITY) ? -INFINITY : -l * 1.4426950408889634f;
                } else {
                    dl_s[ib & 1][sid - 64] = (qi < S) ? -delta[bh * S + qi] * scale : 0.f;
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(smem_u32(&bar_s), ib & 1);
            tc_fence_after();
            if (ib > 0) {     // the dV / dK MMAs of block ib-1 must have consumed the P^T / dS^T tiles
                mbar_wait(smem_u32(&bar_o), (ib - 1) & 1);
                tc_fence_after();
            }
            // per-element predicate only where the (key tile, query block) pair is not uniformly allowed
            bool need_mask = false;
            if (mask_mode == VB_MASK_PREFIX_LM) need_mask = !((j0 + BK2 <= x_len) || (i0 >= x_len && j0 + BK2 - 1 <= i0));
            const float* ls = lse_s[ib & 1];
            const float* ds_ = dl_s[ib & 1];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t sv[32], dv[32];
                tmem_ld_32x32(lane_addr + half * 32, sv);
                tmem_ld_32x32(lane_addr + 64 + half * 32, dv);
                tmem_ld_wait();
                if (half == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bar_sfree));
                }
                uint32_t pk[16], dk_[16];
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    const float4 l4 = *reinterpret_cast<const float4*>(ls + half * 32 + c);
                    const float4 d4 = *reinterpret_cast<const float4*>(ds_ + half * 32 + c);
                    // packed fp32x2: t = s scale log2e + (-lse2_q), g = dp scale + (-delta_q scale), dS^T = p g
                    float2 t01 = __ffma2_rn(make_float2(__uint_as_float(sv[c]), __uint_as_float(sv[c + 1])), sc2, make_float2(l4.x, l4.y));
                    float2 t23 = __ffma2_rn(make_float2(__uint_as_float(sv[c + 2]), __uint_as_float(sv[c + 3])), sc2, make_float2(l4.z, l4.w));
                    const float2 g01 = __ffma2_rn(make_float2(__uint_as_float(dv[c]), __uint_as_float(dv[c + 1])), ss2, make_float2(d4.x, d4.y));
                    const float2 g23 = __ffma2_rn(make_float2(__uint_as_float(dv[c + 2]), __uint_as_float(dv[c + 3])), ss2, make_float2(d4.z, d4.w));
                    if (!key_ok) { t01 = make_float2(-INFINITY, -INFINITY); t23 = t01; }
                    else if (need_mask) {
                        const int qi = i0 + half * 32 + c;
                        const bool pre = kj < x_len;
                        if (!(pre || (qi >= x_len && kj <= qi))) t01.x = -INFINITY;
                        if (!(pre || (qi + 1 >= x_len && kj <= qi + 1))) t01.y = -INFINITY;
                        if (!(pre || (qi + 2 >= x_len && kj <= qi + 2))) t23.x = -INFINITY;
                        if (!(pre || (qi + 3 >= x_len && kj <= qi + 3))) t23.y = -INFINITY;
                    }
                    const float2 p01 = make_float2(ex2(t01.x), ex2(t01.y)), p23 = make_float2(ex2(t23.x), ex2(t23.y));
                    const float2 s01 = __fmul2_rn(p01, g01), s23 = __fmul2_rn(p23, g23);
                    pk[c >> 1] = pack_bf16x2(p01.x, p01.y);
                    pk[(c >> 1) + 1] = pack_bf16x2(p23.x, p23.y);
                    dk_[c >> 1] = pack_bf16x2(s01.x, s01.y);
                    dk_[(c >> 1) + 1] = pack_bf16x2(s23.x, s23.y);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t off = static_cast<uint32_t>((half * 4 + c) ^ (r & 7)) << 4;
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pt_row + off), "r"(pk[c * 4]), "r"(pk[c * 4 + 1]),
                                 "r"(pk[c * 4 + 2]), "r"(pk[c * 4 + 3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst_row + off), "r"(dk_[c * 4]), "r"(dk_[c * 4 + 1]),
                                 "r"(dk_[c * 4 + 2]), "r"(dk_[c * 4 + 3]) : "memory");
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_p));
        }
        __nv_bfloat16* krow = dqkv + (static_cast<int64_t>(row0) + kj) * 3 * d + d + h * DH;
        if (nb > 0) {
            mbar_wait(smem_u32(&bar_o), (nb - 1) & 1);
            tc_fence_after();
        }
#pragma unroll
        for (int part = 0; part < 2; ++part) {        // part 0: dV (TMEM [128,192)) -> column block 2d; part 1: dK -> column block d
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t ov[32];
                if (nb > 0) {
                    tmem_ld_32x32(lane_addr + 128 + part * 64 + half * 32, ov);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) ov[e] = 0u;
                }
                if (kj < S) {
                    __nv_bfloat16* dst = krow + (part == 0 ? d : 0) + half * 32;
#pragma unroll
                    for (int e = 0; e < 32; e += 8) {
                        uint4 w;
                        w.x = pack_bf16x2(__uint_as_float(ov[e]), __uint_as_float(ov[e + 1]));
                        w.y = pack_bf16x2(__uint_as_float(ov[e + 2]), __uint_as_float(ov[e + 3]));
                        w.z = pack_bf16x2(__uint_as_float(ov[e + 4]), __uint_as_float(ov[e + 5]));
                        w.w = pack_bf16x2(__uint_as_float(ov[e + 6]), __uint_as_float(ov[e + 7]));
                        *reinterpret_cast<uint4*>(dst + e) = w;
                    }
                }
            }
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

bool vb_attention_bwd_tc_enabled() {      // VALLE_B200_ATTN_BWD_TC=0: keep the mma.sync dQ kernel
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VALLE_B200_ATTN_BWD_TC");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

// dq part of dqkv (bf16 [B*S][3][H][64]) and delta (fp32 [B][H][S]) from qkv, o, dO and the forward's lse
int vb_attention_bwd_dq_tc(const void* qkv, const void* o, const void* dO, void* dqkv, const float* lse, float* delta, int B, int S,
                           int H, int mask_mode, const int32_t* x_lens, const int32_t* kv_lens, cudaStream_t st) {
    const int64_t d = static_cast<int64_t>(H) * DH;
    const int64_t rows = static_cast<int64_t>(B) * S;
    CUtensorMap tq, tkv, tdo;
    int rc;
    if ((rc = vb_make_tmap_bf16_2d(&tq, qkv, rows, 3 * d, 3 * d, BQ, DH)) != VB_OK) return rc;
    if ((rc = vb_make_tmap_bf16_2d(&tkv, qkv, rows, 3 * d, 3 * d, BKV, DH)) != VB_OK) return rc;
    if ((rc = vb_make_tmap_bf16_2d(&tdo, dO, rows, d, d, BQ, DH)) != VB_OK) return rc;
    static bool configured = false;
    if (!configured) {
        VB_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    dim3 grid(static_cast<unsigned>(vb_ceil_div(S, BQ)), H, B);
    const float scale = 1.0f / sqrtf(static_cast<float>(DH));
    VB_CUDA(vb_launch(false, attn_bwd_dq_tc_kernel, grid, dim3(THREADS), SMEM_BYTES, st, tq, tkv, tdo,
                      static_cast<const __nv_bfloat16*>(o), static_cast<const __nv_bfloat16*>(dO),
                      static_cast<__nv_bfloat16*>(dqkv), lse, delta, S, H, mask_mode, x_lens, kv_lens, scale));
    return VB_OK;
}

// dk, dv parts of dqkv from qkv, dO, lse and delta (written by vb_attention_bwd_dq_tc)
int vb_attention_bwd_dkv_tc(const void* qkv, const void* dO, void* dqkv, const float* lse, const float* delta, int B, int S, int H,
                            int mask_mode, const int32_t* x_lens, const int32_t* kv_lens, cudaStream_t st) {
    const int64_t d = static_cast<int64_t>(H) * DH;
    const int64_t rows = static_cast<int64_t>(B) * S;
    CUtensorMap tkt, tqb, tdob;
    int rc;
    if ((rc = vb_make_tmap_bf16_2d(&tkt, qkv, rows, 3 * d, 3 * d, BK2, DH)) != VB_OK) return rc;
    if ((rc = vb_make_tmap_bf16_2d(&tqb, qkv, rows, 3 * d, 3 * d, BI, DH)) != VB_OK) return rc;
    if ((rc = vb_make_tmap_bf16_2d(&tdob, dO, rows, d, d, BI, DH)) != VB_OK) return rc;
    static bool configured = false;
    if (!configured) {
        VB_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
        configured = true;
    }
    dim3 grid(static_cast<unsigned>(vb_ceil_div(S, BK2)), H, B);
    const float scale = 1.0f / sqrtf(static_cast<float>(DH));
    VB_CUDA(vb_launch(false, attn_bwd_dkv_tc_kernel, grid, dim3(THREADS), SMEM2_BYTES, st, tkt, tqb, tdob,
                      static_cast<__nv_bfloat16*>(dqkv), lse, delta, S, H, mask_mode, x_lens, kv_lens, scale));
    return VB_OK;
}
