// General attention (any Dh <= 256, any strides, fp32 or bf16 I/O, fp32 math): one warp per query row, online
// softmax over 32-key chunks.  This is the reference-faithful kernel behind the module-level API (arbitrary
// materialised masks, modules.py:160-167) and the fp32 validation mode; the throughput kernels are
// attn_decode.cu (paged decode, HBM-bound) and attn_prefill_tc.cu (tcgen05 flash attention).
#include <math.h>

#include "common.cuh"

namespace {

constexpr int WARPS = 4;
constexpr int MAX_DH = 256;

template <typename T, typename TO>
__global__ void __launch_bounds__(WARPS * 32) attention_simt_kernel(
    const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, int64_t q_sb, int64_t q_sh, int64_t q_ss,
    int64_t k_sb, int64_t k_sh, int64_t k_ss, int64_t v_sb, int64_t v_sh, int64_t v_ss, TO* __restrict__ o, int64_t o_sb,
    int64_t o_ss, int H, int Sq, int Sk, int Dh, int mask_mode, int q_pos0, const int32_t* __restrict__ x_lens,
    const int32_t* __restrict__ kv_lens, const uint8_t* __restrict__ mask, int64_t m_sb, int64_t m_sh, int64_t m_sq,
    float scale) {
    __shared__ float qs[WARPS][MAX_DH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * WARPS + warp;
    const int h = blockIdx.y, b = blockIdx.z;
    if (i >= Sq) return;
    const T* qr = q + b * q_sb + h * q_sh + static_cast<int64_t>(i) * q_ss;
    for (int e = lane; e < Dh; e += 32) qs[warp][e] = to_f32<T>(qr[e]) ;
    __syncwarp();
    const int kv_len = kv_lens ? min(kv_lens[b], Sk) : Sk;
    const int x_len = x_lens ? x_lens[b] : 0;
    const int qi = i + q_pos0;
    const T* kb = k + b * k_sb + h * k_sh;
    const T* vb = v + b * v_sb + h * v_sh;
    const uint8_t* mrow = (mask_mode == VB_MASK_EXPLICIT) ? mask + b * m_sb + h * m_sh + static_cast<int64_t>(i) * m_sq : nullptr;
    float m = -INFINITY, l = 0.f;
    float acc[MAX_DH / 32];
#pragma unroll
    for (int u = 0; u < MAX_DH / 32; ++u) acc[u] = 0.f;
    for (int j0 = 0; j0 < kv_len; j0 += 32) {
        const int j = j0 + lane;
        bool ok = j < kv_len;
        if (ok && mask_mode == VB_MASK_PREFIX_LM) ok = (j < x_len) || (qi >= x_len && j <= qi);
        if (ok && mask_mode == VB_MASK_EXPLICIT) ok = (mrow[j] == 0);
        float s = -INFINITY;
        if (ok) {
            const T* kr = kb + static_cast<int64_t>(j) * k_ss;
            float d = 0.f;
            for (int e = 0; e < Dh; ++e) d = fmaf(qs[warp][e], to_f32<T>(kr[e]), d);
            s = d * scale;
        }
        const float m_new = fmaxf(m, warp_max(s));
        if (m_new == -INFINITY) continue;   // whole chunk masked and nothing seen yet
        const float corr = (m == -INFINITY) ? 0.f : expf(m - m_new);
        const float pj = ok ? expf(s - m_new) : 0.f;
        l = l * corr + warp_sum(pj);
#pragma unroll
        for (int u = 0; u < MAX_DH / 32; ++u) acc[u] *= corr;
        const int cnt = min(32, kv_len - j0);
        for (int t = 0; t < cnt; ++t) {
            const float pt = __shfl_sync(0xffffffffu, pj, t);
            if (pt == 0.f) continue;
            const T* vr = vb + static_cast<int64_t>(j0 + t) * v_ss;
#pragma unroll
            for (int u = 0; u < MAX_DH / 32; ++u) {
                const int e = u * 32 + lane;
                if (e < Dh) acc[u] = fmaf(pt, to_f32<T>(vr[e]), acc[u]);
            }
        }
        m = m_new;
    }
    const float inv = (l > 0.f) ? 1.f / l : 0.f;
    TO* orow = o + b * o_sb + static_cast<int64_t>(i) * o_ss + h * Dh;
#pragma unroll
    for (int u = 0; u < MAX_DH / 32; ++u) {
        const int e = u * 32 + lane;
        if (e < Dh) orow[e] = from_f32<TO>(acc[u] * inv);
    }
}

}  // namespace

extern "C" int vb_attention(const void* q, const void* k, const void* v, int dtype, int64_t q_sb, int64_t q_sh, int64_t q_ss,
                            int64_t k_sb, int64_t k_sh, int64_t k_ss, int64_t v_sb, int64_t v_sh, int64_t v_ss, void* o,
                            int o_dtype, int64_t o_sb, int64_t o_ss, int B, int H, int Sq, int Sk, int Dh, int mask_mode,
                            int q_pos0, const int32_t* x_lens, const int32_t* kv_lens, const uint8_t* mask, int64_t m_sb,
                            int64_t m_sh, int64_t m_sq, void* stream) {
    VB_REQUIRE(q && k && v && o, VB_ERR_BAD_ARG, "vb_attention: null pointer");
    VB_REQUIRE(Dh >= 1 && Dh <= MAX_DH, VB_ERR_UNSUPPORTED, "vb_attention: head_dim %d not in [1,%d]", Dh, MAX_DH);
    VB_REQUIRE(B >= 0 && H >= 1 && Sq >= 0 && Sk >= 0 && H <= 65535 && B <= 65535, VB_ERR_BAD_ARG, "vb_attention: bad shape");
    VB_REQUIRE(mask_mode != VB_MASK_EXPLICIT || mask != nullptr, VB_ERR_BAD_ARG, "vb_attention: explicit mask is null");
    VB_REQUIRE(mask_mode != VB_MASK_PREFIX_LM || x_lens != nullptr, VB_ERR_BAD_ARG, "vb_attention: prefix-LM needs x_lens");
    if (B == 0 || Sq == 0) return VB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid(static_cast<unsigned>(vb_ceil_div(Sq, WARPS)), H, B);
    const float scale = 1.0f / sqrtf(static_cast<float>(Dh));
#define ATT(T, TO)                                                                                                       \
    attention_simt_kernel<T, TO><<<grid, WARPS * 32, 0, st>>>(static_cast<const T*>(q), static_cast<const T*>(k),        \
        static_cast<const T*>(v), q_sb, q_sh, q_ss, k_sb, k_sh, k_ss, v_sb, v_sh, v_ss, static_cast<TO*>(o), o_sb, o_ss, \
        H, Sq, Sk, Dh, mask_mode, q_pos0, x_lens, kv_lens, mask, m_sb, m_sh, m_sq, scale)
    if (dtype == VB_F32 && o_dtype == VB_F32) ATT(float, float);
    else if (dtype == VB_BF16 && o_dtype == VB_BF16) ATT(__nv_bfloat16, __nv_bfloat16);
    else if (dtype == VB_BF16 && o_dtype == VB_F32) ATT(__nv_bfloat16, float);
    else if (dtype == VB_F32 && o_dtype == VB_BF16) ATT(float, __nv_bfloat16);
    else VB_REQUIRE(false, VB_ERR_BAD_ARG, "vb_attention: bad dtype");
#undef ATT
    VB_LAUNCH_CHECK();
    return VB_OK;
}
