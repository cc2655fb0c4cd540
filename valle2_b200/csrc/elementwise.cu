// HBM-bound row kernels of the hot path: embedding-sum + PE (K1/K2), split-K reduce + bias + residual +
// (Ada)LayerNorm (K3), split-K reduce + bias + erf-GELU (K8), KV scatter into the paged pool (K5), and the
// device-side beam bookkeeping (K11).  All are coalesced, 16-byte vectorised where the shape allows.
#include <stdlib.h>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// K1/K2  embedding sum + sinusoidal PE
// ------------------------------------------------------------------------------------------------
__global__ void embed_sum_pe_kernel(const int32_t* __restrict__ ids, const float* __restrict__ tables,
                                    const float* __restrict__ pe, float* __restrict__ out, int T, int Q, int V, int d,
                                    int t_split, int nq_a, int nq_b, int pos_offset, const int32_t* __restrict__ pos_b,
                                    int max_len, int64_t out_rows_per_batch, int64_t out_row_offset) {
    pdl_trigger();
    pdl_wait();
    const int t = blockIdx.x, b = blockIdx.y;
    const int nq = (t < t_split) ? nq_a : nq_b;
    int pos = (pos_b ? pos_b[b] : pos_offset) + t;
    pos = min(max(pos, 0), max_len - 1);
    const int32_t* id_row = ids + (static_cast<int64_t>(b) * T + t) * Q;
    float* o = out + (static_cast<int64_t>(b) * out_rows_per_batch + out_row_offset + t) * d;
    const float* pe_row = pe + static_cast<int64_t>(pos) * d;
    for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < nq; ++j) {  // left-to-right, same order as the reference's += loop
            int id = id_row[j];
            id = min(max(id, 0), V - 1);
            const float4 e = *reinterpret_cast<const float4*>(tables + (static_cast<int64_t>(j) * V + id) * d + c);
            acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
        }
        const float4 p = *reinterpret_cast<const float4*>(pe_row + c);
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
        *reinterpret_cast<float4*>(o + c) = acc;
    }
}

extern "C" int vb_embed_sum_pe(const int32_t* ids, const float* tables, const float* pe, float* out, int B, int T, int Q,
                               int V, int d, int t_split, int nq_a, int nq_b, int pos_offset, const int32_t* pos_b,
                               int max_len, int64_t out_rows_per_batch, int64_t out_row_offset, void* stream) {
    VB_REQUIRE(ids && tables && pe && out, VB_ERR_BAD_ARG, "vb_embed_sum_pe: null pointer");
    VB_REQUIRE(B >= 0 && T >= 0 && Q >= 1 && d > 0 && d % 4 == 0, VB_ERR_BAD_ARG, "vb_embed_sum_pe: bad shape B=%d T=%d Q=%d d=%d", B, T, Q, d);
    VB_REQUIRE(nq_a >= 0 && nq_a <= Q && nq_b >= 0 && nq_b <= Q, VB_ERR_BAD_ARG, "vb_embed_sum_pe: nq out of range");
    if (B == 0 || T == 0) return VB_OK;
    VB_REQUIRE(B <= 65535, VB_ERR_BAD_ARG, "vb_embed_sum_pe: B too large");
    int threads = min(256, max(32, d / 4));
    VB_CUDA(vb_launch(T == 1, embed_sum_pe_kernel, dim3(T, B), dim3(threads), 0, static_cast<cudaStream_t>(stream), ids, tables, pe, out,
                      T, Q, V, d, t_split, nq_a, nq_b, pos_offset, pos_b, max_len, out_rows_per_batch, out_row_offset));
    return VB_OK;
}

// Fused embedding-sum + PE + first (Ada)LayerNorm: one warp per row, the row stays in registers between the sum and the
// normalisation, so the fp32 residual row is written once and never read back for layer 0's norm1.  Same lane layout and the same
// order of operations as residual_layernorm_kernel below, hence bit-identical to vb_embed_sum_pe followed by vb_residual_layernorm.
template <typename TY, int NV>  // NV = d / 128 float4 chunks per lane
__global__ void __launch_bounds__(256) embed_sum_pe_norm_kernel(const int32_t* __restrict__ ids, const float* __restrict__ tables,
                                                                const float* __restrict__ pe, float* __restrict__ out, int B, int T, int Q,
                                                                int V, int d, int t_split, int nq_a, int nq_b, int pos_offset,
                                                                const int32_t* __restrict__ pos_b, int max_len,
                                                                int64_t out_rows_per_batch, int64_t out_row_offset,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float eps, TY* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= static_cast<int64_t>(B) * T) return;
    const int b = static_cast<int>(r / T), t = static_cast<int>(r % T);
    const int nq = (t < t_split) ? nq_a : nq_b;
    int pos = (pos_b ? pos_b[b] : pos_offset) + t;
    pos = min(max(pos, 0), max_len - 1);
    const int32_t* id_row = ids + (static_cast<int64_t>(b) * T + t) * Q;
    const int64_t orow = static_cast<int64_t>(b) * out_rows_per_batch + out_row_offset + t;
    const float* pe_row = pe + static_cast<int64_t>(pos) * d;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < nq; ++j) {  // left-to-right, same order as the reference's += loop
        int id = id_row[j];
        id = min(max(id, 0), V - 1);
        const float* e_row = tables + (static_cast<int64_t>(j) * V + id) * d;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 e = *reinterpret_cast<const float4*>(e_row + (i * 32 + lane) * 4);
            v[i].x += e.x; v[i].y += e.y; v[i].z += e.z; v[i].w += e.w;
        }
    }
    float* o = out + orow * d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        const float4 p = *reinterpret_cast<const float4*>(pe_row + c);
        v[i].x += p.x; v[i].y += p.y; v[i].z += p.z; v[i].w += p.w;
        *reinterpret_cast<float4*>(o + c) = v[i];
    }
    TY* yr = y + orow * d;
    if (gamma == nullptr) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if constexpr (sizeof(TY) == 4) *reinterpret_cast<float4*>(yr + c) = v[i];
            else *reinterpret_cast<uint2*>(yr + c) = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
        }
        return;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, b2 = v[i].y - mean, c2 = v[i].z - mean, e2 = v[i].w - mean;
        q += (a * a + b2 * b2) + (c2 * c2 + e2 * e2);
    }
    const float rstd = rsqrtf(warp_sum(q) / d + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        const float4 g = *reinterpret_cast<const float4*>(gamma + c);
        const float4 bt = *reinterpret_cast<const float4*>(beta + c);
        const float o0 = (v[i].x - mean) * rstd * g.x + bt.x, o1 = (v[i].y - mean) * rstd * g.y + bt.y;
        const float o2 = (v[i].z - mean) * rstd * g.z + bt.z, o3 = (v[i].w - mean) * rstd * g.w + bt.w;
        if constexpr (sizeof(TY) == 4) *reinterpret_cast<float4*>(yr + c) = make_float4(o0, o1, o2, o3);
        else *reinterpret_cast<uint2*>(yr + c) = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
    }
}

extern "C" int vb_embed_sum_pe_norm(const int32_t* ids, const float* tables, const float* pe, float* out, int B, int T, int Q,
                                    int V, int d, int t_split, int nq_a, int nq_b, int pos_offset, const int32_t* pos_b,
                                    int max_len, int64_t out_rows_per_batch, int64_t out_row_offset, const float* gamma,
                                    const float* beta, float eps, void* y, int y_dtype, void* stream) {
    VB_REQUIRE(ids && tables && pe && out && y, VB_ERR_BAD_ARG, "vb_embed_sum_pe_norm: null pointer");
    VB_REQUIRE((gamma == nullptr) == (beta == nullptr), VB_ERR_BAD_ARG, "vb_embed_sum_pe_norm: gamma and beta go together");
    VB_REQUIRE(B >= 0 && T >= 0 && Q >= 1, VB_ERR_BAD_ARG, "vb_embed_sum_pe_norm: bad shape B=%d T=%d Q=%d", B, T, Q);
    VB_REQUIRE(d == 256 || d == 512 || d == 1024, VB_ERR_UNSUPPORTED, "vb_embed_sum_pe_norm: d must be 256, 512 or 1024 (got %d)", d);
    VB_REQUIRE(nq_a >= 0 && nq_a <= Q && nq_b >= 0 && nq_b <= Q, VB_ERR_BAD_ARG, "vb_embed_sum_pe_norm: nq out of range");
    VB_REQUIRE(y_dtype == VB_F32 || y_dtype == VB_BF16, VB_ERR_BAD_ARG, "vb_embed_sum_pe_norm: bad y_dtype");
    if (B == 0 || T == 0) return VB_OK;
    const int64_t rows = static_cast<int64_t>(B) * T;
    const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define EPN(TY, NV)                                                                                                          \
    embed_sum_pe_norm_kernel<TY, NV><<<grid, 256, 0, st>>>(ids, tables, pe, out, B, T, Q, V, d, t_split, nq_a, nq_b, pos_offset, \
                                                           pos_b, max_len, out_rows_per_batch, out_row_offset, gamma, beta, eps, \
                                                           static_cast<TY*>(y))
    if (y_dtype == VB_BF16) {
        if (d == 256) EPN(__nv_bfloat16, 2); else if (d == 512) EPN(__nv_bfloat16, 4); else EPN(__nv_bfloat16, 8);
    } else {
        if (d == 256) EPN(float, 2); else if (d == 512) EPN(float, 4); else EPN(float, 8);
    }
#undef EPN
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// K3  [split-K reduce + bias + residual] + LayerNorm;  one warp per row, row kept in registers
// ------------------------------------------------------------------------------------------------
template <typename TY, int NV>  // NV = d / 128 float4 chunks per lane
__global__ void __launch_bounds__(256) residual_layernorm_kernel(float* __restrict__ x, const float* __restrict__ part,
                                                                 int n_part, int64_t part_stride,
                                                                 const float* __restrict__ bias,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, TY* __restrict__ y,
                                                                 int64_t R, int d, float eps) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    float* xr = x + r * d;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
    if (n_part > 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            float4 acc = bias ? *reinterpret_cast<const float4*>(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s = 0; s < n_part; ++s) {
                const float4 p = *reinterpret_cast<const float4*>(part + s * part_stride + r * d + c);
                acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
            }
            v[i].x += acc.x; v[i].y += acc.y; v[i].z += acc.z; v[i].w += acc.w;
            *reinterpret_cast<float4*>(xr + c) = v[i];
        }
    }
    if (y == nullptr) return;
    if (gamma == nullptr) {   // no norm: y = cast(x)  (the stack has no final LayerNorm, reference K-2)
        TY* yc = y + r * d;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if constexpr (sizeof(TY) == 4) *reinterpret_cast<float4*>(yc + c) = v[i];
            else *reinterpret_cast<uint2*>(yc + c) = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
        }
        return;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
        q += (a * a + b * b) + (c * c + e * e);
    }
    const float rstd = rsqrtf(warp_sum(q) / d + eps);
    TY* yr = y + r * d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        const float4 g = *reinterpret_cast<const float4*>(gamma + c);
        const float4 bt = *reinterpret_cast<const float4*>(beta + c);
        const float o0 = (v[i].x - mean) * rstd * g.x + bt.x, o1 = (v[i].y - mean) * rstd * g.y + bt.y;
        const float o2 = (v[i].z - mean) * rstd * g.z + bt.z, o3 = (v[i].w - mean) * rstd * g.w + bt.w;
        if constexpr (sizeof(TY) == 4) {
            *reinterpret_cast<float4*>(yr + c) = make_float4(o0, o1, o2, o3);
        } else {
            uint2 pk = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
            *reinterpret_cast<uint2*>(yr + c) = pk;
        }
    }
}

// generic d (any multiple of 1): one warp per row, row re-read from global (L1/L2 resident)
template <typename TY>
__global__ void __launch_bounds__(256) residual_layernorm_generic_kernel(float* __restrict__ x, const float* __restrict__ part,
                                                                         int n_part, int64_t part_stride,
                                                                         const float* __restrict__ bias,
                                                                         const float* __restrict__ gamma,
                                                                         const float* __restrict__ beta, TY* __restrict__ y,
                                                                         int64_t R, int d, float eps) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    float* xr = x + r * d;
    if (n_part > 0) {
        for (int c = lane; c < d; c += 32) {
            float acc = bias ? bias[c] : 0.f;
            for (int s = 0; s < n_part; ++s) acc += part[s * part_stride + r * d + c];
            xr[c] += acc;
        }
        __syncwarp();
    }
    if (y == nullptr) return;
    if (gamma == nullptr) {
        for (int c = lane; c < d; c += 32) y[r * d + c] = from_f32<TY>(xr[c]);
        return;
    }
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += xr[c];
    const float mean = warp_sum(s) / d;
    float q = 0.f;
    for (int c = lane; c < d; c += 32) {
        const float a = xr[c] - mean;
        q += a * a;
    }
    const float rstd = rsqrtf(warp_sum(q) / d + eps);
    for (int c = lane; c < d; c += 32) y[r * d + c] = from_f32<TY>((xr[c] - mean) * rstd * gamma[c] + beta[c]);
}

// Few rows (decode: R = batch): one 256-thread CTA per row so that the split-K slices are read with many independent
// 16-byte loads in flight instead of one warp walking them serially.  d <= 4096, d % 4 == 0.
template <typename TY>
__global__ void __launch_bounds__(256) residual_layernorm_block_kernel(float* __restrict__ x, const float* __restrict__ part,
                                                                       int n_part, int64_t part_stride,
                                                                       const float* __restrict__ bias,
                                                                       const float* __restrict__ gamma,
                                                                       const float* __restrict__ beta, TY* __restrict__ y,
                                                                       int d, float eps, unsigned long long* __restrict__ dbg) {
    __shared__ float red[8];
    // optional timeline (vb_residual_layernorm_set_debug): %globaltimer [row][4] = {start, dependency resolved, row loaded, stored}
#define LN_STAMP(ev) do { if (dbg != nullptr && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); dbg[blockIdx.x * 4 + (ev)] = t_; } } while (0)
    LN_STAMP(0);
    pdl_trigger();
    const int64_t r = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* xr = x + r * d;
    constexpr int MAXC = 4;
    // gamma / beta / bias are weights: fetched BEFORE the dependency resolves (they used to be read after the two block
    // reductions -- one more L2 round trip on the critical path of every decode-step LayerNorm)
    float4 g4[MAXC], b4[MAXC], bias4[MAXC];
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c4 = tid + i * 256;
        g4[i] = b4[i] = bias4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 < (d >> 2)) {
            if (gamma != nullptr && y != nullptr) {
                g4[i] = *reinterpret_cast<const float4*>(gamma + c4 * 4);
                b4[i] = *reinterpret_cast<const float4*>(beta + c4 * 4);
            }
            if (n_part > 0 && bias != nullptr) bias4[i] = *reinterpret_cast<const float4*>(bias + c4 * 4);
        }
    }
    pdl_wait();
    LN_STAMP(1);
    float4 v[MAXC];
    const int nchunk = d >> 2;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c4 = tid + i * 256;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 < nchunk) {
            const int c = c4 * 4;
            float4 a = *reinterpret_cast<const float4*>(xr + c);
            if (n_part > 0) {
                float4 acc = bias4[i];
                const float* pp = part + r * d + c;
                int s = 0;
                // Sixteen / eight slices at a time with EVERY load issued before the first add (summed in slice order).  The
                // empty asm consumes all loaded registers at once: without it ptxas interleaves load and add to save
                // registers and the slices arrive as a chain of dependent L2 round trips (3.6 us for this kernel in a B=32
                // step with 16 slices -- tools/step_timeline.py).
                for (; s + 16 <= n_part; s += 16) {
                    float4 p[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) p[k] = __ldcg(reinterpret_cast<const float4*>(pp + (s + k) * part_stride));
                    asm volatile("" ::"f"(p[0].x), "f"(p[1].x), "f"(p[2].x), "f"(p[3].x), "f"(p[4].x), "f"(p[5].x), "f"(p[6].x), "f"(p[7].x),
                                 "f"(p[8].x), "f"(p[9].x), "f"(p[10].x), "f"(p[11].x), "f"(p[12].x), "f"(p[13].x), "f"(p[14].x), "f"(p[15].x));
#pragma unroll
                    for (int k = 0; k < 16; ++k) { acc.x += p[k].x; acc.y += p[k].y; acc.z += p[k].z; acc.w += p[k].w; }
                }
                for (; s + 8 <= n_part; s += 8) {
                    float4 p[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) p[k] = __ldcg(reinterpret_cast<const float4*>(pp + (s + k) * part_stride));
                    asm volatile("" ::"f"(p[0].x), "f"(p[1].x), "f"(p[2].x), "f"(p[3].x), "f"(p[4].x), "f"(p[5].x), "f"(p[6].x), "f"(p[7].x));
#pragma unroll
                    for (int k = 0; k < 8; ++k) { acc.x += p[k].x; acc.y += p[k].y; acc.z += p[k].z; acc.w += p[k].w; }
                }
                for (; s + 4 <= n_part; s += 4) {   // four independent loads in flight, summed in slice order
                    const float4 p0 = *reinterpret_cast<const float4*>(pp + (s + 0) * part_stride);
                    const float4 p1 = *reinterpret_cast<const float4*>(pp + (s + 1) * part_stride);
                    const float4 p2 = *reinterpret_cast<const float4*>(pp + (s + 2) * part_stride);
                    const float4 p3 = *reinterpret_cast<const float4*>(pp + (s + 3) * part_stride);
                    acc.x += p0.x; acc.y += p0.y; acc.z += p0.z; acc.w += p0.w;
                    acc.x += p1.x; acc.y += p1.y; acc.z += p1.z; acc.w += p1.w;
                    acc.x += p2.x; acc.y += p2.y; acc.z += p2.z; acc.w += p2.w;
                    acc.x += p3.x; acc.y += p3.y; acc.z += p3.z; acc.w += p3.w;
                }
                for (; s < n_part; ++s) {
                    const float4 p0 = *reinterpret_cast<const float4*>(pp + s * part_stride);
                    acc.x += p0.x; acc.y += p0.y; acc.z += p0.z; acc.w += p0.w;
                }
                a.x += acc.x; a.y += acc.y; a.z += acc.z; a.w += acc.w;
                *reinterpret_cast<float4*>(xr + c) = a;
            }
            v[i] = a;
        }
    }
    LN_STAMP(2);
    if (y == nullptr) return;
    TY* yr = y + r * d;
    if (gamma == nullptr) {
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c4 = tid + i * 256;
            if (c4 < nchunk) {
                if constexpr (sizeof(TY) == 4) *reinterpret_cast<float4*>(yr + c4 * 4) = v[i];
                else *reinterpret_cast<uint2*>(yr + c4 * 4) = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
            }
        }
        return;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);   // chunks past d hold zeros
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w];
    const float mean = tot / d;
    __syncthreads();
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        if (tid + i * 256 < nchunk) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
            q += (a * a + b * b) + (c * c + e * e);
        }
    }
    q = warp_sum(q);
    if (lane == 0) red[warp] = q;
    __syncthreads();
    float qt = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) qt += red[w];
    const float rstd = rsqrtf(qt / d + eps);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c4 = tid + i * 256;
        if (c4 < nchunk) {
            const int c = c4 * 4;
            const float4 g = g4[i];
            const float4 bt = b4[i];
            const float o0 = (v[i].x - mean) * rstd * g.x + bt.x, o1 = (v[i].y - mean) * rstd * g.y + bt.y;
            const float o2 = (v[i].z - mean) * rstd * g.z + bt.z, o3 = (v[i].w - mean) * rstd * g.w + bt.w;
            if constexpr (sizeof(TY) == 4) *reinterpret_cast<float4*>(yr + c) = make_float4(o0, o1, o2, o3);
            else *reinterpret_cast<uint2*>(yr + c) = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
        }
    }
    LN_STAMP(3);
#undef LN_STAMP
}

// Decode rows with many split-K slices: one row spread over a CLUSTER of four CTAs (a quarter of the columns each), so that
// the 36-68 KB of slices of a row arrive through four SMs' L2 ports instead of one (the one-CTA-per-row kernel spends
// 1.3-2.0 us of its ~3 us waiting for them -- tools/step_timeline.py).  Inside a CTA the slices are split over 256 / nch
// thread groups (nch = d / 16 float4 chunks per CTA), summed per group in slice order and then across groups in group order
// (deterministic); the row statistics (mean, M2 of a quarter row) are exchanged through distributed shared memory and
// combined with Chan's equal-count formula.  d in {256, 512, 1024}.
template <typename TY>
__global__ void __launch_bounds__(256) residual_layernorm_cluster_kernel(float* __restrict__ x, const float* __restrict__ part,
                                                                         int n_part, int64_t part_stride,
                                                                         const float* __restrict__ bias,
                                                                         const float* __restrict__ gamma,
                                                                         const float* __restrict__ beta, TY* __restrict__ y,
                                                                         int d, float eps) {
    __shared__ float4 gsum[256];
    __shared__ float wred[2][4];
    __shared__ __align__(8) float2 cstat[4];          // (mean, M2) of each CTA's quarter row, written by every rank
    const int rank = blockIdx.x;                      // cluster rank == blockIdx.x (cluster = 4 x 1 x 1)
    const int64_t r = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nch = d >> 4;                           // float4 chunks of this CTA's quarter row: 64, 32 or 16
    const int ng = 256 / nch;                         // slice groups
    const int chunk = tid % nch, grp = tid / nch;
    const int c = rank * (d >> 2) + chunk * 4;        // first column of this thread's chunk
    pdl_trigger();
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = g4, bias4 = g4;
    if (grp == 0) {                                   // weights: before the dependency resolves
        if (gamma != nullptr && y != nullptr) {
            g4 = *reinterpret_cast<const float4*>(gamma + c);
            b4 = *reinterpret_cast<const float4*>(beta + c);
        }
        if (bias != nullptr) bias4 = *reinterpret_cast<const float4*>(bias + c);
    }
    pdl_wait();
    float* xr = x + r * d;
    const float* pp = part + r * d + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 xv = acc;
    if (grp == 0) xv = *reinterpret_cast<const float4*>(xr + c);
    {
        float4 p[4];                                  // up to four slices of this group in flight (16 slices, 4 groups)
        int s = grp;
        for (; s + 3 * ng < n_part; s += 4 * ng) {
#pragma unroll
            for (int k = 0; k < 4; ++k) p[k] = __ldcg(reinterpret_cast<const float4*>(pp + static_cast<int64_t>(s + k * ng) * part_stride));
#pragma unroll
            for (int k = 0; k < 4; ++k) { acc.x += p[k].x; acc.y += p[k].y; acc.z += p[k].z; acc.w += p[k].w; }
        }
        for (; s < n_part; s += ng) {
            const float4 q = __ldcg(reinterpret_cast<const float4*>(pp + static_cast<int64_t>(s) * part_stride));
            acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
        }
    }
    gsum[tid] = acc;
    __syncthreads();
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grp == 0) {                                   // nch threads (1/2, 1 or 2 warps) finish the row quarter
        float4 t = bias4;
        for (int g = 0; g < ng; ++g) { const float4 q = gsum[g * nch + chunk]; t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w; }
        a = make_float4(xv.x + t.x, xv.y + t.y, xv.z + t.z, xv.w + t.w);
        *reinterpret_cast<float4*>(xr + c) = a;
    }
    if (y == nullptr) return;                         // uniform: residual update only
    TY* yr = y + r * d;
    if (gamma == nullptr) {                           // plain cast
        if (grp == 0) {
            if constexpr (sizeof(TY) == 4) *reinterpret_cast<float4*>(yr + c) = a;
            else *reinterpret_cast<uint2*>(yr + c) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
        }
        return;
    }
    // quarter-row statistics: the first nch threads hold the values (whole warps: nch is 16, 32 or 64)
    const int nw = (nch + 31) >> 5;                   // warps that hold data: 1 or 2
    const float cnt = static_cast<float>(d >> 2);
    float sm = (grp == 0) ? (a.x + a.y) + (a.z + a.w) : 0.f;
    if (warp < nw) { sm = warp_sum(sm); if (lane == 0) wred[0][warp] = sm; }
    __syncthreads();
    const float mean_c = ((nw == 2) ? wred[0][0] + wred[0][1] : wred[0][0]) / cnt;
    float m2 = 0.f;
    if (grp == 0) { const float e0 = a.x - mean_c, e1 = a.y - mean_c, e2 = a.z - mean_c, e3 = a.w - mean_c; m2 = (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3); }
    if (warp < nw) { m2 = warp_sum(m2); if (lane == 0) wred[1][warp] = m2; }
    __syncthreads();
    if (tid < 4) {                                    // thread t publishes this CTA's (mean, M2) into rank t's cstat[rank]
        const float m2_c = (nw == 2) ? wred[1][0] + wred[1][1] : wred[1][0];
        uint32_t ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(&cstat[rank])), "r"(tid));
        asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(ra), "f"(mean_c), "f"(m2_c) : "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (grp == 0) {
        const float2 s0 = cstat[0], s1 = cstat[1], s2 = cstat[2], s3 = cstat[3];
        const float mean = ((s0.x + s1.x) + (s2.x + s3.x)) * 0.25f;
        const float dev = ((s0.x - mean) * (s0.x - mean) + (s1.x - mean) * (s1.x - mean)) + ((s2.x - mean) * (s2.x - mean) + (s3.x - mean) * (s3.x - mean));
        const float var = (((s0.y + s1.y) + (s2.y + s3.y)) + cnt * dev) / static_cast<float>(d);
        const float rstd = rsqrtf(var + eps);
        const float o0 = (a.x - mean) * rstd * g4.x + b4.x, o1 = (a.y - mean) * rstd * g4.y + b4.y;
        const float o2 = (a.z - mean) * rstd * g4.z + b4.z, o3 = (a.w - mean) * rstd * g4.w + b4.w;
        if constexpr (sizeof(TY) == 4) *reinterpret_cast<float4*>(yr + c) = make_float4(o0, o1, o2, o3);
        else *reinterpret_cast<uint2*>(yr + c) = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
    }
}

static unsigned long long* g_ln_dbg = nullptr;
extern "C" int vb_residual_layernorm_set_debug(void* buf) {   /* device buffer of R * 4 uint64 stamps (decode-shape launches), or NULL */
    g_ln_dbg = static_cast<unsigned long long*>(buf);
    return VB_OK;
}

template <typename TY>
static int launch_rln(float* x, const float* part, int n_part, int64_t part_stride, const float* bias, const float* gamma,
                      const float* beta, TY* y, int64_t R, int d, float eps, cudaStream_t st) {
    static const bool ln_cluster = !(getenv("VALLE_B200_LN_CLUSTER") != nullptr && getenv("VALLE_B200_LN_CLUSTER")[0] == '0');
    // The 4-CTA cluster per row is a latency form (one column per thread, the row statistics meet in DSMEM); above 160 rows
    // the CTA-per-row form with float4 loads moves the slices faster: decode step at batch 256 / 192 1.976 / 1.593 vs 2.034 /
    // 1.647 ms, at batch 160 / 128 1.397 / 1.209 vs 1.386 / 1.199 (profiles/r02d_ab_ln_cluster.jsonl)
    if (ln_cluster && g_ln_dbg == nullptr && R <= 160 && n_part >= 4 && (d == 256 || d == 512 || d == 1024)) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(4, static_cast<unsigned>(R));
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        int na = 0;
        if (vb_pdl_enabled()) {
            attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 4; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
        cfg.attrs = attr;
        cfg.numAttrs = na;
        VB_CUDA(cudaLaunchKernelEx(&cfg, residual_layernorm_cluster_kernel<TY>, x, part, n_part, part_stride, bias, gamma, beta, y, d, eps));
        return VB_OK;
    }
    if (R <= 1024 && d % 4 == 0 && d <= 4096) {
        VB_CUDA(vb_launch(true, residual_layernorm_block_kernel<TY>, dim3(static_cast<unsigned>(R)), dim3(256), 0, st, x, part, n_part,
                          part_stride, bias, gamma, beta, y, d, eps, g_ln_dbg));
        return VB_OK;
    }
    const int warps = 8;
    const dim3 grid(static_cast<unsigned>(vb_ceil_div(R, warps)));
#define RLN(NV) VB_CUDA(vb_launch(false, residual_layernorm_kernel<TY, NV>, grid, dim3(warps * 32), 0, st, x, part, n_part, part_stride, bias, gamma, beta, y, R, d, eps))
    switch (d) {
        case 128: RLN(1); break;
        case 256: RLN(2); break;
        case 512: RLN(4); break;
        case 1024: RLN(8); break;
        default:
            VB_CUDA(vb_launch(false, residual_layernorm_generic_kernel<TY>, grid, dim3(warps * 32), 0, st, x, part, n_part, part_stride, bias,
                              gamma, beta, y, R, d, eps));
    }
#undef RLN
    return VB_OK;
}

extern "C" int vb_residual_layernorm(float* x, const float* part, int n_part, int64_t part_stride, const float* bias,
                                     const float* gamma, const float* beta, void* y, int y_dtype, int64_t R, int d,
                                     float eps, void* stream) {
    VB_REQUIRE(x != nullptr && R >= 0 && d > 0, VB_ERR_BAD_ARG, "vb_residual_layernorm: bad args");
    VB_REQUIRE(n_part == 0 || part != nullptr, VB_ERR_BAD_ARG, "vb_residual_layernorm: n_part > 0 needs part");
    VB_REQUIRE(y == nullptr || (gamma != nullptr) == (beta != nullptr), VB_ERR_BAD_ARG,
               "vb_residual_layernorm: gamma and beta must both be given (LayerNorm) or both be null (plain cast)");
    VB_REQUIRE(y_dtype == VB_F32 || y_dtype == VB_BF16, VB_ERR_BAD_ARG, "vb_residual_layernorm: bad y_dtype");
    if (R == 0) return VB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (y_dtype == VB_F32)
        return launch_rln<float>(x, part, n_part, part_stride, bias, gamma, beta, static_cast<float*>(y), R, d, eps, st);
    return launch_rln<__nv_bfloat16>(x, part, n_part, part_stride, bias, gamma, beta, static_cast<__nv_bfloat16*>(y), R, d, eps, st);
}

// ------------------------------------------------------------------------------------------------
// K8  split-K reduce + bias + erf-GELU
// ------------------------------------------------------------------------------------------------
template <typename TY>
__global__ void reduce_bias_act_kernel(const float* __restrict__ part, int n_part, int64_t part_stride,
                                       const float* __restrict__ bias, int gelu, TY* __restrict__ y, int64_t total4, int N) {
    pdl_trigger();
    // the bias of this thread's first element is a weight: fetched before the dependency resolves
    const int64_t i_first = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    float4 bias_first = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias && i_first < total4) bias_first = *reinterpret_cast<const float4*>(bias + static_cast<int>((i_first * 4) % N));
    pdl_wait();
    for (int64_t i = i_first; i < total4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t e = i * 4;
        const int n = static_cast<int>(e % N);
        float4 acc = (i == i_first) ? bias_first : (bias ? *reinterpret_cast<const float4*>(bias + n) : make_float4(0.f, 0.f, 0.f, 0.f));
        int s = 0;
        for (; s + 4 <= n_part; s += 4) {       // four slices with every load issued before the first add (see the LayerNorm kernel)
            float4 p[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) p[k] = __ldcg(reinterpret_cast<const float4*>(part + (s + k) * part_stride + e));
            asm volatile("" ::"f"(p[0].x), "f"(p[1].x), "f"(p[2].x), "f"(p[3].x));
#pragma unroll
            for (int k = 0; k < 4; ++k) { acc.x += p[k].x; acc.y += p[k].y; acc.z += p[k].z; acc.w += p[k].w; }
        }
        for (; s < n_part; ++s) {
            const float4 p = *reinterpret_cast<const float4*>(part + s * part_stride + e);
            acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
        }
        if (gelu) {     // bf16 output: the 1.5e-7 erf approximation is far below its rounding; fp32 output keeps erff
            if constexpr (sizeof(TY) == 2) {
                const float2 g01 = gelu_erf_fast2(make_float2(acc.x, acc.y)), g23 = gelu_erf_fast2(make_float2(acc.z, acc.w));
                acc = make_float4(g01.x, g01.y, g23.x, g23.y);
            } else {
                acc.x = gelu_erf(acc.x); acc.y = gelu_erf(acc.y); acc.z = gelu_erf(acc.z); acc.w = gelu_erf(acc.w);
            }
        }
        if constexpr (sizeof(TY) == 4) {
            *reinterpret_cast<float4*>(y + e) = acc;
        } else {
            *reinterpret_cast<uint2*>(y + e) = make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
        }
    }
}

extern "C" int vb_reduce_bias_act(const float* part, int n_part, int64_t part_stride, const float* bias, int gelu, void* y,
                                  int y_dtype, int64_t R, int N, void* stream) {
    VB_REQUIRE(part && y && n_part >= 1 && N > 0 && N % 4 == 0 && R >= 0, VB_ERR_BAD_ARG, "vb_reduce_bias_act: bad args");
    if (R == 0) return VB_OK;
    const int64_t total4 = R * N / 4;
    const int threads = 256;
    const int64_t want_blocks = vb_ceil_div(total4, threads), cap_blocks = static_cast<int64_t>(vb_sm_count()) * 8;
    const int blocks = static_cast<int>(want_blocks < cap_blocks ? want_blocks : cap_blocks);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (y_dtype == VB_F32)
        VB_CUDA(vb_launch(R <= 1024, reduce_bias_act_kernel<float>, dim3(blocks), dim3(threads), 0, st, part, n_part, part_stride, bias, gelu, static_cast<float*>(y), total4, N));
    else if (y_dtype == VB_BF16)
        VB_CUDA(vb_launch(R <= 1024, reduce_bias_act_kernel<__nv_bfloat16>, dim3(blocks), dim3(threads), 0, st, part, n_part, part_stride, bias, gelu, static_cast<__nv_bfloat16*>(y), total4, N));
    else
        VB_REQUIRE(false, VB_ERR_BAD_ARG, "vb_reduce_bias_act: bad y_dtype");
    return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// K5  prefill K/V -> paged pool   pool: [page][2][H][64][Dh]
// ------------------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void kv_scatter_kernel(const TS* __restrict__ qkv, TD* __restrict__ pool, const int32_t* __restrict__ block_table,
                                  int max_pages, const int32_t* __restrict__ kv_lens, int S, int H, int Dh) {
    pdl_trigger();
    pdl_wait();
    const int s = blockIdx.x, b = blockIdx.y;
    if (s >= kv_lens[b]) return;
    const int page = block_table[b * max_pages + (s >> 6)];
    const int slot = s & 63;
    const int d = H * Dh;
    const TS* src = qkv + (static_cast<int64_t>(b) * S + s) * 3 * d;
    for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) {
        const int kv = c / d, hd = c % d, h = hd / Dh, e = hd % Dh;
        const float val = to_f32<TS>(src[d + c]);
        const int col = (sizeof(TD) == 2 && Dh == 64) ? vb_pool_col_bf16(slot, e) : e;    // bf16 pool rows are chunk-swizzled
        pool[(((static_cast<int64_t>(page) * 2 + kv) * H + h) * 64 + slot) * Dh + col] = from_f32<TD>(val);
    }
}

extern "C" int vb_kv_scatter_paged(const void* qkv, int qkv_dtype, void* pool, int pool_dtype, const int32_t* block_table,
                                   int max_pages, const int32_t* kv_lens, int B, int S, int H, int Dh, void* stream) {
    VB_REQUIRE(qkv && pool && block_table && kv_lens, VB_ERR_BAD_ARG, "vb_kv_scatter_paged: null pointer");
    VB_REQUIRE(B >= 0 && S >= 0 && H > 0 && Dh > 0 && B <= 65535, VB_ERR_BAD_ARG, "vb_kv_scatter_paged: bad shape");
    if (B == 0 || S == 0) return VB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid(S, B);
    const int threads = 256;
#define KVS(TS, TD) VB_CUDA(vb_launch(false, kv_scatter_kernel<TS, TD>, grid, dim3(threads), 0, st, static_cast<const TS*>(qkv), static_cast<TD*>(pool), block_table, max_pages, kv_lens, S, H, Dh))
    if (qkv_dtype == VB_F32 && pool_dtype == VB_F32) KVS(float, float);
    else if (qkv_dtype == VB_BF16 && pool_dtype == VB_BF16) KVS(__nv_bfloat16, __nv_bfloat16);
    else if (qkv_dtype == VB_F32 && pool_dtype == VB_BF16) KVS(float, __nv_bfloat16);
    else VB_REQUIRE(false, VB_ERR_UNSUPPORTED, "vb_kv_scatter_paged: dtype combination %d -> %d", qkv_dtype, pool_dtype);
#undef KVS
    return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// KV prefetch into L2.  The decode step alternates an HBM-bound kernel (paged attention: ~100 MB of KV per layer at
// B=32) with a latency-bound chain of small GEMM / row kernels that leaves HBM almost idle.  This kernel is launched on a
// side stream while that chain runs and asks the memory system to pull the NEXT layer's cached pages into L2
// (cp.async.bulk.prefetch.L2: no destination, fire and forget), so that the next attention kernel finds part of its
// pages on chip.  A page is one contiguous [2][H][64][Dh] block; it is requested in pieces of <= 64 KB.
// ------------------------------------------------------------------------------------------------
__global__ void kv_prefetch_l2_kernel(const uint8_t* __restrict__ pool, int64_t page_bytes, const int32_t* __restrict__ block_table,
                                      int max_pages, const int32_t* __restrict__ seq_lens, int page_lo_pct, int page_hi_pct) {
    const int b = blockIdx.y, p = blockIdx.x;
    const int pages_total = (seq_lens[b] + 63) / 64;
    const int lo = pages_total * page_lo_pct / 100, hi = (pages_total * page_hi_pct + 99) / 100;
    if (p < lo || p >= min(hi, pages_total)) return;
    const int page = block_table[static_cast<int64_t>(b) * max_pages + p];
    const int64_t piece = 65536;
    const int64_t off = static_cast<int64_t>(threadIdx.x) * piece;
    if (off >= page_bytes) return;
    const uint32_t bytes = static_cast<uint32_t>(min(piece, page_bytes - off));
    const uint8_t* src = pool + static_cast<int64_t>(page) * page_bytes + off;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

extern "C" int vb_kv_prefetch_l2(const void* pool, int pool_dtype, const int32_t* block_table, int max_pages,
                                 const int32_t* seq_lens, int B, int H, int Dh, int page_lo_pct, int page_hi_pct, void* stream) {
    VB_REQUIRE(pool && block_table && seq_lens, VB_ERR_BAD_ARG, "vb_kv_prefetch_l2: null pointer");
    VB_REQUIRE(B >= 1 && B <= 65535 && H >= 1 && Dh >= 1 && max_pages >= 1, VB_ERR_BAD_ARG, "vb_kv_prefetch_l2: bad shape");
    VB_REQUIRE(page_lo_pct >= 0 && page_hi_pct <= 100 && page_lo_pct <= page_hi_pct, VB_ERR_BAD_ARG, "vb_kv_prefetch_l2: bad page range");
    const int64_t page_bytes = static_cast<int64_t>(2) * H * 64 * Dh * (pool_dtype == VB_BF16 ? 2 : 4);
    VB_REQUIRE(page_bytes % 16 == 0 && page_bytes <= 32 * 65536, VB_ERR_UNSUPPORTED, "vb_kv_prefetch_l2: page of %lld bytes", (long long)page_bytes);
    if (page_lo_pct == page_hi_pct) return VB_OK;
    VB_CUDA(vb_launch(false, kv_prefetch_l2_kernel, dim3(max_pages, B), dim3(32), 0, static_cast<cudaStream_t>(stream),
                      static_cast<const uint8_t*>(pool), page_bytes, block_table, max_pages, seq_lens, page_lo_pct, page_hi_pct));
    return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// K11  beam bookkeeping (valle_ar.py:167-171) -- one CTA, no host sync
// ------------------------------------------------------------------------------------------------
__global__ void ar_bookkeeping_kernel(const int32_t* __restrict__ sample, const float* __restrict__ logprob,
                                      int32_t* __restrict__ last, float* __restrict__ sum_logprobs,
                                      int32_t* __restrict__ codes_out, int64_t codes_stride, int32_t* __restrict__ seq_lens,
                                      int32_t* __restrict__ audio_pos, int32_t* __restrict__ state, int B, int eos) {
    __shared__ int not_eos;
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x == 0) not_eos = 0;
    __syncthreads();
    const int step = state[0];
    int local = 0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const int prev = last[b];
        const bool done = (prev == eos);
        if (!done) sum_logprobs[b] += logprob[b];
        const int tok = done ? eos : sample[b];
        if (tok != eos) local = 1;
        if (step < codes_stride) codes_out[b * codes_stride + step] = tok;
        last[b] = tok;
        seq_lens[b] += 1;
        audio_pos[b] += 1;
    }
    if (local) atomicOr(&not_eos, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!not_eos && state[1] < 0) state[1] = step;
        state[0] = step + 1;
    }
}

extern "C" int vb_ar_bookkeeping(const int32_t* sample, const float* logprob, int32_t* last, float* sum_logprobs,
                                 int32_t* codes_out, int64_t codes_stride, int32_t* seq_lens, int32_t* audio_pos,
                                 int32_t* state, int B, int eos, void* stream) {
    VB_REQUIRE(sample && logprob && last && sum_logprobs && codes_out && seq_lens && audio_pos && state, VB_ERR_BAD_ARG,
               "vb_ar_bookkeeping: null pointer");
    VB_REQUIRE(B > 0, VB_ERR_BAD_ARG, "vb_ar_bookkeeping: B must be > 0");
    VB_CUDA(vb_launch(true, ar_bookkeeping_kernel, dim3(1), dim3(256), 0, static_cast<cudaStream_t>(stream), sample, logprob, last,
                      sum_logprobs, codes_out, codes_stride, seq_lens, audio_pos, state, B, eos));
    return VB_OK;
}
