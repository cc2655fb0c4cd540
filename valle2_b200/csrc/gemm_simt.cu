// fp32 SIMT GEMM for the "fp32 validation mode" (logits within 1e-5 of the CPU oracle, greedy tokens bit-exact).
// y[M,N] = epilogue(x[M,K] . w[N,K]^T); both operands K-contiguous (nn.Linear layout).  64x64 tile, BK=16,
// 256 threads, 4x4 micro-tile, plain fp32 FMA accumulation in increasing-k order (deterministic).
// This is the correctness path; the throughput path is gemm_tc.cu (tcgen05).
#include "common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename TY>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                        int64_t ldw, const float* __restrict__ bias,
                                                        const float* __restrict__ residual, int64_t ldr, TY* __restrict__ y,
                                                        int64_t ldy, int M, int N, int K, int epilogue) {
    __shared__ float xs[TK][TM + 4];
    __shared__ float ws[TK][TN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 4 (m) x 4 (n)
    float acc[4][4] = {};
    // loader mapping: 64 rows x 16 k = 1024 floats per operand, 256 threads x float4 along k
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    for (int k0 = 0; k0 < K; k0 += TK) {
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), wv = make_float4(0.f, 0.f, 0.f, 0.f);
        const int gm = m0 + lr, gn = n0 + lr, gk = k0 + lk;
        if (gm < M) {
            if (gk + 3 < K) xv = *reinterpret_cast<const float4*>(x + gm * ldx + gk);
            else {
                float t[4] = {0.f, 0.f, 0.f, 0.f};
                for (int i = 0; i < 4; ++i) if (gk + i < K) t[i] = x[gm * ldx + gk + i];
                xv = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
        if (gn < N) {
            if (gk + 3 < K) wv = *reinterpret_cast<const float4*>(w + gn * ldw + gk);
            else {
                float t[4] = {0.f, 0.f, 0.f, 0.f};
                for (int i = 0; i < 4; ++i) if (gk + i < K) t[i] = w[gn * ldw + gk + i];
                wv = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
        xs[lk + 0][lr] = xv.x; xs[lk + 1][lr] = xv.y; xs[lk + 2][lr] = xv.z; xs[lk + 3][lr] = xv.w;
        ws[lk + 0][lr] = wv.x; ws[lk + 1][lr] = wv.y; ws[lk + 2][lr] = wv.z; ws[lk + 3][lr] = wv.w;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&xs[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&ws[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (epilogue != VB_EPI_NONE) v += bias[n];
            if (epilogue == VB_EPI_BIAS_GELU) v = gelu_erf(v);
            if (epilogue == VB_EPI_BIAS_RESIDUAL) v += residual[m * ldr + n];
            y[m * ldy + n] = from_f32<TY>(v);
        }
    }
}

}  // namespace

int vb_linear_simt(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, const float* residual,
                   int64_t ldr, void* y, int y_dtype, int64_t ldy, int64_t M, int64_t N, int64_t K, int epilogue,
                   cudaStream_t st) {
    VB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 && ldx % 4 == 0 &&
                   ldw % 4 == 0,
               VB_ERR_BAD_ARG, "vb_linear(fp32): operands must be 16-byte aligned with pitch %% 4 == 0");
    dim3 grid(static_cast<unsigned>(vb_ceil_div(N, TN)), static_cast<unsigned>(vb_ceil_div(M, TM)));
    VB_REQUIRE(grid.y <= 65535, VB_ERR_UNSUPPORTED, "vb_linear(fp32): M too large for this kernel (%lld)", (long long)M);
    if (y_dtype == VB_F32)
        gemm_simt_kernel<float><<<grid, 256, 0, st>>>(x, ldx, w, ldw, bias, residual, ldr, static_cast<float*>(y), ldy, (int)M, (int)N, (int)K, epilogue);
    else
        gemm_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, ldx, w, ldw, bias, residual, ldr, static_cast<__nv_bfloat16*>(y), ldy, (int)M, (int)N, (int)K, epilogue);
    VB_LAUNCH_CHECK();
    return VB_OK;
}
