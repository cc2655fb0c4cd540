// vb_linear: dtype dispatch between the fp32 SIMT GEMM (validation mode) and the tcgen05 bf16 GEMM.
#include "common.cuh"

int vb_linear_simt(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, const float* residual,
                   int64_t ldr, void* y, int y_dtype, int64_t ldy, int64_t M, int64_t N, int64_t K, int epilogue,
                   cudaStream_t st);
int vb_linear_tc(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, const float* residual,
                 int64_t ldr, void* y, int y_dtype, int64_t ldy, int64_t M, int64_t N, int64_t K, int epilogue,
                 cudaStream_t st, float inv_temp = 1.f, int rng_on = 0, unsigned long long rng_key = 0ull);

int vb_linear_tc_t(const void* x, int64_t ldx, int x_mn, const void* w, int64_t ldw, int w_mn, const float* bias,
                   const float* residual, int64_t ldr, void* y, int y_dtype, int64_t ldy, int64_t M, int64_t N, int64_t K,
                   int epilogue, cudaStream_t st);

extern "C" int vb_linear_t(const void* x, int64_t ldx, int x_transposed, const void* w, int64_t ldw, int w_transposed,
                           const float* bias, const float* residual, int64_t ldr, void* y, int y_dtype, int64_t ldy, int64_t M,
                           int64_t N, int64_t K, int epilogue, void* stream) {
    VB_REQUIRE(x && w && y, VB_ERR_BAD_ARG, "vb_linear_t: null pointer");
    VB_REQUIRE(M >= 0 && N >= 1 && K >= 1, VB_ERR_BAD_ARG, "vb_linear_t: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    VB_REQUIRE(epilogue >= VB_EPI_NONE && epilogue <= VB_EPI_BIAS_RESIDUAL, VB_ERR_BAD_ARG, "vb_linear_t: bad epilogue %d", epilogue);
    VB_REQUIRE(epilogue == VB_EPI_NONE || bias != nullptr, VB_ERR_BAD_ARG, "vb_linear_t: epilogue %d needs a bias", epilogue);
    VB_REQUIRE(epilogue != VB_EPI_BIAS_RESIDUAL || residual != nullptr, VB_ERR_BAD_ARG, "vb_linear_t: residual is null");
    VB_REQUIRE(y_dtype == VB_F32 || y_dtype == VB_BF16, VB_ERR_BAD_ARG, "vb_linear_t: bad y_dtype");
    VB_REQUIRE((x_transposed || w_transposed) ? N > 128 : true, VB_ERR_UNSUPPORTED, "vb_linear_t: transposed operands need N > 128 (got %lld)", (long long)N);
    if (M == 0) return VB_OK;
    return vb_linear_tc_t(x, ldx, x_transposed, w, ldw, w_transposed, bias, residual, ldr, y, y_dtype, ldy, M, N, K, epilogue,
                          static_cast<cudaStream_t>(stream));
}

extern "C" int vb_linear(const void* x, int x_dtype, int64_t ldx, const void* w, int w_dtype, int64_t ldw, const float* bias,
                         const float* residual, int64_t ldr, void* y, int y_dtype, int64_t ldy, int64_t M, int64_t N,
                         int64_t K, int epilogue, void* stream) {
    VB_REQUIRE(x && w && y, VB_ERR_BAD_ARG, "vb_linear: null pointer");
    VB_REQUIRE(M >= 0 && N >= 1 && K >= 1, VB_ERR_BAD_ARG, "vb_linear: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    VB_REQUIRE(epilogue >= VB_EPI_NONE && epilogue <= VB_EPI_BIAS_RESIDUAL, VB_ERR_BAD_ARG, "vb_linear: bad epilogue %d", epilogue);
    VB_REQUIRE(epilogue == VB_EPI_NONE || bias != nullptr, VB_ERR_BAD_ARG, "vb_linear: epilogue %d needs a bias", epilogue);
    VB_REQUIRE(epilogue != VB_EPI_BIAS_RESIDUAL || residual != nullptr, VB_ERR_BAD_ARG, "vb_linear: residual is null");
    VB_REQUIRE(y_dtype == VB_F32 || y_dtype == VB_BF16, VB_ERR_BAD_ARG, "vb_linear: bad y_dtype");
    if (M == 0) return VB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (x_dtype == VB_F32 && w_dtype == VB_F32)
        return vb_linear_simt(static_cast<const float*>(x), ldx, static_cast<const float*>(w), ldw, bias, residual, ldr, y,
                              y_dtype, ldy, M, N, K, epilogue, st);
    if (x_dtype == VB_BF16 && w_dtype == VB_BF16)
        return vb_linear_tc(x, ldx, w, ldw, bias, residual, ldr, y, y_dtype, ldy, M, N, K, epilogue, st);
    VB_REQUIRE(false, VB_ERR_UNSUPPORTED, "vb_linear: dtype combination x=%d w=%d", x_dtype, w_dtype);
}

namespace {
// keys[m] = order-preserving value bits << 32 | ~column (gemm_tc.cu, VB_EPI_ARGMAX) -> token; the key is reset for the next call
__global__ void argmax_unpack_kernel(unsigned long long* __restrict__ keys, int32_t* __restrict__ out_tok, int64_t rows_per_batch,
                                     int64_t batch_stride, int64_t row_stride, int64_t M) {
    const int64_t m = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const unsigned long long key = keys[m];
    keys[m] = 0ull;
    const int32_t tok = key ? static_cast<int32_t>(0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull)) : 0;
    out_tok[(m / rows_per_batch) * batch_stride + (m % rows_per_batch) * row_stride] = tok;
}
}  // namespace

static int linear_pick(const void* x, int64_t ldx, const void* w, int64_t ldw, unsigned long long* keys, int32_t* out_tok,
                       int64_t rows_per_batch, int64_t batch_stride, int64_t row_stride, int64_t M, int64_t N, int64_t K,
                       float inv_temp, int rng_on, unsigned long long rng_key, void* stream) {
    VB_REQUIRE(x && w && keys && out_tok, VB_ERR_BAD_ARG, "vb_linear_argmax / vb_linear_categorical: null pointer");
    VB_REQUIRE(M >= 0 && N >= 1 && K >= 1 && rows_per_batch >= 1, VB_ERR_BAD_ARG, "vb_linear_argmax / vb_linear_categorical: bad shape M=%lld N=%lld K=%lld rows_per_batch=%lld",
               (long long)M, (long long)N, (long long)K, (long long)rows_per_batch);
    if (M == 0) return VB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rc = vb_linear_tc(x, ldx, w, ldw, nullptr, nullptr, 0, keys, VB_F32, N, M, N, K, VB_EPI_ARGMAX, st, inv_temp, rng_on, rng_key);
    if (rc != VB_OK) return rc;
    argmax_unpack_kernel<<<static_cast<unsigned>((M + 255) / 256), 256, 0, st>>>(keys, out_tok, rows_per_batch, batch_stride, row_stride, M);
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_linear_argmax(const void* x, int64_t ldx, const void* w, int64_t ldw, unsigned long long* keys, int32_t* out_tok,
                                int64_t rows_per_batch, int64_t batch_stride, int64_t row_stride, int64_t M, int64_t N, int64_t K,
                                void* stream) {
    return linear_pick(x, ldx, w, ldw, keys, out_tok, rows_per_batch, batch_stride, row_stride, M, N, K, 1.f, 0, 0ull, stream);
}

extern "C" int vb_linear_categorical(const void* x, int64_t ldx, const void* w, int64_t ldw, unsigned long long* keys, int32_t* out_tok,
                                     int64_t rows_per_batch, int64_t batch_stride, int64_t row_stride, int64_t M, int64_t N, int64_t K,
                                     float temperature, uint64_t seed, int step, void* stream) {
    VB_REQUIRE(temperature > 0.f, VB_ERR_BAD_ARG, "vb_linear_categorical: temperature must be positive (got %f)", (double)temperature);
    const unsigned long long key = static_cast<unsigned long long>(seed) * 0xD1342543DE82EF95ull +
                                   static_cast<unsigned long long>(static_cast<unsigned>(step)) * 0xA0761D6478BD642Full + 0x2545F4914F6CDD1Dull;
    return linear_pick(x, ldx, w, ldw, keys, out_tok, rows_per_batch, batch_stride, row_stride, M, N, K, 1.f / temperature, 1, key, stream);
}
