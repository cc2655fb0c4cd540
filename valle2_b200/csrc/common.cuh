// Shared helpers for libvalle_b200.so: error reporting, dtype conversion, sm_100a PTX wrappers
// (mbarrier, TMA, tcgen05/TMEM).  Everything here is device- or host-inline; no global state
// except the thread-local error string defined in api.cu.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/valle_b200.h"

// ------------------------------------------------------------------------------------------------
// host: errors
// ------------------------------------------------------------------------------------------------
void vb_set_error(const char* fmt, ...);

#define VB_REQUIRE(cond, code, ...)                 \
    do {                                            \
        if (!(cond)) {                              \
            vb_set_error(__VA_ARGS__);              \
            return (code);                          \
        }                                           \
    } while (0)

#define VB_CUDA(expr)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            vb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                         __LINE__);                                                          \
            return VB_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

#define VB_LAUNCH_CHECK() VB_CUDA(cudaGetLastError())

static inline int64_t vb_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int vb_sm_count();   // cached cudaDevAttrMultiProcessorCount of the current device
bool vb_pdl_enabled();   // VALLE_B200_PDL != 0 (default on)
// Experiment (VALLE_B200_CARVEOUT=1, off by default): ask for the maximum shared-memory carveout on the first launch of
// every PDL kernel, on the theory that kernels of one dependent chain can only be co-resident when they agree on the
// L1 / shared split.  Measured: no gain in the full step (243.6 vs 245.1 us at B=1) and the GEMM chain alone gets slower
// (135 -> 181 us) with the minimal L1 -- tools/step_breakdown.py.
void vb_prefer_max_carveout(const void* kern);

// Kernel launch with the programmatic-dependent-launch attribute (captured as a programmatic edge in CUDA graphs).
template <typename... Exp, typename... Act>
static inline cudaError_t vb_launch(bool pdl, void (*kern)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Act&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && vb_pdl_enabled()) ? 1 : 0;
    if (pdl) vb_prefer_max_carveout(reinterpret_cast<const void*>(kern));
    return cudaLaunchKernelEx(&cfg, kern, static_cast<Exp>(args)...);
}

// ------------------------------------------------------------------------------------------------
// device: scalar conversion
// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// exact (erf) GELU, nn.GELU() default -- reference modules.py:216
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// erf-GELU with erf from Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7, far below the bf16 output rounding); two MUFU
// ops + ~12 FMA-pipe ops instead of the ~30-instruction erff().  The fp32 validation path keeps erff (gemm_simt.cu).
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
    const float erf_abs = fmaf(-poly * t, e, 1.0f);
    const float erf_x = copysignf(erf_abs, x);
    return 0.5f * x * (1.0f + erf_x);
}

// the same erf-GELU on two values with packed fp32x2 arithmetic (FFMA2 / FMUL2): ~10 issue slots per element instead
// of ~20 -- the FFN1 epilogue of the large-M GEMM is issue-bound (8 epilogue warps finish 128 x 256 values per tile)
__device__ __forceinline__ float2 gelu_erf_fast2(float2 x) {
    const float2 z = __fmul2_rn(make_float2(fabsf(x.x), fabsf(x.y)), make_float2(0.70710678118654752440f, 0.70710678118654752440f));
    const float2 den = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, make_float2(1.0f, 1.0f));
    const float2 t = make_float2(__fdividef(1.0f, den.x), __fdividef(1.0f, den.y));
    float2 poly = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
    poly = __ffma2_rn(poly, t, make_float2(1.421413741f, 1.421413741f));
    poly = __ffma2_rn(poly, t, make_float2(-0.284496736f, -0.284496736f));
    poly = __ffma2_rn(poly, t, make_float2(0.254829592f, 0.254829592f));
    const float2 arg = __fmul2_rn(__fmul2_rn(z, z), make_float2(-1.4426950408889634f, -1.4426950408889634f));
    float2 e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(arg.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(arg.y));
    const float2 q = __fmul2_rn(__fmul2_rn(poly, t), e);
    const float2 erf_abs = __ffma2_rn(q, make_float2(-1.0f, -1.0f), make_float2(1.0f, 1.0f));
    const float2 erf_x = make_float2(copysignf(erf_abs.x, x.x), copysignf(erf_abs.y, x.y));
    const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
    return __ffma2_rn(h, erf_x, h);
}

// erf-GELU on two values WITHOUT special-function ops: erf(x / sqrt 2) = t * P(t^2), t = min(|x|, 4.4), P of degree 9 in t^2
// (minimax fit, |erf error| <= 5.3e-6 on [0, 4.4]; beyond 4.4 the clamp leaves 1 - erf = 1.1e-5), sign restored afterwards.
// GELU error in fp32 Horner arithmetic <= 5.6e-5 absolute -- two orders below the bf16 rounding of the output.  17 issue slots
// per pair on the FMA / ALU pipes instead of 22 + 4 MUFU: in the FFN1 epilogue of the large-M GEMM the four MUFU ops per pair
// alone cost 4 200 XU-pipe cycles per 128 x 256 tile, half of the tile's main loop.
__device__ __forceinline__ float2 gelu_erf_poly2(float2 x) {
    const float2 t = make_float2(fminf(fabsf(x.x), 4.4f), fminf(fabsf(x.y), 4.4f));
    const float2 u = __fmul2_rn(t, t);
    float2 p = __ffma2_rn(make_float2(-4.048130320538634e-12f, -4.048130320538634e-12f), u, make_float2(4.569462586090367e-10f, 4.569462586090367e-10f));
    p = __ffma2_rn(p, u, make_float2(-2.30715997417974e-08f, -2.30715997417974e-08f));
    p = __ffma2_rn(p, u, make_float2(6.947192900952359e-07f, 6.947192900952359e-07f));
    p = __ffma2_rn(p, u, make_float2(-1.4089750038692728e-05f, -1.4089750038692728e-05f));
    p = __ffma2_rn(p, u, make_float2(0.00020676301210187376f, 0.00020676301210187376f));
    p = __ffma2_rn(p, u, make_float2(-0.0022991118021309376f, -0.0022991118021309376f));
    p = __ffma2_rn(p, u, make_float2(0.019814016297459602f, 0.019814016297459602f));
    p = __ffma2_rn(p, u, make_float2(-0.1328718215227127f, -0.1328718215227127f));
    p = __ffma2_rn(p, u, make_float2(0.7978581190109253f, 0.7978581190109253f));
    const float2 e = __fmul2_rn(t, p);
    const float2 erf_x = make_float2(copysignf(e.x, x.x), copysignf(e.y, x.y));
    const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
    return __ffma2_rn(h, erf_x, h);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// KV pool, bf16, head_dim 64: the eight 16-byte chunks of a token row are stored XOR-swizzled by (token & 7), so that a
// (page, K|V, head) chunk copied linearly into shared memory can be read with ldmatrix without bank conflicts (eight
// consecutive token rows hit eight different 16-byte bank groups).  Element (slot, e) of a chunk lives at column:
__host__ __device__ __forceinline__ int vb_pool_col_bf16(int slot, int e) { return ((((e >> 3) ^ (slot & 7)) << 3) | (e & 7)); }

// ------------------------------------------------------------------------------------------------
// device: sm_100a PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost arrive becomes a device trap (CUDA error) after a few seconds instead of a hung GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 22)) __trap();
    }
}
// Same, with a short sleep between polls: for single-lane producer / MMA-issuer roles whose polling would otherwise
// steal issue slots from the math warps that share their scheduler (visible as "branch resolving" stalls in ncu).
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(32);
        if (++spins > (1u << 22)) __trap();
    }
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-serialization attribute may start
// while its predecessor in the stream is still running.  pdl_trigger() lets OUR successor start early; pdl_wait()
// blocks until the predecessor grid has completed and its writes are visible.  Before pdl_wait() a kernel may only
// touch immutable data (weights, tensor maps) and its own shared memory / TMEM.
// -DVB_PDL_DEPTH1 (experiment, tools/build_variant.sh): every kernel releases its successor only AFTER its own wait, so
// at most two kernels of a chain are resident at a time (the default lets a whole stack of future kernels sit on the SMs
// spinning in griddepcontrol.wait, which starves a second, independent chain of registers and shared memory).
#ifdef VB_PDL_DEPTH1
__device__ __forceinline__ void pdl_trigger() {}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n griddepcontrol.launch_dependents;" ::: "memory"); }
#else
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// TMA: 2D/3D tiled load global -> shared, completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// TMA store shared -> global (bulk-group completion): the box of `map` at coordinates (c0, c1, c2) is written from dense smem rows
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest N bulk groups of this thread have finished READING their shared-memory source
template <int N> __device__ __forceinline__ void bulk_wait_group_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// 1D bulk copy global -> shared (no tensor map), bytes % 16 == 0, 16 B aligned both sides
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// tcgen05 / TMEM
template <int kCols> __device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "n"(kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int kCols> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], kind::f16 (bf16/fp16 inputs, fp32 accumulate), single CTA
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 columns of fp32: thread t of the warp gets lane (quadrant*32 + t), 32 consecutive columns
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (sm_100): 128-byte swizzle, version 1.
//   bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
// K-major operand tile [rows][64 bf16] written by TMA SWIZZLE_128B: 8-row groups are 1024 B apart (SBO), LBO unused.
// MN-major operand tile [k rows][64 bf16 along MN]: 8-k groups are 1024 B apart (SBO), LBO = pitch between 64-wide MN atoms.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

// UMMA instruction descriptor, kind::f16, bf16 x bf16 -> fp32.
//   [4,6) c_format=1 (f32) | [7,10) a_format=1 (bf16) | [10,13) b_format=1 | bit15 a_major | bit16 b_major (0 = K-major)
//   [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
           (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// host: TMA tensor maps (driver entry point resolved at run time; the library does not link libcuda)
// ------------------------------------------------------------------------------------------------
// 2D bf16 row-major [rows][cols] with pitch ld (elements); box = [box_rows][64 cols], SWIZZLE_128B, zero OOB fill.
int vb_make_tmap_bf16_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols);
// 3D bf16 [d2][d1][d0] (d0 contiguous), strides in elements; box = [1][box1][box0]
int vb_make_tmap_bf16_3d(CUtensorMap* map, const void* base, int64_t d0, int64_t d1, int64_t d2, int64_t s1, int64_t s2,
                         int box0, int box1);
// 3D fp32 [d2][d1][d0], strides in elements; box = [1][box1][box0], no swizzle (TMA store destination)
int vb_make_tmap_f32_3d(CUtensorMap* map, const void* base, int64_t d0, int64_t d1, int64_t d2, int64_t s1, int64_t s2, int box0,
                        int box1);
