// Micro-probe: issue rate of tcgen05.mma kind::f16 for the shapes the attention kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I valle2_b200/csrc tools/umma_probe.cu -o /tmp/umma_probe
// One elected thread per CTA issues R MMAs (M=128, N, K=16; operands = whatever is in shared memory), commits, waits, and
// reports clock64 cycles per MMA.  Variants: chained into one accumulator / alternating between two; A from shared memory or
// from TMEM; one or two CTAs per SM (grid 148 / 296).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace {

__device__ __forceinline__ void umma_f16_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// mode 0: A smem, one accumulator; 1: A smem, two accumulators alternating every 4 MMAs; 2: A from TMEM, one accumulator;
// 3: attention pattern: 4 x (A smem, N) into D0 then 4 x (A TMEM, N) into D1; 4: B MN-major (the V operand), one accumulator
// out[cta] = {cycles until the last MMA was ISSUED, cycles until all had COMPLETED}
template <int N, int MODE, int M>
__global__ void __launch_bounds__(64) probe_kernel(int R, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = base, b_smem = base + 16384;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        fence_mbar_init();
    }
    if (threadIdx.x >> 5 == 1) tmem_alloc<512>(smem_u32(&tmem_slot));
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t t0 = tmem_slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t IDESC = umma_idesc_bf16(M, N, 0, MODE == 4 ? 1 : 0);
        uint32_t phase = 0;
        long long best = 1ll << 60, best_issue = 0;
        const uint64_t da0 = umma_desc_sw128(a_smem, 16, 1024);
        const uint64_t db0 = MODE == 4 ? umma_desc_sw128(b_smem, 1024, 1024) : umma_desc_sw128(b_smem, 16, 1024);
        for (int rep = 0; rep < 5; ++rep) {
            const long long c0 = clock64();
            for (int i = 0; i < R; i += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int kk = u & 3;
                    const uint64_t da = da0 + kk * 2;                                   // +32 bytes along K
                    const uint64_t db = MODE == 4 ? db0 + kk * 128 : db0 + kk * 2;      // MN-major: +16 rows of 128 bytes
                    if (MODE == 0 || MODE == 4) umma_f16(t0, da, db, IDESC, 1);
                    else if (MODE == 1) umma_f16(t0 + ((u >> 2) & 1) * 256, da, db, IDESC, 1);
                    else if (MODE == 2) umma_f16_ta(t0, t0 + 256 + kk * 8, db, IDESC, 1);
                    else {
                        if ((u >> 2) & 1) umma_f16_ta(t0 + 256, t0 + 384 + kk * 8, db, IDESC, 1);
                        else umma_f16(t0, da, db, IDESC, 1);
                    }
                }
            }
            const long long ci = clock64();
            umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), phase);
            phase ^= 1;
            const long long c1 = clock64();
            if (c1 - c0 < best) { best = c1 - c0; best_issue = ci - c0; }
        }
        out[blockIdx.x * 2] = best_issue;
        out[blockIdx.x * 2 + 1] = best;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x >> 5 == 1) {
        tc_fence_after();
        tmem_dealloc<512>(t0);
    }
}

// TMEM read bandwidth: `active` warps (one per lane quadrant / SM sub-partition) each issue R x tcgen05.ld 32x32b.x32 (4 KB)
__global__ void __launch_bounds__(128) ldtm_probe_kernel(int active, int R, long long* out, uint32_t* sink) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t t0 = tmem_slot + (static_cast<uint32_t>(warp * 32) << 16);
    uint32_t acc = 0;
    long long best = 1ll << 60;
    if (warp < active) {
        for (int rep = 0; rep < 5; ++rep) {
            const long long c0 = clock64();
            for (int i = 0; i < R; ++i) {
                uint32_t v[32];
                tmem_ld_32x32(t0 + (i & 7) * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 32; ++k) acc ^= v[k];
            }
            const long long c1 = clock64();
            if (c1 - c0 < best) best = c1 - c0;
        }
        if ((threadIdx.x & 31) == 0) out[blockIdx.x * 4 + warp] = best;
        if (acc == 0x12345678u) sink[0] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_slot);
    }
}

// MUFU.EX2 rate: `warps` warps per CTA (one CTA per SM), 8 independent chains of ex2 per thread
__global__ void mufu_probe_kernel(int R, long long* out, float* sink) {
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = -0.001f * (threadIdx.x + k);
    __syncthreads();
    const long long c0 = clock64();
    for (int i = 0; i < R; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
    }
    const long long c1 = clock64();
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a += x[k];
    if (a == 123.f) sink[0] = a;
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 32 + (threadIdx.x >> 5)] = c1 - c0;
}
// packed ex2: two exponentials per instruction (bf16x2 / f16x2)
template <int KIND>
__global__ void mufu2_probe_kernel(int R, long long* out, uint32_t* sink) {
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (KIND == 0 ? 0xbc00bc00u : 0xa000a000u) + threadIdx.x + k;
    __syncthreads();
    const long long c0 = clock64();
    for (int i = 0; i < R; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (KIND == 0) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[k]));
            else asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[k]));
        }
    }
    const long long c1 = clock64();
    uint32_t a = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) a ^= x[k];
    if (a == 0x12345u) sink[0] = a;
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 32 + (threadIdx.x >> 5)] = c1 - c0;
}
template <int KIND> void run_mufu2(int warps, long long* d_out, uint32_t* d_sink) {
    const int R = 512;
    mufu2_probe_kernel<KIND><<<148, warps * 32>>>(R, d_out, d_sink);
    cudaDeviceSynchronize();
    std::vector<long long> h(148 * 32);
    cudaMemcpy(h.data(), d_out, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) for (int w = 0; w < warps; ++w) mx = h[i * 32 + w] > mx ? h[i * 32 + w] : mx;
    printf("ex2 %s: %2d warps per SM: %.2f cycles per warp instruction per warp -> %.1f exponentials/cycle/SM\n", KIND == 0 ? "bf16x2" : "f16x2 ",
           warps, static_cast<double>(mx) / (R * 8), 2 * warps * 32.0 * R * 8 / static_cast<double>(mx));
}

void run_mufu(int warps, long long* d_out, float* d_sink) {
    const int R = 512;
    mufu_probe_kernel<<<148, warps * 32>>>(R, d_out, d_sink);
    cudaDeviceSynchronize();
    std::vector<long long> h(148 * 32);
    cudaMemcpy(h.data(), d_out, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) for (int w = 0; w < warps; ++w) mx = h[i * 32 + w] > mx ? h[i * 32 + w] : mx;
    printf("MUFU.EX2: %2d warps per SM: %.2f cycles per warp instruction per warp -> %.1f lanes/cycle/SM\n", warps,
           static_cast<double>(mx) / (R * 8), warps * 32.0 * R * 8 / static_cast<double>(mx));
}

void run_ldtm(int active, long long* d_out, uint32_t* d_sink) {
    const int R = 256;
    ldtm_probe_kernel<<<148, 128>>>(active, R, d_out, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ldtm probe: %s\n", cudaGetErrorString(e)); exit(1); }
    std::vector<long long> h(148 * 4);
    cudaMemcpy(h.data(), d_out, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) for (int w = 0; w < active; ++w) mx = h[i * 4 + w] > mx ? h[i * 4 + w] : mx;
    printf("tcgen05.ld 32x32b.x32 (4 KB per warp), %d warp(s) per SM: %.1f cycles per load (dependent: load, wait, consume) -> %.0f B/cycle/SM\n",
           active, static_cast<double>(mx) / R, active * 4096.0 * R / static_cast<double>(mx));
}

template <int N, int MODE, int M = 128> void run(int grid, int R, long long* d_out) {
    const int smem = 16384 + 32768 + 1024;
    cudaFuncSetAttribute(probe_kernel<N, MODE, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<N, MODE, M><<<grid, 64, smem>>>(R, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d mode=%d: %s\n", N, MODE, cudaGetErrorString(e)); exit(1); }
    std::vector<long long> h(grid * 2);
    cudaMemcpy(h.data(), d_out, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1ll << 60, is = 0;
    for (int i = 0; i < grid; ++i) { const long long v = h[i * 2 + 1]; mx = v > mx ? v : mx; if (v < mn) { mn = v; is = h[i * 2]; } }
    const char* names[] = {"A smem, one accumulator", "A smem, two accumulators", "A tmem, one accumulator", "4 x smem-A + 4 x tmem-A", "A smem, B MN-major"};
    printf("M=%3d N=%3d  grid %3d  %-26s  cycles per MMA: issued %.1f  completed %.1f (slowest CTA %.1f)   math at peak: %d\n", M, N, grid,
           names[MODE], static_cast<double>(is) / R, static_cast<double>(mn) / R, static_cast<double>(mx) / R, N * M / 256);
}

template <int MODE> void run_all(int grid, int R, long long* d_out) {
    run<32, MODE>(grid, R, d_out);
    run<64, MODE>(grid, R, d_out);
    run<128, MODE>(grid, R, d_out);
    run<256, MODE>(grid, R, d_out);
}

}  // namespace

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 148 * 32 * sizeof(long long));
    uint32_t* d_sink;
    cudaMalloc(&d_sink, 64);
    for (int active : {1, 2, 4}) run_ldtm(active, d_out, d_sink);
    for (int warps : {1, 4, 8, 16}) run_mufu(warps, d_out, reinterpret_cast<float*>(d_sink));
    for (int warps : {4, 8}) { run_mufu2<0>(warps, d_out, d_sink); run_mufu2<1>(warps, d_out, d_sink); }
    const int R = 512;
    for (int grid : {148}) {
        run_all<0>(grid, R, d_out);
        run_all<1>(grid, R, d_out);
        run_all<2>(grid, R, d_out);
        run_all<3>(grid, R, d_out);
        run_all<4>(grid, R, d_out);
        run<64, 0, 64>(grid, R, d_out);
        run<128, 0, 64>(grid, R, d_out);
        run<256, 0, 64>(grid, R, d_out);
    }
    return 0;
}
