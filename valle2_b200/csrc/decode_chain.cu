// Persistent "chain" kernel for the decode step: everything between two attention kernels of the AR decoder --
//   out-proj GEMM -> [split-K reduce + bias + residual + LayerNorm] -> FFN1 GEMM -> [reduce + bias + erf-GELU] ->
//   FFN2 GEMM -> [reduce + bias + residual + LayerNorm of the next layer] -> QKV GEMM of the next layer
// -- runs as ONE launch of #SM co-resident CTAs that step through the phases with a grid barrier between them, instead
// of seven dependent kernel launches.  Each hop then costs a ~1 us barrier instead of a kernel tear-down + start-up,
// and the weight stream never stops: warp 0 keeps requesting weight tiles (TMA) for the coming GEMM phases while the
// other warps are still in an earlier phase or waiting at a barrier (the ring of 8 stages decouples them).
//
//   warp 0      weight producer (one elected lane): TMA loads of W tiles for every GEMM phase, in phase order
//   warp 1      MMA issuer (tcgen05.mma, swap-AB: weight rows on MMA-M, batch on MMA-N) + TMEM allocator
//   warp 2-3    activation loaders: x[B][64] bf16 k-blocks -> 128B-swizzled smem (generic loads: the producer of the
//               activations is another CTA of this same kernel, one grid barrier earlier)
//   warp 4-7    epilogue: TMEM -> fp32 split-K slices [split][B][N]  (same slices, same order as vb_linear_decode)
//   warps 1-7   row phases (LayerNorm / GELU / cast) and the grid barrier
// Numerics are bit-identical to the multi-kernel path (same split sizes, same reduction order).
#include <string.h>

#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int STAGES = 8;
constexpr int THREADS = 256;
constexpr int WORKERS = THREADS - 32;   // warps 1..7 take part in row phases and barriers
constexpr int A_BYTES = BM * BK * 2;

enum { PH_GEMM = 0, PH_LN = 1, PH_ACT = 2 };

struct Phase {
    int type;
    // GEMM: part[s][m][n] = x[m, kslice] . w[n, kslice]
    int map;                 // index into the weight tensor maps
    const void* x;           // bf16 [B][K]
    int N, K, n_split, kb_per_split, tiles_a;
    float* part;
    int64_t part_stride;
    // LN  : x32 += bias + sum part (n_part slices); y = LN(x32) as bf16 (gamma == null: plain cast; y == null: none)
    // ACT : y = gelu(sum part + bias) as bf16, rows of width N
    float* x32;
    const float* in_part;
    int n_part;
    int64_t in_stride;
    const float* bias;
    const float* gamma;
    const float* beta;
    void* y;
    int d;
    float eps;
};

struct ChainParams {
    int n_phase;
    int B;
    unsigned* gbar;
    unsigned long long* dbg;   // optional: [cta][phase][4] globaltimer stamps (vb_decode_chain_set_debug)
    Phase ph[8];
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct Maps { CUtensorMap m[4]; };

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory"); }

// grid-wide barrier over the worker warps of all CTAs (all CTAs are co-resident: grid <= #SMs, 1 CTA/SM)
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target) {
    __threadfence();
    worker_sync();
    if (threadIdx.x == 32) {
        atomicAdd(ctr, 1u);
        unsigned spins = 0;
        while (ld_acquire(ctr) < target) {
            if (++spins > (1u << 24)) __trap();
        }
        __threadfence();
    }
    worker_sync();
}

__device__ __forceinline__ float worker_sum(float v, float* red) {   // sum over the 7 worker warps
    v = warp_sum(v);
    const int w = (threadIdx.x >> 5) - 1;
    if ((threadIdx.x & 31) == 0) red[w] = v;
    worker_sync();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < WORKERS / 32; ++i) t += red[i];
    worker_sync();
    return t;
}

template <int BN>
__global__ void __launch_bounds__(THREADS, 1) decode_chain_kernel(const __grid_constant__ Maps maps, const ChainParams p) {
    constexpr int B_BYTES = BN * BK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : 256;
    constexpr uint32_t IDESC = umma_idesc_bf16(BM, BN, 0, 0);
    constexpr int LOADERS = 64;   // warps 2-3

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float red[8];

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.m[i]);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1 + 2);     // weight producer (expect_tx) + two loader warps
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&tfull_bar[a]), 1);
            mbar_init(smem_u32(&tempty_bar[a]), 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(&tmem_base_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ------------- weight producer: runs ahead of the phases, throttled only by the ring -------------
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int pi = 0; pi < p.n_phase; ++pi) {
                const Phase& ph = p.ph[pi];
                if (ph.type != PH_GEMM) continue;
                const int total = ph.tiles_a * ph.n_split, kb_total = (ph.K + BK - 1) / BK;
                for (int t = blockIdx.x; t < total; t += G) {
                    const int ta = t % ph.tiles_a, split = t / ph.tiles_a;
                    const int kb0 = split * ph.kb_per_split, kb1 = min(kb0 + ph.kb_per_split, kb_total);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                        const uint32_t fb = smem_u32(&full_bar[stage]);
                        mbar_expect_tx(fb, A_BYTES);
                        tma_load_2d(smem_base + stage * STAGE_BYTES, &maps.m[ph.map], fb, kb * BK, ta * BM);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else {
        // ------------- workers: everything that depends on data produced by earlier kernels / phases -------------
        pdl_wait();
        pdl_trigger();          // late trigger: the successor (decode attention) may prefetch KV pages
        int stage = 0, local = 0;   // ring position / accumulator parity, advanced identically by every role
        uint32_t phase = 0;
        unsigned bar_count = 0;
        for (int pi = 0; pi < p.n_phase; ++pi) {
            const Phase& ph = p.ph[pi];
            unsigned long long* dbg = p.dbg ? p.dbg + (static_cast<size_t>(blockIdx.x) * 8 + pi) * 4 : nullptr;
            if (dbg && threadIdx.x == 32) dbg[0] = gtimer();
            if (ph.type == PH_GEMM) {
                const int total = ph.tiles_a * ph.n_split, kb_total = (ph.K + BK - 1) / BK;
                for (int t = blockIdx.x; t < total; t += G, ++local) {
                    const int ta = t % ph.tiles_a, split = t / ph.tiles_a;
                    const int kb0 = split * ph.kb_per_split, kb1 = min(kb0 + ph.kb_per_split, kb_total);
                    const int acc = local & 1;
                    const uint32_t acc_phase = (local >> 1) & 1;
                    if (warp == 2 || warp == 3) {
                        // activation loaders: x[m][kb*64 .. +64] -> smem B tile, K-major rows of 128 B, 128B swizzle
                        const int lt = threadIdx.x - 64;
                        int st = stage;
                        uint32_t phs = phase;
                        const __nv_bfloat16* xsrc = static_cast<const __nv_bfloat16*>(ph.x);
                        for (int kb = kb0; kb < kb1; ++kb) {
                            mbar_wait(smem_u32(&empty_bar[st]), phs ^ 1);
                            uint8_t* sb = smem_gen + st * STAGE_BYTES + A_BYTES;
#pragma unroll
                            for (int c = lt; c < BN * 8; c += LOADERS) {
                                const int m = c >> 3, ch = c & 7;
                                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                                const int k = kb * BK + ch * 8;
                                if (m < p.B && k < ph.K) v = __ldcg(reinterpret_cast<const uint4*>(xsrc + static_cast<int64_t>(m) * ph.K + k));   // L2: written by another CTA
                                *reinterpret_cast<uint4*>(sb + m * 128 + ((ch ^ (m & 7)) << 4)) = v;
                            }
                            fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(smem_u32(&full_bar[st]));
                            if (++st == STAGES) { st = 0; phs ^= 1; }
                        }
                    } else if (warp == 1) {
                        if (elect_one()) {
                            mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
                            tc_fence_after();
                            const uint32_t d_tmem = tmem_base + acc * BN;
                            int st = stage;
                            uint32_t phs = phase;
                            for (int kb = kb0; kb < kb1; ++kb) {
                                mbar_wait(smem_u32(&full_bar[st]), phs);
                                tc_fence_after();
                                const uint32_t sa = smem_base + st * STAGE_BYTES;
#pragma unroll
                                for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                                    const uint64_t da = umma_desc_sw128(sa + kk * UMMA_K * 2, 16, 1024);
                                    const uint64_t db = umma_desc_sw128(sa + A_BYTES + kk * UMMA_K * 2, 16, 1024);
                                    umma_f16(d_tmem, da, db, IDESC, (kb > kb0 || kk > 0) ? 1u : 0u);
                                }
                                umma_commit(smem_u32(&empty_bar[st]));
                                if (++st == STAGES) { st = 0; phs ^= 1; }
                            }
                            umma_commit(smem_u32(&tfull_bar[acc]));
                        }
                        __syncwarp();
                    } else if (warp >= 4) {
                        const int q = warp & 3;
                        mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
                        tc_fence_after();
                        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
                        const int row = ta * BM + q * 32 + lane;
                        float* dst = ph.part + split * ph.part_stride + row;
                        constexpr int CH = (BN >= 32) ? 32 : 16;
#pragma unroll 1
                        for (int c0 = 0; c0 < BN; c0 += CH) {
                            if (c0 >= p.B) break;
                            uint32_t v[CH];
                            if constexpr (CH == 32) tmem_ld_32x32(t_addr + c0, reinterpret_cast<uint32_t(&)[32]>(v));
                            else tmem_ld_32x16(t_addr + c0, reinterpret_cast<uint32_t(&)[16]>(v));
                            tmem_ld_wait();
                            if (row < ph.N) {
#pragma unroll
                                for (int j = 0; j < CH; ++j) {
                                    const int m = c0 + j;
                                    if (m < p.B) dst[static_cast<int64_t>(m) * ph.N] = __uint_as_float(v[j]);
                                }
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
                    }
                    // every worker role advances the ring position by this tile's k-blocks
                    for (int kb = kb0; kb < kb1; ++kb)
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            } else if (ph.type == PH_LN) {
                // rows blockIdx.x, blockIdx.x + G, ...: x32 += bias + sum(part); y = LN(x32) / cast(x32)
                const int wt = threadIdx.x - 32;
                const int d = ph.d, nchunk = d >> 2;
                for (int r = blockIdx.x; r < p.B; r += G) {
                    float* xr = ph.x32 + static_cast<int64_t>(r) * d;
                    float4 v[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int c4 = wt + i * WORKERS;
                        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c4 < nchunk) {
                            const int c = c4 * 4;
                            float4 a = __ldcg(reinterpret_cast<const float4*>(xr + c));
                            if (ph.n_part > 0) {
                                float4 accv = ph.bias ? *reinterpret_cast<const float4*>(ph.bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                                const float* pp = ph.in_part + static_cast<int64_t>(r) * d + c;
                                int s = 0;
                                for (; s + 4 <= ph.n_part; s += 4) {
                                    const float4 p0 = __ldcg(reinterpret_cast<const float4*>(pp + (s + 0) * ph.in_stride));
                                    const float4 p1 = __ldcg(reinterpret_cast<const float4*>(pp + (s + 1) * ph.in_stride));
                                    const float4 p2 = __ldcg(reinterpret_cast<const float4*>(pp + (s + 2) * ph.in_stride));
                                    const float4 p3 = __ldcg(reinterpret_cast<const float4*>(pp + (s + 3) * ph.in_stride));
                                    accv.x += p0.x; accv.y += p0.y; accv.z += p0.z; accv.w += p0.w;
                                    accv.x += p1.x; accv.y += p1.y; accv.z += p1.z; accv.w += p1.w;
                                    accv.x += p2.x; accv.y += p2.y; accv.z += p2.z; accv.w += p2.w;
                                    accv.x += p3.x; accv.y += p3.y; accv.z += p3.z; accv.w += p3.w;
                                }
                                for (; s < ph.n_part; ++s) {
                                    const float4 p0 = __ldcg(reinterpret_cast<const float4*>(pp + s * ph.in_stride));
                                    accv.x += p0.x; accv.y += p0.y; accv.z += p0.z; accv.w += p0.w;
                                }
                                a.x += accv.x; a.y += accv.y; a.z += accv.z; a.w += accv.w;
                                *reinterpret_cast<float4*>(xr + c) = a;
                            }
                            v[i] = a;
                        }
                    }
                    if (ph.y == nullptr) continue;
                    __nv_bfloat16* yr = static_cast<__nv_bfloat16*>(ph.y) + static_cast<int64_t>(r) * d;
                    if (ph.gamma == nullptr) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int c4 = wt + i * WORKERS;
                            if (c4 < nchunk) *reinterpret_cast<uint2*>(yr + c4 * 4) = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
                        }
                        continue;
                    }
                    float s = 0.f;
#pragma unroll
                    for (int i = 0; i < 2; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
                    const float mean = worker_sum(s, red) / d;
                    float qv = 0.f;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        if (wt + i * WORKERS < nchunk) {
                            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
                            qv += (a * a + b * b) + (c * c + e * e);
                        }
                    }
                    const float rstd = rsqrtf(worker_sum(qv, red) / d + ph.eps);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int c4 = wt + i * WORKERS;
                        if (c4 < nchunk) {
                            const int c = c4 * 4;
                            const float4 g = *reinterpret_cast<const float4*>(ph.gamma + c);
                            const float4 bt = *reinterpret_cast<const float4*>(ph.beta + c);
                            const float o0 = (v[i].x - mean) * rstd * g.x + bt.x, o1 = (v[i].y - mean) * rstd * g.y + bt.y;
                            const float o2 = (v[i].z - mean) * rstd * g.z + bt.z, o3 = (v[i].w - mean) * rstd * g.w + bt.w;
                            *reinterpret_cast<uint2*>(yr + c) = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
                        }
                    }
                }
            } else {   // PH_ACT
                const int64_t total4 = static_cast<int64_t>(p.B) * ph.N / 4;
                const int wt = threadIdx.x - 32;
                for (int64_t i = static_cast<int64_t>(blockIdx.x) * WORKERS + wt; i < total4; i += static_cast<int64_t>(G) * WORKERS) {
                    const int64_t e = i * 4;
                    const int n = static_cast<int>(e % ph.N);
                    float4 accv = ph.bias ? *reinterpret_cast<const float4*>(ph.bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int s = 0; s < ph.n_part; ++s) {
                        const float4 pv = __ldcg(reinterpret_cast<const float4*>(ph.in_part + s * ph.in_stride + e));
                        accv.x += pv.x; accv.y += pv.y; accv.z += pv.z; accv.w += pv.w;
                    }
                    accv.x = gelu_erf(accv.x); accv.y = gelu_erf(accv.y); accv.z = gelu_erf(accv.z); accv.w = gelu_erf(accv.w);
                    *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(ph.y) + e) = make_uint2(pack_bf16x2(accv.x, accv.y), pack_bf16x2(accv.z, accv.w));
                }
            }
            if (dbg && threadIdx.x == 4 * 32) dbg[1] = gtimer();      // an epilogue warp: its part of the phase is done
            if (pi + 1 < p.n_phase) {
                ++bar_count;
                __threadfence();
                worker_sync();
                if (dbg && threadIdx.x == 32) dbg[2] = gtimer();       // whole CTA arrived
                grid_barrier(p.gbar, bar_count * G);
                if (dbg && threadIdx.x == 32) dbg[3] = gtimer();       // barrier released
            }
        }
        // leave the counter at zero for the next launch: the last CTA to check out resets it
        __threadfence();
        worker_sync();
        if (threadIdx.x == 32) {
            const unsigned done = atomicAdd(p.gbar, 1u);
            if (done == (bar_count + 1) * G - 1) atomicExch(p.gbar, 0u);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

template <int BN>
int launch_chain(const Maps& maps, const ChainParams& p, int grid, cudaStream_t st) {
    constexpr int SMEM = STAGES * (A_BYTES + BN * BK * 2) + 1024;
    static bool configured = false;
    auto kern = decode_chain_kernel<BN>;
    if (!configured) {
        VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    VB_CUDA(vb_launch(true, kern, dim3(grid), dim3(THREADS), SMEM, st, maps, p));
    return VB_OK;
}

}  // namespace

static unsigned long long* g_chain_dbg = nullptr;
extern "C" int vb_decode_chain_set_debug(void* buf) {   /* device buffer of #SM * 8 * 4 uint64, or NULL to switch off */
    g_chain_dbg = static_cast<unsigned long long*>(buf);
    return VB_OK;
}

extern "C" int vb_decode_chain(const vb_chain_phase* phases, int n_phase, int B, void* grid_barrier_counter, void* stream) {
    VB_REQUIRE(phases && grid_barrier_counter, VB_ERR_BAD_ARG, "vb_decode_chain: null pointer");
    VB_REQUIRE(n_phase >= 1 && n_phase <= 8, VB_ERR_BAD_ARG, "vb_decode_chain: 1..8 phases (got %d)", n_phase);
    VB_REQUIRE(B >= 1 && B <= 64, VB_ERR_UNSUPPORTED, "vb_decode_chain: batch must be in [1,64] (got %d)", B);
    Maps maps;
    memset(&maps, 0, sizeof(maps));
    ChainParams p;
    memset(&p, 0, sizeof(p));
    p.n_phase = n_phase;
    p.B = B;
    p.gbar = static_cast<unsigned*>(grid_barrier_counter);
    p.dbg = g_chain_dbg;
    int n_maps = 0, rc;
    for (int i = 0; i < n_phase; ++i) {
        const vb_chain_phase& s = phases[i];
        Phase& d = p.ph[i];
        d.type = s.type;
        if (s.type == VB_PHASE_GEMM) {
            VB_REQUIRE(s.x && s.w && s.out_part && s.K % 8 == 0 && s.N >= 1, VB_ERR_BAD_ARG, "vb_decode_chain: bad GEMM phase %d", i);
            VB_REQUIRE(n_maps < 4, VB_ERR_UNSUPPORTED, "vb_decode_chain: at most 4 GEMM phases");
            if ((rc = vb_make_tmap_bf16_2d(&maps.m[n_maps], s.w, s.N, s.K, s.K, BM, BK)) != VB_OK) return rc;
            d.map = n_maps++;
            d.x = s.x; d.N = s.N; d.K = s.K;
            d.tiles_a = (s.N + BM - 1) / BM;
            const int kb_total = (s.K + BK - 1) / BK;
            d.n_split = vb_linear_decode_splits(s.N, s.K, s.max_split);
            d.kb_per_split = (kb_total + d.n_split - 1) / d.n_split;
            d.part = s.out_part; d.part_stride = s.out_part_stride;
        } else if (s.type == VB_PHASE_LN) {
            VB_REQUIRE(s.x32 && s.d % 4 == 0 && s.d <= 8 * WORKERS && (s.n_part == 0 || s.in_part), VB_ERR_BAD_ARG, "vb_decode_chain: bad LN phase %d", i);
            d.x32 = s.x32; d.in_part = s.in_part; d.n_part = s.n_part; d.in_stride = s.in_part_stride; d.bias = s.bias;
            d.gamma = s.gamma; d.beta = s.beta; d.y = s.y; d.d = s.d; d.eps = s.eps;
        } else if (s.type == VB_PHASE_ACT) {
            VB_REQUIRE(s.in_part && s.y && s.N % 4 == 0 && s.n_part >= 1, VB_ERR_BAD_ARG, "vb_decode_chain: bad ACT phase %d", i);
            d.in_part = s.in_part; d.n_part = s.n_part; d.in_stride = s.in_part_stride; d.bias = s.bias; d.y = s.y; d.N = s.N;
        } else {
            VB_REQUIRE(false, VB_ERR_BAD_ARG, "vb_decode_chain: unknown phase type %d", s.type);
        }
    }
    for (int i = n_maps; i < 4; ++i) maps.m[i] = maps.m[0];
    VB_REQUIRE(n_maps >= 1, VB_ERR_BAD_ARG, "vb_decode_chain: needs at least one GEMM phase");
    const int grid = vb_sm_count();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (B <= 16) return launch_chain<16>(maps, p, grid, st);
    if (B <= 32) return launch_chain<32>(maps, p, grid, st);
    return launch_chain<64>(maps, p, grid, st);
}
