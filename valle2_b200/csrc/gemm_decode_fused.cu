// Decode-shape GEMM with the split-K reduction, LayerNorm and the epilogue fused in:  y[B<=64][N] = epi(A[B][K] . W[N][K]^T)
//
// Same decomposition as the slice kernel (gemm_tc.cu, swap-AB): a CTA multiplies a 128-row slab of W (MMA-M) by the
// batch (MMA-N = 16/32/64) over ONE slice of K, so N/128 x splits ~ #SMs CTAs each stream an equal share of the weights.
// What is new is what happens around the MMAs:
//   * the CTAs that share a weight slab form a thread-block CLUSTER along K.  Their partial accumulators go TMEM ->
//     shared memory and are summed through distributed shared memory in rank order (deterministic); every rank finalises
//     128/C of the slab's rows and applies the epilogue -- bias, erf-GELU -> bf16, or "x += . + bias" in place on the fp32
//     residual stream.  No fp32 slices in HBM, no separate reduce kernel in the step's dependency chain.
//   * LayerNorm runs "on load" (a_dtype = fp32): each rank reads only ITS K-slice of the fp32 residual rows, the ranks
//     exchange per-row partial (mean, M2) through DSMEM and combine them with Chan's formula (ranks of one cluster cover
//     a whole row between them), then each normalises its slice straight into the 128B-swizzled operand tile.  This
//     removes the LN kernels (modules.py:271/276 norm1/norm2) from the chain.
//   * the rank's weight tiles are requested by TMA before the programmatic-dependency wait (they are immutable).
// One decoder layer becomes QKV(LN1) -> attention -> out-proj(+residual) -> FFN1(LN2, GELU) -> FFN2(+residual): five
// dependent kernels instead of eight.
//
//   warp 0   producer: weight TMA (pre-wait), activation TMA (bf16 A)     warp 1   MMA issuer + TMEM owner
//   warp 2-5 LayerNorm-on-load, TMEM -> partials, DSMEM reduce + epilogue (TMEM lane quadrant = warp & 3)
// Replaces the nn.Linear call sites modules.py:146 (qkv), :171 (out), :220-221 (ffn), valle_ar.py:158 (proj) on the
// KV-cached decode path.
#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int MAX_KB = 8;             // k-blocks per CTA
constexpr int THREADS = 192;
constexpr int WORKERS = 128;          // warps 2..5
constexpr int MAX_B = 64;
constexpr int W_BYTES = BM * BK * 2;  // 16 KB weight tile
constexpr int MAXV = 16;              // float4 registers per lane that hold LayerNorm rows across the cluster exchange

struct FusedParams {
    int B, N, K;
    int kb_total, kb_per_cta, cluster;
    int a_mode;                       // 0: bf16 rows by TMA; 1: fp32 rows, LayerNorm (gamma) or plain cast on load
    const float* a32;
    int64_t lda32;
    const float* gamma;
    const float* beta;
    float eps;
    int epilogue;
    const float* bias;
    void* y;
    int y_bf16;
    int64_t ldy;
    int late_trigger;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_addr, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
    return remote;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t remote) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
    return v;
}
__device__ __forceinline__ float2 ld_dsmem_f32x2(uint32_t remote) {
    float2 v;
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(remote) : "memory");
    return v;
}

// one finished dot product: row b (batch), column n (output feature)
__device__ __forceinline__ void emit(const FusedParams& p, int b, int n, float v) {
    if (p.epilogue != VB_EPI_NONE) v += __ldg(p.bias + n);
    if (p.epilogue == VB_EPI_BIAS_GELU) v = gelu_erf_fast(v);
    const int64_t off = static_cast<int64_t>(b) * p.ldy + n;
    if (p.y_bf16) {
        static_cast<__nv_bfloat16*>(p.y)[off] = __float2bfloat16_rn(v);
    } else {
        float* dst = static_cast<float*>(p.y) + off;
        if (p.epilogue == VB_EPI_BIAS_RESIDUAL) v += *dst;
        *dst = v;
    }
}

// LayerNorm (or plain cast when gamma == null) of this rank's K-slice [col0, col0 + kcols) of the fp32 rows, written as
// bf16 into the 128B-swizzled activation tiles.  Worker warp wi owns rows wi, wi + 4, ...; a lane holds NCH float4 chunks of
// a row (columns (lane + 32 ch) * 4 of the slice).  Rows stay in registers across the cluster exchange of the statistics
// when they fit (RPG rows); otherwise they are read again from L2.
template <int NCH>
__device__ __forceinline__ void ln_on_load(const FusedParams& p, float2* stats, uint8_t* smem_gen, int stage_bytes, int col0,
                                           int kcols, int wi, int lane, int C, bool ln_exchange) {
    constexpr int RPG = MAXV / NCH;                       // rows per register group
    const int rows_w = (p.B - wi + 3) >> 2;
    const bool single = rows_w <= RPG;
    const bool norm = p.gamma != nullptr;
    float4 v[MAXV];
    auto load_group = [&](int g0) {
#pragma unroll
        for (int r = 0; r < RPG; ++r)
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int cs = (lane + 32 * ch) * 4;
                v[r * NCH + ch] = (g0 + r < rows_w && cs < kcols)
                    ? __ldcg(reinterpret_cast<const float4*>(p.a32 + static_cast<int64_t>(wi + 4 * (g0 + r)) * p.lda32 + col0 + cs))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
            }
    };
    if (norm) {
        // pass A: (mean, M2) of the slice, two-pass from registers
        for (int g0 = 0; g0 < rows_w; g0 += RPG) {
            load_group(g0);
#pragma unroll
            for (int r = 0; r < RPG; ++r) {
                if (g0 + r >= rows_w) break;
                float s = 0.f;
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) s += (v[r * NCH + ch].x + v[r * NCH + ch].y) + (v[r * NCH + ch].z + v[r * NCH + ch].w);
                const float mean = kcols > 0 ? warp_sum(s) / kcols : 0.f;    // a rank past the end of K: count 0
                float q = 0.f;
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    if ((lane + 32 * ch) * 4 < kcols) {
                        const float4 t = v[r * NCH + ch];
                        const float a = t.x - mean, b = t.y - mean, c = t.z - mean, e = t.w - mean;
                        q += (a * a + b * b) + (c * c + e * e);
                    }
                }
                q = warp_sum(q);
                if (lane == 0) stats[wi + 4 * (g0 + r)] = make_float2(mean, q);
            }
        }
    }
    __syncwarp();
    if (ln_exchange) cluster_sync_all();          // every rank's slice statistics are in its shared memory
    // pass B: combine the ranks' statistics (Chan), normalise, store
    for (int g0 = 0; g0 < rows_w; g0 += RPG) {
        if (!single || !norm) load_group(g0);
        float2 st[RPG];
        if (norm) {
#pragma unroll
            for (int r = 0; r < RPG; ++r) {                  // all remote reads first, then the arithmetic
                st[r] = make_float2(0.f, 0.f);
                if (g0 + r < rows_w && lane < C) {
                    const uint32_t la = smem_u32(&stats[wi + 4 * (g0 + r)]);
                    st[r] = (C > 1) ? ld_dsmem_f32x2(dsmem_addr(la, lane)) : stats[wi + 4 * (g0 + r)];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RPG; ++r) {
            if (g0 + r >= rows_w) break;
            const int row = wi + 4 * (g0 + r);
            float mean = 0.f, rstd = 1.f;
            if (norm) {
                const int slice = p.kb_per_cta * BK;
                float tot = 0.f;
                for (int l = 0; l < C; ++l) tot += __shfl_sync(0xffffffffu, st[r].x, l) * max(0, min(slice, p.K - l * slice));
                mean = tot / p.K;
                float m2 = 0.f;
                for (int l = 0; l < C; ++l) {
                    const float dm = __shfl_sync(0xffffffffu, st[r].x, l) - mean;
                    m2 += __shfl_sync(0xffffffffu, st[r].y, l) + max(0, min(slice, p.K - l * slice)) * dm * dm;
                }
                rstd = rsqrtf(m2 / p.K + p.eps);
            }
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int cs = (lane + 32 * ch) * 4;
                if (cs < kcols) {
                    float4 o = v[r * NCH + ch];
                    if (norm) {
                        const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + col0 + cs));
                        const float4 bt = __ldg(reinterpret_cast<const float4*>(p.beta + col0 + cs));
                        o.x = (o.x - mean) * rstd * g.x + bt.x; o.y = (o.y - mean) * rstd * g.y + bt.y;
                        o.z = (o.z - mean) * rstd * g.z + bt.z; o.w = (o.w - mean) * rstd * g.w + bt.w;
                    }
                    const int kb = cs >> 6, cin = cs & 63;
                    uint8_t* dst = smem_gen + kb * stage_bytes + W_BYTES + row * 128 + (((cin >> 3) ^ (row & 7)) << 4) + ((cin >> 2) & 1) * 8;
                    *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
                }
            }
        }
    }
}

template <int BN>
__global__ void __launch_bounds__(THREADS, 1) gemm_decode_fused_kernel(const __grid_constant__ CUtensorMap tm_a,
                                                                       const __grid_constant__ CUtensorMap tm_w,
                                                                       const FusedParams p) {
    constexpr int A_BYTES = BN * BK * 2;          // activation tile (MMA B operand): BN batch rows x 64 bf16
    constexpr int STAGE_BYTES = W_BYTES + A_BYTES;
    constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    constexpr uint32_t IDESC = umma_idesc_bf16(BM, BN, 0, 0);

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t wbar[MAX_KB];
    __shared__ __align__(8) uint64_t abar[MAX_KB];
    __shared__ __align__(8) uint64_t d_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(8) float2 stats[MAX_B];  // per batch row: (mean, M2) of this rank's K-slice

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = p.cluster;
    const int crank = (C > 1) ? static_cast<int>(cluster_ctarank()) : 0;
    const int tile = blockIdx.x / C;
    const int n0 = tile * BM;
    const int kb0 = crank * p.kb_per_cta;
    const int nkb = max(0, min(p.kb_per_cta, p.kb_total - kb0));
    const bool ln_exchange = (p.a_mode == 1) && (C > 1) && (p.gamma != nullptr);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_w);
        if (p.a_mode == 0) tma_prefetch_desc(&tm_a);
        for (int i = 0; i < MAX_KB; ++i) {
            mbar_init(smem_u32(&wbar[i]), 1);
            mbar_init(smem_u32(&abar[i]), p.a_mode == 0 ? 1 : WORKERS / 32);
        }
        mbar_init(smem_u32(&d_bar), 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(&tmem_base_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (!p.late_trigger) pdl_trigger();
    if (warp == 0) {
        const bool leader = elect_one();
        if (leader) {
            for (int kb = 0; kb < nkb; ++kb) {            // weights: immutable, requested before the dependency wait
                const uint32_t wb = smem_u32(&wbar[kb]);
                mbar_expect_tx(wb, W_BYTES);
                tma_load_2d(smem_base + kb * STAGE_BYTES, &tm_w, wb, (kb0 + kb) * BK, n0);
            }
        }
        __syncwarp();
        pdl_wait();
        if (p.late_trigger) pdl_trigger();
        if (leader && p.a_mode == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const uint32_t ab = smem_u32(&abar[kb]);
                mbar_expect_tx(ab, A_BYTES);
                tma_load_2d(smem_base + kb * STAGE_BYTES + W_BYTES, &tm_a, ab, (kb0 + kb) * BK, 0);
            }
        }
        __syncwarp();
        if (ln_exchange) cluster_sync_all();
    } else if (warp == 1) {
        pdl_wait();
        if (p.late_trigger) pdl_trigger();
        if (ln_exchange) cluster_sync_all();
        if (elect_one()) {
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(smem_u32(&wbar[kb]), 0);
                mbar_wait(smem_u32(&abar[kb]), 0);
                tc_fence_after();
                const uint32_t sa = smem_base + kb * STAGE_BYTES;
#pragma unroll
                for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                    const uint64_t da = umma_desc_sw128(sa + kk * UMMA_K * 2, 16, 1024);
                    const uint64_t db = umma_desc_sw128(sa + W_BYTES + kk * UMMA_K * 2, 16, 1024);
                    umma_f16(tmem_base, da, db, IDESC, (kb > 0 || kk > 0) ? 1u : 0u);
                }
            }
            if (nkb > 0) umma_commit(smem_u32(&d_bar));
            else mbar_arrive(smem_u32(&d_bar));          // a rank past the end of K has nothing in flight
        }
        __syncwarp();
    } else {
        const int wi = warp - 2;
        pdl_wait();
        if (p.late_trigger) pdl_trigger();
        if (p.a_mode == 1) {
            const int kcols = max(0, min(nkb * BK, p.K - kb0 * BK));       // K % 64 == 0: whole float4 chunks
            if (kcols <= 128) ln_on_load<1>(p, stats, smem_gen, STAGE_BYTES, kb0 * BK, kcols, wi, lane, C, ln_exchange);
            else if (kcols <= 256) ln_on_load<2>(p, stats, smem_gen, STAGE_BYTES, kb0 * BK, kcols, wi, lane, C, ln_exchange);
            else ln_on_load<4>(p, stats, smem_gen, STAGE_BYTES, kb0 * BK, kcols, wi, lane, C, ln_exchange);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0)
                for (int kb = 0; kb < nkb; ++kb) mbar_arrive(smem_u32(&abar[kb]));
        }
        // ---------------- accumulator: TMEM -> final values (no cluster) or -> shared-memory partial ----------------
        const int q = warp & 3;
        const int nl = q * 32 + lane;                    // weight row inside the slab (TMEM lane)
        mbar_wait(smem_u32(&d_bar), 0);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        float* pbuf = reinterpret_cast<float*>(smem_gen);           // [b][128] fp32; the operand tiles are dead now
        constexpr int CH = (BN >= 32) ? 32 : 16;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += CH) {
            if (c0 >= p.B) break;
            uint32_t acc[CH];
            if constexpr (CH == 32) tmem_ld_32x32(t_addr + c0, reinterpret_cast<uint32_t(&)[32]>(acc));
            else tmem_ld_32x16(t_addr + c0, reinterpret_cast<uint32_t(&)[16]>(acc));
            tmem_ld_wait();
            if (C == 1) {
                if (n0 + nl < p.N) {
#pragma unroll
                    for (int j = 0; j < CH; ++j)
                        if (c0 + j < p.B) emit(p, c0 + j, n0 + nl, __uint_as_float(acc[j]));
                }
            } else {
                const uint32_t keep = nkb > 0 ? 0xffffffffu : 0u;
#pragma unroll
                for (int j = 0; j < CH; ++j)
                    if (c0 + j < p.B) pbuf[(c0 + j) * BM + nl] = __uint_as_float(acc[j] & keep);
            }
        }
        __syncwarp();
    }
    if (C > 1) {
        cluster_sync_all();            // every rank's partial is in its shared memory
        if (warp >= 2) {
            // rank r finalises weight rows [128 r / C, 128 (r+1) / C) of the slab for every batch row
            const int n_lo = (BM * crank) / C, n_hi = (BM * (crank + 1)) / C, nr = n_hi - n_lo;
            const int total = p.B * nr;
            for (int e = threadIdx.x - 64; e < total; e += WORKERS) {
                const int b = e / nr, n = n_lo + (e - b * nr);
                if (n0 + n >= p.N) continue;
                const uint32_t la = smem_base + static_cast<uint32_t>(b * BM + n) * 4u;
                float t[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) t[r] = (r < C) ? ld_dsmem_f32(dsmem_addr(la, r)) : 0.f;   // all loads in flight
                float acc = 0.f;
#pragma unroll
                for (int r = 0; r < 16; ++r) acc += t[r];                                              // fixed rank order
                emit(p, b, n0 + n, acc);
            }
        }
        __syncwarp();
        cluster_sync_all();            // nobody leaves (and frees its shared memory) while a peer may still read it
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

template <int BN>
int launch_fused(const CUtensorMap& ta, const CUtensorMap& tw, const FusedParams& p, int tiles, cudaStream_t st) {
    const int stage = W_BYTES + BN * BK * 2;
    const int smem = max(max(p.kb_per_cta, 1) * stage, BN * BM * 4) + 1024;
    static int configured = 0;
    static bool nonportable = false;
    auto kern = gemm_decode_fused_kernel<BN>;
    if (smem > configured) {
        VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    if (p.cluster > 8 && !nonportable) {
        VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        nonportable = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tiles * p.cluster);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (vb_pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (p.cluster > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = p.cluster;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    VB_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tw, p));
    return VB_OK;
}

}  // namespace

// cluster size along K chosen so that (N/128 slabs) x cluster fills the SMs: any size up to 8, or 16
extern "C" int vb_linear_decode_fused_cluster(int N, int K) {
    const int tiles = (N + BM - 1) / BM, kb_total = (K + BK - 1) / BK;
    int c = vb_sm_count() / tiles;
    c = max(1, min(c, kb_total));
    if (c >= 16 && tiles * 16 <= 128) c = 16;       // 8 GPCs hold one 16-CTA cluster each
    else c = min(c, 8);
    while ((kb_total + c - 1) / c > MAX_KB && c < 16) c = (c < 8) ? c + 1 : 16;
    return c;
}

extern "C" int vb_linear_decode_fused(const void* a, int a_dtype, int64_t lda, const float* gamma, const float* beta, float eps,
                                      const void* w, int64_t ldw, const float* bias, void* y, int y_dtype, int64_t ldy,
                                      int B, int N, int K, int epilogue, int cluster_k, int flags, void* stream) {
    VB_REQUIRE(a && w && y, VB_ERR_BAD_ARG, "vb_linear_decode_fused: null pointer");
    VB_REQUIRE(B >= 1 && B <= MAX_B, VB_ERR_UNSUPPORTED, "vb_linear_decode_fused: batch must be in [1,%d] (got %d)", MAX_B, B);
    VB_REQUIRE(N >= 1 && K >= 8 && K % 8 == 0, VB_ERR_UNSUPPORTED, "vb_linear_decode_fused: N >= 1 and K %% 8 == 0 required (N=%d K=%d)", N, K);
    VB_REQUIRE(epilogue >= VB_EPI_NONE && epilogue <= VB_EPI_BIAS_RESIDUAL, VB_ERR_BAD_ARG, "vb_linear_decode_fused: bad epilogue %d", epilogue);
    VB_REQUIRE(epilogue == VB_EPI_NONE || bias != nullptr, VB_ERR_BAD_ARG, "vb_linear_decode_fused: epilogue %d needs a bias", epilogue);
    VB_REQUIRE(y_dtype == VB_F32 || y_dtype == VB_BF16, VB_ERR_BAD_ARG, "vb_linear_decode_fused: bad y_dtype");
    VB_REQUIRE(epilogue != VB_EPI_BIAS_RESIDUAL || y_dtype == VB_F32, VB_ERR_BAD_ARG,
               "vb_linear_decode_fused: the residual epilogue updates an fp32 y in place");
    VB_REQUIRE(a_dtype == VB_BF16 || a_dtype == VB_F32, VB_ERR_BAD_ARG, "vb_linear_decode_fused: bad a_dtype");
    FusedParams p{};
    p.B = B; p.N = N; p.K = K;
    p.kb_total = (K + BK - 1) / BK;
    p.a_mode = (a_dtype == VB_F32) ? 1 : 0;
    const int cluster = cluster_k > 0 ? cluster_k : vb_linear_decode_fused_cluster(N, K);
    VB_REQUIRE((cluster >= 1 && cluster <= 8) || cluster == 16, VB_ERR_UNSUPPORTED, "vb_linear_decode_fused: cluster_k must be 1..8 or 16 (got %d)", cluster);
    p.cluster = cluster;
    p.kb_per_cta = (p.kb_total + cluster - 1) / cluster;
    VB_REQUIRE(p.kb_per_cta <= MAX_KB, VB_ERR_UNSUPPORTED, "vb_linear_decode_fused: K=%d too long for cluster_k=%d (at most %d columns per CTA)", K, cluster, MAX_KB * BK);
    if (p.a_mode == 1) {
        VB_REQUIRE(K % 64 == 0, VB_ERR_UNSUPPORTED, "vb_linear_decode_fused: fp32 rows (LayerNorm on load) need K %% 64 == 0 (K=%d)", K);
        VB_REQUIRE((gamma == nullptr) == (beta == nullptr), VB_ERR_BAD_ARG, "vb_linear_decode_fused: gamma and beta go together");
        VB_REQUIRE(lda % 4 == 0, VB_ERR_BAD_ARG, "vb_linear_decode_fused: lda must be a multiple of 4 for fp32 rows");
        p.a32 = static_cast<const float*>(a); p.lda32 = lda; p.gamma = gamma; p.beta = beta; p.eps = eps;
    }
    p.epilogue = epilogue; p.bias = bias; p.y = y; p.y_bf16 = (y_dtype == VB_BF16); p.ldy = ldy;
    p.late_trigger = (flags & VB_FLAG_LATE_TRIGGER) ? 1 : 0;
    const int tiles = (N + BM - 1) / BM;
    const int bn = B <= 16 ? 16 : (B <= 32 ? 32 : 64);
    CUtensorMap ta, tw;
    int rc;
    if ((rc = vb_make_tmap_bf16_2d(&tw, w, N, K, ldw, BM, BK)) != VB_OK) return rc;
    if (p.a_mode == 0) {
        if ((rc = vb_make_tmap_bf16_2d(&ta, a, B, K, lda, bn, BK)) != VB_OK) return rc;
    } else {
        ta = tw;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (bn == 16) return launch_fused<16>(ta, tw, p, tiles, st);
    if (bn == 32) return launch_fused<32>(ta, tw, p, tiles, st);
    return launch_fused<64>(ta, tw, p, tiles, st);
}
