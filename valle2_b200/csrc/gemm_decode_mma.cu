// Decode-shape linear layers, "rows" form: y[M][N] = epilogue(x[M][K] . w[N][K]^T) for M <= 32 (one decode step of the batch).
//
// Why a second decode GEMM next to the tcgen05 swap-AB kernel (gemm_tc.cu): at M <= 32 the contraction is a weight STREAM
// (304 MB per step for the whole stack) followed by a latency chain -- activations of the step arrive, a few dozen MMAs,
// a store, and the next kernel of the layer waits for all of it.  The tcgen05 form pays TMEM allocation, TMA boxes for 32
// activation rows, commit -> mbarrier -> tcgen05.ld and, above all, split-K slices that a following kernel has to reduce.
// This kernel keeps the whole K of an output element inside ONE CTA whenever K <= 1024, so that bias / GELU / residual run
// in the epilogue and the reduce kernels (and their PDL hops) disappear from the decode chain:
//
//   * weights go HBM -> registers directly, in mma.sync B-fragment order, BEFORE the kernel waits on its predecessor
//     (programmatic dependent launch): thread (g, t) of a warp owns 16-byte pieces of weight row n0 + g.  The k index
//     inside each 64-element group is permuted (thread t takes elements 16t..16t+15) -- legal because A and B use the same
//     permutation -- which makes every load a full 16-byte vector and every weight row a 128-byte contiguous request;
//   * the step's activations (M x Kc bf16) come L2 -> shared memory with 16-byte cp.async, every warp copying its own
//     k-slice in commit groups per (128-byte k group, m16 tile) -- a warp starts the MMAs of a tile as soon as that tile
//     has landed, while the rest is still in flight (256 small cp.async.bulk copies per CTA cost ~9 cycles of TMA issue
//     each and delivered the last byte 2 300 cycles after the dependency resolved -- tools/rows_timeline.py);
//   * eight warps split Kc; the partial 32 x (8 NT) accumulators are summed through shared memory in warp order
//     (deterministic) by 256 threads that then apply the epilogue and write 16-byte vectors.
//
//   * LayerNorm on load (vb_linear_decode_rows_ln, M <= 8): the CTA reads the fp32 residual rows themselves (4 KB per
//     row), every warp normalises its k-slice -- row statistics are combined across the eight warps with Chan's formula
//     (per-warp mean and M2 over equal counts) -- and writes bf16 rows into shared memory for the MMAs: the separate
//     LayerNorm kernel and its PDL hop disappear (5 kernels per layer instead of 8 at small batch);
//   * whole K inside one CTA up to K = 4096 when the batch is small enough for the activations to fit (FFN2 at M <= 8).
//
// Where it wins (tools/rows_timeline.py, tools/layer_chain.py, bench.py --batch): every CTA reads the WHOLE (M, K)
// activation matrix from L2, 128 CTAs at once: 0.3 us after the dependency resolves at M = 1, 1.4 us at M = 32 (8.4 MB
// of L2 -> SM traffic, ~3 100 B/cycle chip-wide; replicating the rows in HBM does not help).  Below ~16 rows it beats the
// split-K tcgen05 kernel + reduce kernels, above it does not -- ARDecoder picks per batch size.
//
// Replaces on the decode path: modules.py:146 (qkv), :171 (out + residual :274), :220-221 (linear_1 + GELU, linear_2 +
// residual :278), valle_ar.py:158 (proj).
#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kRedPitch = 40;   // floats; pitch % 32 == 8 keeps the 8-byte fragment stores conflict-free
constexpr int kStatBytes = 512; // LN variant: per-warp (mean, M2) of up to 8 rows
constexpr int kLnRows = 8;      // LN-on-load handles up to 8 batch rows (their fp32 k-slices live in registers)

struct RowsParams {
    const __nv_bfloat16* x;   // bf16 activations (plain variant)
    const float* x32;         // fp32 residual rows (LN variant)
    int64_t ldx;
    const float* gamma;       // LN variant; NULL = plain cast
    const float* beta;
    float eps;
    const __nv_bfloat16* w;
    int64_t ldw;
    const float* bias;
    void* y;
    int64_t ldy;
    int64_t split_stride;
    int M, N, Kc;
    int epi, y_bf16, late_trigger, vec_ok;
    unsigned long long* dbg;   // optional timeline stamps [cta][16] (vb_linear_decode_rows_set_debug)
};

// events 0..7 as %globaltimer (comparable across SMs), the same events as SM cycle counts in 8..15; thread 0 only
__device__ __forceinline__ void stamp(const RowsParams& p, int ev) {
    if (p.dbg != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        unsigned long long* d = p.dbg + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 16;
        d[ev] = t;
        d[8 + ev] = static_cast<unsigned long long>(clock64());
    }
}

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// streamed once per decode step: do not keep the line in L1
__device__ __forceinline__ uint4 ldg_weights(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// element offset of 16-byte vector j of thread t inside the warp's k-slice (the shared k permutation of A and B)
template <int KCH> __device__ __forceinline__ int vec_off(int j, int t) {
    if constexpr (KCH % 2 == 0) return (j >> 1) * 64 + t * 16 + (j & 1) * 8;
    else return j * 32 + t * 8;
}

// MT: m16 tiles of batch rows (M <= 16 MT);  NT: n8 tiles of weight rows per CTA;  KCH: 32-element k chunks per warp;
// LR > 0: activations are fp32 rows normalised on load (M <= LR <= 8, MT == 1, KCH <= 4, one CTA covers the whole row);
// the per-row work is ~80 instructions per thread, so LR is the smallest of 1, 2, 4, 8 that holds M.
template <int MT, int NT, int KCH, int LR>
__global__ void __launch_bounds__(kThreads, 2) linear_decode_rows_kernel(const RowsParams p) {
    constexpr bool LN = LR > 0;        // LR: batch rows normalised on load (1, 2, 4 or 8 >= M); 0: bf16 activations
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* buf = smem + kStatBytes;                                   // activations, later the reduction buffer
    constexpr int KW = KCH * 32;
    constexpr int JG = (KCH % 2 == 0) ? 2 : 1;                          // chunks per copy group (paired: 128 B per row)
    constexpr int NG = KCH / JG;                                        // k groups per warp
    constexpr int SEGS = JG * 4;                                        // 16-byte segments per row and group
    constexpr int RPI = 32 / SEGS;                                      // rows covered by one warp-wide cp.async
    constexpr bool PIPE = (MT > 1);                                     // stage the activation copy in commit groups
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int pitch = p.Kc * 2 + 16;                                    // bytes; % 128 == 16: conflict-free 16-byte reads
    const int n0 = blockIdx.x * (NT * 8);
    const int64_t k0 = static_cast<int64_t>(blockIdx.y) * p.Kc + warp * KW;

    stamp(p, 0);
    if (!p.late_trigger) pdl_trigger();
    // ---- prologue (before the dependency resolves): everything immutable -- weights -> registers, bias, gamma / beta
    uint4 wr[NT][KCH];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int row = min(n0 + nt * 8 + g, p.N - 1);
        const __nv_bfloat16* wp = p.w + static_cast<int64_t>(row) * p.ldw + k0;
#pragma unroll
        for (int j = 0; j < KCH; ++j) wr[nt][j] = ldg_weights(wp + vec_off<KCH>(j, t));
    }
    // the float4 group of the output tile this thread finishes after the k-slices are summed
    constexpr int kGroups = MT * 16 * NT * 2;
    static_assert(kGroups <= kThreads, "one output group per thread");
    const int om = tid / (NT * 2), oc = (tid % (NT * 2)) * 4, on = n0 + oc;
    const bool o_live = tid < kGroups && om < p.M && on < p.N;
    const bool o_full = p.vec_ok && on + 3 < p.N;
    float4 obias = make_float4(0.f, 0.f, 0.f, 0.f), ores = make_float4(0.f, 0.f, 0.f, 0.f);
    if (o_live && p.epi != VB_EPI_NONE) {
        if (o_full) obias = *reinterpret_cast<const float4*>(p.bias + on);
        else {
            obias.x = p.bias[on];
            if (on + 1 < p.N) obias.y = p.bias[on + 1];
            if (on + 2 < p.N) obias.z = p.bias[on + 2];
            if (on + 3 < p.N) obias.w = p.bias[on + 3];
        }
    }
    float lg[LN ? KCH : 1], lb[LN ? KCH : 1];                           // LN: this lane's KCH columns of gamma / beta
    if constexpr (LN) {
#pragma unroll
        for (int e = 0; e < KCH; ++e) {
            lg[e] = p.gamma ? p.gamma[k0 + lane * KCH + e] : 1.f;
            lb[e] = p.gamma ? p.beta[k0 + lane * KCH + e] : 0.f;
        }
    }
    stamp(p, 1);
    pdl_wait();
    stamp(p, 2);
    if (p.late_trigger) pdl_trigger();

    const uint32_t abase = smem_u32(buf);
    if constexpr (LN) {
        // ---- fp32 residual rows: this warp's k-slice of every row -> registers -> LayerNorm -> bf16 rows in shared memory
        static_assert(!LN || (KCH <= 4 && MT == 1), "LN on load: K <= 1024, M <= 8");
        constexpr int LRC = LN ? LR : 1;
        float xv[LRC][KCH];
#pragma unroll
        for (int r = 0; r < LRC; ++r) {
            const float* xp = p.x32 + static_cast<int64_t>(min(r, p.M - 1)) * p.ldx + k0 + lane * KCH;
            if constexpr (KCH == 4) {
                const float4 v = *reinterpret_cast<const float4*>(xp);
                xv[r][0] = v.x; xv[r][1] = v.y; xv[r][2] = v.z; xv[r][3] = v.w;
            } else if constexpr (KCH == 2) {
                const float2 v = *reinterpret_cast<const float2*>(xp);
                xv[r][0] = v.x; xv[r][1] = v.y;
            } else {
                xv[r][0] = *xp;
            }
        }
        float2* stats = reinterpret_cast<float2*>(smem);               // [warp][row] = (mean, M2) of KW elements
        if (p.gamma != nullptr) {
            // per-lane (mean, M2) of its KCH values, then a butterfly over the warp that merges pairs with Chan's formula
            // (equal counts n: mean = (a + b) / 2, M2 = M2a + M2b + (a - b)^2 n / 2).  Level-outer / row-inner order keeps
            // eight independent shuffles in flight (row-outer order serialised 80 dependent shuffles: +2 000 cycles).
            float mw[LRC], qw[LRC];
#pragma unroll
            for (int r = 0; r < LRC; ++r) {
                float s = 0.f;
#pragma unroll
                for (int e = 0; e < KCH; ++e) s += xv[r][e];
                mw[r] = s * (1.0f / KCH);
                qw[r] = 0.f;
#pragma unroll
                for (int e = 0; e < KCH; ++e) qw[r] += (xv[r][e] - mw[r]) * (xv[r][e] - mw[r]);
            }
            float half_n = 0.5f * KCH;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
                for (int r = 0; r < LRC; ++r) {     // all eight rows, no branch: rows >= M hold copies of row M-1
                    const float mo = __shfl_xor_sync(0xffffffffu, mw[r], o), qo = __shfl_xor_sync(0xffffffffu, qw[r], o);
                    const float dlt = mw[r] - mo;
                    qw[r] = (qw[r] + qo) + dlt * dlt * half_n;
                    mw[r] = 0.5f * (mw[r] + mo);
                }
                half_n *= 2.f;
            }
            if (lane == 0) {
#pragma unroll
                for (int r = 0; r < LRC; ++r) stats[warp * kLnRows + r] = make_float2(mw[r], qw[r]);
            }
            __syncthreads();
            // lane r combines row r over the eight warps (Chan, equal counts): mean = avg(mean_w),
            // M2 = sum M2_w + KW * sum (mean_w - mean)^2
            float mean_l = 0.f, rstd_l = 0.f;
            if (lane < LRC) {   // rows >= M hold copies of row M-1
                float2 sw[kWarps];
#pragma unroll
                for (int w = 0; w < kWarps; ++w) sw[w] = stats[w * kLnRows + lane];
                float mean = 0.f;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) mean += sw[w].x;
                mean *= (1.0f / kWarps);
                float m2 = 0.f, dev = 0.f;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) { m2 += sw[w].y; dev += (sw[w].x - mean) * (sw[w].x - mean); }
                m2 += KW * dev;
                mean_l = mean;
                rstd_l = rsqrtf(m2 / (kWarps * KW) + p.eps);
            }
#pragma unroll
            for (int r = 0; r < LRC; ++r) {
                const float mean = __shfl_sync(0xffffffffu, mean_l, r), rstd = __shfl_sync(0xffffffffu, rstd_l, r);
#pragma unroll
                for (int e = 0; e < KCH; ++e) xv[r][e] = (xv[r][e] - mean) * rstd * lg[e] + lb[e];
            }
        }
#pragma unroll
        for (int r = 0; r < LRC; ++r) {
            if (r < p.M) {
                const uint32_t dst = abase + r * pitch + (warp * KW + lane * KCH) * 2;
                if constexpr (KCH == 4)
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(dst), "r"(pack_bf16x2(xv[r][0], xv[r][1])), "r"(pack_bf16x2(xv[r][2], xv[r][3])) : "memory");
                else if constexpr (KCH == 2)
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(dst), "r"(pack_bf16x2(xv[r][0], xv[r][1])) : "memory");
                else
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(dst), "h"(__bfloat16_as_ushort(__float2bfloat16_rn(xv[r][0]))) : "memory");
            }
        }
        __syncwarp();
    } else {
        // ---- bf16 activations: every warp copies ITS k-slice of all batch rows (cp.async, L2 -> shared).  More than 16
        //      rows: one commit group per (k group, m16 tile), so that the MMAs of a tile start while the rest is still in
        //      flight.  Up to 16 rows the copy is latency-, not bandwidth-bound: one group, one wait, and the MMA loop
        //      below is free of barriers (the per-group waits kept ptxas from hoisting the shared-memory loads: 67 cycles
        //      per k chunk at M = 1).
        const int seg = lane % SEGS, r_in = lane / SEGS;
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int it = 0; it < 16 / RPI; ++it) {
                    const int row = mt * 16 + it * RPI + r_in;
                    if (row < p.M) {
                        const int e = gi * (JG * 32) + seg * 8;          // element offset inside the warp's k-slice
                        cp_async16(abase + row * pitch + (warp * KW + e) * 2, p.x + static_cast<int64_t>(row) * p.ldx + k0 + e);
                    }
                }
                if constexpr (PIPE) cp_async_commit();
            }
        }
        if constexpr (!PIPE) cp_async_commit();
    }
    // the residual rows are written by the predecessor chain: fetch them now, use them in the epilogue
    float* yp32 = static_cast<float*>(p.y) + blockIdx.y * p.split_stride + static_cast<int64_t>(om) * p.ldy + on;
    if (o_live && p.epi == VB_EPI_BIAS_RESIDUAL) {
        if (o_full) ores = *reinterpret_cast<const float4*>(yp32);
        else {
            ores.x = yp32[0];
            if (on + 1 < p.N) ores.y = yp32[1];
            if (on + 2 < p.N) ores.z = yp32[2];
            if (on + 3 < p.N) ores.w = yp32[3];
        }
    }

    // SETS independent accumulator sets: one accumulator per (mt, nt) would make the 2 KCH MMAs of a tile one dependent
    // chain (33 cycles each: 1 050 cycles for FFN2 with the whole K = 4096 at NT = 1)
    constexpr int SETS = (MT * NT >= 8) ? 1 : ((MT * NT >= 3) ? 2 : ((MT * NT == 2) ? 4 : 8));
    float acc[SETS][MT][NT][4];
#pragma unroll
    for (int s_ = 0; s_ < SETS; ++s_)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[s_][mt][nt][i] = 0.f;

    if constexpr (!LN && !PIPE) {
        cp_async_wait<0>();
        __syncwarp();
    }
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            if constexpr (!LN && PIPE) {
                // groups still allowed in flight after (gi, mt) has landed (resolved at compile time: full unroll)
                const int pending = NG * MT - 1 - (gi * MT + mt);
                switch (pending) {
                    case 0: cp_async_wait<0>(); break;
                    case 1: cp_async_wait<1>(); break;
                    case 2: cp_async_wait<2>(); break;
                    default: cp_async_wait<3>(); break;
                }
                __syncwarp();
            }
            if (gi == 0 && mt == 0) stamp(p, 3);
            // only rows < M exist in shared memory; the MMA rows above them read row M-1 again (their outputs are dropped)
            const uint32_t row_lo = abase + min(mt * 16 + g, p.M - 1) * pitch, row_hi = abase + min(mt * 16 + g + 8, p.M - 1) * pitch;
#pragma unroll
            for (int jj = 0; jj < JG; ++jj) {
                const int j = gi * JG + jj;
                const uint32_t koff = static_cast<uint32_t>(warp * KW + vec_off<KCH>(j, t)) * 2;
                const uint4 a_lo = lds128(row_lo + koff);
                const uint4 a_hi = lds128(row_hi + koff);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    mma_16816(acc[(2 * j) % SETS][mt][nt], a_lo.x, a_hi.x, a_lo.y, a_hi.y, wr[nt][j].x, wr[nt][j].y);
                    mma_16816(acc[(2 * j + 1) % SETS][mt][nt], a_lo.z, a_hi.z, a_lo.w, a_hi.w, wr[nt][j].z, wr[nt][j].w);
                }
            }
        }
    }

    // ---- sum the eight k-slices through shared memory (the activation buffer is dead once every warp is here)
    stamp(p, 4);
#pragma unroll
    for (int s_ = 1; s_ < SETS; ++s_)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[0][mt][nt][i] += acc[s_][mt][nt][i];
    __syncthreads();
    float* red = reinterpret_cast<float*>(buf);
    {
        float* mine = red + warp * (MT * 16 * kRedPitch);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                *reinterpret_cast<float2*>(mine + (mt * 16 + g) * kRedPitch + nt * 8 + 2 * t) = make_float2(acc[0][mt][nt][0], acc[0][mt][nt][1]);
                *reinterpret_cast<float2*>(mine + (mt * 16 + g + 8) * kRedPitch + nt * 8 + 2 * t) = make_float2(acc[0][mt][nt][2], acc[0][mt][nt][3]);
            }
    }
    __syncthreads();
    stamp(p, 5);
    if (o_live) {
        float4 part[kWarps];
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
            part[w] = *reinterpret_cast<const float4*>(red + w * (MT * 16 * kRedPitch) + om * kRedPitch + oc);
        float v[4] = {part[0].x, part[0].y, part[0].z, part[0].w};
#pragma unroll
        for (int w = 1; w < kWarps; ++w) { v[0] += part[w].x; v[1] += part[w].y; v[2] += part[w].z; v[3] += part[w].w; }
        v[0] += obias.x; v[1] += obias.y; v[2] += obias.z; v[3] += obias.w;       // zeros without an epilogue
        if (p.epi == VB_EPI_BIAS_GELU) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = gelu_erf_fast(v[e]);
        }
        v[0] += ores.x; v[1] += ores.y; v[2] += ores.z; v[3] += ores.w;           // zeros unless VB_EPI_BIAS_RESIDUAL
        if (p.y_bf16) {
            __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(p.y) + static_cast<int64_t>(om) * p.ldy + on;
            if (o_full) *reinterpret_cast<uint2*>(yp) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
            else
#pragma unroll
                for (int e = 0; e < 4; ++e) if (on + e < p.N) yp[e] = __float2bfloat16_rn(v[e]);
        } else {
            if (o_full) *reinterpret_cast<float4*>(yp32) = make_float4(v[0], v[1], v[2], v[3]);
            else
#pragma unroll
                for (int e = 0; e < 4; ++e) if (on + e < p.N) yp32[e] = v[e];
        }
    }
    stamp(p, 6);
}

unsigned long long* g_rows_dbg = nullptr;
constexpr int kMaxSmem = 100 * 1024;

// K per CTA: 8 warps x 32 KCH elements.  want_split >= 1: Kc in {1024, 512, 256}, as large as the split count allows.
// want_split == 0: the whole K in one CTA whenever the M activation rows fit in shared memory (K <= 4096).
int pick_kc(int64_t K, int want_split, int M) {
    if (K % 256 != 0) return 0;
    if (want_split == 0) {
        for (int kc : {4096, 2048}) {
            if (K == kc && static_cast<int64_t>(M > 0 ? M : 1) * (kc * 2 + 16) + kStatBytes <= kMaxSmem) return kc;
        }
        want_split = 1;
    }
    for (int kc : {1024, 512, 256}) {
        if (K % kc != 0) continue;
        if (K / kc >= want_split) return kc;
    }
    return 256;
}

template <int MT, int NT, int KCH, int LN>
int launch_rows(const RowsParams& p, int n_split, cudaStream_t st) {
    const int acts = p.M * (p.Kc * 2 + 16), red = kWarps * MT * 16 * kRedPitch * 4;
    const size_t smem = kStatBytes + static_cast<size_t>(acts > red ? acts : red);
    VB_REQUIRE(smem <= static_cast<size_t>(kMaxSmem), VB_ERR_UNSUPPORTED, "vb_linear_decode_rows: %zu bytes of shared memory", smem);
    static bool attr_done = false;      // per instantiation
    if (!attr_done) {
        VB_CUDA(cudaFuncSetAttribute(linear_decode_rows_kernel<MT, NT, KCH, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        attr_done = true;
    }
    const int n_tiles = static_cast<int>(vb_ceil_div(p.N, 8));
    const dim3 grid(static_cast<unsigned>(vb_ceil_div(n_tiles, NT)), static_cast<unsigned>(n_split));
    VB_CUDA(vb_launch(true, linear_decode_rows_kernel<MT, NT, KCH, LN>, grid, dim3(kThreads), smem, st, p));
    return VB_OK;
}

template <int MT, int NT, int LN>
int launch_rows_k(const RowsParams& p, int n_split, cudaStream_t st) {
    switch (p.Kc) {
        case 256: return launch_rows<MT, NT, 1, LN>(p, n_split, st);
        case 512: return launch_rows<MT, NT, 2, LN>(p, n_split, st);
        case 1024: return launch_rows<MT, NT, 4, LN>(p, n_split, st);
        default: break;
    }
    if constexpr (!LN && NT <= 2) {
        if (p.Kc == 2048) return launch_rows<MT, NT, 8, 0>(p, n_split, st);
    }
    if constexpr (!LN && NT == 1) {
        if (p.Kc == 4096) return launch_rows<MT, 1, 16, 0>(p, n_split, st);
    }
    vb_set_error("vb_linear_decode_rows: no kernel for Kc = %d with %d n8 tiles per CTA", p.Kc, NT);
    return VB_ERR_UNSUPPORTED;
}

template <int MT, int LN>
int launch_rows_n(const RowsParams& p, int nt, int n_split, cudaStream_t st) {
    switch (nt) {
        case 1: return launch_rows_k<MT, 1, LN>(p, n_split, st);
        case 2: return launch_rows_k<MT, 2, LN>(p, n_split, st);
        case 3: return launch_rows_k<MT, 3, LN>(p, n_split, st);
        default: return launch_rows_k<MT, 4, LN>(p, n_split, st);
    }
}

int n8_tiles_per_cta(int64_t N, int n_split, int kc) {
    const int64_t n_tiles = vb_ceil_div(N, 8);
    int nt = static_cast<int>(vb_ceil_div(n_tiles * n_split, vb_sm_count()));
    const int cap = kc > 2048 ? 1 : (kc > 1024 ? 2 : 4);     // weight registers: NT * Kc / 256 vectors of 16 bytes
    return nt < 1 ? 1 : (nt > cap ? cap : nt);
}

int fill_common(RowsParams& p, const void* w, int64_t ldw, const float* bias, void* y, int y_dtype, int64_t ldy,
                int64_t split_stride, int M, int64_t N, int kc, int epilogue, int flags) {
    VB_REQUIRE(w && y && M >= 0 && N > 0, VB_ERR_BAD_ARG, "vb_linear_decode_rows: bad args");
    VB_REQUIRE(ldw % 8 == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0, VB_ERR_BAD_ARG,
               "vb_linear_decode_rows: w must be 16-byte aligned with a pitch that is a multiple of 8");
    VB_REQUIRE(epilogue >= VB_EPI_NONE && epilogue <= VB_EPI_BIAS_RESIDUAL, VB_ERR_BAD_ARG, "vb_linear_decode_rows: bad epilogue");
    VB_REQUIRE(epilogue == VB_EPI_NONE || bias != nullptr, VB_ERR_BAD_ARG, "vb_linear_decode_rows: epilogue needs bias");
    VB_REQUIRE(y_dtype == VB_F32 || y_dtype == VB_BF16, VB_ERR_BAD_ARG, "vb_linear_decode_rows: bad y_dtype");
    VB_REQUIRE(epilogue != VB_EPI_BIAS_RESIDUAL || y_dtype == VB_F32, VB_ERR_BAD_ARG, "vb_linear_decode_rows: residual needs fp32 y");
    p.x = nullptr;
    p.x32 = nullptr;
    p.gamma = p.beta = nullptr;
    p.eps = 0.f;
    p.w = static_cast<const __nv_bfloat16*>(w);
    p.ldw = ldw;
    p.bias = bias;
    p.y = y;
    p.ldy = ldy;
    p.split_stride = split_stride;
    p.M = M;
    p.N = static_cast<int>(N);
    p.Kc = kc;
    p.epi = epilogue;
    p.y_bf16 = (y_dtype == VB_BF16);
    p.late_trigger = (flags & VB_FLAG_LATE_TRIGGER) ? 1 : 0;
    p.dbg = g_rows_dbg;
    const int esz = p.y_bf16 ? 2 : 4;
    p.vec_ok = (N % 4 == 0) && (ldy % 4 == 0) && (split_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(y) % (4 * esz)) == 0) &&
               (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0);
    return VB_OK;
}

}  // namespace

extern "C" int vb_linear_decode_rows_set_debug(void* buf) {   /* device buffer of grid * 16 uint64 stamps, or NULL */
    g_rows_dbg = static_cast<unsigned long long*>(buf);
    return VB_OK;
}

extern "C" int vb_linear_decode_rows_splits(int M, int64_t K, int want_split) {
    const int kc = pick_kc(K, want_split < 0 ? 1 : want_split, M);
    return kc == 0 ? 0 : static_cast<int>(K / kc);
}

extern "C" int vb_linear_decode_rows(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, void* y,
                                     int y_dtype, int64_t ldy, int64_t split_stride, int M, int64_t N, int64_t K, int epilogue,
                                     int want_split, int flags, int* n_split_out, void* stream) {
    VB_REQUIRE(x != nullptr && K > 0, VB_ERR_BAD_ARG, "vb_linear_decode_rows: bad args");
    VB_REQUIRE(M <= 16, VB_ERR_UNSUPPORTED, "vb_linear_decode_rows: M = %d > 16 (use vb_decode_gemm / vb_linear_decode)", M);
    VB_REQUIRE(K % 256 == 0, VB_ERR_UNSUPPORTED, "vb_linear_decode_rows: K = %lld is not a multiple of 256", (long long)K);
    VB_REQUIRE(ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, VB_ERR_BAD_ARG,
               "vb_linear_decode_rows: x must be 16-byte aligned with a pitch that is a multiple of 8");
    const int kc = pick_kc(K, want_split < 0 ? 1 : want_split, M);
    const int n_split = static_cast<int>(K / kc);
    VB_REQUIRE(n_split == 1 || (epilogue == VB_EPI_NONE && y_dtype == VB_F32), VB_ERR_BAD_ARG,
               "vb_linear_decode_rows: K = %lld needs %d slices; split-K output is fp32 slices without an epilogue", (long long)K, n_split);
    RowsParams p;
    const int rc = fill_common(p, w, ldw, bias, y, y_dtype, ldy, split_stride, M, N, kc, epilogue, flags);
    if (rc != VB_OK) return rc;
    if (n_split_out) *n_split_out = n_split;
    if (M == 0) return VB_OK;
    p.x = static_cast<const __nv_bfloat16*>(x);
    p.ldx = ldx;
    const int nt = n8_tiles_per_cta(N, n_split, kc);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return launch_rows_n<1, 0>(p, nt, n_split, st);   // one m16 tile; 17..32 rows (a second m16 tile) measured slower than the tcgen05 forms
}

extern "C" int vb_linear_decode_rows_ln(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                                        const void* w, int64_t ldw, const float* bias, void* y, int y_dtype, int64_t ldy, int M,
                                        int64_t N, int64_t K, int epilogue, int flags, void* stream) {
    VB_REQUIRE(x != nullptr, VB_ERR_BAD_ARG, "vb_linear_decode_rows_ln: bad args");
    VB_REQUIRE(M <= kLnRows, VB_ERR_UNSUPPORTED, "vb_linear_decode_rows_ln: M = %d > %d rows", M, kLnRows);
    VB_REQUIRE(K == 256 || K == 512 || K == 1024, VB_ERR_UNSUPPORTED, "vb_linear_decode_rows_ln: K = %lld (256, 512 or 1024: one CTA normalises whole rows)", (long long)K);
    VB_REQUIRE((gamma != nullptr) == (beta != nullptr), VB_ERR_BAD_ARG, "vb_linear_decode_rows_ln: gamma and beta must both be given (LayerNorm) or both be null (plain cast)");
    VB_REQUIRE(ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, VB_ERR_BAD_ARG, "vb_linear_decode_rows_ln: x must be 16-byte aligned, pitch % 4 == 0");
    RowsParams p;
    const int rc = fill_common(p, w, ldw, bias, y, y_dtype, ldy, 0, M, N, static_cast<int>(K), epilogue, flags);
    if (rc != VB_OK) return rc;
    if (M == 0) return VB_OK;
    p.x32 = x;
    p.ldx = ldx;
    p.gamma = gamma;
    p.beta = beta;
    p.eps = eps;
    const int nt = n8_tiles_per_cta(N, 1, static_cast<int>(K));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (M <= 1) return launch_rows_n<1, 1>(p, nt, 1, st);
    if (M <= 2) return launch_rows_n<1, 2>(p, nt, 1, st);
    if (M <= 4) return launch_rows_n<1, 4>(p, nt, 1, st);
    return launch_rows_n<1, 8>(p, nt, 1, st);
}
