// Decode-time attention over the paged KV pool (one new token per sequence) -- the HBM-bound kernel of the AR path.
//
//   pool (per layer): [page][2 (K,V)][H][64 tokens][64 dims]   -> for one (b,h) a page is one contiguous 8 KB (bf16)
//   chunk of K and one of V.  grid = (n_tsplit, H, B): flash-decoding split over the sequence; each CTA streams its
//   pages with 1D bulk-async copies (cp.async.bulk -> UBLKCP) into a 4-deep shared-memory ring guarded by mbarriers
//   (one producer warp, four consumer warps), so the bytes in flight do not cost registers.  Consumers read 16-byte
//   vectors (8 lanes cover one 128-byte K row: coalesced + bank-conflict free), reduce q.k with warp shuffles, keep an
//   online softmax per warp, and accumulate p.V in registers.  Warps merge through shared memory, splits merge in the
//   last-arriving CTA of each (b,h) (atomic ticket) -- no second launch.
//   The new token's q/k/v arrive as split-K partials of the QKV GEMM and are reduced here in fixed order; k/v are
//   appended to the pool by the CTA that owns the last split.
// Replaces modules.py:146-167 on the cached path (qkv head split, torch.cat cache growth, SDPA with one query).
#include <math.h>

#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int DH = 64;
constexpr int PAGE = 64;
constexpr int NSTAGE = 3;
constexpr int CONSUMER_WARPS = 4;
constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;

template <typename T> struct PoolTraits;
template <> struct PoolTraits<__nv_bfloat16> {
    static constexpr int LPT = 8;   // lanes per token row (128 B / 16 B)
    static constexpr int EPL = 8;   // elements per lane
};
template <> struct PoolTraits<float> {
    static constexpr int LPT = 16;
    static constexpr int EPL = 4;
};

template <typename T, int EPL> __device__ __forceinline__ void load_vec(const uint8_t* smem_ptr, float (&out)[EPL]) {
    const uint4 raw = *reinterpret_cast<const uint4*>(smem_ptr);
    if constexpr (sizeof(T) == 2) {
        out[0] = bf16_lo(raw.x); out[1] = bf16_hi(raw.x); out[2] = bf16_lo(raw.y); out[3] = bf16_hi(raw.y);
        out[4] = bf16_lo(raw.z); out[5] = bf16_hi(raw.z); out[6] = bf16_lo(raw.w); out[7] = bf16_hi(raw.w);
    } else {
        out[0] = __uint_as_float(raw.x); out[1] = __uint_as_float(raw.y);
        out[2] = __uint_as_float(raw.z); out[3] = __uint_as_float(raw.w);
    }
}

template <typename T, typename TO>
__global__ void __launch_bounds__(THREADS) attn_decode_kernel(const float* __restrict__ qkv_part, int n_part,
                                                              int64_t part_stride, T* __restrict__ pool,
                                                              const int32_t* __restrict__ block_table, int max_pages,
                                                              const int32_t* __restrict__ seq_lens, TO* __restrict__ o,
                                                              float* __restrict__ ws_o, float* __restrict__ ws_ml,
                                                              unsigned* __restrict__ counters, int H, int n_tsplit,
                                                              float scale_log2e, int prefetch_kv) {
    constexpr int LPT = PoolTraits<T>::LPT, EPL = PoolTraits<T>::EPL;
    constexpr int TPP = 32 / LPT;                       // tokens per warp pass
    constexpr int TOK_PER_WARP = PAGE / CONSUMER_WARPS; // 16
    constexpr int NPASS = TOK_PER_WARP / TPP;
    constexpr int ROW_BYTES = DH * sizeof(T);
    constexpr int CHUNK_BYTES = PAGE * ROW_BYTES;       // one (page, K or V, head) chunk
    constexpr int STAGE_BYTES = 2 * CHUNK_BYTES;

    extern __shared__ __align__(128) uint8_t ring[];
    __shared__ __align__(8) uint64_t full_bar[NSTAGE];
    __shared__ __align__(8) uint64_t empty_bar[NSTAGE];
    __shared__ float q_s[DH], k_s[DH], v_s[DH];
    __shared__ float w_acc[CONSUMER_WARPS][DH];
    __shared__ float w_ml[CONSUMER_WARPS][2];
    __shared__ int is_last;

    const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d_model = H * DH;
    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), CONSUMER_WARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int32_t* bt = block_table + static_cast<int64_t>(b) * max_pages;
    const bool owns_new = (split == n_tsplit - 1);

    if (warp == CONSUMER_WARPS) {
        // ---------------- producer: starts streaming KV pages as early as the caller allows ----------------
        if (lane == 0) {
            if (!prefetch_kv) pdl_wait();     // with VB_FLAG_PREFETCH_KV the cached pages are known to be final already
            const int n_old = seq_lens[b];
            const int pages_total = (n_old + PAGE - 1) / PAGE;
            const int pp = (pages_total + n_tsplit - 1) / n_tsplit;
            const int p0 = min(split * pp, pages_total), p1 = min(p0 + pp, pages_total);
            int stage = 0;
            uint32_t phase = 0;
            for (int p = p0; p < p1; ++p) {
                mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                const uint32_t fb = smem_u32(&full_bar[stage]);
                mbar_expect_tx(fb, STAGE_BYTES);
                const int page = bt[p];
                const T* ksrc = pool + (static_cast<int64_t>(page) * 2 * H + h) * PAGE * DH;
                const uint32_t dst = smem_u32(ring) + stage * STAGE_BYTES;
                bulk_load_1d(dst, ksrc, CHUNK_BYTES, fb);
                bulk_load_1d(dst + CHUNK_BYTES, ksrc + static_cast<int64_t>(H) * PAGE * DH, CHUNK_BYTES, fb);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            if (prefetch_kv) pdl_wait();      // every thread of the CTA passes the dependency before it exits
        }
    }

    int n_old = 0, p0 = 0, p1 = 0;
    if (warp < CONSUMER_WARPS) {
        pdl_wait();     // q/k/v partials of the new token come from the predecessor (QKV GEMM)
        n_old = seq_lens[b];
        const int pages_total = (n_old + PAGE - 1) / PAGE;
        const int pp = (pages_total + n_tsplit - 1) / n_tsplit;
        p0 = min(split * pp, pages_total);
        p1 = min(p0 + pp, pages_total);
        // new-token q (and k, v for the owner of the last split): fixed-order reduction of the split-K partials
        if (threadIdx.x < DH) {
            const int e = threadIdx.x;
            const float* src = qkv_part + static_cast<int64_t>(b) * 3 * d_model + h * DH + e;
            float qv = 0.f, kv = 0.f, vv = 0.f;
            for (int s = 0; s < n_part; ++s) {
                qv += src[s * part_stride];
                if (owns_new) { kv += src[s * part_stride + d_model]; vv += src[s * part_stride + 2 * d_model]; }
            }
            q_s[e] = qv * scale_log2e;
            if (owns_new) {
                const T kq = from_f32<T>(kv), vq = from_f32<T>(vv);   // the cache precision is what later steps will read
                k_s[e] = to_f32<T>(kq);
                v_s[e] = to_f32<T>(vq);
                const int page = bt[n_old / PAGE], slot = n_old % PAGE;
                T* kdst = pool + ((static_cast<int64_t>(page) * 2 * H + h) * PAGE + slot) * DH + (sizeof(T) == 2 ? vb_pool_col_bf16(slot, e) : e);
                kdst[0] = kq;
                kdst[static_cast<int64_t>(H) * PAGE * DH] = vq;
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32) : "memory");   // consumer warps only
    }

    float m_run = -INFINITY, l_run = 0.f;
    float acc[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) acc[i] = 0.f;

    if (warp < CONSUMER_WARPS) {
        const int grp = lane / LPT, c = lane % LPT;
        float qreg[EPL];
#pragma unroll
        for (int i = 0; i < EPL; ++i) qreg[i] = q_s[c * EPL + i];
        int stage = 0;
        uint32_t phase = 0;
        for (int p = p0; p < p1; ++p) {
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            const uint8_t* ks = ring + stage * STAGE_BYTES;
            const uint8_t* vs = ks + CHUNK_BYTES;
            float sc[NPASS];
            float pmax = -INFINITY;
#pragma unroll
            for (int ps = 0; ps < NPASS; ++ps) {
                const int tok = warp * TOK_PER_WARP + ps * TPP + grp;
                float kv[EPL];
                load_vec<T, EPL>(ks + tok * ROW_BYTES + ((sizeof(T) == 2 ? (c ^ (tok & 7)) : c) << 4), kv);
                float d = 0.f;
#pragma unroll
                for (int i = 0; i < EPL; ++i) d = fmaf(qreg[i], kv[i], d);
#pragma unroll
                for (int off = 1; off < LPT; off <<= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
                const bool valid = (p * PAGE + tok) < n_old;
                sc[ps] = valid ? d : -INFINITY;
                pmax = fmaxf(pmax, sc[ps]);
            }
            pmax = warp_max(pmax);
            const float m_new = fmaxf(m_run, pmax);
            if (m_new != -INFINITY) {
                const float corr = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
                l_run *= corr;
#pragma unroll
                for (int i = 0; i < EPL; ++i) acc[i] *= corr;
#pragma unroll
                for (int ps = 0; ps < NPASS; ++ps) {
                    const int tok = warp * TOK_PER_WARP + ps * TPP + grp;
                    const float pj = (sc[ps] == -INFINITY) ? 0.f : exp2f(sc[ps] - m_new);
                    l_run += pj;
                    float vv[EPL];
                    load_vec<T, EPL>(vs + tok * ROW_BYTES + ((sizeof(T) == 2 ? (c ^ (tok & 7)) : c) << 4), vv);
#pragma unroll
                    for (int i = 0; i < EPL; ++i) acc[i] = fmaf(pj, vv[i], acc[i]);
                }
                m_run = m_new;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&empty_bar[stage]));
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        // the new token (position n_old), handled by lane-group 0 of warp 0 of the last split
        if (owns_new && warp == 0) {
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < EPL; ++i) d = fmaf(qreg[i], k_s[c * EPL + i], d);
#pragma unroll
            for (int off = 1; off < LPT; off <<= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
            const float s_new = __shfl_sync(0xffffffffu, d, 0);
            const float m_new = fmaxf(m_run, s_new);
            const float corr = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
            l_run *= corr;
#pragma unroll
            for (int i = 0; i < EPL; ++i) acc[i] *= corr;
            if (grp == 0) {
                const float pj = exp2f(s_new - m_new);
                l_run += pj;
#pragma unroll
                for (int i = 0; i < EPL; ++i) acc[i] = fmaf(pj, v_s[c * EPL + i], acc[i]);
            }
            m_run = m_new;
        }
        // merge the lane groups of the warp (all lanes share m_run); l is replicated across the LPT lanes of a group
#pragma unroll
        for (int off = LPT; off < 32; off <<= 1) {
            l_run += __shfl_xor_sync(0xffffffffu, l_run, off);
#pragma unroll
            for (int i = 0; i < EPL; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
        }
        if (grp == 0) {
#pragma unroll
            for (int i = 0; i < EPL; ++i) w_acc[warp][c * EPL + i] = acc[i];
        }
        if (lane == 0) { w_ml[warp][0] = m_run; w_ml[warp][1] = l_run; }
    }
    __syncthreads();

    // merge the consumer warps: threads 0..63 own one output dim each
    float out_acc = 0.f, out_l = 0.f, out_m = -INFINITY;
    if (threadIdx.x < DH) {
#pragma unroll
        for (int w = 0; w < CONSUMER_WARPS; ++w) out_m = fmaxf(out_m, w_ml[w][0]);
#pragma unroll
        for (int w = 0; w < CONSUMER_WARPS; ++w) {
            const float wt = (w_ml[w][0] == -INFINITY) ? 0.f : exp2f(w_ml[w][0] - out_m);
            out_acc = fmaf(w_acc[w][threadIdx.x], wt, out_acc);
            out_l = fmaf(w_ml[w][1], wt, out_l);
        }
    }
    TO* orow = o + static_cast<int64_t>(b) * d_model + h * DH;
    if (n_tsplit == 1) {
        if (threadIdx.x < DH) orow[threadIdx.x] = from_f32<TO>(out_l > 0.f ? out_acc / out_l : 0.f);
        return;
    }
    const int64_t slot = (static_cast<int64_t>(b) * H + h) * n_tsplit;
    if (threadIdx.x < DH) {
        ws_o[(slot + split) * DH + threadIdx.x] = out_acc;
        if (threadIdx.x == 0) { ws_ml[(slot + split) * 2] = out_m; ws_ml[(slot + split) * 2 + 1] = out_l; }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicInc(&counters[b * H + h], static_cast<unsigned>(n_tsplit - 1));
        is_last = (ticket == static_cast<unsigned>(n_tsplit - 1));
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < DH) {
        float M = -INFINITY;
        for (int s = 0; s < n_tsplit; ++s) M = fmaxf(M, __ldcg(&ws_ml[(slot + s) * 2]));
        float a = 0.f, l = 0.f;
        for (int s = 0; s < n_tsplit; ++s) {
            const float ms = __ldcg(&ws_ml[(slot + s) * 2]);
            const float wt = (ms == -INFINITY) ? 0.f : exp2f(ms - M);
            a = fmaf(__ldcg(&ws_o[(slot + s) * DH + threadIdx.x]), wt, a);
            l = fmaf(__ldcg(&ws_ml[(slot + s) * 2 + 1]), wt, l);
        }
        orow[threadIdx.x] = from_f32<TO>(l > 0.f ? a / l : 0.f);
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Tensor-core variant for the bf16 pool.  The SIMT kernel above spends ~1200 warp-instructions per 16 KB page pair and
// its consumers, not HBM, set the pace (ncu: 19 % of the warp samples wait for data, the rest is shuffle/FMA chains).
// Here ONE warp owns a whole page: S = q.K^T and O += P.V are mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with the
// single query in row 0 of the 16-row A operand, K and V fragments come from ldmatrix on the XOR-swizzled page rows
// (conflict-free), the softmax runs on the four lanes that hold row 0.  ~200 warp-instructions per page, and up to
// NSTAGE pages of a CTA are in different warps at the same time, so ring slots turn over as fast as the copies land.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// D(16x8) += A(16x16) . B(16x8); rows 8..15 of A are zero here (a1 = a3 = 0)
__device__ __forceinline__ void mma_row0(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}

constexpr int MAX_PG_SMEM = 64;     // page ids of a split kept in shared memory (longer splits read the table directly)
// One consumer warp per ring stage: warp w only ever waits on stage w, phase after phase (a warp that could run ahead to
// a later phase of another stage would alias the mbarrier parity).
constexpr int MMA_WARPS = NSTAGE;
constexpr int MMA_THREADS = (MMA_WARPS + 1) * 32;

template <typename TO>
__global__ void __launch_bounds__(MMA_THREADS, 4) attn_decode_mma_kernel(const float* __restrict__ qkv_part, int n_part,
                                                                  int64_t part_stride, __nv_bfloat16* __restrict__ pool,
                                                                  const int32_t* __restrict__ block_table, int max_pages,
                                                                  const int32_t* __restrict__ seq_lens, TO* __restrict__ o,
                                                                  float* __restrict__ ws_o, float* __restrict__ ws_ml,
                                                                  unsigned* __restrict__ counters, int H, int n_tsplit,
                                                                  float scale_log2e, int prefetch_kv,
                                                                  unsigned long long* __restrict__ dbg) {
    typedef __nv_bfloat16 T;
    constexpr int CHUNK_BYTES = PAGE * DH * 2;          // 8 KB: one (page, K or V, head) chunk
    // optional timeline (vb_attn_decode_set_debug): %globaltimer stamps [cta][8] = {start, producer issued its copies,
    // consumers' dependency resolved, q ready, first page landed, pages done, partial written, output written}
#define ATT_STAMP(ev)                                                                                          \
    do {                                                                                                       \
        if (dbg != nullptr) {                                                                                  \
            unsigned long long t_;                                                                             \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                             \
            dbg[((static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (ev)] = t_; \
        }                                                                                                      \
    } while (0)
    constexpr int STAGE_BYTES = 2 * CHUNK_BYTES;

    extern __shared__ __align__(128) uint8_t ring[];
    __shared__ __align__(8) uint64_t full_bar[NSTAGE];
    __shared__ __align__(8) uint64_t empty_bar[NSTAGE];
    __shared__ float q_s[DH], k_s[DH], v_s[DH];
    __shared__ float w_acc[MMA_WARPS][DH];
    __shared__ float w_ml[MMA_WARPS][2];
    __shared__ int pg_s[MAX_PG_SMEM];
    __shared__ int is_last;

    const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d_model = H * DH;
    // prefetch_kv bit 1 (VB_FLAG_LATE_TRIGGER): release the successor only after this kernel's own dependency has resolved.
    // With the early trigger the out-proj / FFN1 CTAs of the same layer become resident (spinning in griddepcontrol.wait)
    // while this kernel's own CTAs are still waiting for registers: at batch 1 some clusters started 2 us late and their
    // pages landed 3 us after the median (tools/step_timeline.py).
    const bool late_trigger = (prefetch_kv & 2) != 0;
    prefetch_kv &= 1;
    if (!late_trigger) pdl_trigger();
    if (threadIdx.x == 0) {
        ATT_STAMP(0);
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);       // one warp consumes a whole stage
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int32_t* bt = block_table + static_cast<int64_t>(b) * max_pages;
    const bool owns_new = (split == n_tsplit - 1);

    if (warp == MMA_WARPS) {
        // ---------------- producer warp: page ids first (one coalesced read), then the bulk copies ----------------
        if (!prefetch_kv) pdl_wait();     // with VB_FLAG_PREFETCH_KV the cached pages are known to be final already
        const int n_old = seq_lens[b];
        const int pages_total = (n_old + PAGE - 1) / PAGE;
        // pages spread evenly over the splits (12 pages over 8 splits: 2,1,2,1,... instead of 2,2,2,2,2,2,0,0)
        const int p0 = (split * pages_total) / n_tsplit, p1 = ((split + 1) * pages_total) / n_tsplit;
        for (int i = lane; i < min(p1 - p0, MAX_PG_SMEM); i += 32) pg_s[i] = bt[p0 + i];
        __syncwarp();
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int p = p0; p < p1; ++p) {
                const int page = (p - p0 < MAX_PG_SMEM) ? pg_s[p - p0] : bt[p];
                mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                const uint32_t fb = smem_u32(&full_bar[stage]);
                mbar_expect_tx(fb, STAGE_BYTES);
                const T* ksrc = pool + (static_cast<int64_t>(page) * 2 * H + h) * PAGE * DH;
                const uint32_t dst = smem_u32(ring) + stage * STAGE_BYTES;
                bulk_load_1d(dst, ksrc, CHUNK_BYTES, fb);
                bulk_load_1d(dst + CHUNK_BYTES, ksrc + static_cast<int64_t>(H) * PAGE * DH, CHUNK_BYTES, fb);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            ATT_STAMP(1);
        }
        __syncwarp();
        if (prefetch_kv) pdl_wait();      // every thread of the CTA passes the dependency before it exits
    }

    float m_run = -INFINITY, l_run = 0.f;
    float oacc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) oacc[j][i] = 0.f;

    if (warp < MMA_WARPS) {
        // Under VB_FLAG_PREFETCH_KV seq_lens and the page table are final before this kernel starts: fetch the sequence
        // length and the page that receives the new token BEFORE the dependency wait (two dependent L2 round trips that
        // otherwise sit between the wait and the q/k/v reduction).
        int n_old = 0, new_page = 0;
        if (prefetch_kv) {
            n_old = seq_lens[b];
            if (owns_new) new_page = bt[n_old / PAGE];
        }
        pdl_wait();     // q/k/v partials of the new token come from the predecessor (QKV GEMM)
        if (late_trigger) pdl_trigger();
        if (threadIdx.x == 0) ATT_STAMP(2);
        if (!prefetch_kv) {
            n_old = seq_lens[b];
            if (owns_new) new_page = bt[n_old / PAGE];
        }
        const int pages_total = (n_old + PAGE - 1) / PAGE;
        // pages spread evenly over the splits (12 pages over 8 splits: 2,1,2,1,... instead of 2,2,2,2,2,2,0,0)
        const int p0 = (split * pages_total) / n_tsplit, p1 = ((split + 1) * pages_total) / n_tsplit;
        // new-token q, k, v (k, v only in the owner of the last split): threads 0..63 own one dim each; fixed-order reduction
        // of the split-K partials with every load issued before the first add
        if (threadIdx.x < DH) {
            const int e = threadIdx.x;
            const float* src = qkv_part + static_cast<int64_t>(b) * 3 * d_model + h * DH + e;
            float tq[8], tk[8], tv[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                tq[s] = (s < n_part) ? __ldcg(src + s * part_stride) : 0.f;
                tk[s] = (owns_new && s < n_part) ? __ldcg(src + s * part_stride + d_model) : 0.f;
                tv[s] = (owns_new && s < n_part) ? __ldcg(src + s * part_stride + 2 * d_model) : 0.f;
            }
            // consume all 24 registers at once: keeps ptxas from interleaving load and add (a chain of dependent L2 round
            // trips: q was ready 1.5-2.0 us after the dependency resolved in a B=32 step with 6 slices)
            asm volatile("" ::"f"(tq[0]), "f"(tq[1]), "f"(tq[2]), "f"(tq[3]), "f"(tq[4]), "f"(tq[5]), "f"(tq[6]), "f"(tq[7]),
                         "f"(tk[0]), "f"(tk[1]), "f"(tk[2]), "f"(tk[3]), "f"(tk[4]), "f"(tk[5]), "f"(tk[6]), "f"(tk[7]),
                         "f"(tv[0]), "f"(tv[1]), "f"(tv[2]), "f"(tv[3]), "f"(tv[4]), "f"(tv[5]), "f"(tv[6]), "f"(tv[7]));
            float aq = 0.f, ak = 0.f, av = 0.f;
#pragma unroll
            for (int s = 0; s < 8; ++s) { aq += tq[s]; ak += tk[s]; av += tv[s]; }
            for (int s = 8; s < n_part; ++s) {
                aq += __ldcg(src + s * part_stride);
                if (owns_new) { ak += __ldcg(src + s * part_stride + d_model); av += __ldcg(src + s * part_stride + 2 * d_model); }
            }
            q_s[e] = aq * scale_log2e;
            if (owns_new) {
                const T kq = __float2bfloat16_rn(ak), vq = __float2bfloat16_rn(av);   // the cache precision is what later steps read
                k_s[e] = __bfloat162float(kq);
                v_s[e] = __bfloat162float(vq);
                const int page = new_page, slot = n_old % PAGE;
                T* kdst = pool + ((static_cast<int64_t>(page) * 2 * H + h) * PAGE + slot) * DH + vb_pool_col_bf16(slot, e);
                kdst[0] = kq;
                kdst[static_cast<int64_t>(H) * PAGE * DH] = vq;
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(MMA_WARPS * 32) : "memory");   // consumer warps only
        if (threadIdx.x == 0) ATT_STAMP(3);

        const int g = lane >> 2, t = lane & 3;
        // A fragments of q (row 0 of the 16-row tile): k-step kk covers dims 16kk .. 16kk+15
        uint32_t qa0[4], qa2[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            qa0[kk] = (g == 0) ? pack_bf16x2(q_s[16 * kk + 2 * t], q_s[16 * kk + 2 * t + 1]) : 0u;
            qa2[kk] = (g == 0) ? pack_bf16x2(q_s[16 * kk + 8 + 2 * t], q_s[16 * kk + 9 + 2 * t]) : 0u;
        }
        const int lm = lane >> 3, lr = lane & 7;      // ldmatrix: lane supplies row lr of matrix lm
        for (int ip = warp; ip < p1 - p0; ip += MMA_WARPS) {
            const int stage = ip % NSTAGE;
            const uint32_t phase = (ip / NSTAGE) & 1;
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            if (threadIdx.x == 0 && ip == 0) ATT_STAMP(4);
            const uint32_t kbase = smem_u32(ring) + stage * STAGE_BYTES;
            const uint32_t vbase = kbase + CHUNK_BYTES;
            const int tok0 = (p0 + ip) * PAGE;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {          // two half pages of 32 tokens: halves the live accumulator registers
                // ---- S = q . K^T : 4 token tiles x 4 dim steps ----
                float sacc[4][4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = 4 * hf + jj;
                    sacc[jj][0] = sacc[jj][1] = sacc[jj][2] = sacc[jj][3] = 0.f;
#pragma unroll
                    for (int kp = 0; kp < 2; ++kp) {  // two k-steps per ldmatrix.x4: dim chunks 4kp .. 4kp+3
                        uint32_t r0, r1, r2, r3;
                        ldsm_x4(kbase + (8 * j + lr) * 128 + (((4 * kp + lm) ^ lr) << 4), r0, r1, r2, r3);
                        mma_row0(sacc[jj], qa0[2 * kp], qa2[2 * kp], r0, r1);
                        mma_row0(sacc[jj], qa0[2 * kp + 1], qa2[2 * kp + 1], r2, r3);
                    }
                }
                // ---- online softmax on row 0 (lanes 0..3 hold it; the other lanes carry zeros through the same code) ----
                float pmax = -INFINITY;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int tk = tok0 + 8 * (4 * hf + jj) + 2 * t;
                    if (tk >= n_old) sacc[jj][0] = -INFINITY;
                    if (tk + 1 >= n_old) sacc[jj][1] = -INFINITY;
                    pmax = fmaxf(pmax, fmaxf(sacc[jj][0], sacc[jj][1]));
                }
                pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, 1));
                pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, 2));
                const float m_new = fmaxf(m_run, pmax);          // finite for row 0: token 0 of a page is always cached
                const float corr = exp2f(m_run - m_new);
                l_run *= corr;
                uint32_t pa[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float p0f = exp2f(sacc[jj][0] - m_new), p1f = exp2f(sacc[jj][1] - m_new);
                    l_run += p0f + p1f;
                    pa[jj] = pack_bf16x2(p0f, p1f);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    oacc[j][0] *= corr;
                    oacc[j][1] *= corr;
                }
                m_run = m_new;
                // ---- O += P . V : 2 token steps x 8 dim tiles ----
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2) {
                    const int kk = 2 * hf + k2;
#pragma unroll
                    for (int jp = 0; jp < 4; ++jp) {  // two dim tiles per ldmatrix.x4.trans
                        uint32_t r0, r1, r2, r3;
                        const int tok = 16 * kk + (lm & 1) * 8 + lr;
                        ldsm_x4_t(vbase + tok * 128 + (((2 * jp + (lm >> 1)) ^ lr) << 4), r0, r1, r2, r3);
                        mma_row0(oacc[2 * jp], pa[2 * k2], pa[2 * k2 + 1], r0, r1);
                        mma_row0(oacc[2 * jp + 1], pa[2 * k2], pa[2 * k2 + 1], r2, r3);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&empty_bar[stage]));
        }
        if (threadIdx.x == 0) ATT_STAMP(5);
        // row 0 lives in lanes 0..3: dims 8j + 2t, 8j + 2t + 1; l is a per-lane partial sum
        l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
        l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
        if (g == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                w_acc[warp][8 * j + 2 * t] = oacc[j][0];
                w_acc[warp][8 * j + 2 * t + 1] = oacc[j][1];
            }
        }
        if (lane == 0) { w_ml[warp][0] = m_run; w_ml[warp][1] = l_run; }
    }
    __syncthreads();

    // merge the consumer warps and (owner of the last split) the new token: threads 0..63 own one output dim each
    float out_acc = 0.f, out_l = 0.f, out_m = -INFINITY;
    if (threadIdx.x < DH) {
        float s_new = -INFINITY;
        if (owns_new) {
            s_new = 0.f;
#pragma unroll 8
            for (int e = 0; e < DH; ++e) s_new = fmaf(q_s[e], k_s[e], s_new);
        }
        out_m = s_new;
#pragma unroll
        for (int w = 0; w < MMA_WARPS; ++w) out_m = fmaxf(out_m, w_ml[w][0]);
#pragma unroll
        for (int w = 0; w < MMA_WARPS; ++w) {
            const float wt = (w_ml[w][0] == -INFINITY) ? 0.f : exp2f(w_ml[w][0] - out_m);
            out_acc = fmaf(w_acc[w][threadIdx.x], wt, out_acc);
            out_l = fmaf(w_ml[w][1], wt, out_l);
        }
        if (owns_new) {
            const float wt = exp2f(s_new - out_m);
            out_acc = fmaf(v_s[threadIdx.x], wt, out_acc);
            out_l += wt;
        }
    }
    TO* orow = o + static_cast<int64_t>(b) * d_model + h * DH;
    if (n_tsplit == 1) {
        if (threadIdx.x < DH) orow[threadIdx.x] = from_f32<TO>(out_l > 0.f ? out_acc / out_l : 0.f);
        if (threadIdx.x == 0) ATT_STAMP(7);
        return;
    }
    if (counters == nullptr) {
        // Launched as a thread-block cluster of n_tsplit CTAs along x (one cluster per (b, h)): the splits are merged in
        // CTA 0 through distributed shared memory -- one remote store + one cluster barrier instead of store, fence, ticket
        // atomic, fence and an L2 round trip for the partials (1.5-3 us of a 9 us launch at batch 1).  Same order and
        // arithmetic as the ticket path: bit-identical results.
        __shared__ float c_o[8][DH];
        __shared__ __align__(8) float c_ml[8][2];
        if (threadIdx.x < DH) {
            uint32_t ra;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(&c_o[split][threadIdx.x])), "r"(0));
            asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(out_acc) : "memory");
            if (threadIdx.x == 0) {
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(&c_ml[split][0])), "r"(0));
                asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(ra), "f"(out_m), "f"(out_l) : "memory");
            }
        }
        if (threadIdx.x == 0) ATT_STAMP(6);
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        if (split != 0) return;
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        if (threadIdx.x < DH) {
            float M = -INFINITY;
            for (int s = 0; s < n_tsplit; ++s) M = fmaxf(M, c_ml[s][0]);
            float a = 0.f, l = 0.f;
            for (int s = 0; s < n_tsplit; ++s) {
                const float wt = (c_ml[s][0] == -INFINITY) ? 0.f : exp2f(c_ml[s][0] - M);
                a = fmaf(c_o[s][threadIdx.x], wt, a);
                l = fmaf(c_ml[s][1], wt, l);
            }
            orow[threadIdx.x] = from_f32<TO>(l > 0.f ? a / l : 0.f);
            if (threadIdx.x == 0) ATT_STAMP(7);
        }
        return;
    }
    const int64_t slot = (static_cast<int64_t>(b) * H + h) * n_tsplit;
    if (threadIdx.x < DH) {
        ws_o[(slot + split) * DH + threadIdx.x] = out_acc;
        if (threadIdx.x == 0) { ws_ml[(slot + split) * 2] = out_m; ws_ml[(slot + split) * 2 + 1] = out_l; }
    }
    if (threadIdx.x == 0) ATT_STAMP(6);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicInc(&counters[b * H + h], static_cast<unsigned>(n_tsplit - 1));
        is_last = (ticket == static_cast<unsigned>(n_tsplit - 1));
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < DH) {
        // every partial is requested before the first one is used: one L2 round trip instead of n_tsplit dependent ones
        // (the rolled loop cost 5 us of the 9.5 us this kernel took in a batch-1 step -- tools/step_timeline.py)
        constexpr int MAXS = 8;
        float ms[MAXS], ls[MAXS], os[MAXS];
#pragma unroll
        for (int s = 0; s < MAXS; ++s) {
            const bool on = s < n_tsplit;
            ms[s] = on ? __ldcg(&ws_ml[(slot + s) * 2]) : -INFINITY;
            ls[s] = on ? __ldcg(&ws_ml[(slot + s) * 2 + 1]) : 0.f;
            os[s] = on ? __ldcg(&ws_o[(slot + s) * DH + threadIdx.x]) : 0.f;
        }
        float M = -INFINITY;
#pragma unroll
        for (int s = 0; s < MAXS; ++s) M = fmaxf(M, ms[s]);
        for (int s = MAXS; s < n_tsplit; ++s) M = fmaxf(M, __ldcg(&ws_ml[(slot + s) * 2]));
        float a = 0.f, l = 0.f;
#pragma unroll
        for (int s = 0; s < MAXS; ++s) {        // same order and arithmetic as the rolled loop: results are bit-identical
            const float wt = (ms[s] == -INFINITY) ? 0.f : exp2f(ms[s] - M);
            a = fmaf(os[s], wt, a);
            l = fmaf(ls[s], wt, l);
        }
        for (int s = MAXS; s < n_tsplit; ++s) {
            const float m2 = __ldcg(&ws_ml[(slot + s) * 2]);
            const float wt = (m2 == -INFINITY) ? 0.f : exp2f(m2 - M);
            a = fmaf(__ldcg(&ws_o[(slot + s) * DH + threadIdx.x]), wt, a);
            l = fmaf(__ldcg(&ws_ml[(slot + s) * 2 + 1]), wt, l);
        }
        orow[threadIdx.x] = from_f32<TO>(l > 0.f ? a / l : 0.f);
        if (threadIdx.x == 0) ATT_STAMP(7);
    }
}

#undef ATT_STAMP
}  // namespace

static unsigned long long* g_attn_dbg = nullptr;
static bool attn_cluster_enabled() {     // VALLE_B200_ATTN_CLUSTER=0: merge the splits through the global ticket instead
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VALLE_B200_ATTN_CLUSTER");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}
extern "C" int vb_attn_decode_set_debug(void* buf) {   /* device buffer of grid * 8 uint64 stamps, or NULL */
    g_attn_dbg = static_cast<unsigned long long*>(buf);
    return VB_OK;
}

extern "C" int64_t vb_attn_decode_ws_bytes(int B, int H, int n_tsplit) {
    const int64_t bh = static_cast<int64_t>(B) * H;
    // [counters: bh u32 (rounded to 256 B)] [ws_ml: bh*ns*2 f32] [ws_o: bh*ns*64 f32]
    return ((bh * 4 + 255) / 256) * 256 + bh * n_tsplit * 2 * 4 + bh * n_tsplit * DH * 4;
}

extern "C" int vb_attn_decode_paged(const float* qkv_part, int n_part, int64_t part_stride, void* pool, int pool_dtype,
                                    const int32_t* block_table, int max_pages, const int32_t* seq_lens, void* o,
                                    int o_dtype, int B, int H, int Dh, int n_tsplit, int flags, void* ws, void* stream) {
    VB_REQUIRE(qkv_part && pool && block_table && seq_lens && o, VB_ERR_BAD_ARG, "vb_attn_decode_paged: null pointer");
    VB_REQUIRE(Dh == DH, VB_ERR_UNSUPPORTED, "vb_attn_decode_paged: head_dim must be 64 (got %d)", Dh);
    VB_REQUIRE(B >= 1 && B <= 65535 && H >= 1 && H <= 65535 && n_part >= 1 && n_tsplit >= 1, VB_ERR_BAD_ARG,
               "vb_attn_decode_paged: bad shape");
    VB_REQUIRE(n_tsplit == 1 || ws != nullptr, VB_ERR_BAD_ARG, "vb_attn_decode_paged: workspace required for n_tsplit > 1");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t bh = static_cast<int64_t>(B) * H;
    unsigned* counters = static_cast<unsigned*>(ws);   // must be zero-initialised once; self-resetting afterwards
    float* ws_ml = ws ? reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + ((bh * 4 + 255) / 256) * 256) : nullptr;
    float* ws_o = ws ? ws_ml + bh * n_tsplit * 2 : nullptr;
    const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(DH));
    dim3 grid(n_tsplit, H, B);
#define DEC(T, TO, SMEM)                                                                                              \
    {                                                                                                                 \
        auto kern = attn_decode_kernel<T, TO>;                                                                        \
        static bool configured = false;                                                                               \
        if (!configured) {                                                                                            \
            VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));                   \
            configured = true;                                                                                        \
        }                                                                                                             \
        VB_CUDA(vb_launch(true, kern, grid, dim3(THREADS), SMEM, st, qkv_part, n_part, part_stride, static_cast<T*>(pool),   \
                          block_table, max_pages, seq_lens, static_cast<TO*>(o), ws_o, ws_ml, counters, H, n_tsplit,  \
                          scale_log2e, (flags & VB_FLAG_PREFETCH_KV) ? 1 : 0));                                       \
    }
#define DECM(TO, SMEM)                                                                                                \
    {                                                                                                                 \
        auto kern = attn_decode_mma_kernel<TO>;                                                                       \
        static bool configured = false;                                                                               \
        if (!configured) {                                                                                            \
            VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));                   \
            configured = true;                                                                                        \
        }                                                                                                             \
        if (n_tsplit > 1 && n_tsplit <= 8 && attn_cluster_enabled() && !(flags & VB_FLAG_ATTN_TICKET)) {                                                \
            /* one cluster per (b, h): splits merged through DSMEM, counters == nullptr selects that path */          \
            cudaLaunchConfig_t cfg = {};                                                                              \
            cfg.gridDim = grid; cfg.blockDim = dim3(MMA_THREADS); cfg.dynamicSmemBytes = SMEM; cfg.stream = st;       \
            cudaLaunchAttribute attr[2];                                                                              \
            int na = 0;                                                                                               \
            if (vb_pdl_enabled()) {                                                                                   \
                attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                     \
                attr[na].val.programmaticStreamSerializationAllowed = 1;                                              \
                ++na;                                                                                                 \
            }                                                                                                         \
            attr[na].id = cudaLaunchAttributeClusterDimension;                                                        \
            attr[na].val.clusterDim.x = n_tsplit; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;       \
            ++na;                                                                                                     \
            cfg.attrs = attr; cfg.numAttrs = na;                                                                      \
            VB_CUDA(cudaLaunchKernelEx(&cfg, kern, qkv_part, n_part, part_stride, static_cast<__nv_bfloat16*>(pool),  \
                                       block_table, max_pages, seq_lens, static_cast<TO*>(o), ws_o, ws_ml,            \
                                       static_cast<unsigned*>(nullptr), H, n_tsplit, scale_log2e,                     \
                                       ((flags & VB_FLAG_PREFETCH_KV) ? 1 : 0) | ((flags & VB_FLAG_LATE_TRIGGER) ? 2 : 0), g_attn_dbg)); \
        } else {                                                                                                      \
            VB_CUDA(vb_launch(true, kern, grid, dim3(MMA_THREADS), SMEM, st, qkv_part, n_part, part_stride,           \
                              static_cast<__nv_bfloat16*>(pool), block_table, max_pages, seq_lens, static_cast<TO*>(o), \
                              ws_o, ws_ml, counters, H, n_tsplit, scale_log2e,                                        \
                              ((flags & VB_FLAG_PREFETCH_KV) ? 1 : 0) | ((flags & VB_FLAG_LATE_TRIGGER) ? 2 : 0),     \
                              g_attn_dbg));                                                                           \
        }                                                                                                             \
    }
    const bool simt = (flags & VB_FLAG_ATTN_SIMT) != 0;
    if (pool_dtype == VB_BF16 && o_dtype == VB_BF16 && !simt) DECM(__nv_bfloat16, NSTAGE * 2 * PAGE * DH * 2)
    else if (pool_dtype == VB_BF16 && o_dtype == VB_F32 && !simt) DECM(float, NSTAGE * 2 * PAGE * DH * 2)
    else if (pool_dtype == VB_BF16 && o_dtype == VB_BF16) DEC(__nv_bfloat16, __nv_bfloat16, NSTAGE * 2 * PAGE * DH * 2)
    else if (pool_dtype == VB_BF16 && o_dtype == VB_F32) DEC(__nv_bfloat16, float, NSTAGE * 2 * PAGE * DH * 2)
    else if (pool_dtype == VB_F32 && o_dtype == VB_F32) DEC(float, float, NSTAGE * 2 * PAGE * DH * 4)
    else VB_REQUIRE(false, VB_ERR_UNSUPPORTED, "vb_attn_decode_paged: dtype combination pool=%d o=%d", pool_dtype, o_dtype);
#undef DEC
#undef DECM
    return VB_OK;
}
