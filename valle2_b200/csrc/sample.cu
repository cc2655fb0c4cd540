// Fused logits reduction -> temperature -> top-k -> top-p -> draw -> log-prob (K9-K12): one CTA per row, the
// whole vocabulary (<= 4096, here 1025 / 1024) lives in shared memory.  Semantics follow valle/models/utils.py:46-68
// with transformers 4.38.2 top_k_top_p_filtering:
//   top-k : keep scores >= k-th largest (ties kept)            (radix select on order-preserving keys)
//   top-p : ascending (value, index) order, cumulative softmax, drop cum <= 1-p, always keep the largest
//   draw  : inverse CDF in index order with an explicit uniform (the reference uses torch.multinomial, whose RNG
//           stream is not reproducible outside torch); top_k == 1 -> lowest-index maximum (greedy)
//   logp  : log_softmax of the FILTERED logits at the draw (utils.py:65-66)
#include <math.h>

#include "common.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int MAX_V = 4096;

__device__ __forceinline__ uint32_t order_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int i = 1; i < THREADS / 32; ++i) r = fmaxf(r, red[i]);
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) r += red[i];
    __syncthreads();
    return r;
}

// One row: the whole CTA (THREADS threads) works on it.  The result is returned to EVERY thread (tok, logprob).
// u_given < 0: draw u from hash(seed, step, row_id) instead.
__device__ void sample_row(const float* __restrict__ src, int n_part, int64_t part_stride, int V, float temperature,
                           int top_k, float top_p, float u_given, uint64_t seed, int step, int row_id, int& tok_out,
                           float& logprob_out) {
    __shared__ float val[MAX_V];
    __shared__ uint16_t kept_idx[MAX_V];
    __shared__ uint8_t alive[MAX_V];
    __shared__ float red[THREADS / 32];
    __shared__ unsigned hist[256];
    __shared__ unsigned sel_prefix, sel_k, n_kept_s;
    __shared__ int pick;
    __shared__ float warp_tot[THREADS / 32];
    __shared__ float res_lp;

    const int tid = threadIdx.x;
    for (int i = tid; i < V; i += THREADS) {
        float a = 0.f;
        for (int s = 0; s < n_part; ++s) a += src[s * part_stride + i];
        val[i] = a / temperature;                                     // utils.py:59-60
    }
    __syncthreads();

    if (top_k == 1) {
        // greedy: lowest-index maximum; log-prob under the filtered distribution = -log(#ties)
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = tid; i < V; i += THREADS) {
            if (val[i] > best) { best = val[i]; bi = i; }
        }
        const float gmax = block_reduce_max(best, red);
        int cnt = 0;
        for (int i = tid; i < V; i += THREADS) cnt += (val[i] == gmax);
        if (tid == 0) pick = 0x7fffffff;
        __syncthreads();
        if (best == gmax && bi != 0x7fffffff) atomicMin(&pick, bi);
        const float ties = block_reduce_sum(static_cast<float>(cnt), red);
        __syncthreads();
        tok_out = pick;
        logprob_out = -logf(ties);
        __syncthreads();                                              // pick may be rewritten by the next call
        return;
    }

    // ---- top-k threshold by radix select (4 x 8 bits, from the most significant byte) ----
    float thr = -INFINITY;
    if (top_k > 0 && top_k < V) {
        if (tid == 0) { sel_prefix = 0; sel_k = static_cast<unsigned>(top_k); }
        __syncthreads();
        for (int pass = 3; pass >= 0; --pass) {
            hist[tid] = 0;
            __syncthreads();
            const unsigned prefix = sel_prefix;
            const unsigned himask = (pass == 3) ? 0u : (0xffffffffu << ((pass + 1) * 8));
            for (int i = tid; i < V; i += THREADS) {
                const uint32_t key = order_key(val[i]);
                if ((key & himask) == prefix) atomicAdd(&hist[(key >> (pass * 8)) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                unsigned k = sel_k, bin = 255;
                for (;; --bin) {
                    if (hist[bin] >= k) break;
                    k -= hist[bin];
                    if (bin == 0) break;
                }
                sel_k = k;
                sel_prefix = prefix | (bin << (pass * 8));
            }
            __syncthreads();
        }
        const uint32_t key = sel_prefix;   // key of the k-th largest value
        const uint32_t u = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
        thr = __uint_as_float(u);
    }
    for (int i = tid; i < V; i += THREADS) alive[i] = (val[i] >= thr) ? 1 : 0;
    __syncthreads();

    // ---- top-p over the survivors ----
    if (top_p >= 0.f && top_p < 1.f) {
        float lm = -INFINITY;
        for (int i = tid; i < V; i += THREADS) if (alive[i]) lm = fmaxf(lm, val[i]);
        const float M = block_reduce_max(lm, red);
        float ls = 0.f;
        for (int i = tid; i < V; i += THREADS) if (alive[i]) ls += expf(val[i] - M);
        const float Z = block_reduce_sum(ls, red);
        if (tid == 0) n_kept_s = 0;
        __syncthreads();
        for (int i = tid; i < V; i += THREADS) if (alive[i]) kept_idx[atomicAdd(&n_kept_s, 1u)] = static_cast<uint16_t>(i);
        __syncthreads();
        const int n_kept = static_cast<int>(n_kept_s);
        const float cut = 1.f - top_p;
        for (int a = tid; a < n_kept; a += THREADS) {
            const int i = kept_idx[a];
            const float vi = val[i];
            float cum = 0.f;
            bool is_largest = true;
            for (int bb = 0; bb < n_kept; ++bb) {
                const int j = kept_idx[bb];
                const float vj = val[j];
                const bool le = (vj < vi) || (vj == vi && j <= i);
                if (le) cum += expf(vj - M);
                else is_largest = false;
            }
            if (!is_largest && (cum / Z) <= cut) alive[i] = 2;   // mark, applied after the barrier
        }
        __syncthreads();
        for (int i = tid; i < V; i += THREADS) if (alive[i] == 2) alive[i] = 0;
        __syncthreads();
    }

    // ---- softmax over the survivors, inverse-CDF draw in index order ----
    float lm = -INFINITY;
    for (int i = tid; i < V; i += THREADS) if (alive[i]) lm = fmaxf(lm, val[i]);
    const float M2 = block_reduce_max(lm, red);
    const int chunk = (V + THREADS - 1) / THREADS;
    const int i0 = tid * chunk, i1 = min(i0 + chunk, V);
    float local = 0.f;
    for (int i = i0; i < i1; ++i) if (alive[i]) local += expf(val[i] - M2);
    // block exclusive scan of the per-thread sums
    float incl = local;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, incl, off);
        if ((tid & 31) >= off) incl += t;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    if (tid == 0) pick = 0x7fffffff;
    __syncthreads();
    float base = 0.f, Z2 = 0.f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) {
        if (w < (tid >> 5)) base += warp_tot[w];
        Z2 += warp_tot[w];
    }
    float u = u_given;
    if (u < 0.f) {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (static_cast<uint64_t>(step) * 1000003ull + row_id + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        u = static_cast<float>(z >> 40) * (1.0f / 16777216.0f);
    }
    const float target = u * Z2;
    float run = base + incl - local;
    for (int i = i0; i < i1; ++i) {
        if (!alive[i]) continue;
        run += expf(val[i] - M2);
        if (run > target) { atomicMin(&pick, i); break; }
    }
    __syncthreads();
    if (pick == 0x7fffffff) {   // u*Z2 rounded past the last survivor: take the last surviving index
        __syncthreads();
        if (tid == 0) pick = -1;
        __syncthreads();
        for (int i = i0; i < i1; ++i) if (alive[i]) atomicMax(&pick, i);
        __syncthreads();
    }
    if (tid == 0) res_lp = (val[pick] - M2) - logf(Z2);
    __syncthreads();
    tok_out = pick;
    logprob_out = res_lp;
    __syncthreads();
}

__global__ void __launch_bounds__(THREADS) sample_kernel(const float* __restrict__ logits_part, int n_part,
                                                         int64_t part_stride, int64_t row_stride, int V,
                                                         float temperature, int top_k, float top_p,
                                                         const float* __restrict__ uniforms, uint64_t seed,
                                                         const int32_t* __restrict__ step_ptr, int row_offset,
                                                         int32_t* __restrict__ out_tok, float* __restrict__ out_logprob) {
    pdl_trigger();
    pdl_wait();
    const int r = blockIdx.x;
    int tok;
    float lp;
    sample_row(logits_part + static_cast<int64_t>(r) * row_stride, n_part, part_stride, V, temperature, top_k, top_p,
               uniforms ? uniforms[r] : -1.f, seed, step_ptr ? *step_ptr : 0, r + row_offset, tok, lp);
    if (threadIdx.x == 0) {
        out_tok[r] = tok;
        if (out_logprob) out_logprob[r] = lp;
    }
}

// End of a decode step in ONE launch, one CTA per sequence: sample (valle_ar.py:158-166) -> bookkeeping of the row
// (:167-171; the last CTA to arrive advances {step, stop_step}) -> the NEXT step's input row (:143-144): audio embedding +
// PE as fp32 residual row, its bf16 copy (the QKV GEMM's operand) and its (sum, sum of squares) for the folded LayerNorm.
// state int32 [4] = {step, stop_step, arrivals, any row still running}.
__global__ void __launch_bounds__(THREADS) ar_step_tail_kernel(const float* __restrict__ logits, int64_t row_stride, int V,
                                                               float temperature, int top_k, float top_p,
                                                               const float* __restrict__ uniforms,
                                                               const uint64_t* __restrict__ seed_ptr, int row_offset,
                                                               int32_t* __restrict__ last, float* __restrict__ sum_logprobs,
                                                               int32_t* __restrict__ codes_out, int64_t codes_stride,
                                                               int32_t* __restrict__ seq_lens, int32_t* __restrict__ audio_pos,
                                                               int32_t* __restrict__ state, int B, int eos,
                                                               const float* __restrict__ table, const float* __restrict__ pe,
                                                               int d, float* __restrict__ x, __nv_bfloat16* __restrict__ xb,
                                                               float2* __restrict__ stats) {
    __shared__ float red2[2][THREADS / 32];
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x, tid = threadIdx.x;
    const int step = state[0];
    int tok;
    float lp;
    sample_row(logits + static_cast<int64_t>(b) * row_stride, 1, 0, V, temperature, top_k, top_p, uniforms ? uniforms[b] : -1.f,
               seed_ptr ? *seed_ptr : 0ull, step, b + row_offset, tok, lp);
    const int prev = last[b];
    const bool done = (prev == eos);
    if (done) tok = eos;
    const int pos = audio_pos[b] + 1;
    __syncthreads();                                                  // every thread has read last[b] / audio_pos[b]
    if (tid == 0) {
        if (!done) sum_logprobs[b] += lp;
        if (step < codes_stride) codes_out[b * codes_stride + step] = tok;
        last[b] = tok;
        seq_lens[b] += 1;
        audio_pos[b] = pos;
        if (tok != eos) atomicOr(&state[3], 1);
        __threadfence();
        if (atomicAdd(&state[2], 1) == B - 1) {                      // last row of the batch: advance the step
            __threadfence();
            const int any = atomicExch(&state[3], 0);
            if (!any && state[1] < 0) state[1] = step;
            state[2] = 0;
            state[0] = step + 1;
        }
    }
    // next step's input row
    const float* trow = table + static_cast<int64_t>(tok) * d;
    const float* prow = pe + static_cast<int64_t>(pos) * d;
    float s = 0.f, sq = 0.f;
    for (int i = tid * 4; i < d; i += THREADS * 4) {
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(trow + i));
        const float4 p4 = __ldg(reinterpret_cast<const float4*>(prow + i));
        const float4 v = make_float4(t4.x + p4.x, t4.y + p4.y, t4.z + p4.z, t4.w + p4.w);
        *reinterpret_cast<float4*>(x + static_cast<int64_t>(b) * d + i) = v;
        *reinterpret_cast<uint2*>(xb + static_cast<int64_t>(b) * d + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        s += (v.x + v.y) + (v.z + v.w);
        sq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, sq))));
    }
    s = warp_sum(s);
    sq = warp_sum(sq);
    if ((tid & 31) == 0) { red2[0][tid >> 5] = s; red2[1][tid >> 5] = sq; }
    __syncthreads();
    if (tid == 0) {
        float a = 0.f, c = 0.f;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) { a += red2[0][w]; c += red2[1][w]; }
        stats[b] = make_float2(a, c);
    }
}

}  // namespace

extern "C" int vb_ar_step_tail(const float* logits, int64_t row_stride, int V, float temperature, int top_k, float top_p,
                               const float* uniforms, const uint64_t* seed_ptr, int row_offset, int32_t* last,
                               float* sum_logprobs, int32_t* codes_out, int64_t codes_stride, int32_t* seq_lens,
                               int32_t* audio_pos, int32_t* state, int B, int eos, const float* table, const float* pe,
                               int d, float* x, void* x_bf16, float* stats, void* stream) {
    VB_REQUIRE(logits && last && sum_logprobs && codes_out && seq_lens && audio_pos && state && table && pe && x && x_bf16 && stats,
               VB_ERR_BAD_ARG, "vb_ar_step_tail: null pointer");
    VB_REQUIRE(V >= 1 && V <= MAX_V, VB_ERR_UNSUPPORTED, "vb_ar_step_tail: vocabulary %d not in [1,%d]", V, MAX_V);
    VB_REQUIRE(B >= 1 && B <= 65535 && temperature > 0.f && d >= 4 && d % 4 == 0, VB_ERR_BAD_ARG, "vb_ar_step_tail: bad args");
    VB_CUDA(vb_launch(true, ar_step_tail_kernel, dim3(B), dim3(THREADS), 0, static_cast<cudaStream_t>(stream), logits, row_stride, V,
                      temperature, top_k, top_p, uniforms, seed_ptr, row_offset, last, sum_logprobs, codes_out, codes_stride,
                      seq_lens, audio_pos, state, B, eos, table, pe, d, x, static_cast<__nv_bfloat16*>(x_bf16),
                      reinterpret_cast<float2*>(stats)));
    return VB_OK;
}

extern "C" int vb_sample(const float* logits_part, int n_part, int64_t part_stride, int64_t row_stride, int R, int V,
                         float temperature, int top_k, float top_p, const float* uniforms, uint64_t seed,
                         const int32_t* step_ptr, int row_offset, int32_t* out_tok, float* out_logprob, void* stream) {
    VB_REQUIRE(logits_part && out_tok, VB_ERR_BAD_ARG, "vb_sample: null pointer");
    VB_REQUIRE(V >= 1 && V <= MAX_V, VB_ERR_UNSUPPORTED, "vb_sample: vocabulary %d not in [1,%d]", V, MAX_V);
    VB_REQUIRE(R >= 0 && n_part >= 1 && temperature > 0.f, VB_ERR_BAD_ARG, "vb_sample: bad args");
    if (R == 0) return VB_OK;
    VB_CUDA(vb_launch(R <= 1024, sample_kernel, dim3(R), dim3(THREADS), 0, static_cast<cudaStream_t>(stream), logits_part, n_part, part_stride,
                      row_stride, V, temperature, top_k, top_p, uniforms, seed, step_ptr, row_offset, out_tok, out_logprob));
    return VB_OK;
}
