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

__global__ void __launch_bounds__(THREADS) sample_kernel(const float* __restrict__ logits_part, int n_part,
                                                         int64_t part_stride, int64_t row_stride, int V,
                                                         float temperature, int top_k, float top_p,
                                                         const float* __restrict__ uniforms, uint64_t seed,
                                                         const int32_t* __restrict__ step_ptr, int row_offset,
                                                         int32_t* __restrict__ out_tok, float* __restrict__ out_logprob) {
    __shared__ float val[MAX_V];
    __shared__ uint16_t kept_idx[MAX_V];
    __shared__ uint8_t alive[MAX_V];
    __shared__ float red[THREADS / 32];
    __shared__ unsigned hist[256];
    __shared__ unsigned sel_prefix, sel_k, n_kept_s;
    __shared__ int pick;
    __shared__ float warp_tot[THREADS / 32];

    pdl_trigger();
    pdl_wait();
    const int r = blockIdx.x, tid = threadIdx.x;
    const float* src = logits_part + static_cast<int64_t>(r) * row_stride;
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
        if (tid == 0) {
            out_tok[r] = pick;
            if (out_logprob) out_logprob[r] = -logf(ties);
        }
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
    float u;
    if (uniforms) u = uniforms[r];
    else {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (static_cast<uint64_t>(step_ptr ? *step_ptr : 0) * 1000003ull + (r + row_offset) + 1);
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
    if (tid == 0) {
        const int t = pick;
        out_tok[r] = t;
        if (out_logprob) out_logprob[r] = (val[t] - M2) - logf(Z2);
    }
}

}  // namespace

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
