/*
 * valle_b200.h -- C ABI of libvalle_b200.so (hand-written sm_100a kernels for the Valle2 hot path).
 *
 * The reference (KubiakJakub01/Valle2) has no FFI: its "operator interface" for this path is the
 * set of ATen call sites inside valle/models/{modules,utils,valle_ar,valle_nar}.py.  Each entry
 * point below replaces one (or a fused group) of those call sites; the reference file:line it
 * stands in for is cited next to it.  The Python host mirror (the modules under valle2_b200/models) binds these
 * symbols with ctypes (valle2_b200/_lib.py) -- see INTEGRATION.md for the binding stub.
 *
 * Conventions
 *   - plain C: raw device pointers, sizes, scalars, a CUDA stream passed as void* (cudaStream_t).
 *   - every function returns 0 (VB_OK) or a negative vb_status; vb_last_error_string() explains.
 *   - the caller owns all memory (weights, KV page pool, block tables, workspaces, outputs).
 *   - all launches are asynchronous on the given stream; no hidden host synchronisation, so the
 *     whole decode step can be captured in a CUDA graph.
 *   - dtype codes: VB_F32 = 0, VB_BF16 = 1.  "fp32 validation mode" = every tensor VB_F32 and the
 *     SIMT kernels; "bf16 mode" = bf16 weights/activations/KV with fp32 accumulation (tcgen05).
 */
#ifndef VALLE_B200_H_
#define VALLE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    VB_OK = 0,
    VB_ERR_BAD_ARG = -1,
    VB_ERR_UNSUPPORTED = -2,
    VB_ERR_CUDA = -3
} vb_status;

enum { VB_F32 = 0, VB_BF16 = 1 };

/* GEMM epilogues (vb_linear). */
enum {
    VB_EPI_NONE = 0,          /* y = x W^T                                  modules.py:146 (qkv), valle_ar.py:158 (proj) */
    VB_EPI_BIAS = 1,          /* y = x W^T + b                              modules.py:171 (out), :94-99 (AdaLN project) */
    VB_EPI_BIAS_GELU = 2,     /* y = gelu_erf(x W^T + b)                    modules.py:220-221 (linear_1 + nn.GELU)      */
    VB_EPI_BIAS_RESIDUAL = 3, /* y = res + x W^T + b                        modules.py:277-278 (residual adds)           */
    VB_EPI_ARGMAX = 4         /* internal to vb_linear_argmax: y is never written, the epilogue keeps the row maxima      */
};

/* Attention mask modes (vb_attention). */
enum {
    VB_MASK_NONE = 0,       /* NAR / decode: no mask (valle_nar.py:152-154, modules.py:338); keys >= kv_lens[b] are
                               still skipped when kv_lens is given (ragged-batch extension)                              */
    VB_MASK_PREFIX_LM = 1,  /* utils.py:17-43 build_attn_mask(x_len, y_len) OR key padding (valle_ar.py:69-74),
                               evaluated from (x_lens[b], kv_lens[b]) instead of a materialised (B,H,S,S) tensor        */
    VB_MASK_EXPLICIT = 2    /* arbitrary uint8 mask, nonzero = masked (modules.py:160-164 after merge_masks :175-207)    */
};

/* ---- housekeeping ------------------------------------------------------------------------------------------------ */
int vb_version(void);
const char* vb_last_error_string(void);
/* Fails (VB_ERR_UNSUPPORTED) unless the current device is compute capability 10.x. */
int vb_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* smem_optin_bytes);

/* ---- K1/K2: embedding gather (sum over codebooks) + sinusoidal PE ------------------------------------------------ */
/* out[(b*out_rows_per_batch + out_row_offset + t), :] =
 *       sum_{j < nq(t)} tables[j][ids[b][t][j]]  (left-to-right fp32 sum)  +  pe[pos(b,t)]
 *   nq(t) = nq_a for t < t_split else nq_b;  pos(b,t) = (pos_b ? pos_b[b] : pos_offset) + t
 *   ids: int32 [B][T][Q]; tables: fp32 [Q][V][d]; pe: fp32 [max_len][d]; out: fp32 rows of d.
 * Replaces modules.py:33-37 (TokenEmbedding), :78-80 (PositionalEncoding), valle_ar.py:128-129,143-144,
 * valle_nar.py:131-133,144-148,180-185 (codebook sums). */
int vb_embed_sum_pe(const int32_t* ids, const float* tables, const float* pe, float* out,
                    int B, int T, int Q, int V, int d, int t_split, int nq_a, int nq_b,
                    int pos_offset, const int32_t* pos_b, int max_len,
                    int64_t out_rows_per_batch, int64_t out_row_offset, void* stream);

/* vb_embed_sum_pe + the first (Ada)LayerNorm of the stack in one kernel (valle_nar.py:140-152 into modules.py:271 with the folded
 * AdaLN of modules.py:93-99; valle_ar.py:127-139 likewise): out row as above AND
 *   y[row][:] = (out[row] - mean) * rsqrt(var + eps) * gamma + beta      (gamma = beta = NULL: y = cast(out))
 * with the row kept in registers in between (one warp per row), so layer 0's norm1 never re-reads the residual stream.
 * y: y_dtype [rows][d], indexed like out.  d in {256, 512, 1024}.  Bit-identical to vb_embed_sum_pe + vb_residual_layernorm for
 * more than 1024 rows (the warp-per-row LayerNorm kernel, same lane layout); equal to fp32 rounding below (CTA-per-row kernel). */
int vb_embed_sum_pe_norm(const int32_t* ids, const float* tables, const float* pe, float* out,
                         int B, int T, int Q, int V, int d, int t_split, int nq_a, int nq_b,
                         int pos_offset, const int32_t* pos_b, int max_len,
                         int64_t out_rows_per_batch, int64_t out_row_offset,
                         const float* gamma, const float* beta, float eps, void* y, int y_dtype, void* stream);

/* ---- K3: LayerNorm / folded AdaLN, optionally fused with the split-K reduction + bias + residual ----------------- */
/* if (n_part > 0)  x[r,:] += bias[:] + sum_{s<n_part} part[s*part_stride + r*d + :]   (fixed order, x updated in place)
 * y[r,:] = (x[r,:] - mean) * rsqrt(var + eps) * gamma + beta           (y may be NULL: residual update only;
 *                                                                        gamma = beta = NULL: y = cast(x), no norm)
 * x: fp32 [R][d]; gamma/beta fp32 [d] (AdaLN: pre-folded w*gamma, w*beta+b -- modules.py:93-99); y: y_dtype [R][d].
 * Replaces nn.LayerNorm at modules.py:89,234-235,271,278 and the residual adds :277-278 on the decode path.
 * Decode rows (R <= 1024) with n_part >= 4 and d in {256, 512, 1024} run one row over a cluster of four CTAs. */
int vb_residual_layernorm(float* x, const float* part, int n_part, int64_t part_stride, const float* bias,
                          const float* gamma, const float* beta, void* y, int y_dtype,
                          int64_t R, int d, float eps, void* stream);

/* Profiling aid: decode-shape (R <= 1024) vb_residual_layernorm launches write %globaltimer stamps [row][4] = {start,
 * dependency resolved, row loaded (slices added), stored}; NULL switches it off. */
int vb_residual_layernorm_set_debug(void* buf);

/* y[r, n] = act(sum_s part[s][r][n] + bias[n]); act = gelu_erf if gelu != 0.  y: y_dtype.  (modules.py:220-221) */
int vb_reduce_bias_act(const float* part, int n_part, int64_t part_stride, const float* bias, int gelu,
                       void* y, int y_dtype, int64_t R, int N, void* stream);

/* ---- K4/K7/K8/K9: linear layers --------------------------------------------------------------------------------- */
/* y[M,N] = epilogue(x[M,K] . w[N,K]^T)   (nn.Linear layout: w is (out,in) row-major)
 *   fp32 mode : x,w,y VB_F32 -> SIMT fp32 kernel (bit-reproducible, used for the 1e-5 validation mode)
 *   bf16 mode : x,w VB_BF16 (K % 64 == 0... see vb_linear_query) -> tcgen05/TMEM GEMM fed by TMA, fp32 accumulate;
 *               y VB_BF16 or VB_F32; bias fp32 [N]; residual fp32 [M][N] (may alias y when y is fp32).
 * ldx/ldw/ldy are row pitches in elements. */
int vb_linear(const void* x, int x_dtype, int64_t ldx, const void* w, int w_dtype, int64_t ldw,
              const float* bias, const float* residual, int64_t ldr, void* y, int y_dtype, int64_t ldy,
              int64_t M, int64_t N, int64_t K, int epilogue, void* stream);

/* Fused logits + greedy pick (valle_nar.py:157-160 with argmax sampling; the `logits.argmax(-1)` of a greedy stage):
 *   tok[m] = argmin{ n : (x W^T)[m][n] is the row maximum }      x bf16 [M][K], w bf16 [N][K], fp32 accumulate.
 * The tcgen05 CTA-pair GEMM of vb_linear runs with an epilogue that reads each accumulator row out of tensor memory, keeps its
 * running (value, column) maximum in registers and merges the four 256-column tiles of a row with ONE 64-bit atomicMax on
 * keys[m] = order-preserving bits of the value << 32 | ~column (ties -> lowest column, as vb_sample with top_k == 1): the
 * (M x N) fp32 logits are never written to or read back from HBM.  A second tiny kernel unpacks keys[m] into
 * out_tok[(m / rows_per_batch) * batch_stride + (m % rows_per_batch) * row_stride] (int32; e.g. straight into column n of the
 * (B, T, Q) code tensor) and resets keys[m] to zero for the next call.  keys: M uint64, ZERO before the first call.
 * Needs M >= 1024 and N > 128 (the CTA-pair GEMM); a row whose logits are all -inf / NaN yields token 0.
 * Bit-identical to vb_linear(VB_EPI_NONE, fp32 y) followed by vb_sample(top_k = 1). */
int vb_linear_argmax(const void* x, int64_t ldx, const void* w, int64_t ldw, unsigned long long* keys, int32_t* out_tok,
                     int64_t rows_per_batch, int64_t batch_stride, int64_t row_stride, int64_t M, int64_t N, int64_t K,
                     void* stream);

/* Fused logits + Categorical draw (valle_nar.py:160: `Categorical(logits = logits / temperature).sample()`), same kernel and same
 * arguments as vb_linear_argmax: the epilogue scans logit / temperature + Gumbel noise, and the arg-max of that IS a draw from
 * softmax(logits / temperature) (Gumbel-max trick).  Noise = -log(-log(u)), u = a counter-based hash of (seed, step, row, column),
 * so a call is a pure function of its arguments (step: the stage, so that stages draw independently).  The draws differ from
 * vb_sample's inverse-CDF draws for the same seed (as both differ from torch.multinomial's stream: sampling is compared by
 * distribution); temperature > 0. */
int vb_linear_categorical(const void* x, int64_t ldx, const void* w, int64_t ldw, unsigned long long* keys, int32_t* out_tok,
                          int64_t rows_per_batch, int64_t batch_stride, int64_t row_stride, int64_t M, int64_t N, int64_t K,
                          float temperature, uint64_t seed, int step, void* stream);

/* The same bf16 tcgen05 GEMM with operands optionally given TRANSPOSED in memory (MN-major MMA operands, no copy):
 *   x_transposed: x is [K][M] row-major (pitch ldx) instead of [M][K];  w_transposed: w is [K][N] (pitch ldw) instead of [N][K].
 *   y[M,N] = epilogue(X . W^T) as for vb_linear.  Serves the backward pass of every nn.Linear of the path without a
 *   transpose kernel: dgrad dx = dy . W reads W (N,K) as the transposed "w" of an (M,K_out=K) product, wgrad
 *   dW = dy^T . x reads dy (R,N) and x (R,K) as transposed "x" and "w" of an (N,K) product over R.  N > 128, pitches % 8 == 0. */
int vb_linear_t(const void* x, int64_t ldx, int x_transposed, const void* w, int64_t ldw, int w_transposed, const float* bias,
                const float* residual, int64_t ldr, void* y, int y_dtype, int64_t ldy, int64_t M, int64_t N, int64_t K,
                int epilogue, void* stream);

/* Decode-shape (M <= 1024; above 128 rows the batch is tiled) weight-streaming GEMM, swap-AB on tcgen05 with split-K:
 *   part[s][m][n] = sum_{k in slice s} x[m,k] w[n,k]      s < n_split, fp32, part_stride = elements between slices.
 * Deterministic: consumers (vb_residual_layernorm / vb_reduce_bias_act / vb_attn_decode_paged / vb_sample) add the
 * slices in index order.  x,w bf16.  Returns the split count actually used in *n_split_out (<= max_split). */
int vb_linear_decode_splits(int64_t N, int64_t K, int max_split);   /* pure query: the split count vb_linear_decode uses */
int vb_linear_decode_splits_m(int64_t M, int64_t N, int64_t K, int max_split);   /* the same for M batch rows: above 128 rows the
                                                                                    batch is tiled as well and fewer splits are used */
/* flags: VB_FLAG_LATE_TRIGGER -- programmatic dependent launch: the kernel lets its successor start only after it has
 * itself waited for its predecessor, so the successor's pre-wait code may read anything written before this kernel
 * (used for the QKV GEMM, whose successor vb_attn_decode_paged prefetches KV pages before waiting). */
enum { VB_FLAG_LATE_TRIGGER = 1, VB_FLAG_PREFETCH_KV = 2, VB_FLAG_ATTN_SIMT = 4, VB_FLAG_ATTN_TICKET = 8, VB_FLAG_DG_GLOBAL = 16 };
int vb_linear_decode(const void* x, int64_t ldx, const void* w, int64_t ldw, float* part, int64_t part_stride,
                     int64_t M, int64_t N, int64_t K, int max_split, int flags, int* n_split_out, void* stream);

/* Decode-shape linear layer in "rows" form (csrc/gemm_decode_mma.cu), M <= 16, K % 256 == 0, x and w bf16:
 *   y[M][N] = epilogue(x[M][K] . w[N][K]^T)   -- the whole K of an output element stays inside one CTA when K <= 1024, so
 *   bias / erf-GELU / the residual add run in the epilogue and no reduce kernel follows (modules.py:146, :171+:274,
 *   :220-221, :278; valle_ar.py:158).  Weights stream HBM -> registers in mma.sync fragment order before the kernel waits
 *   on its predecessor (PDL); activations are bulk-copied into shared memory; eight warps split K and are summed in warp
 *   order (deterministic).
 *   epilogue VB_EPI_NONE / VB_EPI_BIAS / VB_EPI_BIAS_GELU: y is VB_F32 or VB_BF16 [M][N] (pitch ldy);
 *   VB_EPI_BIAS_RESIDUAL: y (fp32) += x.w^T + bias in place (the residual stream).
 *   K > 1024 (or want_split > 1): split-K over vb_linear_decode_rows_splits(M, K, want_split) CTAs; then epilogue must be
 *   VB_EPI_NONE and y receives fp32 slices [split][M][N] (split_stride elements apart) for vb_residual_layernorm /
 *   vb_attn_decode_paged / vb_sample to add in index order.  want_split == 0: keep the whole K (<= 4096) in one CTA
 *   whenever the M activation rows fit in shared memory (K = 4096: M <= 8), so that K > 1024 gets an epilogue too.
 *   flags: VB_FLAG_LATE_TRIGGER as for vb_linear_decode. */
int vb_linear_decode_rows_splits(int M, int64_t K, int want_split);   /* pure query; 0 = K unsupported */
int vb_linear_decode_rows(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, void* y, int y_dtype,
                          int64_t ldy, int64_t split_stride, int M, int64_t N, int64_t K, int epilogue, int want_split,
                          int flags, int* n_split_out, void* stream);

/* Same kernel with LayerNorm ON LOAD (M <= 8, K in {256, 512, 1024}): x is the fp32 residual stream [M][K]; every CTA
 * normalises the rows itself, y = epilogue(LayerNorm(x; gamma, beta, eps) . w^T), so the separate LayerNorm kernel of
 * modules.py:271 / :276 (norm1 / norm2) and its dependent launch disappear; gamma = beta = NULL: plain bf16 cast (the AR
 * logits projection, valle_ar.py:158 -- no final norm).  One slice only (no split-K). */
int vb_linear_decode_rows_ln(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, const void* w,
                             int64_t ldw, const float* bias, void* y, int y_dtype, int64_t ldy, int M, int64_t N, int64_t K,
                             int epilogue, int flags, void* stream);

/* Profiling aid: subsequent vb_linear_decode_rows launches write per-CTA stamps [cta][16]: events {kernel start, weights
 * requested, dependency resolved, activations landed, MMAs done, k-slices summed, stores issued, -} as %globaltimer, then
 * the same events as SM cycle counters (uint64, device memory, grid * 16 entries); NULL switches it off. */
int vb_linear_decode_rows_set_debug(void* buf);

/* Profiling aid: subsequent vb_linear_decode launches write %globaltimer stamps [cta][8] = {prologue done, weights requested,
 * dependency resolved, first k-block landed, MMAs issued, accumulator complete, epilogue stores issued, -} followed by the
 * same eight events as SM cycle counters (uint64, device memory, #SM*16 entries); NULL switches it off. */
int vb_linear_decode_set_debug(void* buf);

/* Decode-shape linear layer that finishes inside one launch (csrc/gemm_decode_tc.cu), M <= 256, x and w bf16:
 * swap-AB tcgen05 GEMM with split-K over CTAs, the split-K partials exchanged through an L2-resident workspace and reduced
 * in split order by the same CTAs (deterministic), and the layer's arithmetic in the epilogue.  With it a decoder layer of the
 * batched AR step is QKV -> attention -> out-proj -> FFN1 -> FFN2: five dependent launches, no LayerNorm / reduce kernels.
 *   VB_DG_PLAIN    y32[m][n] = x.w^T (+ bias)                                            valle_ar.py:158 (logits)
 *   VB_DG_LN       y32[m][n] = rstd_m (x.w^T - mean_m colsum[n]) + bias[n]               modules.py:271 + :146 (norm1 + qkv)
 *   VB_DG_LN_GELU  y16[m][n] = gelu_erf(rstd_m (x.w^T - mean_m colsum[n]) + bias[n])     modules.py:278 + :220-221
 *   VB_DG_RESIDUAL xres[m][n] += x.w^T + bias[n]; y16 = bf16(xres) (nullable);           modules.py:171+:274, :221+:278
 *                  stats_out[m][t] = (sum, sum of squares) of xres over the 128 columns of tile t (tiles = ceil(N / 128))
 * LN modes implement LayerNorm(x; gamma, beta) . W^T algebraically: the caller passes w = bf16(gamma (.) W) (columns scaled),
 * colsum[n] = sum_k w[n][k], bias = beta . W^T (+ the layer's bias), x = the RAW residual rows in bf16, and stats_in
 * [M][n_chunks_in] (sum, sum of squares) partials of those rows over K (written by the producing VB_DG_RESIDUAL launch or by
 * vb_ar_step_tail); mean_m / rstd_m are derived from them with eps.
 * Exchange: for M <= 64 and 2 <= n_split <= 16 the tile's splits run as one thread-block cluster and push their partials into
 * the owners' shared memory (DSMEM) + one cluster barrier; otherwise (or with VB_FLAG_DG_GLOBAL) through ws_part + counters.
 * Both sum the partials in split order: bit-identical results.
 *   ws_part:  >= ws_bytes of vb_decode_gemm_plan (shared by all launches of one stream-ordered chain);
 *   counters: >= tiles uint32, zeroed once; a counter array must only ever serve launches of ONE (N, K) shape (arrivals are
 *             counted modulo that shape's n_split);  flags: VB_FLAG_LATE_TRIGGER as for vb_linear_decode. */
enum { VB_DG_PLAIN = 0, VB_DG_LN = 1, VB_DG_LN_GELU = 2, VB_DG_RESIDUAL = 3 };
int vb_decode_gemm_plan(int M, int64_t N, int64_t K, int* tiles, int* n_split, int64_t* ws_bytes);   /* pure query */
int vb_decode_gemm(const void* x, int64_t ldx, const void* w, int64_t ldw, int M, int64_t N, int64_t K, int mode,
                   const float* bias, const float* colsum, const float* stats_in, int n_chunks_in, float eps,
                   float* y32, int64_t ldy32, void* y16, int64_t ldy16, float* xres, int64_t ldxres, float* stats_out,
                   void* ws_part, void* counters, int flags, void* stream);
/* Profiling aid: subsequent vb_decode_gemm launches write %globaltimer stamps [cta][16] = {prologue done, weights requested,
 * dependency resolved, first k-block landed, MMAs issued, accumulator complete, tile's splits all arrived, outputs stored,
 * partial stored, arrived on the counter, ...} (uint64, device memory, grid * 16 entries); NULL switches it off. */
int vb_decode_gemm_set_debug(void* buf);

/* ---- K5/K6: attention -------------------------------------------------------------------------------------------- */
/* General attention over strided q/k/v (element strides), fp32 or bf16 I/O, fp32 math (SIMT).
 *   q: [B][H][Sq][Dh] via (q_sb,q_sh,q_ss); k,v: [B][H][Sk][Dh]; o: [B][Sq][H*Dh] (row pitch o_ss, batch o_sb).
 *   mask_mode VB_MASK_PREFIX_LM: query i (absolute index i + q_pos0) may attend key j iff
 *        j < kv_lens[b]  and  (j < x_lens[b]  or  (i + q_pos0 >= x_lens[b] and j <= i + q_pos0))
 *   mask_mode VB_MASK_EXPLICIT: mask uint8 with strides (m_sb,m_sh,m_sq) and unit key stride, nonzero = masked.
 * Replaces F.scaled_dot_product_attention at modules.py:167 (scale 1/sqrt(Dh)). */
int vb_attention(const void* q, const void* k, const void* v, int dtype,
                 int64_t q_sb, int64_t q_sh, int64_t q_ss, int64_t k_sb, int64_t k_sh, int64_t k_ss,
                 int64_t v_sb, int64_t v_sh, int64_t v_ss, void* o, int o_dtype, int64_t o_sb, int64_t o_ss,
                 int B, int H, int Sq, int Sk, int Dh, int mask_mode, int q_pos0,
                 const int32_t* x_lens, const int32_t* kv_lens,
                 const uint8_t* mask, int64_t m_sb, int64_t m_sh, int64_t m_sq, void* stream);

/* Flash-style tensor-core attention for prefill / NAR over a packed qkv buffer [B][S][3][H][64] (bf16).
 * Same mask semantics as vb_attention (NONE / PREFIX_LM).  o: bf16 [B][S][H*64].
 * lse (nullable): fp32 [B][H][S] receives the log-sum-exp of every query row's scaled scores (+inf for a row that attends
 * nothing) -- saved by the training forward so that vb_attention_bwd does not recompute it. */
int vb_attention_prefill_tc(const void* qkv, void* o, int B, int S, int H, int mask_mode,
                            const int32_t* x_lens, const int32_t* kv_lens, float* lse, void* stream);

/* Profiling aid: subsequent vb_attention_prefill_tc launches write SM cycle stamps of one row thread per CTA,
 * [cta][key block < 32][4] = {S ready, row maximum known, previous PV done, P written} (int64); NULL switches it off. */
int vb_attention_prefill_set_debug(void* buf);

/* Paged KV pool layout (one pool per layer): [page][2 (K,V)][H][page_size=64][Dh=64], dtype f32 or bf16.
 * bf16 pools with Dh = 64 store the eight 16-byte chunks of a token row XOR-swizzled: chunk c of token slot t sits at
 * chunk c ^ (t & 7) (so that a page copied linearly into shared memory is read by ldmatrix without bank conflicts).
 * Copy K,V of a packed qkv buffer [B][S][3][H][Dh] (qkv_dtype) into the pool through block_table[B][max_pages];
 * only positions < kv_lens[b] are written.  Replaces the cache construction at modules.py:151-157. */
int vb_kv_scatter_paged(const void* qkv, int qkv_dtype, void* pool, int pool_dtype, const int32_t* block_table,
                        int max_pages, const int32_t* kv_lens, int B, int S, int H, int Dh, void* stream);

/* One decode step of attention for every sequence (modules.py:146-167 at n=1 with a cache):
 *   q,k,v (new token) = sum_s qkv_part[s][b][3*H*Dh]  (fixed order);  k,v are appended to the pool at position
 *   seq_lens[b];  o[b] = softmax(q K^T / sqrt(Dh)) V over positions 0..seq_lens[b] (inclusive of the new token).
 * Split-T (flash-decoding): grid (n_tsplit, H, B); partials in ws (>= vb_attn_decode_ws_bytes), the last CTA of each
 * (b,h) merges them -- no second launch.  o: o_dtype [B][H*Dh].  Dh must be 64, page_size 64. */
int64_t vb_attn_decode_ws_bytes(int B, int H, int n_tsplit);
/* flags: VB_FLAG_PREFETCH_KV -- the producer warp starts streaming KV pages BEFORE waiting on the predecessor kernel
 * (PDL).  Only legal when seq_lens / block_table / cached positions were written before the predecessor started
 * waiting, i.e. the predecessor is vb_linear_decode(..., VB_FLAG_LATE_TRIGGER) or any non-PDL kernel.
 * VB_FLAG_ATTN_TICKET -- merge the n_tsplit partials through the global-memory ticket even where the thread-block
 * cluster (DSMEM) merge applies (2 <= n_tsplit <= 8); both merges add the partials in split order. */
int vb_attn_decode_paged(const float* qkv_part, int n_part, int64_t part_stride, void* pool, int pool_dtype,
                         const int32_t* block_table, int max_pages, const int32_t* seq_lens,
                         void* o, int o_dtype, int B, int H, int Dh, int n_tsplit, int flags, void* ws, void* stream);

/* Profiling aid: subsequent bf16 vb_attn_decode_paged launches write %globaltimer stamps [cta][8] = {start, producer issued
 * its copies, consumers' dependency resolved, q ready, first page landed, pages done, partial written, output written}
 * (uint64, device memory, grid * 8 entries; 0 = event did not occur in that CTA); NULL switches it off. */
int vb_attn_decode_set_debug(void* buf);

/* Ask the memory system to pull cached KV pages of one layer's pool into L2 (cp.async.bulk.prefetch.L2; a hint, nothing is
 * written).  Pages [pages_total * page_lo_pct / 100, pages_total * page_hi_pct / 100) of every sequence are requested.
 * Launched on a side stream while the latency-bound GEMM chain of the previous layer leaves HBM idle. */
int vb_kv_prefetch_l2(const void* pool, int pool_dtype, const int32_t* block_table, int max_pages, const int32_t* seq_lens,
                      int B, int H, int Dh, int page_lo_pct, int page_hi_pct, void* stream);

/* ---- training step, backward pass (csrc/train.cu) ------------------------------------------------------------------ */
/* The backward GEMMs (dgrad = dY.W, wgrad = dY^T.X) are vb_linear calls on transposed operands; these are the rest.
 * Replaces torch autograd under ValleAR.training_step (valle_ar.py:43-90) / ValleNAR.training_step (valle_nar.py:53-105). */
int vb_transpose(const void* src, int dtype, int64_t rows, int64_t cols, int64_t lds, void* dst, int64_t ldd, void* stream);
/* out[n] (+)= scale * sum_r x[r][n], deterministic */
int vb_colsum(const void* x, int dtype, int64_t R, int N, int64_t ldx, float* out, int accumulate, float scale, void* stream);
/* first level of a two-level column sum over long matrices: part[row_blocks][N]; reduce part with vb_colsum */
int vb_colsum_blocks(const void* x, int dtype, int64_t R, int N, int64_t ldx, float* part, int row_blocks, void* stream);
/* y = gelu_erf(pre);  dpre = dy * gelu_erf'(pre)   (modules.py:216) */
int vb_gelu_fwd(const void* pre, int dtype, void* y, int64_t n, void* stream);
int vb_gelu_bwd(const void* pre, const void* dy, int dtype, void* dpre, int64_t n, void* stream);
/* nn.Dropout of the training step (modules.py:56/78 PositionalEncoding p = 0.1, :215/:221 FeedForward, :234-235/:277-278
 * dropout1 / dropout2).  The keep mask is a counter-based hash of (seed, site, element index): the backward pass applies the
 * SAME call to the gradient instead of reading a stored mask.
 *   vb_dropout:      x[i] = keep(i) ? x[i] / (1 - p) : 0      in place, x fp32 or bf16
 *   vb_dropout_add:  x[i] += keep(i) ? t[i] / (1 - p) : 0     x fp32 residual stream, t fp32 or bf16     (x + dropout(t)) */
int vb_dropout(void* x, int dtype, int64_t n, float p, uint64_t seed, uint64_t site, void* stream);
int vb_dropout_add(float* x, const void* t, int t_dtype, int64_t n, float p, uint64_t seed, uint64_t site, void* stream);
/* LayerNorm backward: dx[R][d] (fp32) += d/dx of LN(x; gamma, beta) applied to dy; dgamma_part / dbeta_part receive
 * vb_layernorm_bwd_blocks(R) partial rows of d floats (reduce them with vb_colsum).  gamma == NULL: the forward was a plain
 * cast, dx += dy. */
int vb_layernorm_bwd_blocks(int64_t R);
int vb_layernorm_bwd(const float* x, const float* gamma, const void* dy, int dy_dtype, float* dx, float* dgamma_part,
                     float* dbeta_part, int64_t R, int d, float eps, void* stream);
/* Attention backward over packed qkv rows [B*S][3][H][64] (dtype f32/bf16): dqkv from dO, o; lse/delta are fp32 [B][H][S]
 * work buffers.  mask_mode VB_MASK_NONE or VB_MASK_PREFIX_LM with the forward's predicate (x_lens, kv_lens).
 * lse_is_input != 0 (bf16 only): lse already holds the forward's log-sum-exp (vb_attention_prefill_tc) and is not recomputed. */
int vb_attention_bwd(const void* qkv, const void* o, const void* dO, void* dqkv, int dtype, float* lse, float* delta, int B,
                     int S, int H, int Dh, int mask_mode, const int32_t* x_lens, const int32_t* kv_lens, int lse_is_input,
                     void* stream);
/* loss_rows[r] = logsumexp(logits[r]) - logits[r][target[r]];  dlogits[r] = (softmax(logits[r]) - onehot) * scale (nullable) */
int vb_cross_entropy(const float* logits, int64_t ld, const int32_t* target, int64_t R, int V, float* loss_rows,
                     float* dlogits, int64_t ldd, float scale, void* stream);
/* grad_tables[j][ids[b][t][j]][:] += dx[row(b,t)][:] for j < nq(t): the transpose of vb_embed_sum_pe (PE has no gradient) */
int vb_embed_bwd(const int32_t* ids, const float* dx, float* grad_tables, int B, int T, int Q, int V, int d, int t_split,
                 int nq_a, int nq_b, int64_t rows_per_batch, int64_t row_offset, void* stream);

/* ---- K9-K12: logits -> temperature -> top-k -> top-p -> sample (+ log-prob) --------------------------------------- */
/* logits[r][:] = sum_s logits_part[s*part_stride + r*row_stride + :], r < R, V <= 4096.
 * utils.py:59-66 + transformers 4.38.2 top_k_top_p_filtering: /temperature; keep >= k-th largest (ties kept);
 * ascending cumulative softmax, drop cum <= 1-p, always keep the largest; softmax; draw; log_softmax at the draw.
 * The draw is inverse-CDF in index order with u = uniforms[r] (if given) else hash(seed, *step_ptr, r + row_offset)
 * (row_offset: position of row 0 inside a larger batch, so that a batch decoded in pieces draws the same numbers);
 * top_k == 1 picks the lowest-index maximum (greedy).  out_tok int32 [R]; out_logprob fp32 [R] (nullable).
 * Also serves valle_nar.py:160 (top_k=0, top_p=1 -> plain Categorical; greedy -> argmax). */
int vb_sample(const float* logits_part, int n_part, int64_t part_stride, int64_t row_stride, int R, int V,
              float temperature, int top_k, float top_p, const float* uniforms, uint64_t seed,
              const int32_t* step_ptr, int row_offset, int32_t* out_tok, float* out_logprob, void* stream);

/* K11: device-side beam bookkeeping of valle_ar.py:167-171 for B rows (no host sync):
 *   sum_logprobs[b] += logprob[b] * (last[b] != eos);  tok = (last[b]==eos) ? eos : sample[b];
 *   if every tok == eos and *stop_step < 0: *stop_step = *step;   codes_out[b][*step] = tok; last[b] = tok;
 *   seq_lens[b] += 1; audio_pos[b] += 1; *step += 1.
 * state: int32 [2] = {step, stop_step}. */
int vb_ar_bookkeeping(const int32_t* sample, const float* logprob, int32_t* last, float* sum_logprobs,
                      int32_t* codes_out, int64_t codes_stride, int32_t* seq_lens, int32_t* audio_pos,
                      int32_t* state, int B, int eos, void* stream);

/* End of a batched decode step in ONE launch, one CTA per sequence (valle_ar.py:158-171 + :143-144 of the next iteration):
 *   tok, logprob = vb_sample's draw from logits[b][0..V)  (u = uniforms[b] if given, else hash(*seed_ptr, step, b + row_offset));
 *   the row's bookkeeping as in vb_ar_bookkeeping; the last CTA to arrive advances step / stop_step;
 *   x[b] = table[tok] + pe[audio_pos[b]] -- the next step's input: fp32 residual row, x_bf16 its bf16 copy, stats[b] its
 *   (sum, sum of squares) for the LayerNorm folded into the QKV GEMM (vb_decode_gemm, n_chunks_in = 1).
 * state: int32 [4] = {step, stop_step, 0, 0} (the last two are scratch, left at zero).  The seed is read from device memory
 * so that a captured step graph serves every seed. */
int vb_ar_step_tail(const float* logits, int64_t row_stride, int V, float temperature, int top_k, float top_p,
                    const float* uniforms, const uint64_t* seed_ptr, int row_offset, int32_t* last, float* sum_logprobs,
                    int32_t* codes_out, int64_t codes_stride, int32_t* seq_lens, int32_t* audio_pos, int32_t* state, int B,
                    int eos, const float* table, const float* pe, int d, float* x, void* x_bf16, float* stats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VALLE_B200_H_ */
