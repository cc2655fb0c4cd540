"""Execution engines for the hot path: packed weights, KV page pool, workspaces, CUDA-graph decode.

``StackWeights``   per-layer weights of a ``Transformer`` stack in the compute dtype; AdaLN scale/shift
                   pre-folded per (layer, norm, stage):  (w*gamma) * xhat + (w*beta + b)   (modules.py:93-99).
``StackRunner``    large-M forward over packed rows (AR prefill, NAR stages, teacher-forced logits).
``ARDecoder``      batched KV-cached decode (valle_ar.py:92-180): prefill -> CUDA-graph decode steps with
                   device-side sampling + EOS bookkeeping (no per-step host sync).
``NARDecoder``     stages 2..Q with full attention and summed codebook embeddings (valle_nar.py:107-165, repaired
                   per SURVEY Appendix A).

Precision: 'bf16' = bf16 weights / activations / KV, fp32 accumulation and fp32 residual stream (tcgen05 GEMMs);
'fp32' = validation mode, everything fp32 on the SIMT kernels.
"""
from __future__ import annotations

import math
import os

import torch

from . import ops
from .ops import MASK_NONE, MASK_PREFIX_LM, PAGE


_NVTX = os.environ.get('VALLE_B200_NVTX', '0') != '0'


class _nvtx:
    """NVTX range around a phase of the path (VALLE_B200_NVTX=1; off by default: no overhead in the timed loops)."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def _cdtype(precision: str) -> torch.dtype:
    return torch.bfloat16 if precision == 'bf16' else torch.float32


class StackWeights:
    def __init__(self, transformer, norm: str, precision: str, stage_embs: list[torch.Tensor] | None = None):
        self.precision = precision
        cd = _cdtype(precision)
        self.layers = []
        for layer in transformer.layers:
            a, f = layer.self_attn, layer.ffn
            L = {
                'wqkv': a.qkv.weight.detach().to(cd).contiguous(),
                'wo': a.out.weight.detach().to(cd).contiguous(),
                'bo': a.out.bias.detach().float().contiguous(),
                'w1': f.linear_1.weight.detach().to(cd).contiguous(),
                'b1': f.linear_1.bias.detach().float().contiguous(),
                'w2': f.linear_2.weight.detach().to(cd).contiguous(),
                'b2': f.linear_2.bias.detach().float().contiguous(),
            }
            for name in ('norm1', 'norm2'):
                nm = getattr(layer, name)
                if norm == 'LayerNorm':
                    L[name] = (nm.weight.detach().float().reshape(1, -1).contiguous(),
                               nm.bias.detach().float().reshape(1, -1).contiguous(), nm.eps)
                else:
                    assert stage_embs is not None, 'AdaptiveLayerNorm needs stage embeddings'
                    d = nm.d_model
                    e = torch.cat([s.detach().float().reshape(1, d) for s in stage_embs], 0).contiguous()
                    wb = ops.linear(e, nm.project_layer.weight.detach().float().contiguous(),
                                    nm.project_layer.bias.detach().float().contiguous())      # (n_stage, 2d) fp32 SIMT
                    w, b = wb[:, :d], wb[:, d:]
                    g0, b0 = nm.norm.weight.detach().float(), nm.norm.bias.detach().float()
                    # the fold is two elementwise products at weight-load time (exact up to 1 ulp)
                    L[name] = ((w * g0).contiguous(), (w * b0 + b).contiguous(), nm.eps)
            self.layers.append(L)
        l0 = self.layers[0]
        self.d = l0['wo'].shape[0]
        self.F = l0['w1'].shape[0]


class StackRunner:
    """Large-M forward: x (R,d) fp32 residual stream updated in place."""

    def __init__(self, weights: StackWeights, n_heads: int):
        self.w = weights
        self.H = n_heads
        self.cd = _cdtype(weights.precision)
        self._ws = {}

    def _buffers(self, R: int, device):
        key = (R, device)
        if key not in self._ws:
            d, F = self.w.d, self.w.F
            self._ws = {key: {                       # keep only the latest size resident
                'h': torch.empty(R, d, device=device, dtype=self.cd),
                'qkv': torch.empty(R, 3 * d, device=device, dtype=self.cd),
                'o': torch.empty(R, d, device=device, dtype=self.cd),
                'f': torch.empty(R, F, device=device, dtype=self.cd),
            }}
        return self._ws[key]

    def fused_first_norm_ok(self) -> bool:
        return self.w.d in (256, 512, 1024)

    def first_norm_buffer(self, R: int, device) -> torch.Tensor:
        return self._buffers(R, device)['h']

    def first_norm_args(self, stage: int = 0) -> dict:
        """gamma / beta / eps of layer 0's norm1 (the stage's folded AdaLN affine), as keyword arguments of ops.embed_sum_pe."""
        g, b, eps = self.w.layers[0]['norm1']
        return {'gamma': g[min(stage, g.shape[0] - 1)], 'beta': b[min(stage, b.shape[0] - 1)], 'eps': eps}

    def forward(self, x: torch.Tensor, B: int, S: int, *, mask_mode: int, x_lens=None, kv_lens=None, stage: int = 0,
                kv_pools: torch.Tensor | None = None, block_table: torch.Tensor | None = None,
                use_tc_attention: bool | None = None, first_norm_done: bool = False) -> torch.Tensor:
        """``first_norm_done``: layer 0's norm1 output already sits in ``first_norm_buffer(R)`` (written by the fused
        embedding-sum + PE + LayerNorm kernel, ``first_norm_args``), so its LayerNorm launch is skipped."""
        R, d = x.shape
        if use_tc_attention is None:    # tcgen05 flash attention whenever the shape allows (bf16, head_dim 64)
            use_tc_attention = (self.cd == torch.bfloat16 and d // self.H == 64
                                and os.environ.get('VALLE_B200_TC_ATTN', '1') != '0')
        assert R == B * S
        buf = self._buffers(R, x.device)
        h, qkv, o, f = buf['h'], buf['qkv'], buf['o'], buf['f']
        H, Dh = self.H, d // self.H
        for li, L in enumerate(self.w.layers):
            g, b, eps = L['norm1']
            if not (li == 0 and first_norm_done):
                ops.residual_layernorm(x, g[min(stage, g.shape[0] - 1)], b[min(stage, b.shape[0] - 1)], h, eps=eps)
            ops.linear(h, L['wqkv'], out=qkv)
            if kv_pools is not None:
                ops.kv_scatter_paged(qkv, kv_pools[li], block_table, kv_lens, B, S, H, Dh)
            ops.attention_packed(qkv, o, B, S, H, mask_mode=mask_mode, x_lens=x_lens, kv_lens=kv_lens,
                                 use_tc=use_tc_attention)
            ops.linear(o, L['wo'], L['bo'], residual=x, out=x)
            g, b, eps = L['norm2']
            ops.residual_layernorm(x, g[min(stage, g.shape[0] - 1)], b[min(stage, b.shape[0] - 1)], h, eps=eps)
            ops.linear(h, L['w1'], L['b1'], gelu=True, out=f)
            ops.linear(f, L['w2'], L['b2'], residual=x, out=x)
        return x


def _as_i64(seed: int) -> int:
    """The 64 seed bits as a signed int64 (what a torch.int64 tensor stores)."""
    seed &= (1 << 64) - 1
    return seed - (1 << 64) if seed >= (1 << 63) else seed


def _i32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.int32).contiguous()


class ARDecoder:
    """Batched KV-cached decode.  The batch may be cut into ``n_sub`` independent sub-batches (whole utterances) that run
    on parallel branches of the step's CUDA graph: while one sub-batch is in its HBM-bound attention kernel the other is in
    its latency-bound GEMM chain, so the two kinds of kernel overlap on the GPU.  Every sub-batch owns its workspaces;
    KV pool, block table, counters and outputs are row-slices of the full-batch tensors."""

    def __init__(self, model, precision: str):
        cfg = model.config
        assert cfg.norm == 'LayerNorm', 'ValleAR runs only with norm=LayerNorm (reference defect A-10)'
        self.cfg = cfg
        self.precision = precision
        self.cd = _cdtype(precision)
        self.device = model.device
        self.H = cfg.n_heads
        self.d = cfg.d_model
        self.Dh = self.d // self.H
        if self.Dh != 64:
            # the paged decode attention (csrc/attn_decode.cu) is built for 64-wide heads: one KV page row = one 128-byte line.
            # The default VALL-E (1024 / 16) and the tiny test config (256 / 4) both have them; the module-level forwards
            # (MultiHeadAttention.forward, csrc/attn_simt.cu) take any head_dim <= 256.
            raise ValueError(f'KV-cached generation needs d_model / n_heads == 64 (got {self.d} / {self.H} = {self.Dh}); '
                             'see DESIGN.md section 7')
        self.V = cfg.num_audio_tokens + 1
        self.weights = StackWeights(model.transformer, cfg.norm, precision)
        self.runner = StackRunner(self.weights, self.H)
        self.tok_table = model.tokens_emb.weight.detach().float().unsqueeze(0).contiguous()
        self.aud_table = model.audio_emb.weight.detach().float().unsqueeze(0).contiguous()
        self.pe_t = model.tokens_position_emb.pe.detach().float().reshape(-1, self.d).contiguous()
        self.pe_a = model.audio_position_emb.pe.detach().float().reshape(-1, self.d).contiguous()
        self.wproj = model.proj.weight.detach().to(self.cd).contiguous()
        self._state = None
        self._graph = None
        self._graph_key = None
        self._streams = []
        self.replays = 0                       # CUDA-graph replays of the decode step so far (launch accounting)
        self._phase_events = None
        self._seed_dev = torch.zeros(1, device=self.device, dtype=torch.int64)     # read by vb_ar_step_tail: graphs are seed-independent
        if precision == 'bf16':
            self._fold_layernorm(model)
        self.page_permutation_seed = None      # tests: scatter the logical pages over the pool
        # tunables (env overrides are for experiments; the defaults are the measured best)
        # decode GEMM form: 'lean' = full-K mma.sync rows kernels with LayerNorm on load (csrc/gemm_decode_mma.cu, <= 8
        # sequences), 'tc' = tcgen05 swap-AB GEMMs with the split-K reduction + folded LayerNorm + epilogues inside the launch
        # (csrc/gemm_decode_tc.cu, 5 launches per layer), 'splitk' = the round-1 form: tcgen05 split-K slices + LayerNorm /
        # GELU-reduce kernels (8 launches per layer; kept as the A/B reference), 'auto' = lean up to 8 sequences, tc above.
        self.fused_embed_norm = True   # prefill: embedding + PE + layer 0's norm1 in one kernel (tests: A/B)
        self.decode_gemm = os.environ.get('VALLE_B200_DECODE_GEMM', 'auto')
        # lean path, >= 4 sequences: the attention kernel releases its successors only after its own wait (see
        # csrc/attn_decode.cu; measured -2 % step time at B = 4..8, +2 % at B = 1, tools/step_breakdown.py)
        self.attn_late = ops.FLAG_LATE_TRIGGER if os.environ.get('VALLE_B200_ATTN_LATE', '1') != '0' else 0
        self.attn_late_splitk = ops.FLAG_LATE_TRIGGER if os.environ.get('VALLE_B200_ATTN_LATE_SPLITK', '0') != '0' else 0   # A/B
        self.qkv_late_all = os.environ.get('VALLE_B200_QKV_LATE_ALL', '0') != '0'      # A/B: late PDL trigger in every layer
        # tc path, bit mask over (out-proj, FFN1, FFN2): release the successor only after the own dependency wait, so that at
        # most two kernels of the chain are resident at a time (their CTAs then all find a slot before their predecessor ends)
        self.dg_late = int(os.environ.get('VALLE_B200_DG_LATE', '0'))
        self.n_sub_override = int(os.environ.get('VALLE_B200_SUBBATCH', '0'))
        self.attn_ctas = int(os.environ.get('VALLE_B200_ATTN_CTAS', '0'))
        self.n_tsplit_override = 0             # tests: pin the flash-decoding split
        self.attn_flags = ops.FLAG_ATTN_SIMT if os.environ.get('VALLE_B200_ATTN_SIMT', '0') != '0' else 0

    def _fold_layernorm(self, model):
        """Weight-load-time fold of norm1 / norm2 into the QKV / FFN1 GEMMs of the decode path (csrc/gemm_decode_tc.cu):
            LN(x) . W^T = rstd (x . (gamma (.) W)^T - mean c) + beta . W^T,    c[n] = sum_k bf16(gamma[k] W[n][k])
        The scaled weights are rounded to bf16 ONCE from the fp32 master weights; c is summed from the rounded values (what
        the MMA multiplies); beta . W^T (+ the layer's bias) stays fp32.  Elementwise products and row sums only."""
        for L, layer in zip(self.weights.layers, model.transformer.layers):
            for wkey, lin, nkey, bkey, out in (('wqkv_s', layer.self_attn.qkv, 'norm1', None, ('c_qkv', 'b_qkv')),
                                               ('w1_s', layer.ffn.linear_1, 'norm2', 'b1', ('c_1', 'b_1'))):
                g, b, _ = L[nkey]
                w32 = lin.weight.detach().float()
                ws = (w32 * g[0][None, :]).to(self.cd).contiguous()
                L[wkey] = ws
                L[out[0]] = ws.float().sum(dim=1).contiguous()
                fold_b = (w32 * b[0][None, :]).sum(dim=1)
                L[out[1]] = (fold_b + L[bkey] if bkey else fold_b).contiguous()

    # ------------------------------------------------------------------------------------------
    def _n_sub(self, B: int) -> int:
        if self.precision != 'bf16':
            return 1
        if self.n_sub_override > 0:
            return max(1, min(self.n_sub_override, B))
        return 1

    def _lean_ok(self, sub: dict) -> bool:
        """Five kernels per layer (small batch): LayerNorm on load inside the QKV / FFN1 GEMMs, residual adds and GELU in
        the epilogues, FFN2 with its whole K = F in one CTA (csrc/gemm_decode_mma.cu).  Every CTA reads the whole fp32
        residual matrix, which is what limits it to <= 8 sequences."""
        if self.decode_gemm not in ('auto', 'lean') or self.precision != 'bf16':
            return False
        if sub['B'] > (8 if self.decode_gemm == 'lean' else 7):
            return False
        return (self.d in (256, 512, 1024) and self.weights.F % 256 == 0
                and ops.linear_decode_rows_splits(self.weights.F, 0, sub['B']) == 1)

    def _tc_ok(self, sub: dict) -> bool:
        """Five launches per layer on the tcgen05 decode GEMM with in-kernel split-K reduction (csrc/gemm_decode_tc.cu).
        'auto' takes it for 8..32 sequences (measured, tools/ab_r02_batches.sh: B=8 0.305 vs 0.316 ms lean, B=16 0.355 vs
        0.384 ms split-K, B=32 0.46 vs 0.505; above 32 rows the exchange + epilogue inside the launch grow with the batch
        and the separate reduce kernels win: B=64 0.85 vs 0.75 ms, B=256 2.63 vs 2.16)."""
        if self.precision != 'bf16' or 'dg' not in sub or self._lean_ok(sub):
            return False
        if self.decode_gemm == 'tc':
            return sub['B'] <= 256
        return self.decode_gemm == 'auto' and sub['B'] <= 32

    def _make_sub(self, st: dict, b0: int, b1: int, state: torch.Tensor) -> dict:
        """Workspaces of one sub-batch (rows b0..b1 of the batch) + row-slice views of the shared decode state."""
        dev, d, F, H, V = self.device, self.d, self.weights.F, self.H, self.V
        B = b1 - b0
        sub = {'B': B, 'b0': b0, 'state': state}
        for k in ('block_table', 'seq_lens', 'audio_pos', 'last', 'sum_logprobs', 'codes_out'):
            sub[k] = st[k][b0:b1]
        sub['sample'] = torch.zeros(B, device=dev, dtype=torch.int32)
        sub['logprob'] = torch.zeros(B, device=dev, dtype=torch.float32)
        sub['x'] = torch.zeros(B, d, device=dev, dtype=torch.float32)
        sub['lg'] = torch.zeros(B, V, device=dev, dtype=torch.float32)
        if self.precision == 'bf16':
            sub['h'] = torch.zeros(B, d, device=dev, dtype=self.cd)
            sub['o'] = torch.zeros(B, d, device=dev, dtype=self.cd)
            sub['f'] = torch.zeros(B, F, device=dev, dtype=self.cd)
            sub['xb'] = torch.zeros(B, d, device=dev, dtype=self.cd)          # bf16 copy of the residual rows (tc path)
            sub['qkv32'] = torch.zeros(B, 3 * d, device=dev, dtype=torch.float32)
            if self.decode_gemm == 'splitk' or (self.decode_gemm == 'auto' and B > 32) or B > 256:
                ms = 32
                ns = {k: ops.linear_decode_splits(n, kk, ms, B) for k, (n, kk) in
                      {'qkv': (3 * d, d), 'o': (d, d), 'f1': (F, d), 'f2': (d, F), 'lg': (V, d)}.items()}
                sub['ns'] = ns
                sub['p_qkv'] = torch.zeros(ns['qkv'], B, 3 * d, device=dev, dtype=torch.float32)
                sub['p_o'] = torch.zeros(ns['o'], B, d, device=dev, dtype=torch.float32)
                sub['p_f1'] = torch.zeros(ns['f1'], B, F, device=dev, dtype=torch.float32)
                sub['p_f2'] = torch.zeros(ns['f2'], B, d, device=dev, dtype=torch.float32)
                sub['p_lg'] = torch.zeros(ns['lg'], B, V, device=dev, dtype=torch.float32)
            else:
                # tc path: plan of every GEMM shape (tiles, n_split), the shared exchange buffer, one counter array per shape
                dg = {k: ops.decode_gemm_plan(B, n, kk) for k, (n, kk) in
                      {'qkv': (3 * d, d), 'o': (d, d), 'f1': (F, d), 'f2': (d, F), 'lg': (V, d)}.items()}
                sub['dg'] = dg
                sub['dg_ws'] = torch.zeros(max(v['ws_bytes'] for v in dg.values()) // 4, device=dev, dtype=torch.float32)
                sub['dg_ctr'] = torch.zeros(len(dg), 256, device=dev, dtype=torch.int32)
                chunks = max(dg['o']['tiles'], dg['f2']['tiles'], 1)
                sub['stats'] = torch.zeros(B * chunks * 2, device=dev, dtype=torch.float32)
            if self.d % 256 == 0 and F % 256 == 0 and B <= 8:
                sub['r_qkv'] = torch.zeros(1, B, 3 * d, device=dev, dtype=torch.float32)
            if 'stats' not in sub:
                sub['stats'] = torch.zeros(B * 2, device=dev, dtype=torch.float32)
        else:
            sub['h'] = torch.zeros(B, d, device=dev, dtype=torch.float32)
            sub['qkv'] = torch.zeros(B, 3 * d, device=dev, dtype=torch.float32)
            sub['o'] = torch.zeros(B, d, device=dev, dtype=torch.float32)
            sub['f'] = torch.zeros(B, F, device=dev, dtype=torch.float32)
        # flash-decoding split: enough CTAs to fill the GPU, never more splits than pages
        sm = ops.device_info()['sm_count']
        # ~3.5 CTAs per SM keep enough pages in flight and enough warps issuing (measured: 512 CTAs at B=32, H=16)
        want = self.attn_ctas if self.attn_ctas > 0 else int(3.46 * sm)
        n_ts = self.n_tsplit_override or max(1, min(8, st['max_pages'], math.ceil(want / (B * H))))
        sub['n_tsplit'] = n_ts
        sub['attn_ws'] = torch.zeros(ops.attn_decode_ws_bytes(B, H, n_ts) // 4 + 64, device=dev, dtype=torch.int32)
        return sub

    def _alloc_key(self, B: int, max_pages: int, max_new: int):
        return (B, max_pages, max_new, self._n_sub(B), self.n_tsplit_override, self.attn_ctas, self.page_permutation_seed,
                self.decode_gemm, self.precision)

    def _alloc(self, B: int, max_ctx: int, max_new: int):
        """Decode state for a batch: KV page pools, page table, counters, workspaces.  A request of the same shape as the
        previous one REUSES the buffers (and with them the captured step graph, see generate): only the counters are
        reset -- stale pages are never read because every kernel is bounded by seq_lens.  Re-allocating 1.9 GB of pools
        and re-capturing per call cost 5-270 ms of a 400 ms request (tools/e2e_phases.py: the graph's private pool is
        released with cudaFree when the old graph dies)."""
        dev, H, L = self.device, self.H, len(self.weights.layers)
        max_pages = (max_ctx + PAGE - 1) // PAGE + 1
        key = self._alloc_key(B, max_pages, max_new)
        st = self._state
        if st is not None and st.get('key') == key:
            st['seq_lens'].zero_(); st['audio_pos'].zero_(); st['last'].zero_(); st['sum_logprobs'].zero_()
            st['codes_out'].zero_()
            st['state'].zero_(); st['state'][:, 1].fill_(-1)
            for sub in st['subs']:
                if 'dg_ctr' in sub:
                    sub['dg_ctr'].zero_()
            return st
        st = {'B': B, 'max_pages': max_pages, 'max_new': max_new, 'key': key}
        st['pools'] = torch.zeros(L, B * max_pages, 2, H, PAGE, self.Dh, device=dev, dtype=self.cd)
        if self.page_permutation_seed is None:
            table = torch.arange(B * max_pages, dtype=torch.int32)
        else:
            table = torch.randperm(B * max_pages, generator=torch.Generator().manual_seed(self.page_permutation_seed)).to(torch.int32)
        st['block_table'] = table.view(B, max_pages).contiguous().to(dev)
        st['seq_lens'] = torch.zeros(B, device=dev, dtype=torch.int32)
        st['audio_pos'] = torch.zeros(B, device=dev, dtype=torch.int32)
        st['last'] = torch.zeros(B, device=dev, dtype=torch.int32)
        st['sum_logprobs'] = torch.zeros(B, device=dev, dtype=torch.float32)
        st['codes_out'] = torch.zeros(B, max_new, device=dev, dtype=torch.int32)
        n_sub = self._n_sub(B)
        # per sub-batch {step, stop_step, arrivals, any row running} (the last two: scratch of vb_ar_step_tail)
        st['state'] = torch.tensor([[0, -1, 0, 0]] * n_sub, device=dev, dtype=torch.int32)
        bounds = [(B * i) // n_sub for i in range(n_sub + 1)]
        self._state = None
        self._graph = None                 # release the old pools / graph before the new ones are allocated
        self._graph_key = None
        st['subs'] = [self._make_sub(st, bounds[i], bounds[i + 1], st['state'][i]) for i in range(n_sub)]
        while len(self._streams) < n_sub:
            self._streams.append(torch.cuda.Stream(device=dev))
        self._state = st
        return st

    # ------------------------------------------------------------------------------------------
    def _for_each_sub(self, fn):
        """Run fn(sub) for every sub-batch; more than one -> parallel streams forked from / joined to the current one
        (captured as parallel branches when the step is recorded into a CUDA graph)."""
        subs = self._state['subs']
        if len(subs) == 1:
            fn(subs[0])
            return
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        for sub, s in zip(subs, self._streams):
            s.wait_event(fork)
            with torch.cuda.stream(s):
                fn(sub)
                done = torch.cuda.Event()
                done.record(s)
            cur.wait_event(done)

    def _dg(self, sub: dict, kind: str, x: torch.Tensor, w: torch.Tensor, mode: int, **kw):
        """One vb_decode_gemm launch of the tc path; kind picks the counter array (one per GEMM shape)."""
        idx = ('qkv', 'o', 'f1', 'f2', 'lg').index(kind)
        ops.decode_gemm(x, w, mode, ws=sub['dg_ws'], counters=sub['dg_ctr'][idx], **kw)

    def _tail(self, sub: dict, samp: dict, uniforms: torch.Tensor | None, eos: int):
        """sample -> bookkeeping -> next step's input row (fp32, bf16 copy, row statistics) in one launch."""
        B = sub['B']
        if uniforms is not None:
            uniforms = uniforms[sub['b0']:sub['b0'] + B].contiguous()
        ops.ar_step_tail(sub['lg'], self.V, temperature=samp['temperature'], top_k=samp['top_k'], top_p=samp['top_p'],
                         uniforms=uniforms, seed=self._seed_dev, row_offset=sub['b0'], last=sub['last'],
                         sum_logprobs=sub['sum_logprobs'], codes_out=sub['codes_out'], seq_lens=sub['seq_lens'],
                         audio_pos=sub['audio_pos'], state=sub['state'], eos=eos, table=self.aud_table[0], pe=self.pe_a,
                         x=sub['x'], xb=sub['xb'], stats=sub['stats'])

    def _logits_sample_book(self, sub: dict, x_rows: torch.Tensor, samp: dict, uniforms: torch.Tensor | None, eos: int,
                            logits_done: bool = False):
        """x_rows (B,d) fp32 final hidden rows of the sub-batch -> logits -> sample -> bookkeeping (-> next input row)."""
        B, V = sub['B'], self.V
        if self._lean_ok(sub):
            ops.linear_decode_rows_ln(x_rows, self.wproj, sub['lg'])                  # plain cast on load (no final norm, K-2)
            self._tail(sub, samp, uniforms, eos)
            return
        if self._tc_ok(sub):
            if not logits_done:                 # first token: the prefill's last hidden rows, cast to bf16
                ops.residual_layernorm(x_rows, None, None, sub['h'])
                self._dg(sub, 'lg', sub['h'], self.wproj, ops.DG_PLAIN, y32=sub['lg'])
            self._tail(sub, samp, uniforms, eos)
            return
        if self.precision == 'bf16':
            if not logits_done:
                ops.residual_layernorm(x_rows, None, None, sub['h'])                  # cast to bf16 (no final norm, K-2)
                ops.linear_decode(sub['h'], self.wproj, sub['p_lg'], B * V, 32)
            lg, n_part, pstride = sub['p_lg'], sub['ns']['lg'], B * V
        else:
            ops.linear(x_rows, self.wproj, out=sub['lg'])
            lg, n_part, pstride = sub['lg'], 1, 0
        if uniforms is not None:
            uniforms = uniforms[sub['b0']:sub['b0'] + B].contiguous()
        ops.sample(lg, n_part, pstride, V, B, V, temperature=samp['temperature'], top_k=samp['top_k'],
                   top_p=samp['top_p'], out_tok=sub['sample'], out_logprob=sub['logprob'], uniforms=uniforms,
                   seed=samp['seed'], step_ptr=sub['state'], row_offset=sub['b0'])
        ops.ar_bookkeeping(sub['sample'], sub['logprob'], sub['last'], sub['sum_logprobs'], sub['codes_out'],
                           sub['seq_lens'], sub['audio_pos'], sub['state'], eos)

    def first_token(self, samp: dict, uniforms: torch.Tensor | None, eos: int):
        """Sample the first generated token of every sequence from the prefill's last hidden rows."""
        st = self._state
        self._seed_dev.fill_(_as_i64(samp['seed']))
        self._for_each_sub(lambda sub: self._logits_sample_book(
            sub, st['x_last'][sub['b0']:sub['b0'] + sub['B']], samp, uniforms, eos))

    def decode_step(self, samp: dict, uniforms: torch.Tensor | None, eos: int):
        """One decode step of the whole batch (all sub-batches)."""
        self._for_each_sub(lambda sub: self._decode_step(sub, samp, uniforms, eos))

    def _decode_step(self, sub: dict, samp: dict, uniforms: torch.Tensor | None, eos: int):
        st = self._state
        B, d, H, Dh, F = sub['B'], self.d, self.H, self.Dh, self.weights.F
        x = sub['x']
        layers = self.weights.layers
        if self._lean_ok(sub):
            # the step's input row x was written by the previous launch of vb_ar_step_tail (first_token or the last step)
            qkv32 = sub['r_qkv'][0]
            for li, L in enumerate(layers):
                g, b, eps = L['norm1']
                # Layer 0 releases the attention kernel only after its own dependency has resolved (late trigger): attention
                # reads seq_lens and KV pages BEFORE it waits, and the previous step's bookkeeping must have finished by then.
                # Deeper layers release it at once -- seq_lens is final since the step began and the layer's pages were
                # appended a whole step ago -- so its page copies are in flight while the QKV GEMM still runs.
                ops.linear_decode_rows_ln(x, L['wqkv'], qkv32, gamma=g[0], beta=b[0], eps=eps,
                                          flags=ops.FLAG_LATE_TRIGGER if (li == 0 or self.qkv_late_all) else 0)
                ops.attn_decode_paged(qkv32, 1, 0, st['pools'][li], sub['block_table'], sub['seq_lens'], sub['o'],
                                      B, H, Dh, sub['n_tsplit'], sub['attn_ws'],
                                      ops.FLAG_PREFETCH_KV | self.attn_flags | (self.attn_late if B >= 4 else 0))
                ops.linear_decode_rows(sub['o'], L['wo'], x, bias=L['bo'], residual=True)
                g, b, eps = L['norm2']
                ops.linear_decode_rows_ln(x, L['w1'], sub['f'], gamma=g[0], beta=b[0], eps=eps, bias=L['b1'], gelu=True)
                ops.linear_decode_rows(sub['f'], L['w2'], x, bias=L['b2'], residual=True, want_split=0)
            self._logits_sample_book(sub, x, samp, uniforms, eos)
            return
        if self._tc_ok(sub):
            # Five launches per layer (csrc/gemm_decode_tc.cu): every GEMM reduces its split-K partials itself; norm1 / norm2
            # are folded into the QKV / FFN1 GEMMs (pre-scaled weights + row statistics written by the producer of the rows);
            # out-proj and FFN2 add into the fp32 residual stream and leave its bf16 copy + statistics for the next GEMM.
            dg, xb, stats = sub['dg'], sub['xb'], sub['stats']
            ch_o, ch_f2 = dg['o']['tiles'], dg['f2']['tiles']
            for li, L in enumerate(layers):
                eps1, eps2 = L['norm1'][2], L['norm2'][2]
                self._dg(sub, 'qkv', xb, L['wqkv_s'], ops.DG_LN, bias=L['b_qkv'], colsum=L['c_qkv'], stats_in=stats,
                         n_chunks_in=1 if li == 0 else ch_f2, eps=eps1, y32=sub['qkv32'],
                         flags=ops.FLAG_LATE_TRIGGER if (li == 0 or self.qkv_late_all) else 0)
                ops.attn_decode_paged(sub['qkv32'], 1, 0, st['pools'][li], sub['block_table'], sub['seq_lens'], sub['o'],
                                      B, H, Dh, sub['n_tsplit'], sub['attn_ws'],
                                      ops.FLAG_PREFETCH_KV | self.attn_flags | self.attn_late_splitk)
                self._dg(sub, 'o', sub['o'], L['wo'], ops.DG_RESIDUAL, bias=L['bo'], xres=x, y16=xb, stats_out=stats,
                         flags=ops.FLAG_LATE_TRIGGER if self.dg_late & 1 else 0)
                self._dg(sub, 'f1', xb, L['w1_s'], ops.DG_LN_GELU, bias=L['b_1'], colsum=L['c_1'], stats_in=stats,
                         n_chunks_in=ch_o, eps=eps2, y16=sub['f'], flags=ops.FLAG_LATE_TRIGGER if self.dg_late & 2 else 0)
                self._dg(sub, 'f2', sub['f'], L['w2'], ops.DG_RESIDUAL, bias=L['b2'], xres=x, y16=xb, stats_out=stats,
                         flags=ops.FLAG_LATE_TRIGGER if self.dg_late & 4 else 0)
            self._dg(sub, 'lg', xb, self.wproj, ops.DG_PLAIN, y32=sub['lg'])            # no final norm (K-2)
            self._logits_sample_book(sub, x, samp, uniforms, eos, logits_done=True)
            return
        ops.embed_sum_pe(sub['last'].view(B, 1, 1), self.aud_table, self.pe_a, x, pos_b=sub['audio_pos'])
        if self.precision == 'bf16':
            ns = sub['ns']
            for li, L in enumerate(layers):
                g, b, eps = L['norm1']
                if li == 0:
                    ops.residual_layernorm(x, g[0], b[0], sub['h'], eps=eps)
                else:
                    ops.residual_layernorm(x, g[0], b[0], sub['h'], part=sub['p_f2'], n_part=ns['f2'],
                                           part_stride=B * d, bias=layers[li - 1]['b2'], eps=eps)
                ops.linear_decode(sub['h'], L['wqkv'], sub['p_qkv'], B * 3 * d, 32, ops.FLAG_LATE_TRIGGER)
                ops.attn_decode_paged(sub['p_qkv'], ns['qkv'], B * 3 * d, st['pools'][li], sub['block_table'],
                                      sub['seq_lens'], sub['o'], B, H, Dh, sub['n_tsplit'], sub['attn_ws'],
                                      ops.FLAG_PREFETCH_KV | self.attn_flags | self.attn_late_splitk)
                ops.linear_decode(sub['o'], L['wo'], sub['p_o'], B * d, 32)
                g, b, eps = L['norm2']
                ops.residual_layernorm(x, g[0], b[0], sub['h'], part=sub['p_o'], n_part=ns['o'], part_stride=B * d,
                                       bias=L['bo'], eps=eps)
                ops.linear_decode(sub['h'], L['w1'], sub['p_f1'], B * F, 32)
                ops.reduce_bias_act(sub['p_f1'], ns['f1'], B * F, L['b1'], True, sub['f'])
                ops.linear_decode(sub['f'], L['w2'], sub['p_f2'], B * d, 32)
            ops.residual_layernorm(x, None, None, None, part=sub['p_f2'], n_part=ns['f2'], part_stride=B * d,
                                   bias=layers[-1]['b2'])
        else:
            for li, L in enumerate(layers):
                g, b, eps = L['norm1']
                ops.residual_layernorm(x, g[0], b[0], sub['h'], eps=eps)
                ops.linear(sub['h'], L['wqkv'], out=sub['qkv'])
                ops.attn_decode_paged(sub['qkv'], 1, 0, st['pools'][li], sub['block_table'], sub['seq_lens'], sub['o'],
                                      B, H, Dh, sub['n_tsplit'], sub['attn_ws'])
                ops.linear(sub['o'], L['wo'], L['bo'], residual=x, out=x)
                g, b, eps = L['norm2']
                ops.residual_layernorm(x, g[0], b[0], sub['h'], eps=eps)
                ops.linear(sub['h'], L['w1'], L['b1'], gelu=True, out=sub['f'])
                ops.linear(sub['f'], L['w2'], L['b2'], residual=x, out=x)
        self._logits_sample_book(sub, x, samp, uniforms, eos)

    @property
    def last_phases(self) -> dict:
        """Device time of the last generate() call split into prefill (+ first token) and the decode loop, from CUDA events
        recorded in stream order (synchronises on the last one)."""
        if self._phase_events is None:
            return {}
        ev = self._phase_events
        ev[2].synchronize()
        return {'ar_prefill_ms': ev[0].elapsed_time(ev[1]), 'ar_decode_ms': ev[1].elapsed_time(ev[2])}

    def step_logits(self) -> torch.Tensor:
        """(B, V) fp32 logits of the most recent first_token / decode_step, for parity tests: where the logits GEMM leaves
        split-K slices they are summed in index order, exactly as the sampling kernel does."""
        rows = []
        for sub in self._state['subs']:
            if 'ns' not in sub or self._lean_ok(sub):
                rows.append(sub['lg'].clone())
            else:
                acc = sub['p_lg'][0].clone()
                for s in range(1, sub['ns']['lg']):
                    acc += sub['p_lg'][s]
                rows.append(acc)
        return torch.cat(rows, 0)

    def launches_per_step(self) -> int:
        """Kernel launches of one decode step (all sub-batches) -- the bench's gpu_launches claim."""
        L = len(self.weights.layers)
        total = 0
        for sub in self._state['subs']:
            if self._lean_ok(sub) or self._tc_ok(sub):
                total += 5 * L + 1 + 1              # 5 per layer, logits, sample + bookkeeping + next input row
            elif self.precision == 'bf16':
                total += 1 + 8 * L + 1 + 2 + 2      # embed, 8 per layer, x-update, cast + logits, sample + bookkeeping
            else:
                total += 1 + 7 * L + 1 + 2
        return total

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def prefill(self, tokens: torch.Tensor, codes: torch.Tensor, *, code_lens=None, max_new: int):
        """tokens (B,Tx), codes (B,P) incl. BOS.  Lays every sequence out as [text | audio] with the audio part
        starting at column Tx (the reference's layout, valle_ar.py:147).  Ragged PROMPTS pass code_lens.  Ragged TEXT has no
        length argument on purpose: the reference never masks text padding (SURVEY K-4: padded phoneme positions, token id 0,
        are attended), so a batch of utterances with different text lengths is decoded with its padding attended, exactly as
        the reference's teacher-forced step does; callers who want per-utterance text masking must batch equal lengths."""
        dev = self.device
        B, Tx = tokens.shape
        P = codes.shape[1]
        S = Tx + P
        st = self._alloc(B, S + max_new, max_new)
        tok_i = _i32(tokens, dev).view(B, Tx, 1)
        cod_i = _i32(codes, dev).view(B, P, 1)
        x = torch.empty(B * S, self.d, device=dev, dtype=torch.float32)
        # embedding + PE (valle_ar.py:127-139) and layer 0's norm1 (modules.py:271) in one kernel per segment
        fuse = self.fused_embed_norm and self.runner.fused_first_norm_ok()
        nrm = dict(norm_y=self.runner.first_norm_buffer(B * S, dev), **self.runner.first_norm_args(0)) if fuse else {}
        ops.embed_sum_pe(tok_i, self.tok_table, self.pe_t, x, out_rows_per_batch=S, out_row_offset=0, **nrm)
        ops.embed_sum_pe(cod_i, self.aud_table, self.pe_a, x, out_rows_per_batch=S, out_row_offset=Tx, **nrm)
        xl = torch.full((B,), Tx, device=dev, dtype=torch.int32)
        cl = torch.full((B,), P, device=dev, dtype=torch.int32) if code_lens is None else _i32(code_lens, dev)
        kv_lens = (xl + cl).contiguous()
        self.runner.forward(x, B, S, mask_mode=MASK_PREFIX_LM, x_lens=xl, kv_lens=kv_lens, kv_pools=st['pools'],
                            block_table=st['block_table'], first_norm_done=fuse)
        st['seq_lens'].copy_(kv_lens - 1)
        st['audio_pos'].copy_(cl - 1)
        rows = x.view(B, S, self.d)
        idx = (kv_lens - 1).long()
        st['x_last'] = rows[torch.arange(B, device=dev), idx].contiguous()
        st['last'].copy_(cod_i.view(B, P)[torch.arange(B, device=dev), (cl - 1).long()])
        return st

    @torch.no_grad()
    def generate(self, tokens: torch.Tensor, codes: torch.Tensor, *, max_new: int, top_k: int, top_p: float,
                 temperature: float, code_lens=None, uniforms: torch.Tensor | None = None, seed: int = 0,
                 ignore_eos: bool = False, use_graph: bool = True, poll_every: int = 16):
        """Returns (codes_out (B,n_steps) int32, sum_logprobs (B,), n_steps).  Semantics: valle_ar.py:141-171."""
        samp = {'temperature': temperature, 'top_k': top_k, 'top_p': top_p, 'seed': seed}
        eos = -1 if ignore_eos else self.cfg.num_audio_tokens
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        with _nvtx('valle_b200.ar.prefill'):
            st = self.prefill(tokens, codes, code_lens=code_lens, max_new=max_new)
            self.first_token(samp, None if uniforms is None else uniforms[0], eos)
        ev[1].record()
        graph = None
        step = 1
        if use_graph and uniforms is None and max_new > 2:
            # the captured step bakes in the state's buffers and the sampling scalars: reuse it while they are the same
            # (the seed too where the separate sampling kernel takes it as a launch argument; vb_ar_step_tail reads it from memory)
            baked_seed = any(not (self._lean_ok(s_) or self._tc_ok(s_)) for s_ in st['subs'])
            gkey = (st['key'], temperature, top_k, top_p, seed if baked_seed else None, eos)
            if self._graph is not None and self._graph_key == gkey:
                graph = self._graph
            else:
                self._graph = None
                self.decode_step(samp, None, eos)           # warm-up (also loads modules, sets func attributes)
                step += 1
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self.decode_step(samp, None, eos)
                # capture does not execute: the captured step still has to be replayed for this position
                self._graph, self._graph_key = graph, gkey
        else:
            self._graph, self._graph_key = None, None
        if _NVTX:
            torch.cuda.nvtx.range_push('valle_b200.ar.decode')
        while step < max_new:
            if graph is not None:
                graph.replay()
                self.replays += 1
            else:
                self.decode_step(samp, None if uniforms is None else uniforms[step], eos)
            step += 1
            if not ignore_eos and (step % poll_every == 0):
                if int(st['state'][:, 1].min().item()) >= 0:      # every sub-batch has seen all of its rows stop
                    break
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        ev[2].record()
        self._phase_events = ev
        state = st['state'].tolist()
        s_now = state[0][0]
        stops = [s[1] for s in state]
        s_stop = max(stops) if min(stops) >= 0 else -1
        n = s_stop if s_stop >= 0 else min(s_now, max_new)
        return st['codes_out'][:, :n].clone(), st['sum_logprobs'].clone(), n      # the state's buffers are reused by the next call


class NARDecoder:
    def __init__(self, model, precision: str):
        cfg = model.config
        assert cfg.norm == 'AdaptiveLayerNorm', 'ValleNAR conditions on the stage through AdaptiveLayerNorm'
        self.cfg = cfg
        self.precision = precision
        self.cd = _cdtype(precision)
        self.device = model.device
        self.H, self.d, self.Q = cfg.n_heads, cfg.d_model, cfg.num_quantizers
        stage_embs = [m.weight for m in model.stage_embs]
        self.weights = StackWeights(model.transformer, cfg.norm, precision, stage_embs)
        self.runner = StackRunner(self.weights, self.H)
        self.tok_table = model.tokens_emb.weight.detach().float().unsqueeze(0).contiguous()
        self.code_tables = torch.stack([m.weight.detach().float() for m in model.codes_embs]).contiguous()
        self.pe_t = model.tokens_position_emb.pe.detach().float().reshape(-1, self.d).contiguous()
        self.pe_a = model.audio_position_emb.pe.detach().float().reshape(-1, self.d).contiguous()
        self.wproj = [m.weight.detach().to(self.cd).contiguous() for m in model.proj_layers]
        self.fused_embed_norm = True   # embedding-sum + PE + the stage's first AdaLN in one kernel (tests: A/B)
        self.fused_argmax = True       # bf16 stages of >= 1024 target rows: logits GEMM + pick / draw in one kernel (tests: A/B)

    @torch.no_grad()
    def generate(self, prompt_tokens: torch.Tensor, prompt_codes: torch.Tensor, target_tokens: torch.Tensor,
                 first_layer: torch.Tensor, *, greedy: bool = True, temperature: float = 1.0, seed: int = 0,
                 return_logits: bool = False, use_tc_attention: bool | None = None,
                 target_lens: torch.Tensor | None = None):
        """Batched stages 2..Q.  prompt_tokens (B,Tp), prompt_codes (B,Tc,Q), target_tokens (B,Tt),
        first_layer (B,T) -> (B,T,Q) int64.  ``target_lens`` (B,) marks ragged targets (extension): keys past
        Tx+Tc+target_lens[b] are not attended and rows past it are padding in the result."""
        dev, d, Q = self.device, self.d, self.Q
        B, Tc, Qc = prompt_codes.shape
        assert Qc == Q
        tokens = torch.cat([prompt_tokens, target_tokens], dim=1)
        Tx, T = tokens.shape[1], first_layer.shape[1]
        S = Tx + Tc + T
        tok_i = _i32(tokens, dev).view(B, Tx, 1)
        ids = torch.zeros(B, Tc + T, Q, device=dev, dtype=torch.int32)
        ids[:, :Tc] = prompt_codes.to(dev)
        ids[:, Tc:, 0] = first_layer.to(dev)
        x = torch.empty(B * S, d, device=dev, dtype=torch.float32)
        V = self.cfg.num_audio_tokens
        sampled = torch.empty(B * T, device=dev, dtype=torch.int32)
        step = torch.zeros(1, device=dev, dtype=torch.int32)
        trace = []
        kv_lens = None
        keys = None
        if target_lens is not None:
            kv_lens = (_i32(target_lens, dev) + (Tx + Tc)).contiguous()
        for n in range(1, Q):
            if _NVTX:
                torch.cuda.nvtx.range_push(f'valle_b200.nar.stage{n}')
            # embedding sums + PE (valle_nar.py:140-152) and the stage's first AdaLN (modules.py:93-99, :271) in one kernel per
            # segment: the residual rows are written once, layer 0's norm1 comes out of the same registers
            fuse = self.fused_embed_norm and self.runner.fused_first_norm_ok()
            nrm = dict(norm_y=self.runner.first_norm_buffer(B * S, dev), **self.runner.first_norm_args(n - 1)) if fuse else {}
            ops.embed_sum_pe(tok_i, self.tok_table, self.pe_t, x, out_rows_per_batch=S, out_row_offset=0, **nrm)
            ops.embed_sum_pe(ids, self.code_tables, self.pe_a, x, t_split=Tc, nq_a=Q, nq_b=n,
                             out_rows_per_batch=S, out_row_offset=Tx, **nrm)
            self.runner.forward(x, B, S, mask_mode=MASK_NONE, kv_lens=kv_lens, stage=n - 1,
                                use_tc_attention=use_tc_attention, first_norm_done=fuse)
            tgt = x.view(B, S, d)[:, Tx + Tc:].reshape(B * T, d)          # strided gather (memory plumbing)
            if self.fused_argmax and self.precision == 'bf16' and not return_logits and ops.linear_argmax_ok(B * T, V):
                # fused logits + pick (valle_nar.py:157-160): the (B T x 1024) fp32 logits are never written; the row maxima --
                # of the logits (greedy) or of logits / temperature + Gumbel noise (the reference's Categorical draw) -- leave
                # the GEMM epilogue as packed keys and land in column n of the code tensor
                hb = torch.empty(B * T, d, device=dev, dtype=self.cd)
                ops.residual_layernorm(tgt, None, None, hb)
                if keys is None:
                    keys = torch.zeros(B * T, device=dev, dtype=torch.int64)
                ops.linear_argmax(hb, self.wproj[n - 1], keys, ids[:, Tc:, n], rows_per_batch=T, batch_stride=(Tc + T) * Q,
                                  row_stride=Q, temperature=None if greedy else temperature, seed=seed, step=n)
                if _NVTX:
                    torch.cuda.nvtx.range_pop()
                continue
            if self.precision == 'bf16':
                hb = torch.empty(B * T, d, device=dev, dtype=self.cd)
                ops.residual_layernorm(tgt, None, None, hb)
                logits = ops.linear(hb, self.wproj[n - 1], out_dtype=torch.float32)
            else:
                logits = ops.linear(tgt, self.wproj[n - 1])
            if return_logits:
                trace.append(logits.view(B, T, V).clone())
            step.fill_(n)
            if greedy:
                ops.sample(logits, 1, 0, V, B * T, V, temperature=1.0, top_k=1, top_p=1.0, out_tok=sampled)
            else:
                ops.sample(logits, 1, 0, V, B * T, V, temperature=temperature, top_k=0, top_p=1.0, out_tok=sampled,
                           seed=seed, step_ptr=step)
            ids[:, Tc:, n] = sampled.view(B, T)
            if _NVTX:
                torch.cuda.nvtx.range_pop()
        out = ids[:, Tc:].long()
        return (out, trace) if return_logits else out
