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

    def forward(self, x: torch.Tensor, B: int, S: int, *, mask_mode: int, x_lens=None, kv_lens=None, stage: int = 0,
                kv_pools: torch.Tensor | None = None, block_table: torch.Tensor | None = None,
                use_tc_attention: bool | None = None) -> torch.Tensor:
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
        self.page_permutation_seed = None      # tests: scatter the logical pages over the pool
        # tunables (env overrides are for experiments; the defaults are the measured best)
        self.use_chain = os.environ.get('VALLE_B200_CHAIN', '0') != '0'
        self.use_fused = os.environ.get('VALLE_B200_FUSED', '0') != '0'
        # cluster size along K of the fused decode GEMMs (0 = fill the SMs), csrc/gemm_decode_fused.cu
        self.fused_cluster = {'qkv': 0, 'o': 0, 'f1': 0, 'f2': 0, 'lg': 0}
        for k, v in [kv.split('=') for kv in os.environ.get('VALLE_B200_FUSED_CLUSTER', '').split(';') if kv]:
            self.fused_cluster[k] = int(v)
        # decode GEMM form: 'rows' = full-K mma.sync kernel with fused epilogues (csrc/gemm_decode_mma.cu, sub-batches of
        # <= 32 sequences), 'splitk' = tcgen05 swap-AB slices + reduce kernels (csrc/gemm_tc.cu), 'auto' = by batch size,
        # or five letters [rs] for qkv, out-proj, FFN1, FFN2, logits (see _mix)
        self.decode_gemm = os.environ.get('VALLE_B200_DECODE_GEMM', 'auto')
        self.rows_qkv_split = int(os.environ.get('VALLE_B200_ROWS_QKV_SPLIT', '1'))
        # lean path, >= 4 sequences: the attention kernel releases its successors only after its own wait (see
        # csrc/attn_decode.cu; measured -2 % step time at B = 4..8, +2 % at B = 1, tools/step_breakdown.py)
        self.attn_late = ops.FLAG_LATE_TRIGGER if os.environ.get('VALLE_B200_ATTN_LATE', '1') != '0' else 0
        self.attn_late_splitk = ops.FLAG_LATE_TRIGGER if os.environ.get('VALLE_B200_ATTN_LATE_SPLITK', '0') != '0' else 0   # A/B
        self.qkv_late_all = os.environ.get('VALLE_B200_QKV_LATE_ALL', '0') != '0'      # A/B: late PDL trigger in every layer
        self.n_sub_override = int(os.environ.get('VALLE_B200_SUBBATCH', '0'))
        self.attn_ctas = int(os.environ.get('VALLE_B200_ATTN_CTAS', '0'))
        self.n_tsplit_override = 0             # tests: pin the flash-decoding split
        # percentage range of every sequence's cached pages of the NEXT layer pulled into L2 while the GEMM chain runs
        pf = os.environ.get('VALLE_B200_KV_PREFETCH', '0,0').split(',')
        self.kv_prefetch = (int(pf[0]), int(pf[1]))
        self._pf_stream = None
        self.attn_flags = ops.FLAG_ATTN_SIMT if os.environ.get('VALLE_B200_ATTN_SIMT', '0') != '0' else 0

    # ------------------------------------------------------------------------------------------
    def _n_sub(self, B: int) -> int:
        if self.precision != 'bf16':
            return 1
        if self.n_sub_override > 0:
            return max(1, min(self.n_sub_override, B))
        return 1

    def _chain_ok(self, sub: dict) -> bool:
        # Concurrent chain kernels (one per sub-batch) spin on grid barriers and cannot share an SM: two of them could
        # each hold part of the GPU and wait for the rest forever, so the chain runs only when the batch is not split.
        return (self.precision == 'bf16' and self.use_chain and sub['B'] <= 64 and len(self._state['subs']) == 1)

    def _mix(self, sub: dict) -> dict:
        """Form of each decode GEMM for this sub-batch: 'r' rows (mma.sync, full K per CTA), 's' tcgen05 split-K slices.
        In the rows form every CTA reads the whole (B, K) activation matrix from L2 -- 0.3 us at B = 1, 1.4 us at B = 32
        (tools/rows_timeline.py) -- so 'auto' uses it (as the lean path) up to 8 sequences and slices above; every mix
        was measured with tools/layer_chain.py and tools/ab_mix.sh."""
        dg = self.decode_gemm
        if dg in ('auto', 'lean'):      # small batches take the lean path (_lean_ok) before this is consulted
            dg = 'splitk'
        if dg == 'rows':
            dg = 'rrrrr'
        elif dg == 'splitk':
            dg = 'sssss'
        assert len(dg) == 5 and set(dg) <= {'r', 's'}, 'VALLE_B200_DECODE_GEMM: auto | rows | splitk | 5 x [rs] (qkv,o,f1,f2,logits)'
        return dict(zip(('qkv', 'o', 'f1', 'f2', 'lg'), dg))

    def _lean_ok(self, sub: dict) -> bool:
        """Five kernels per layer (small batch): LayerNorm on load inside the QKV / FFN1 GEMMs, residual adds and GELU in
        the epilogues, FFN2 with its whole K = F in one CTA (csrc/gemm_decode_mma.cu).  Every CTA reads the whole fp32
        residual matrix, which is what limits it to <= 8 sequences."""
        dg = self.decode_gemm
        if dg not in ('auto', 'lean') or self.precision != 'bf16' or self.use_fused or self.use_chain:
            return False
        return (sub['B'] <= 8 and self.d in (256, 512, 1024) and self.weights.F % 256 == 0
                and ops.linear_decode_rows_splits(self.weights.F, 0, sub['B']) == 1)

    def _rows_ok(self, sub: dict) -> bool:
        """The mixed decode path (any GEMM in rows form) applies to this sub-batch."""
        return (self.precision == 'bf16' and not self.use_fused and not self.use_chain and 'nr' in sub
                and 'r' in self._mix(sub).values())

    def _fused_ok(self, sub: dict) -> bool:
        return self.precision == 'bf16' and self.use_fused and sub['B'] <= 64 and self.d % 64 == 0

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
        if self.precision == 'bf16':
            ms = 32
            ns = {k: ops.linear_decode_splits(n, kk, ms) for k, (n, kk) in
                  {'qkv': (3 * d, d), 'o': (d, d), 'f1': (F, d), 'f2': (d, F), 'lg': (V, d)}.items()}
            sub['ns'] = ns
            sub['h'] = torch.zeros(B, d, device=dev, dtype=self.cd)
            sub['o'] = torch.zeros(B, d, device=dev, dtype=self.cd)
            sub['f'] = torch.zeros(B, F, device=dev, dtype=self.cd)
            sub['p_qkv'] = torch.zeros(ns['qkv'], B, 3 * d, device=dev, dtype=torch.float32)
            sub['p_o'] = torch.zeros(ns['o'], B, d, device=dev, dtype=torch.float32)
            sub['p_f1'] = torch.zeros(ns['f1'], B, F, device=dev, dtype=torch.float32)
            sub['p_f2'] = torch.zeros(ns['f2'], B, d, device=dev, dtype=torch.float32)
            sub['p_lg'] = torch.zeros(ns['lg'], B, V, device=dev, dtype=torch.float32)
            sub['gbar'] = torch.zeros(64, device=dev, dtype=torch.int32)     # grid-barrier counter of the chain kernel
            sub['qkv32'] = torch.zeros(B, 3 * d, device=dev, dtype=torch.float32)
            if self.d % 256 == 0 and F % 256 == 0 and B <= 32:
                nr = {'qkv': ops.linear_decode_rows_splits(d, self.rows_qkv_split, B), 'f2': ops.linear_decode_rows_splits(F, 1, B)}
                sub['nr'] = nr
                sub['r_qkv'] = torch.zeros(nr['qkv'], B, 3 * d, device=dev, dtype=torch.float32)
                sub['r_f2'] = torch.zeros(nr['f2'], B, d, device=dev, dtype=torch.float32)
            sub['lg'] = torch.zeros(B, V, device=dev, dtype=torch.float32)
        else:
            sub['h'] = torch.zeros(B, d, device=dev, dtype=torch.float32)
            sub['qkv'] = torch.zeros(B, 3 * d, device=dev, dtype=torch.float32)
            sub['o'] = torch.zeros(B, d, device=dev, dtype=torch.float32)
            sub['f'] = torch.zeros(B, F, device=dev, dtype=torch.float32)
            sub['lg'] = torch.zeros(B, V, device=dev, dtype=torch.float32)
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
                self.decode_gemm, self.use_chain, self.use_fused, self.rows_qkv_split, self.precision)

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
            st['state'][:, 0].zero_(); st['state'][:, 1].fill_(-1)
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
        st['state'] = torch.tensor([[0, -1]] * n_sub, device=dev, dtype=torch.int32)     # per sub-batch {step, stop_step}
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

    def _logits_sample_book(self, sub: dict, x_rows: torch.Tensor, samp: dict, uniforms: torch.Tensor | None, eos: int,
                            logits_done: bool = False):
        """x_rows (B,d) fp32 final hidden rows of the sub-batch -> logits -> sample -> bookkeeping."""
        B, V = sub['B'], self.V
        if self._lean_ok(sub):
            ops.linear_decode_rows_ln(x_rows, self.wproj, sub['lg'])                  # plain cast on load (no final norm, K-2)
            lg, n_part, pstride = sub['lg'], 1, 0
        elif self._rows_ok(sub):
            if not logits_done:
                ops.residual_layernorm(x_rows, None, None, sub['h'])                  # cast to bf16 (no final norm, K-2)
            if self._mix(sub)['lg'] == 'r':
                ops.linear_decode_rows(sub['h'], self.wproj, sub['lg'])
                lg, n_part, pstride = sub['lg'], 1, 0
            else:
                ops.linear_decode(sub['h'], self.wproj, sub['p_lg'], B * V, 32)
                lg, n_part, pstride = sub['p_lg'], sub['ns']['lg'], B * V
        elif self._fused_ok(sub):
            ops.linear_decode_fused(x_rows, self.wproj, sub['lg'], cluster_k=self.fused_cluster['lg'])   # plain cast on load (no final norm, K-2)
            lg, n_part, pstride = sub['lg'], 1, 0
        elif self.precision == 'bf16':
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
        self._for_each_sub(lambda sub: self._logits_sample_book(
            sub, st['x_last'][sub['b0']:sub['b0'] + sub['B']], samp, uniforms, eos))

    def decode_step(self, samp: dict, uniforms: torch.Tensor | None, eos: int):
        """One decode step of the whole batch (all sub-batches)."""
        self._for_each_sub(lambda sub: self._decode_step(sub, samp, uniforms, eos))

    def _decode_step(self, sub: dict, samp: dict, uniforms: torch.Tensor | None, eos: int):
        st = self._state
        B, d, H, Dh, F = sub['B'], self.d, self.H, self.Dh, self.weights.F
        x = sub['x']
        ops.embed_sum_pe(sub['last'].view(B, 1, 1), self.aud_table, self.pe_a, x, pos_b=sub['audio_pos'])
        layers = self.weights.layers
        if self._lean_ok(sub):
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
        if self._rows_ok(sub):
            # Per GEMM one of two forms (self._mix(sub)): 'r' = rows kernel (csrc/gemm_decode_mma.cu: whole K inside one CTA,
            # so out-proj adds into the residual stream and FFN1 applies bias + GELU in the epilogue -- no reduce kernel; only
            # FFN2, K = F > 1024, leaves slices); 's' = tcgen05 swap-AB split-K slices (csrc/gemm_tc.cu) that the next
            # LayerNorm / GELU-reduce / attention kernel adds in index order.
            mix, nr, ns = self._mix(sub), sub['nr'], sub['ns']
            f2_direct = mix['f2'] == 'r' and nr['f2'] == 1
            if mix['f2'] == 'r':
                f2_part, f2_n = sub['r_f2'], nr['f2']
            else:
                f2_part, f2_n = sub['p_f2'], ns['f2']
            for li, L in enumerate(layers):
                g, b, eps = L['norm1']
                if li == 0 or f2_direct:
                    ops.residual_layernorm(x, g[0], b[0], sub['h'], eps=eps)
                else:
                    ops.residual_layernorm(x, g[0], b[0], sub['h'], part=f2_part, n_part=f2_n, part_stride=B * d,
                                           bias=layers[li - 1]['b2'], eps=eps)
                if mix['qkv'] == 'r':
                    ops.linear_decode_rows(sub['h'], L['wqkv'], sub['r_qkv'] if nr['qkv'] > 1 else sub['r_qkv'][0],
                                           want_split=self.rows_qkv_split, flags=ops.FLAG_LATE_TRIGGER)
                    qkv_part, qkv_n = sub['r_qkv'], nr['qkv']
                else:
                    ops.linear_decode(sub['h'], L['wqkv'], sub['p_qkv'], B * 3 * d, 32, ops.FLAG_LATE_TRIGGER)
                    qkv_part, qkv_n = sub['p_qkv'], ns['qkv']
                ops.attn_decode_paged(qkv_part, qkv_n, B * 3 * d, st['pools'][li], sub['block_table'],
                                      sub['seq_lens'], sub['o'], B, H, Dh, sub['n_tsplit'], sub['attn_ws'],
                                      ops.FLAG_PREFETCH_KV | self.attn_flags | self.attn_late_splitk)
                g, b, eps = L['norm2']
                if mix['o'] == 'r':
                    ops.linear_decode_rows(sub['o'], L['wo'], x, bias=L['bo'], residual=True)
                    ops.residual_layernorm(x, g[0], b[0], sub['h'], eps=eps)
                else:
                    ops.linear_decode(sub['o'], L['wo'], sub['p_o'], B * d, 32)
                    ops.residual_layernorm(x, g[0], b[0], sub['h'], part=sub['p_o'], n_part=ns['o'], part_stride=B * d,
                                           bias=L['bo'], eps=eps)
                if mix['f1'] == 'r':
                    ops.linear_decode_rows(sub['h'], L['w1'], sub['f'], bias=L['b1'], gelu=True)
                else:
                    ops.linear_decode(sub['h'], L['w1'], sub['p_f1'], B * F, 32)
                    ops.reduce_bias_act(sub['p_f1'], ns['f1'], B * F, L['b1'], True, sub['f'])
                if f2_direct:
                    ops.linear_decode_rows(sub['f'], L['w2'], x, bias=L['b2'], residual=True)
                elif mix['f2'] == 'r':
                    ops.linear_decode_rows(sub['f'], L['w2'], sub['r_f2'])
                else:
                    ops.linear_decode(sub['f'], L['w2'], sub['p_f2'], B * d, 32)
            if f2_direct:
                ops.residual_layernorm(x, None, None, sub['h'])
            else:       # last FFN2 slices + bias into the residual stream, and the bf16 cast for the logits projection
                ops.residual_layernorm(x, None, None, sub['h'], part=f2_part, n_part=f2_n, part_stride=B * d,
                                       bias=layers[-1]['b2'])
            self._logits_sample_book(sub, x, samp, uniforms, eos, logits_done=True)
            return
        if self._fused_ok(sub):
            # five dependent kernels per layer: split-K is reduced inside a thread-block cluster, LayerNorm runs on load
            # inside the QKV / FFN1 GEMMs, bias + residual / GELU in the epilogues (csrc/gemm_decode_fused.cu)
            cl = self.fused_cluster
            for li, L in enumerate(layers):
                g, b, eps = L['norm1']
                ops.linear_decode_fused(x, L['wqkv'], sub['qkv32'], gamma=g[0], beta=b[0], eps=eps, cluster_k=cl['qkv'],
                                        flags=ops.FLAG_LATE_TRIGGER)
                ops.attn_decode_paged(sub['qkv32'], 1, 0, st['pools'][li], sub['block_table'], sub['seq_lens'], sub['o'],
                                      B, H, Dh, sub['n_tsplit'], sub['attn_ws'], ops.FLAG_PREFETCH_KV | self.attn_flags | self.attn_late_splitk)
                ops.linear_decode_fused(sub['o'], L['wo'], x, bias=L['bo'], residual=True, cluster_k=cl['o'])
                g, b, eps = L['norm2']
                ops.linear_decode_fused(x, L['w1'], sub['f'], bias=L['b1'], gelu=True, gamma=g[0], beta=b[0], eps=eps,
                                        cluster_k=cl['f1'])
                ops.linear_decode_fused(sub['f'], L['w2'], x, bias=L['b2'], residual=True, cluster_k=cl['f2'])
            self._logits_sample_book(sub, x, samp, uniforms, eos)
            return
        if self._chain_ok(sub):
            # persistent chain kernels: everything between two attention kernels is ONE launch (csrc/decode_chain.cu)
            ns = sub['ns']
            g, b, eps = layers[0]['norm1']
            ops.decode_chain([ops.chain_ln(x, g[0], b[0], sub['h'], eps=eps),
                              ops.chain_gemm(sub['h'], layers[0]['wqkv'], sub['p_qkv'], B * 3 * d)], B, sub['gbar'])
            for li, L in enumerate(layers):
                ops.attn_decode_paged(sub['p_qkv'], ns['qkv'], B * 3 * d, st['pools'][li], sub['block_table'],
                                      sub['seq_lens'], sub['o'], B, H, Dh, sub['n_tsplit'], sub['attn_ws'],
                                      ops.FLAG_PREFETCH_KV | self.attn_flags | self.attn_late_splitk)
                g2, b2, eps2 = L['norm2']
                ph = [ops.chain_gemm(sub['o'], L['wo'], sub['p_o'], B * d),
                      ops.chain_ln(x, g2[0], b2[0], sub['h'], part=sub['p_o'], n_part=ns['o'], part_stride=B * d,
                                   bias=L['bo'], eps=eps2),
                      ops.chain_gemm(sub['h'], L['w1'], sub['p_f1'], B * F),
                      ops.chain_act(sub['p_f1'], ns['f1'], B * F, L['b1'], sub['f']),
                      ops.chain_gemm(sub['f'], L['w2'], sub['p_f2'], B * d)]
                if li + 1 < len(layers):
                    nxt = layers[li + 1]
                    g1, b1, eps1 = nxt['norm1']
                    ph += [ops.chain_ln(x, g1[0], b1[0], sub['h'], part=sub['p_f2'], n_part=ns['f2'], part_stride=B * d,
                                        bias=L['b2'], eps=eps1),
                           ops.chain_gemm(sub['h'], nxt['wqkv'], sub['p_qkv'], B * 3 * d)]
                else:       # no final norm (K-2): cast the hidden rows and run the logits projection
                    ph += [ops.chain_ln(x, None, None, sub['h'], part=sub['p_f2'], n_part=ns['f2'], part_stride=B * d,
                                        bias=L['b2']),
                           ops.chain_gemm(sub['h'], self.wproj, sub['p_lg'], B * self.V)]
                ops.decode_chain(ph, B, sub['gbar'])
            self._logits_sample_book(sub, x, samp, uniforms, eos, logits_done=True)
            return
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
                self._prefetch_next_kv(sub, li)
                ops.linear_decode(sub['o'], L['wo'], sub['p_o'], B * d, 32)
                g, b, eps = L['norm2']
                ops.residual_layernorm(x, g[0], b[0], sub['h'], part=sub['p_o'], n_part=ns['o'], part_stride=B * d,
                                       bias=L['bo'], eps=eps)
                ops.linear_decode(sub['h'], L['w1'], sub['p_f1'], B * F, 32)
                ops.reduce_bias_act(sub['p_f1'], ns['f1'], B * F, L['b1'], True, sub['f'])
                ops.linear_decode(sub['f'], L['w2'], sub['p_f2'], B * d, 32)
            ops.residual_layernorm(x, None, None, None, part=sub['p_f2'], n_part=ns['f2'], part_stride=B * d,
                                   bias=layers[-1]['b2'])
            self._prefetch_join(sub)
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

    def _prefetch_next_kv(self, sub: dict, li: int):
        """After layer li's attention has been launched: on a side stream (ordered after that attention), ask for the
        next layer's pages (layer 0 of the next step after the last layer).  Joined in _decode_step's caller."""
        lo, hi = self.kv_prefetch
        if hi <= lo or self.precision != 'bf16':
            return
        st = self._state
        if self._pf_stream is None:
            self._pf_stream = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        self._pf_stream.wait_event(ev)
        nxt = (li + 1) % len(self.weights.layers)
        with torch.cuda.stream(self._pf_stream):
            ops.kv_prefetch_l2(st['pools'][nxt], sub['block_table'], sub['seq_lens'], sub['B'], self.H, self.Dh, lo, hi)
        sub['_pf_used'] = True

    def _prefetch_join(self, sub: dict):
        if sub.pop('_pf_used', False):
            ev = torch.cuda.Event()
            ev.record(self._pf_stream)
            torch.cuda.current_stream().wait_event(ev)

    def launches_per_step(self) -> int:
        """Kernel launches of one decode step (all sub-batches) -- the bench's gpu_launches claim."""
        L = len(self.weights.layers)
        subs = self._state['subs']
        total = 0
        for sub in subs:
            if self._lean_ok(sub):
                total += 1 + 5 * L + 1 + 2          # embed, 5 per layer, logits, sample + bookkeeping
            elif self._rows_ok(sub):
                # embed, 7 per layer (+1 GELU-reduce when FFN1 leaves slices), final reduce + cast, logits, sample + bookkeeping
                total += 1 + (7 + (self._mix(sub)['f1'] == 's')) * L + 1 + 1 + 2
            elif self._fused_ok(sub):
                total += 1 + 5 * L + 1 + 2          # embed, 5 per layer, logits, sample + bookkeeping
            elif self._chain_ok(sub):
                total += 1 + 1 + 2 * L + 2          # embed, first chain, (attention + chain) per layer, sample, bookkeeping
            elif self.precision == 'bf16':
                total += 1 + 8 * L + 1 + 2 + 2      # embed, 8 per layer, x-update, cast + logits, sample + bookkeeping
                if self.kv_prefetch[1] > self.kv_prefetch[0]:
                    total += L                      # one L2 prefetch launch per layer (side stream)
            else:
                total += 1 + 7 * L + 1 + 2
        return total

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def prefill(self, tokens: torch.Tensor, codes: torch.Tensor, *, x_lens=None, code_lens=None, max_new: int):
        """tokens (B,Tx), codes (B,P) incl. BOS.  Lays every sequence out as [text | audio] with the audio part
        starting at column Tx (the reference's layout, valle_ar.py:147); ragged batches pass x_lens / code_lens
        (then text padding between x_lens[b] and Tx is attended exactly like the reference does, K-4)."""
        dev = self.device
        B, Tx = tokens.shape
        P = codes.shape[1]
        S = Tx + P
        st = self._alloc(B, S + max_new, max_new)
        tok_i = _i32(tokens, dev).view(B, Tx, 1)
        cod_i = _i32(codes, dev).view(B, P, 1)
        x = torch.empty(B * S, self.d, device=dev, dtype=torch.float32)
        ops.embed_sum_pe(tok_i, self.tok_table, self.pe_t, x, out_rows_per_batch=S, out_row_offset=0)
        ops.embed_sum_pe(cod_i, self.aud_table, self.pe_a, x, out_rows_per_batch=S, out_row_offset=Tx)
        xl = torch.full((B,), Tx, device=dev, dtype=torch.int32)
        cl = torch.full((B,), P, device=dev, dtype=torch.int32) if code_lens is None else _i32(code_lens, dev)
        kv_lens = (xl + cl).contiguous()
        self.runner.forward(x, B, S, mask_mode=MASK_PREFIX_LM, x_lens=xl, kv_lens=kv_lens, kv_pools=st['pools'],
                            block_table=st['block_table'])
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
        st = self.prefill(tokens, codes, code_lens=code_lens, max_new=max_new)
        self.first_token(samp, None if uniforms is None else uniforms[0], eos)
        graph = None
        step = 1
        if use_graph and uniforms is None and max_new > 2:
            # the captured step bakes in the state's buffers and the sampling scalars: reuse it while they are the same
            gkey = (st['key'], temperature, top_k, top_p, seed, eos)
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
        while step < max_new:
            if graph is not None:
                graph.replay()
            else:
                self.decode_step(samp, None if uniforms is None else uniforms[step], eos)
            step += 1
            if not ignore_eos and (step % poll_every == 0):
                if int(st['state'][:, 1].min().item()) >= 0:      # every sub-batch has seen all of its rows stop
                    break
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
        if target_lens is not None:
            kv_lens = (_i32(target_lens, dev) + (Tx + Tc)).contiguous()
        for n in range(1, Q):
            ops.embed_sum_pe(tok_i, self.tok_table, self.pe_t, x, out_rows_per_batch=S, out_row_offset=0)
            ops.embed_sum_pe(ids, self.code_tables, self.pe_a, x, t_split=Tc, nq_a=Q, nq_b=n,
                             out_rows_per_batch=S, out_row_offset=Tx)
            self.runner.forward(x, B, S, mask_mode=MASK_NONE, kv_lens=kv_lens, stage=n - 1,
                                use_tc_attention=use_tc_attention)
            tgt = x.view(B, S, d)[:, Tx + Tc:].reshape(B * T, d)          # strided gather (memory plumbing)
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
        out = ids[:, Tc:].long()
        return (out, trace) if return_logits else out
