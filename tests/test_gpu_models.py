"""Model-level parity on a B200, through the reference-facing API (valle.models.*) and the C ABI underneath:
  - fp32 validation mode: logits within 1e-5 (relative to the logit scale) of the CPU oracle / frozen reference
    vectors, greedy codec tokens bit-exact;
  - bf16 mode: logits within 1e-2 relative; tokens compared up to the first step whose reference top-1/top-2
    margin is below the bf16 noise (random-init margins go down to 1e-3).
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import valle2_b200  # noqa: E402
from oracle import synth  # noqa: E402
from oracle import valle_oracle as vo  # noqa: E402
from oracle.valle_oracle import OracleConfig  # noqa: E402

T = torch.from_numpy


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def make_cfg(oc: OracleConfig, tmp_path, **kw):
    from valle.config import ConfigValle
    fields = {k: getattr(oc, k) for k in OracleConfig.__dataclass_fields__}
    fields.update(kw)
    return ConfigValle(dropout=0.0, ckpt_path=tmp_path / 'c', log_path=tmp_path / 'l', **fields)


def build(kind, oc, tmp_path, seed):
    from valle.models import get_model_class
    model = get_model_class(kind)(make_cfg(oc, tmp_path)).eval()
    shapes = synth.ar_state_shapes(oc) if kind == 'ValleAR' else synth.nar_state_shapes(oc)
    sd = synth.synth_state_dict(shapes, seed)
    missing, unexpected = model.load_state_dict(sd, strict=True)
    return model.cuda(), sd


def oracle_nar_stage_logits(sd, oc, tokens, prompt_codes, columns, n):
    """Stage-n logits (T, 1024) of the repaired NAR loop (oracle.nar_generate, valle_nar.py:131-157) with the first n
    codebooks of the target GIVEN (columns (T, >=n)) instead of sampled stage by stage."""
    d = oc.d_model
    Tc, Q = prompt_codes.shape
    emb_prompt = torch.zeros(Tc, d)
    for j in range(Q):
        emb_prompt = emb_prompt + sd[f'codes_embs.{j}.word_embeddings.weight'][prompt_codes[:, j]]
    emb_out = torch.zeros(columns.shape[0], d)
    for j in range(n):
        emb_out = emb_out + sd[f'codes_embs.{j}.word_embeddings.weight'][columns[:, j]]
    x_tok = vo.add_pe(vo.embed(sd['tokens_emb.word_embeddings.weight'], tokens[None]), vo._pe_for(sd, 'tokens_position_emb.pe', d))
    x_aud = vo.add_pe(torch.cat([emb_prompt, emb_out], 0)[None], vo._pe_for(sd, 'audio_position_emb.pe', d))
    h, _ = vo.transformer(torch.cat([x_tok, x_aud], 1), sd, oc, embedding=sd[f'stage_embs.{n - 1}.word_embeddings.weight'])
    return h[0, tokens.shape[0] + Tc:] @ sd[f'proj_layers.{n - 1}.weight'].t()


@pytest.fixture(autouse=True)
def _restore_precision():
    prev = valle2_b200.get_precision()
    yield
    valle2_b200.set_precision(prev)


@pytest.mark.parametrize('precision,tol', [('fp32', 2e-5), ('bf16', 1e-2)])
def test_modules_vs_reference_vectors(golden, precision, tol, tmp_path):
    from valle.models.modules import (AdaptiveLayerNorm, FeedForward, MultiHeadAttention, PositionalEncoding,
                                      Transformer)
    valle2_b200.set_precision(precision)
    g = golden('modules')
    d, H = 64, 4
    x = T(g['mha_x']).cuda()
    S = x.shape[1]
    mha = MultiHeadAttention(d_model=d, n_heads=H)
    mha.load_state_dict(synth.synth_state_dict({'qkv.weight': (3 * d, d), 'out.weight': (d, d), 'out.bias': (d,)}, 7))
    mha = mha.cuda()
    causal = torch.triu(torch.ones(S, S), diagonal=1).cuda()
    y, (k, v) = mha(x, attn_mask=causal, padding_mask=T(g['mha_pad']).cuda(), use_cache=True)
    assert y.shape == x.shape and k.shape == (x.shape[0], H, S, d // H)
    assert rel_err(y, T(g['mha_y'])) < tol and rel_err(k, T(g['mha_k'])) < tol and rel_err(v, T(g['mha_v'])) < tol
    y2, none_kv = mha(x)
    assert none_kv is None and rel_err(y2, T(g['mha_y_nomask'])) < tol
    y1, (k1, _) = mha(T(g['mha_x1']).cuda(), kv_cache=(k, v), use_cache=True)
    assert rel_err(y1, T(g['mha_y1'])) < tol and rel_err(k1, T(g['mha_k1'])) < tol

    ffn = FeedForward(d, 4 * d, dropout=0.0)
    ffn.load_state_dict(synth.synth_state_dict({'linear_1.weight': (4 * d, d), 'linear_1.bias': (4 * d,),
                                                'linear_2.weight': (d, 4 * d), 'linear_2.bias': (d,)}, 8))
    assert rel_err(ffn.cuda()(x), T(g['ffn_y'])) < tol

    ada = AdaptiveLayerNorm(d)
    ada.load_state_dict(synth.synth_state_dict({'project_layer.weight': (2 * d, d), 'project_layer.bias': (2 * d,),
                                                'norm.weight': (d,), 'norm.bias': (d,)}, 9))
    emb = T(g['ada_emb']).cuda()
    assert rel_err(ada.cuda()(x, emb), T(g['ada_y'])) < 2e-5      # norm kernels are fp32 in both modes
    assert rel_err(PositionalEncoding(d).cuda().eval()(x), T(g['pe_y'])) < 1e-6

    from valle.models.utils import build_attn_mask
    for norm, tag in (('LayerNorm', 'ln'), ('AdaptiveLayerNorm', 'ada')):
        oc = OracleConfig(num_layers=2, d_model=d, n_heads=H, dim_feedforward=4 * d, norm=norm)
        tr = Transformer(make_cfg(oc, tmp_path)).eval()
        shapes = {}
        for i in range(2):
            shapes.update(synth._layer_shapes(oc, f'layers.{i}.'))
        tr.load_state_dict(synth.synth_state_dict(shapes, 10))
        tr = tr.cuda()
        e = emb if norm != 'LayerNorm' else None
        mask = build_attn_mask(2, 4, 'cuda')
        yf, kv = tr(x, attn_mask=mask, embedding=e, use_cache=True)
        ys, kv2 = tr(torch.cat([x, T(g['mha_x1']).cuda()], 1), attn_mask=mask, embedding=e, kv_cache=kv, use_cache=True)
        yp, empty = tr(x, embedding=e)
        assert empty == ()
        assert rel_err(yf, T(g[f'tr_{tag}_full'])) < tol and rel_err(ys, T(g[f'tr_{tag}_step'])) < tol
        assert rel_err(yp, T(g[f'tr_{tag}_plain'])) < tol and rel_err(kv2[-1][0], T(g[f'tr_{tag}_k_last'])) < tol


def test_reference_shape_tests_on_gpu():
    """tests/test_modules.py:7-30 of the reference, with CUDA tensors."""
    from valle.models.modules import MultiHeadAttention
    for d_model, n_heads, batch_size, seq_len in [(512, 8, 4, 5), (256, 4, 8, 10), (128, 2, 16, 20)]:
        attention = MultiHeadAttention(d_model=d_model, n_heads=n_heads).cuda()
        head_dim = d_model // n_heads
        assert attention.head_dim == head_dim
        x = torch.randn(batch_size, seq_len, d_model, device='cuda')
        mask = torch.triu(torch.ones(seq_len, seq_len, device='cuda'), diagonal=1)
        output, kv = attention(x, attn_mask=mask, use_cache=True)
        k, v = kv
        assert output.shape == (batch_size, seq_len, d_model)
        assert k.shape == (batch_size, n_heads, seq_len, head_dim) and v.shape == k.shape
        assert torch.isfinite(output).all()


def test_ar_tiny_fp32_token_exact(golden, tmp_path):
    valle2_b200.set_precision('fp32')
    g = golden('ar_tiny')
    inp = synth.tiny_inputs(0)
    for beams in (1, 2):
        oc = synth.tiny_config('LayerNorm', num_beams=beams)
        model, sd = build('ValleAR', oc, tmp_path, 0)
        for use_graph in (False, True):
            codes = model.generate(inp['prompt_tokens'].cuda(), inp['prompt_codes'].cuda(), inp['target_tokens'].cuda(),
                                   use_graph=use_graph)
            assert codes.dtype == torch.int64 and codes.dim() == 1
            assert np.array_equal(codes.cpu().numpy(), g[f'gen_b{beams}_codes']), (beams, use_graph)
    # teacher-forced logits + loss against the executed reference
    oc = synth.tiny_config('LayerNorm')
    model, sd = build('ValleAR', oc, tmp_path, 0)
    batch = {k[3:]: T(v) for k, v in g.items() if k.startswith('tf_') and k not in ('tf_logits', 'tf_loss')}
    logits = model.forward_logits(batch)
    valid = T(g['tf_logits'])
    assert rel_err(logits, valid) < 1e-5
    loss = model.training_step(batch)
    assert abs(loss.item() - float(g['tf_loss'])) < 1e-4
    # early EOS: loop break + EOS stripping (valle_ar.py:167-178)
    w = sd['proj.weight'].clone()
    w[oc.eos_token] = T(g['eos_row'])
    with torch.no_grad():
        model.proj.weight.copy_(w.cuda())
    codes = model.generate(inp['prompt_tokens'].cuda(), inp['prompt_codes'].cuda(), inp['target_tokens'].cuda())
    assert np.array_equal(codes.cpu().numpy(), g['gen_eos_codes'])


def test_ar_tiny_injected_uniforms_match_oracle(tmp_path):
    """Stochastic decoding with injected draws: 3 beams diverge, finish at different steps; fp32 mode vs oracle."""
    valle2_b200.set_precision('fp32')
    inp = synth.tiny_inputs(1)
    oc = synth.tiny_config('LayerNorm', num_beams=3, top_k=20, tok_p=0.9, temperature=0.8, max_audio_len=24)
    model, sd = build('ValleAR', oc, tmp_path, 3)
    u = torch.rand(oc.max_audio_len, 3, generator=torch.Generator().manual_seed(5))
    ref = vo.ar_generate(sd, oc, inp['prompt_tokens'], inp['prompt_codes'], inp['target_tokens'], uniforms=u)
    out = model.generate(inp['prompt_tokens'].cuda(), inp['prompt_codes'].cuda(), inp['target_tokens'].cuda(),
                         uniforms=u.cuda())
    assert np.array_equal(out.cpu().numpy(), ref.numpy())


def test_ar_tiny_bf16(golden, tmp_path):
    valle2_b200.set_precision('bf16')
    g = golden('ar_tiny')
    oc = synth.tiny_config('LayerNorm')
    model, sd = build('ValleAR', oc, tmp_path, 0)
    batch = {k[3:]: T(v) for k, v in g.items() if k.startswith('tf_') and k not in ('tf_logits', 'tf_loss')}
    logits = model.forward_logits(batch)
    assert rel_err(logits, T(g['tf_logits'])) < 1e-2
    inp = synth.tiny_inputs(0)
    codes = model.generate(inp['prompt_tokens'].cuda(), inp['prompt_codes'].cuda(), inp['target_tokens'].cuda())
    ref = g['gen_b1_codes']
    ref_logits = T(g['gen_b1_logits'])[:, 0]                       # (steps, V+1) reference logits of the greedy path
    top2 = ref_logits.topk(2, dim=-1).values
    margin = (top2[:, 0] - top2[:, 1])
    noise = 2e-2 * ref_logits.abs().max()
    first_fragile = int((margin < noise).nonzero()[0]) if (margin < noise).any() else len(ref)
    got = codes.cpu().numpy()
    n = min(first_fragile, len(got), len(ref))
    assert np.array_equal(got[:n], ref[:n]), f'bf16 greedy diverged before the first fragile step {first_fragile}'
    # graph replay == eager launches (bit-identical kernels, deterministic split-K)
    codes2 = model.generate(inp['prompt_tokens'].cuda(), inp['prompt_codes'].cuda(), inp['target_tokens'].cuda(),
                            use_graph=False)
    assert torch.equal(codes, codes2)


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-5), ('bf16', 1e-2)])
def test_nar_tiny(golden, precision, tol, tmp_path):
    valle2_b200.set_precision(precision)
    g = golden('nar_tiny')
    inp = synth.tiny_inputs(0)
    oc = synth.tiny_config('AdaptiveLayerNorm')
    model, sd = build('ValleNAR', oc, tmp_path, 1)
    args = [inp[k].cuda() for k in ('prompt_tokens', 'prompt_codes', 'target_tokens', 'first_layer')]
    out = model.generate(*args, greedy=True)
    assert out.shape == (11, 8) and out.dtype == torch.int64
    ref = T(g['nar_codes'])
    if precision == 'fp32':
        assert torch.equal(out.cpu(), ref)
    else:
        # bf16: a stage's input is the previous stages' OUTPUT, so after the first near-tie the two runs see different
        # inputs.  Checked per stage against the oracle driven with the codebooks this run produced: logits within tol,
        # and the arg-max equal wherever the oracle's top-1/top-2 margin is outside that tolerance.
        assert torch.equal(out.cpu()[:, 0], ref[:, 0])
        n_clear = 0
        for n in range(1, 8):
            lg = oracle_nar_stage_logits(sd, oc, torch.cat([inp['prompt_tokens'], inp['target_tokens']]), inp['prompt_codes'],
                                         out.cpu(), n)
            top2 = lg.topk(2, dim=-1).values
            clear = (top2[:, 0] - top2[:, 1]) > 2 * tol * lg.abs().max()
            n_clear += int(clear.sum())
            assert (out.cpu()[clear, n] == lg.argmax(-1)[clear]).all(), n
        assert n_clear >= 7 * 11 // 2
    # stage logits with the reference's codes as context (teacher forcing inside generate is not exposed; use
    # the batched engine with return_logits on the reference's own first column -> stage 1 logits are comparable)
    eng = model._engine()
    _, trace = eng.generate(args[0][None], args[1][None], args[2][None], args[3][None], greedy=True, return_logits=True)
    assert rel_err(trace[0][0], T(g['nar_logits'])[0]) < tol
    if precision == 'fp32':
        for n in range(7):
            assert rel_err(trace[n][0], T(g['nar_logits'])[n]) < tol
    # teacher-forced stage + _prepare_audio_codes
    batch = {'codes': T(g['nar_tf_codes']), 'tokens': T(g['nar_tf_tokens']),
             'tokens_lens': torch.full((2,), 5), 'codes_lens': torch.full((2,), 12)}
    logits, prefix_len = model.forward_logits(batch, int(g['nar_tf_layer']))
    assert prefix_len == int(g['nar_tf_prefix_len'])
    assert rel_err(logits, T(g['nar_tf_logits'])) < tol
    y_emb, _ = model._prepare_audio_codes(T(g['nar_tf_codes']).cuda(), int(g['nar_tf_layer']))
    assert rel_err(y_emb, T(g['nar_tf_yemb'])) < 1e-6
    loss = model.training_step(batch, layer=int(g['nar_tf_layer']))
    assert abs(loss.item() - float(g['nar_tf_loss'])) < (1e-4 if precision == 'fp32' else 5e-2)


def test_ar_batch_matches_single(tmp_path):
    """generate_batch (extension) over B different utterances == B single-utterance greedy decodes (fp32 mode)."""
    valle2_b200.set_precision('fp32')
    oc = synth.tiny_config('LayerNorm', num_beams=1, max_audio_len=12)
    model, sd = build('ValleAR', oc, tmp_path, 0)
    g = torch.Generator().manual_seed(77)
    B, Tx, Pc = 3, 9, 6
    tokens = torch.randint(0, 256, (B, Tx), generator=g)
    pcodes = torch.randint(0, 1024, (B, Pc, 8), generator=g)
    codes = torch.cat([torch.full((B, 1), oc.bos_token), pcodes[:, :, 0]], 1)
    out, n = model.generate_batch(tokens.cuda(), codes.cuda(), max_new=12)
    for b in range(B):
        ref = vo.ar_generate(sd, oc, tokens[b], pcodes[b], None)
        got = out[b].cpu().long()
        got = got[got != oc.eos_token]
        assert np.array_equal(got.numpy()[: len(ref)], ref.numpy())


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_ar_large_short_decode_vs_oracle(precision, tmp_path):
    """BASELINE config 2 architecture (12 layers, d=1024, 16 heads): prefill + 6 greedy steps, 2 beams."""
    valle2_b200.set_precision(precision)
    oc = synth.large_config('LayerNorm', num_beams=2, max_audio_len=6)
    model, sd = build('ValleAR', oc, tmp_path, 11)
    g = torch.Generator().manual_seed(3)
    pt, pc, tt = torch.randint(0, 256, (20,), generator=g), torch.randint(0, 1024, (70, 8), generator=g), \
        torch.randint(0, 256, (30,), generator=g)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref, trace, _, _ = vo.ar_generate(sd, oc, pt, pc, tt, return_trace=True)
    out = model.generate(pt.cuda(), pc.cuda(), tt.cuda())
    if precision == 'fp32':
        assert np.array_equal(out.cpu().numpy(), ref.numpy())
    # teacher-forced logits over the oracle's own sequence
    codes_in = torch.cat([torch.tensor([oc.bos_token]), pc[:, 0], ref[:-1]])[None]
    batch = {'tokens': torch.cat([pt, tt])[None], 'codes': codes_in, 'tokens_lens': torch.tensor([50]),
             'codes_lens': torch.tensor([codes_in.shape[1]])}
    logits = model.forward_logits(batch)[0]
    ref_logits, _ = vo.ar_teacher_forced(sd, oc, batch['tokens'], codes_in, batch['tokens_lens'], batch['codes_lens'])
    assert rel_err(logits, ref_logits[0]) < (1e-5 if precision == 'fp32' else 1e-2)
    # the last len(ref) rows of the teacher-forced logits are the decode-step logits (cached == uncached invariant)
    for s in range(len(ref)):
        assert rel_err(logits[70 + s], trace[s][0]) < (2e-5 if precision == 'fp32' else 1e-2)


def test_nar_fused_logits_argmax_equals_unfused_stages(tmp_path):
    """Greedy bf16 NAR stages of >= 1024 target rows run the logits projection and the pick in ONE kernel (vb_linear_argmax,
    valle_nar.py:157-160); the codes must be bit-identical to the stages that write fp32 logits and pick with vb_sample
    (same GEMM, same accumulation order; 7 dependent stages, so a single different pick would change every later stage).
    Ragged targets included (rows past a target's length are padding in both)."""
    valle2_b200.set_precision('bf16')
    oc = synth.tiny_config('AdaptiveLayerNorm')
    model, sd = build('ValleNAR', oc, tmp_path, 4)
    g = torch.Generator().manual_seed(11)
    B, Tp, Tt, Tc, T = 3, 6, 9, 20, 400                      # B T = 1200 target rows
    pt, tt = torch.randint(0, 256, (B, Tp), generator=g).cuda(), torch.randint(0, 256, (B, Tt), generator=g).cuda()
    pc = torch.randint(0, 1024, (B, Tc, 8), generator=g).cuda()
    first = torch.randint(0, 1024, (B, T), generator=g).cuda()
    lens = torch.tensor([400, 333, 250])
    eng = model._engine()
    assert eng.fused_argmax and eng.fused_embed_norm
    for tl in (None, lens):
        fused = eng.generate(pt, pc, tt, first, greedy=True, target_lens=tl)
        eng.fused_argmax = eng.fused_embed_norm = False          # also: embedding-sum + PE + first AdaLN in one kernel
        try:
            plain = eng.generate(pt, pc, tt, first, greedy=True, target_lens=tl)
        finally:
            eng.fused_argmax = eng.fused_embed_norm = True
        assert fused.shape == (B, T, 8) and fused.dtype == torch.int64
        assert torch.equal(fused, plain)
        assert torch.equal(fused[:, :, 0], first)


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_ar_prefill_fused_embed_norm_equals_unfused(tmp_path, precision):
    """AR prefill with the embedding + PE + layer-0 norm1 kernel (vb_embed_sum_pe_norm) against the two-kernel form: the
    kernels agree bit for bit (above 1024 rows, where the unfused side runs the warp-per-row LayerNorm), so the continuation
    (prefill state + KV pools + 24 cached steps) must be identical, ragged prompts included."""
    valle2_b200.set_precision(precision)
    oc = synth.tiny_config('LayerNorm')
    model, sd = build('ValleAR', oc, tmp_path, 2)
    g = torch.Generator().manual_seed(3)
    B = 5
    tok = torch.randint(0, 256, (B, 13), generator=g).cuda()
    cod = torch.randint(0, 1024, (B, 220), generator=g)          # 5 x 233 rows: the large-M LayerNorm kernel on the unfused side
    cod[:, 0] = oc.bos_token
    lens = torch.tensor([220, 90, 120, 220, 30])
    eng = model._engine()
    assert eng.fused_embed_norm
    fused, n1 = model.generate_batch(tok, cod.cuda(), code_lens=lens, max_new=24, ignore_eos=True)
    eng.fused_embed_norm = False
    try:
        plain, n2 = model.generate_batch(tok, cod.cuda(), code_lens=lens, max_new=24, ignore_eos=True)
    finally:
        eng.fused_embed_norm = True
    assert torch.equal(fused, plain) and torch.equal(torch.as_tensor(n1).cpu(), torch.as_tensor(n2).cpu())


def test_nar_sampled_stages_fused_categorical(tmp_path):
    """ValleNAR.generate as the reference runs it (valle_nar.py:160: a Categorical draw per stage, greedy=False) on the fused
    logits + draw kernel: codes in range, first codebook untouched, reproducible by seed, different for another seed, and --
    because every stage conditions on the previous stages' draws -- statistically indistinguishable from the unfused path
    (vb_linear + vb_sample, other random stream): per-stage agreement with the greedy codes is the same within noise."""
    valle2_b200.set_precision('bf16')
    oc = synth.tiny_config('AdaptiveLayerNorm')
    model, sd = build('ValleNAR', oc, tmp_path, 4)
    g = torch.Generator().manual_seed(12)
    B, Tp, Tt, Tc, T = 4, 6, 9, 20, 512                      # B T = 2048 target rows
    pt, tt = torch.randint(0, 256, (B, Tp), generator=g).cuda(), torch.randint(0, 256, (B, Tt), generator=g).cuda()
    pc = torch.randint(0, 1024, (B, Tc, 8), generator=g).cuda()
    first = torch.randint(0, 1024, (B, T), generator=g).cuda()
    eng = model._engine()
    a = eng.generate(pt, pc, tt, first, greedy=False, temperature=1.0, seed=5)
    assert a.shape == (B, T, 8) and int(a.min()) >= 0 and int(a.max()) < 1024
    assert torch.equal(a[:, :, 0], first)
    assert torch.equal(a, eng.generate(pt, pc, tt, first, greedy=False, temperature=1.0, seed=5))
    b = eng.generate(pt, pc, tt, first, greedy=False, temperature=1.0, seed=6)
    assert (a[:, :, 1:] != b[:, :, 1:]).double().mean().item() > 0.3
    greedy = eng.generate(pt, pc, tt, first, greedy=True)
    eng.fused_argmax = False
    try:
        c = eng.generate(pt, pc, tt, first, greedy=False, temperature=1.0, seed=5)
    finally:
        eng.fused_argmax = True
    # stage 2 (column 1) sees identical inputs in all runs: P(draw == arg-max) must agree between the two samplers
    hit_fused = (a[:, :, 1] == greedy[:, :, 1]).double().mean().item()
    hit_plain = (c[:, :, 1] == greedy[:, :, 1]).double().mean().item()
    assert abs(hit_fused - hit_plain) < 5 * math.sqrt(0.25 / (B * T)) + 0.01, (hit_fused, hit_plain)
