"""Pins the CPU oracle (oracle/valle_oracle.py) against vectors frozen from the EXECUTED reference
(oracle/make_golden.py -> tests/golden/*.npz).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import synth
from oracle import valle_oracle as vo
from oracle.valle_oracle import OracleConfig

T = torch.from_numpy


def close(a, b, tol=2e-5):
    a = a if isinstance(a, torch.Tensor) else T(np.asarray(a))
    b = b if isinstance(b, torch.Tensor) else T(np.asarray(b))
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    ref = b.abs().max().item() + 1e-12
    assert err <= tol * max(1.0, ref), f'max err {err} (ref scale {ref})'


def test_modules(golden):
    g = golden('modules')
    d, H = 64, 4
    x = T(g['mha_x'])
    S = x.shape[1]
    sd = {'a.' + k: v for k, v in synth.synth_state_dict(
        {'qkv.weight': (3 * d, d), 'out.weight': (d, d), 'out.bias': (d,)}, 7).items()}
    causal = torch.triu(torch.ones(S, S), diagonal=1)
    y, (k, v) = vo.multi_head_attention(x, sd, 'a.', H, attn_mask=causal, padding_mask=T(g['mha_pad']),
                                        use_cache=True)
    close(y, g['mha_y']); close(k, g['mha_k']); close(v, g['mha_v'])
    y2, _ = vo.multi_head_attention(x, sd, 'a.', H)
    close(y2, g['mha_y_nomask'])
    y1, (k1, _) = vo.multi_head_attention(T(g['mha_x1']), sd, 'a.', H, kv_cache=(k, v), use_cache=True)
    close(y1, g['mha_y1']); close(k1, g['mha_k1'])

    sdf = {'f.' + k: v for k, v in synth.synth_state_dict(
        {'linear_1.weight': (4 * d, d), 'linear_1.bias': (4 * d,),
         'linear_2.weight': (d, 4 * d), 'linear_2.bias': (d,)}, 8).items()}
    close(vo.feed_forward(x, sdf, 'f.'), g['ffn_y'])

    sda = synth.synth_state_dict({'project_layer.weight': (2 * d, d), 'project_layer.bias': (2 * d,),
                                  'norm.weight': (d,), 'norm.bias': (d,)}, 9)
    emb = T(g['ada_emb'])
    close(vo.adaptive_layer_norm(x, emb, sda['project_layer.weight'], sda['project_layer.bias'],
                                 sda['norm.weight'], sda['norm.bias']), g['ada_y'])
    close(vo.add_pe(x, vo.sinusoidal_pe(5000, d)), g['pe_y'], tol=1e-6)

    for norm, tag in (('LayerNorm', 'ln'), ('AdaptiveLayerNorm', 'ada')):
        oc = OracleConfig(num_layers=2, d_model=d, n_heads=H, dim_feedforward=4 * d, norm=norm)
        shapes = {}
        for i in range(2):
            shapes.update(synth._layer_shapes(oc, f'layers.{i}.'))
        sdt = synth.synth_state_dict(shapes, 10)
        e = emb if norm != 'LayerNorm' else None
        mask = vo.build_attn_mask(2, 4)
        yf, kv = vo.transformer(x, sdt, oc, prefix='', attn_mask=mask, embedding=e, use_cache=True)
        ys, kv2 = vo.transformer(torch.cat([x, T(g['mha_x1'])], 1), sdt, oc, prefix='', attn_mask=mask,
                                 embedding=e, kv_cache=kv, use_cache=True)
        yp, _ = vo.transformer(x, sdt, oc, prefix='', embedding=e)
        close(yf, g[f'tr_{tag}_full']); close(ys, g[f'tr_{tag}_step']); close(yp, g[f'tr_{tag}_plain'])
        close(kv2[-1][0], g[f'tr_{tag}_k_last'])


def test_masks_sampling(golden):
    g = golden('masks_sampling')
    assert np.array_equal(vo.build_attn_mask(3, 4).numpy(), g['attn_mask_3_4'])
    assert np.array_equal(vo.build_pad_mask(torch.tensor([4, 2, 3, 1])).numpy(), g['pad_mask'])
    logits = T(g['samp_logits'])
    for i, (k, p) in enumerate(g['samp_cases']):
        f = vo.top_k_top_p_filter(logits, int(k), float(p))
        ref = T(g[f'samp_filtered_{i}'])
        assert torch.equal(torch.isinf(f), torch.isinf(ref)), f'case {i}'
        assert torch.equal(f[~torch.isinf(f)], ref[~torch.isinf(ref)])
    # log-prob of the reference's own draw under the filtered distribution (utils.py:65-66)
    tok = T(g['samp_tok'])
    f = vo.top_k_top_p_filter(logits / 0.7, 10, 0.8)
    lp = torch.log_softmax(f, -1).gather(1, tok)[:, 0]
    close(lp, g['samp_logprob'], tol=1e-6)
    s, lp1 = vo.topk_sampling(logits, 1, 1.0, 1.0)
    assert np.array_equal(s.numpy(), g['greedy_tok'])
    close(lp1, g['greedy_logprob'], tol=1e-6)
    for i, lpn in enumerate((1.0, 0.0, 2.0)):
        best = vo.get_best_beam(T(g['beam_x']), T(g['beam_slp']), 1024, lpn)
        assert np.array_equal(best.numpy(), g[f'beam_best_{i}'])


def test_injected_uniform_sampling_is_consistent():
    torch.manual_seed(0)
    logits = torch.randn(4, 1025) * 2
    f = vo.top_k_top_p_filter(logits, 20, 0.9)
    p = torch.softmax(f, -1)
    for u in (0.0, 0.3, 0.77, 0.999999):
        s, lp = vo.topk_sampling(logits, 20, 0.9, 1.0, uniforms=torch.full((4,), u))
        assert (p.gather(1, s) > 0).all()
        cdf = p.cumsum(-1)
        hi = cdf.gather(1, s)[:, 0]
        lo = hi - p.gather(1, s)[:, 0]
        assert ((lo <= u + 1e-6) & (u < hi + 1e-6)).all()
        close(lp, torch.log(p.gather(1, s))[:, 0], tol=1e-5)


def test_ar_tiny(golden):
    g = golden('ar_tiny')
    inp = synth.tiny_inputs(0)
    for beams in (1, 2):
        oc = synth.tiny_config('LayerNorm', num_beams=beams)
        sd = synth.synth_state_dict(synth.ar_state_shapes(oc), 0)
        codes, trace, _, _ = vo.ar_generate(sd, oc, inp['prompt_tokens'], inp['prompt_codes'],
                                            inp['target_tokens'], return_trace=True)
        assert np.array_equal(codes.numpy(), g[f'gen_b{beams}_codes'])
        close(torch.stack(trace), g[f'gen_b{beams}_logits'])
    oc = synth.tiny_config('LayerNorm')
    sd = synth.synth_state_dict(synth.ar_state_shapes(oc), 0)
    logits, loss = vo.ar_teacher_forced(sd, oc, T(g['tf_tokens']), T(g['tf_codes']),
                                        T(g['tf_tokens_lens']), T(g['tf_codes_lens']), T(g['tf_target']))
    close(logits, g['tf_logits']); close(loss, g['tf_loss'], tol=1e-6)
    # early EOS
    w = sd['proj.weight'].clone()
    w[oc.eos_token] = T(g['eos_row'])
    sd['proj.weight'] = w
    codes = vo.ar_generate(sd, oc, inp['prompt_tokens'], inp['prompt_codes'], inp['target_tokens'])
    assert np.array_equal(codes.numpy(), g['gen_eos_codes'])
    assert 0 < len(codes) <= 10


def test_nar_tiny(golden):
    g = golden('nar_tiny')
    inp = synth.tiny_inputs(0)
    oc = synth.tiny_config('AdaptiveLayerNorm')
    sd = synth.synth_state_dict(synth.nar_state_shapes(oc), 1)
    codes, trace = vo.nar_generate(sd, oc, inp['prompt_tokens'], inp['prompt_codes'],
                                   inp['target_tokens'], inp['first_layer'], return_trace=True)
    assert np.array_equal(codes.numpy(), g['nar_codes'])
    close(torch.stack(trace), g['nar_logits'])
    y_emb, prefix_len = vo.nar_prepare_audio_codes(sd, oc, T(g['nar_tf_codes']), int(g['nar_tf_layer']))
    assert prefix_len == int(g['nar_tf_prefix_len'])
    close(y_emb, g['nar_tf_yemb'], tol=1e-6)
    B = g['nar_tf_codes'].shape[0]
    logits, loss = vo.nar_teacher_forced(sd, oc, T(g['nar_tf_tokens']), T(g['nar_tf_codes']),
                                         torch.full((B,), 5), torch.full((B,), 12), int(g['nar_tf_layer']))
    close(logits, g['nar_tf_logits']); close(loss, g['nar_tf_loss'], tol=1e-6)


def test_ar_training_gradients_match_the_executed_reference(golden):
    """torch autograd over the oracle's restatement of ValleAR.training_step == autograd through the reference's own
    modules (per-parameter summaries frozen by oracle/make_golden.py:golden_ar_grads).  This pins the gradient oracle
    used by tests/test_gpu_training.py."""
    g, gg = golden('ar_tiny'), golden('ar_tiny_grads')
    oc = synth.tiny_config('LayerNorm')
    sd = synth.synth_state_dict(synth.ar_state_shapes(oc), 0)
    sdg = {k: (v.clone().requires_grad_(True) if not k.endswith('.pe') else v) for k, v in sd.items()}
    _, loss = vo.ar_teacher_forced(sdg, oc, T(g['tf_tokens']), T(g['tf_codes']), T(g['tf_tokens_lens']),
                                   T(g['tf_codes_lens']), T(g['tf_target']))
    loss.backward()
    close(loss.detach(), gg['loss'], tol=1e-6)
    names = [k[6:] for k in gg if k.startswith('stats.')]
    assert len(names) == 2 + 2 * 11 + 1          # two embedding tables, 11 tensors per layer, proj
    for name in names:
        gr = sdg[name].grad.double().flatten()
        stats = np.array([float(gr.norm()), float(gr.sum()), float(gr.abs().max())])
        scale = gg['stats.' + name][2] + 1e-12
        assert np.abs(stats - gg['stats.' + name]).max() / max(scale, gg['stats.' + name][0]) < 1e-4, name
        assert np.abs(gr[:16].numpy() - gg['head.' + name]).max() / scale < 1e-4, name


@pytest.mark.parametrize('case', [
    dict(seed=1, beams=1, layers=2, d=256, heads=4, F=1024, Tp=7, Tc=9, Tt=5, steps=12),
    dict(seed=2, beams=3, layers=1, d=128, heads=2, F=256, Tp=3, Tc=4, Tt=0, steps=9),
    dict(seed=3, beams=2, layers=3, d=192, heads=3, F=384, Tp=11, Tc=1, Tt=8, steps=7),
])
def test_oracle_ar_equals_the_executed_reference_live(tmp_path, case):
    """Beyond the frozen fixtures: the oracle against the executed reference itself (oracle/_ref, unmodified source) on OTHER
    shapes than the golden files hold -- other depths / widths / head counts, 1-3 beams, a one-frame prompt, no target text --
    greedy tokens equal, per-step logits of every beam within 1e-5, and the teacher-forced logits + loss of a ragged batch
    (valle_ar.py:43-90, :92-180).  Skipped where the reference is not installed."""
    from oracle import ref_shims, synth
    if not ref_shims.reference_available():
        pytest.skip('oracle/_ref not installed')
    c = case
    oc = synth.tiny_config('LayerNorm', num_layers=c['layers'], d_model=c['d'], n_heads=c['heads'], dim_feedforward=c['F'],
                           num_beams=c['beams'], max_audio_len=c['steps'])
    sd = synth.synth_state_dict(synth.ar_state_shapes(oc), c['seed'])
    g = torch.Generator().manual_seed(100 + c['seed'])
    pt = torch.randint(0, 256, (c['Tp'],), generator=g)
    pc = torch.randint(0, 1024, (c['Tc'], 8), generator=g)
    tt = torch.randint(0, 256, (c['Tt'],), generator=g) if c['Tt'] else None
    B, Tx, Ty = 3, 6, 10
    batch = {'tokens': torch.randint(0, 256, (B, Tx), generator=g), 'tokens_lens': torch.tensor([6, 4, 5]),
             'codes': torch.randint(0, 1024, (B, Ty), generator=g), 'codes_lens': torch.tensor([10, 7, 9]),
             'target': torch.randint(0, 1025, (B, Ty), generator=g)}
    out, trace, _, _ = vo.ar_generate(sd, oc, pt, pc, tt, return_trace=True)
    tf_logits, tf_loss = vo.ar_teacher_forced(sd, oc, batch['tokens'], batch['codes'], batch['tokens_lens'], batch['codes_lens'],
                                              batch['target'])
    valle = ref_shims.import_reference()
    try:
        kw = {k: getattr(oc, k) for k in type(oc).__dataclass_fields__}
        cfg = valle.config.ConfigValle(dropout=0.0, ckpt_path=str(tmp_path / 'c'), log_path=str(tmp_path / 'l'), **kw)
        model = valle.models.ValleAR(cfg).eval()
        model.load_state_dict(sd, strict=True)
        ref_trace, ref_tf = [], []
        hook = model.proj.register_forward_hook(lambda m, i, o: ref_trace.append(o[:, -1].clone()))
        with torch.no_grad():
            ref_out = model.generate(pt, pc, tt)
        hook.remove()
        hook = model.proj.register_forward_hook(lambda m, i, o: ref_tf.append(o.clone()))
        with torch.no_grad():
            ref_loss = model.training_step({k: v.clone() for k, v in batch.items()})
        hook.remove()
    finally:
        ref_shims.release_reference()
    assert torch.equal(out, ref_out)
    assert len(trace) == len(ref_trace)
    for a, b in zip(trace, ref_trace):
        close(a, b.numpy(), tol=1e-5)
    close(tf_logits, ref_tf[0].numpy(), tol=1e-5)
    assert abs(float(tf_loss) - float(ref_loss)) < 1e-5


@pytest.mark.parametrize('norm', ['AdaptiveLayerNorm', 'LayerNorm'])
@pytest.mark.parametrize('L,d,H,B,S', [(1, 96, 3, 2, 7), (3, 128, 8, 4, 13)])
def test_oracle_transformer_equals_the_executed_reference_live(tmp_path, norm, L, d, H, B, S):
    """The oracle's Transformer / EncoderLayer / (Adaptive)LayerNorm / MHA / FFN stack (modules.py:84-352) against the executed
    reference modules on other shapes than the fixtures: odd head counts, key-padding masks, the stage embedding of AdaLN, and a
    KV-cached continuation of two more positions (what the NAR stages and the AR loop are made of)."""
    from oracle import ref_shims
    if not ref_shims.reference_available():
        pytest.skip('oracle/_ref not installed')
    oc = OracleConfig(num_layers=L, d_model=d, n_heads=H, dim_feedforward=2 * d, norm=norm)
    shapes = {}
    for i in range(L):
        shapes.update(synth._layer_shapes(oc, f'layers.{i}.'))
    sd = synth.synth_state_dict(shapes, 20 + L)
    g = torch.Generator().manual_seed(d + S)
    x = torch.randn(B, S, d, generator=g)
    x2 = torch.randn(B, 2, d, generator=g)
    emb = torch.randn(1, d, generator=g) if norm == 'AdaptiveLayerNorm' else None
    lens = torch.randint(S // 2, S + 1, (B,), generator=g)
    lens[0] = S
    pad = vo.build_pad_mask(lens)
    y, kv = vo.transformer(x, sd, oc, prefix='', padding_mask=pad, embedding=emb, use_cache=True)
    y2, _ = vo.transformer(torch.cat([x, x2], 1), sd, oc, prefix='', embedding=emb, kv_cache=kv, use_cache=True)
    valle = ref_shims.import_reference()
    try:
        kw = {k: getattr(oc, k) for k in type(oc).__dataclass_fields__}
        cfg = valle.config.ConfigValle(dropout=0.0, ckpt_path=str(tmp_path / 'c'), log_path=str(tmp_path / 'l'), **kw)
        tr = valle.models.modules.Transformer(cfg).eval()
        tr.load_state_dict(sd, strict=True)
        with torch.no_grad():
            ry, rkv = tr(x, padding_mask=pad, embedding=emb, use_cache=True)
            ry2, _ = tr(torch.cat([x, x2], 1), embedding=emb, kv_cache=rkv, use_cache=True)
    finally:
        ref_shims.release_reference()
    close(y, ry.numpy(), tol=1e-5)
    close(y2, ry2.numpy(), tol=1e-5)
    close(kv[-1][0], rkv[-1][0].numpy(), tol=1e-5)


def test_oracle_sampling_filter_equals_the_transformers_warpers_live():
    """The third-party arithmetic of the path (transformers' top-k / top-p warpers behind `top_k_top_p_filtering`, utils.py:63):
    the oracle's restatement against the installed library's own warper classes on 60 random (k, p) pairs and three logit
    tables -- smooth, heavy ties (logits rounded to halves) and one dominant class -- the kept SET and the kept values equal."""
    from oracle.ref_shims import _top_k_top_p_filtering
    pytest.importorskip('transformers')
    g = torch.Generator().manual_seed(17)
    smooth = torch.randn(5, 1025, generator=g) * 3
    ties = (torch.randn(5, 1025, generator=g) * 2).mul(2).round().div(2)
    peaked = torch.randn(5, 1025, generator=g)
    peaked[:, 77] += 25.0
    ks = [0, 1, 2, 5, 50, 1024, 1025, 4000]
    for i in range(60):
        k = ks[i % len(ks)]
        p = [1.0, 0.0, 0.999, 0.5][i % 4] if i < 16 else float(torch.rand(1, generator=g))
        for name, lg in (('smooth', smooth), ('ties', ties), ('peaked', peaked)):
            mine = vo.top_k_top_p_filter(lg, k, p)
            ref = _top_k_top_p_filtering(lg.clone(), top_k=k, top_p=p)
            if name == 'ties' and p < 1:
                # the order of equal logits inside torch.sort is the only freedom (the oracle sorts stably, the library does
                # not promise to): the kept COUNT per row and the multiset of kept values must still agree
                assert torch.equal(torch.isinf(mine).sum(-1), torch.isinf(ref).sum(-1)), (name, k, p)
                assert torch.equal(mine.sort(-1).values, ref.sort(-1).values), (name, k, p)
            else:
                assert torch.equal(torch.isinf(mine), torch.isinf(ref)), (name, k, p)
                assert torch.equal(mine[~torch.isinf(mine)], ref[~torch.isinf(ref)]), (name, k, p)
