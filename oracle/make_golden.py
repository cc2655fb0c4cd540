"""TEST INFRASTRUCTURE ONLY -- freeze golden vectors from the EXECUTED reference.

Run in the authoring container (needs ``/root/reference``)::

    cd /root/repo && python -m oracle.make_golden

Writes ``tests/golden/*.npz``.  Weights are not stored: they are regenerated from
``oracle/synth.py`` (seeded, per-key), loaded with ``load_state_dict(strict=True)`` into the
reference's own modules here, and into the oracle / CUDA engine in the tests.
Reference code executed: valle/models/modules.py, valle/models/utils.py, valle/models/valle_ar.py
(unmodified).  ``ValleNAR.generate``/``training_step`` raise upstream, so the NAR
vectors drive the reference's own sub-modules (embeddings, PositionalEncoding, Transformer with
AdaptiveLayerNorm, proj_layers) with the repaired stage loop.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

from . import ref_shims, synth
from .valle_oracle import OracleConfig

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _ref_config(valle, oc: OracleConfig, tmp: str):
    kw = {k: getattr(oc, k) for k in OracleConfig.__dataclass_fields__}
    return valle.config.ConfigValle(dropout=0.0, ckpt_path=os.path.join(tmp, 'c'),
                                    log_path=os.path.join(tmp, 'l'), **kw)


def _np(t):
    return t.detach().cpu().numpy()


def golden_modules(valle, tmp):
    M = valle.models.modules
    out = {}
    torch.manual_seed(0)
    d, H, B, S = 64, 4, 3, 6
    mha = M.MultiHeadAttention(d_model=d, n_heads=H).eval()
    sd = synth.synth_state_dict({'qkv.weight': (3 * d, d), 'out.weight': (d, d), 'out.bias': (d,)}, 7)
    mha.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, S, d, generator=g)
    causal = torch.triu(torch.ones(S, S), diagonal=1)
    pad = torch.tensor([[0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 1], [0, 0, 0, 0, 1, 1]])
    with torch.no_grad():
        y, (k, v) = mha(x, attn_mask=causal, padding_mask=pad, use_cache=True)
        y_nomask, _ = mha(x)
        x1 = torch.randn(B, 1, d, generator=g)
        y1, (k1, v1) = mha(x1, kv_cache=(k, v), use_cache=True)
    out.update(mha_x=_np(x), mha_pad=_np(pad), mha_y=_np(y), mha_k=_np(k), mha_v=_np(v),
               mha_y_nomask=_np(y_nomask), mha_x1=_np(x1), mha_y1=_np(y1), mha_k1=_np(k1))

    ffn = M.FeedForward(d, 4 * d, dropout=0.0).eval()
    sdf = synth.synth_state_dict({'linear_1.weight': (4 * d, d), 'linear_1.bias': (4 * d,),
                                  'linear_2.weight': (d, 4 * d), 'linear_2.bias': (d,)}, 8)
    ffn.load_state_dict(sdf, strict=True)
    with torch.no_grad():
        out['ffn_y'] = _np(ffn(x))

    ada = M.AdaptiveLayerNorm(d).eval()
    sda = synth.synth_state_dict({'project_layer.weight': (2 * d, d), 'project_layer.bias': (2 * d,),
                                  'norm.weight': (d,), 'norm.bias': (d,)}, 9)
    ada.load_state_dict(sda, strict=True)
    emb = torch.randn(1, d, generator=g)
    with torch.no_grad():
        out['ada_emb'] = _np(emb)
        out['ada_y'] = _np(ada(x, emb))

    pe = M.PositionalEncoding(d).eval()
    with torch.no_grad():
        out['pe_y'] = _np(pe(x))

    # full stack, both norms
    for norm in ('LayerNorm', 'AdaptiveLayerNorm'):
        oc = OracleConfig(num_layers=2, d_model=d, n_heads=H, dim_feedforward=4 * d, norm=norm)
        cfg = _ref_config(valle, oc, tmp)
        tr = M.Transformer(cfg).eval()
        shapes = {}
        for i in range(2):
            shapes.update(synth._layer_shapes(oc, f'layers.{i}.'))
        tr.load_state_dict(synth.synth_state_dict(shapes, 10), strict=True)
        e = emb if norm != 'LayerNorm' else None
        mask = valle.models.utils.build_attn_mask(2, 4, 'cpu')
        with torch.no_grad():
            y_full, kv = tr(x, attn_mask=mask, embedding=e, use_cache=True)
            y_step, kv2 = tr(torch.cat([x, x1], dim=1), attn_mask=mask, embedding=e, kv_cache=kv,
                             use_cache=True)
            y_plain, _ = tr(x, embedding=e)
        tag = 'ln' if norm == 'LayerNorm' else 'ada'
        out[f'tr_{tag}_full'] = _np(y_full)
        out[f'tr_{tag}_step'] = _np(y_step)
        out[f'tr_{tag}_plain'] = _np(y_plain)
        out[f'tr_{tag}_k_last'] = _np(kv2[-1][0])
    np.savez_compressed(os.path.join(OUT, 'modules.npz'), **out)


def golden_masks_sampling(valle):
    U = valle.models.utils
    out = {}
    out['attn_mask_3_4'] = _np(U.build_attn_mask(3, 4, 'cpu'))
    out['pad_mask'] = _np(U.build_pad_mask(torch.tensor([4, 2, 3, 1]), 'cpu'))
    from transformers.generation.utils import top_k_top_p_filtering
    g = torch.Generator().manual_seed(21)
    logits = torch.randn(6, 1025, generator=g) * 3.0
    out['samp_logits'] = _np(logits)
    cases = [(50, 1.0), (10, 0.8), (0, 0.5), (1, 1.0), (5, 0.3), (0, 1.0)]
    out['samp_cases'] = np.array(cases, dtype=np.float64)
    for i, (k, p) in enumerate(cases):
        out[f'samp_filtered_{i}'] = _np(top_k_top_p_filtering(logits.clone(), top_k=k, top_p=p))
    # logprob of a given token under the filtered distribution (utils.py:65-66), temperature 0.7
    torch.manual_seed(5)
    tok, lp = U.topk_sampling(logits, top_k=10, tok_p=0.8, temperature=0.7)
    out['samp_tok'] = _np(tok)
    out['samp_logprob'] = _np(lp)
    # greedy through the reference's own sampler
    tok1, lp1 = U.topk_sampling(logits, top_k=1, tok_p=1.0, temperature=1.0)
    out['greedy_tok'] = _np(tok1)
    out['greedy_logprob'] = _np(lp1)
    # get_best_beam
    x = torch.tensor([[1025, 3, 4, 5, 1024, 1024], [1025, 3, 9, 9, 9, 1024], [1025, 1, 2, 3, 4, 5]])
    slp = torch.tensor([-3.0, -3.5, -6.5])
    out['beam_x'] = _np(x)
    out['beam_slp'] = _np(slp)
    for i, lpn in enumerate((1.0, 0.0, 2.0)):
        out[f'beam_best_{i}'] = _np(U.get_best_beam(x, slp, 1024, lpn))
    np.savez_compressed(os.path.join(OUT, 'masks_sampling.npz'), **out)


def golden_ar(valle, tmp):
    out = {}
    inp = synth.tiny_inputs(0)
    for beams in (1, 2):
        oc = synth.tiny_config('LayerNorm', num_beams=beams)
        cfg = _ref_config(valle, oc, tmp)
        model = valle.models.ValleAR(cfg).eval()
        sd = synth.synth_state_dict(synth.ar_state_shapes(oc), 0)
        model.load_state_dict(sd, strict=True)
        trace = []
        hook = model.proj.register_forward_hook(lambda m, i, o: trace.append(o[:, -1].clone()))
        codes = model.generate(inp['prompt_tokens'], inp['prompt_codes'], inp['target_tokens'])
        hook.remove()
        out[f'gen_b{beams}_codes'] = _np(codes)
        out[f'gen_b{beams}_logits'] = _np(torch.stack(trace))          # (steps, B, 1025)
        if beams == 1:
            # teacher-forced forward over a ragged batch (valle_ar.py:43-90)
            g = torch.Generator().manual_seed(31)
            tokens = torch.randint(0, 256, (2, 6), generator=g)
            tokens_lens = torch.tensor([6, 4])
            tokens[1, 4:] = 0
            c = torch.randint(0, 1024, (2, 9), generator=g)
            codes_in = torch.cat([torch.full((2, 1), oc.bos_token), c], dim=1)
            target = torch.cat([c, torch.full((2, 1), oc.eos_token)], dim=1)
            codes_lens = torch.tensor([10, 7])
            codes_in[1, 7:] = 0
            target[1, 7:] = 0
            batch = dict(codes=codes_in, codes_lens=codes_lens, tokens=tokens,
                         tokens_lens=tokens_lens, target=target)
            tf_logits = []
            hook = model.proj.register_forward_hook(lambda m, i, o: tf_logits.append(o.clone()))
            with torch.no_grad():
                loss = model.training_step(batch)
            hook.remove()
            for k, v in batch.items():
                out[f'tf_{k}'] = _np(v)
            out['tf_logits'] = _np(tf_logits[0])
            out['tf_loss'] = _np(loss)
            # early-EOS case: point the EOS row of proj at the step-10 hidden state so that greedy
            # decoding emits EOS no later than step 10 (exercises valle_ar.py:167-171,177-178)
            hs = []
            hook = model.proj.register_forward_hook(lambda m, i, o: hs.append(i[0][:, -1].clone()))
            model.generate(inp['prompt_tokens'], inp['prompt_codes'], inp['target_tokens'])
            hook.remove()
            h10 = hs[10][0]
            top = float((h10 @ sd['proj.weight'].t()).max())
            eos_row = h10 / float(h10 @ h10) * (abs(top) + 2.0)
            out['eos_row'] = _np(eos_row)
            sd2 = dict(sd)
            w = sd['proj.weight'].clone()
            w[oc.eos_token] = eos_row
            sd2['proj.weight'] = w
            model.load_state_dict(sd2, strict=True)
            out['gen_eos_codes'] = _np(model.generate(inp['prompt_tokens'], inp['prompt_codes'],
                                                      inp['target_tokens']))
    np.savez_compressed(os.path.join(OUT, 'ar_tiny.npz'), **out)


def golden_nar(valle, tmp):
    out = {}
    inp = synth.tiny_inputs(0)
    oc = synth.tiny_config('AdaptiveLayerNorm')
    cfg = _ref_config(valle, oc, tmp)
    model = valle.models.ValleNAR(cfg).eval()
    sd = synth.synth_state_dict(synth.nar_state_shapes(oc), 1)
    model.load_state_dict(sd, strict=True)
    pt, pc, tt, fl = inp['prompt_tokens'], inp['prompt_codes'], inp['target_tokens'], inp['first_layer']
    Tc, Q = pc.shape
    with torch.no_grad():
        # valle_nar.py:127-163 with repairs A-5..A-9, using the reference's own sub-modules
        emb_prompt = torch.zeros(Tc, oc.d_model)
        for j in range(Q):
            emb_prompt = emb_prompt + model.codes_embs[j](pc[:, j])
        tokens = torch.cat([pt, tt]).unsqueeze(0)
        Tx = tokens.shape[1]
        x_tok = model.tokens_position_emb(model.tokens_emb(tokens))
        emb_out = torch.zeros(fl.shape[0], oc.d_model)
        cols = [fl]
        logits_all = []
        for n in range(1, Q):
            emb_out = emb_out + model.codes_embs[n - 1](cols[n - 1])
            codes = model.audio_position_emb(torch.cat([emb_prompt, emb_out], dim=0).unsqueeze(0))
            h, _ = model.transformer(torch.cat([x_tok, codes], dim=1),
                                     embedding=model.stage_embs[n - 1].weight)
            logits = model.proj_layers[n - 1](h[:, Tx + Tc:])
            logits_all.append(logits[0])
            cols.append(logits[0].argmax(-1))
        out['nar_codes'] = _np(torch.stack(cols, dim=-1))
        out['nar_logits'] = _np(torch.stack(logits_all))
        # teacher-forced stage (valle_nar.py:53-105, repairs A-1..3) with the reference's
        # unmodified _prepare_audio_codes (:167-188)
        g = torch.Generator().manual_seed(41)
        B, T = 2, 12
        codes_b = torch.randint(0, 1024, (B, T, Q), generator=g)
        tokens_b = torch.randint(0, 256, (B, 5), generator=g)
        layer = 3
        y_emb, prefix_len = model._prepare_audio_codes(codes_b, layer)
        out['nar_tf_yemb'] = _np(y_emb)
        out['nar_tf_prefix_len'] = np.array(prefix_len)
        x_tok = model.tokens_position_emb(model.tokens_emb(tokens_b))
        z, _ = model.transformer(torch.cat([x_tok, model.audio_position_emb(y_emb)], dim=1),
                                 embedding=model.stage_embs[layer - 1].weight)
        lg = model.proj_layers[layer - 1](z[:, 5 + prefix_len:])
        tgt = codes_b[:, prefix_len:, layer]
        loss = torch.nn.functional.cross_entropy(lg.permute(0, 2, 1), tgt)
        out.update(nar_tf_codes=_np(codes_b), nar_tf_tokens=_np(tokens_b), nar_tf_logits=_np(lg),
                   nar_tf_loss=_np(loss), nar_tf_layer=np.array(layer))
    np.savez_compressed(os.path.join(OUT, 'nar_tiny.npz'), **out)


def golden_ar_grads(valle, tmp):
    """Gradients of the reference's own ``ValleAR.training_step`` (autograd through the unmodified modules, eval mode so
    that the hard-wired PE dropout is inactive, K-3) on the ragged teacher-forced batch of ``ar_tiny.npz``.  Stored as
    per-parameter summaries (L2 norm, sum, abs-max, first 16 values) to keep the fixture small."""
    g = np.load(os.path.join(OUT, 'ar_tiny.npz'))
    oc = synth.tiny_config('LayerNorm')
    cfg = _ref_config(valle, oc, tmp)
    model = valle.models.ValleAR(cfg).eval()
    sd = synth.synth_state_dict(synth.ar_state_shapes(oc), 0)
    model.load_state_dict(sd, strict=True)
    batch = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('tf_') and k not in ('tf_logits', 'tf_loss')}
    loss = model.training_step(batch)
    loss.backward()
    out = {'loss': _np(loss)}
    for name, p in model.named_parameters():
        gr = p.grad.detach().double().flatten()
        out['stats.' + name] = np.array([float(gr.norm()), float(gr.sum()), float(gr.abs().max())])
        out['head.' + name] = gr[:16].numpy()
    np.savez_compressed(os.path.join(OUT, 'ar_tiny_grads.npz'), **out)


def golden_collate(valle, tmp):
    """The reference's own ``ValleARCollate`` on three ragged items (collate.py:19-46), and the proof that its
    ``ValleNARCollate`` raises on ragged lengths (SURVEY A-13) -- the repaired layout is pinned by construction instead."""
    from valle import collate as C
    oc = synth.tiny_config('LayerNorm')
    cfg = _ref_config(valle, oc, tmp)
    g = torch.Generator().manual_seed(41)
    items = []
    for T_, Tx_ in ((11, 4), (7, 5), (9, 3)):
        items.append({'codes': torch.randint(0, 1024, (8, T_), generator=g), 'tokens': torch.randint(1, 256, (Tx_,), generator=g)})
    out = {}
    for i, it in enumerate(items):
        out[f'item{i}_codes'], out[f'item{i}_tokens'] = _np(it['codes']), _np(it['tokens'])
    for k, v in C.ValleARCollate(cfg)(items).items():
        out['ar_' + k] = _np(v)
    try:
        C.ValleNARCollate(cfg)(items)
        out['nar_upstream_raises'] = np.array(0)
    except Exception:
        out['nar_upstream_raises'] = np.array(1)
    np.savez_compressed(os.path.join(OUT, 'collate.npz'), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix='valle_golden_')
    cwd = os.getcwd()
    os.chdir(tmp)                      # ConfigValle.__post_init__ mkdirs relative paths (K-10)
    try:
        valle = ref_shims.import_reference()
        torch.set_num_threads(1)       # fixed reduction order for the frozen vectors
        golden_modules(valle, tmp)
        golden_masks_sampling(valle)
        golden_ar(valle, tmp)
        golden_nar(valle, tmp)
        golden_ar_grads(valle, tmp)
        golden_collate(valle, tmp)
    finally:
        os.chdir(cwd)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    sys.exit(main())
