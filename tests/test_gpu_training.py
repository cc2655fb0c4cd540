"""Training step: loss and EVERY parameter gradient of ValleAR.training_step / ValleNAR.training_step computed
by the CUDA stack (valle2_b200/train.py + csrc/train.cu) against torch autograd over the CPU oracle's restatement of the
same step, plus kernel-level checks of the backward kernels.  fp32 validation mode: 1e-4; bf16: 3e-2 of the gradient's
scale (bf16 operands, fp32 accumulation)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import valle2_b200  # noqa: E402
from oracle import synth  # noqa: E402
from oracle import valle_oracle as vo  # noqa: E402
from test_gpu_models import build, rel_err  # noqa: E402
from valle2_b200.train import zero_pe_dropout  # noqa: E402


@pytest.fixture(autouse=True)
def _restore_precision():
    prev = valle2_b200.get_precision()
    yield
    valle2_b200.set_precision(prev)


@pytest.fixture(scope='module')
def ops():
    from valle2_b200 import ops as _ops
    return _ops


def _oracle_grads(sd, loss_fn):
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith('.pe') else v) for k, v in sd.items()}
    loss = loss_fn(sdg)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in sdg.items() if getattr(v, 'grad', None) is not None}


def _ar_batch(oc, B, Tx, Ty, seed, ragged=True):
    g = torch.Generator().manual_seed(seed)
    tokens = torch.randint(0, 256, (B, Tx), generator=g)
    codes = torch.randint(0, 1024, (B, Ty), generator=g)
    codes[:, 0] = oc.bos_token
    target = torch.randint(0, 1025, (B, Ty), generator=g)
    tl = torch.full((B,), Tx)
    cl = torch.full((B,), Ty)
    if ragged and B > 1:
        tl[1:] = torch.randint(max(1, Tx // 2), Tx + 1, (B - 1,), generator=g)
        cl[1:] = torch.randint(max(2, Ty // 2), Ty + 1, (B - 1,), generator=g)
    return {'tokens': tokens, 'codes': codes, 'target': target, 'tokens_lens': tl, 'codes_lens': cl}


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-4), ('bf16', 3e-2)])
@pytest.mark.parametrize('B,Tx,Ty', [(2, 7, 11), (3, 20, 70)])
def test_ar_training_step_gradients_vs_oracle_autograd(tmp_path, precision, tol, B, Tx, Ty):
    valle2_b200.set_precision(precision)
    oc = synth.tiny_config('LayerNorm')
    model, sd = build('ValleAR', oc, tmp_path, 3)
    model.train()
    zero_pe_dropout(model)          # parity against the dropout-free oracle (SURVEY K-3)
    batch = _ar_batch(oc, B, Tx, Ty, 5)
    ref_loss, ref = _oracle_grads(sd, lambda s: vo.ar_teacher_forced(s, oc, batch['tokens'], batch['codes'], batch['tokens_lens'],
                                                                     batch['codes_lens'], batch['target'])[1])
    loss = model.training_step(batch)
    assert loss.requires_grad
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) < (1e-4 if precision == 'fp32' else 5e-2)
    checked = 0
    for name, p in model.named_parameters():
        assert name in ref, name
        assert p.grad is not None, name
        scale = ref[name].abs().max().item() + 1e-12
        err = (p.grad.cpu().double() - ref[name].double()).abs().max().item() / scale
        assert err < tol, (name, err)
        checked += 1
    assert checked == len(ref)


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-4), ('bf16', 3e-2)])
@pytest.mark.parametrize('layer', [1, 4, 7])
def test_nar_training_step_gradients_vs_oracle_autograd(tmp_path, precision, tol, layer):
    """Pinned to torch autograd over the ORACLE's restatement only: the reference's own ValleNAR.training_step raises (SURVEY
    App. A-1..4), so no executed-reference golden can exist for it -- unlike the AR step, which is also checked against gradients
    of the executed reference (test_ar_training_gradients_vs_reference_golden).  The oracle's NAR step is the reference code with
    the four documented repairs; the Transformer / AdaptiveLayerNorm it drives are pinned to the executed reference module by
    module (tests/test_oracle_golden.py::test_modules)."""
    valle2_b200.set_precision(precision)
    oc = synth.tiny_config('AdaptiveLayerNorm')
    model, sd = build('ValleNAR', oc, tmp_path, 4)
    model.train()
    zero_pe_dropout(model)          # parity against the dropout-free oracle (SURVEY K-3)
    g = torch.Generator().manual_seed(6)
    B, Tx, T = 2, 6, 13
    batch = {'tokens': torch.randint(0, 256, (B, Tx), generator=g), 'codes': torch.randint(0, 1024, (B, T, 8), generator=g),
             'tokens_lens': torch.full((B,), Tx), 'codes_lens': torch.full((B,), T)}
    ref_loss, ref = _oracle_grads(sd, lambda s: vo.nar_teacher_forced(s, oc, batch['tokens'], batch['codes'], batch['tokens_lens'],
                                                                      batch['codes_lens'], layer)[1])
    loss = model.training_step(batch, layer=layer)
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) < (1e-4 if precision == 'fp32' else 5e-2)
    n_checked = 0
    for name, p in model.named_parameters():
        if name not in ref:          # parameters of other stages: no gradient on either side
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, name
        scale = ref[name].abs().max().item() + 1e-12
        err = (p.grad.cpu().double() - ref[name].double()).abs().max().item() / scale
        assert err < tol, (name, err)
        n_checked += 1
    assert n_checked >= 12 * 2 + 4


def test_training_step_drives_an_optimizer(tmp_path):
    """configure_optimizers + three AdamW steps on a fixed batch lower the loss (the loop Lightning would run,
    train_model.py:28-35)."""
    valle2_b200.set_precision('fp32')
    oc = synth.tiny_config('LayerNorm')
    model, _ = build('ValleAR', oc, tmp_path, 8)
    model.train()
    zero_pe_dropout(model)          # parity against the dropout-free oracle (SURVEY K-3)
    batch = _ar_batch(oc, 2, 6, 12, 9, ragged=False)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    losses = []
    for _ in range(4):
        opt.zero_grad()
        loss = model.training_step(batch)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0] - 0.05, losses


# ---- kernel-level checks -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dt', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('mode', ['none', 'prefix', 'ragged'])
@pytest.mark.parametrize('B,S,H', [(2, 70, 2), (1, 200, 3), (3, 64, 1), (2, 333, 2)])
def test_attention_bwd_vs_autograd(ops, dt, mode, B, S, H):
    torch.manual_seed(31)
    Dh, d = 64, H * 64
    qkv = (torch.randn(B * S, 3 * d, device='cuda') * 0.5).to(dt)
    do = torch.randn(B * S, d, device='cuda').to(dt)
    x_lens = kv_lens = None
    mask_mode = ops.MASK_NONE
    if mode != 'none':
        mask_mode = ops.MASK_PREFIX_LM
        x_lens = torch.full((B,), S // 3, dtype=torch.int32, device='cuda')
        kv_lens = torch.full((B,), S, dtype=torch.int32, device='cuda')
        if mode == 'ragged':
            kv_lens = torch.randint(S // 2, S + 1, (B,), dtype=torch.int32, device='cuda')
    # reference: dense attention in fp64 with the same predicate
    q64 = qkv.double().view(B, S, 3, H, Dh).requires_grad_(True)
    q, k, v = (q64[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    s = q @ k.transpose(-1, -2) / math.sqrt(Dh)
    i = torch.arange(S, device='cuda')[:, None]
    j = torch.arange(S, device='cuda')[None, :]
    ok = torch.ones(B, 1, S, S, dtype=torch.bool, device='cuda')
    if mode != 'none':
        xl = x_lens.view(B, 1, 1, 1).long()
        ok = ((j < xl) | ((i >= xl) & (j <= i))) & (j < kv_lens.view(B, 1, 1, 1).long())
    s = s.masked_fill(~ok, float('-inf'))
    p = torch.softmax(s, -1)
    p = torch.nan_to_num(p, nan=0.0)
    o_ref = (p @ v).permute(0, 2, 1, 3).reshape(B * S, d)
    o_ref.backward(do.double())
    ref = q64.grad.reshape(B * S, 3 * d)
    o = o_ref.detach().to(dt)
    dqkv = torch.full((B * S, 3 * d), float('nan'), device='cuda').to(dt)
    ops.attention_bwd(qkv, o, do, dqkv, B, S, H, mask_mode=mask_mode, x_lens=x_lens, kv_lens=kv_lens)
    assert rel_err(dqkv.float(), ref) < (1e-4 if dt == torch.float32 else 2e-2)
    if dt == torch.bfloat16:
        # the same gradients from the tcgen05 kernels (csrc/attn_bwd_tc.cu), which take the rows' log-sum-exp as an input
        lse = torch.logsumexp(s.detach(), dim=-1).float().contiguous()                   # (B, H, S); -inf rows do not occur here
        dq2 = torch.full((B * S, 3 * d), float('nan'), device='cuda').to(dt)
        ops.attention_bwd(qkv, o, do, dq2, B, S, H, mask_mode=mask_mode, x_lens=x_lens, kv_lens=kv_lens, lse=lse)
        assert not torch.isnan(dq2.float()).any()
        assert rel_err(dq2.float(), ref) < 2e-2


@pytest.mark.parametrize('dt', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('R,d', [(37, 256), (300, 1024), (5, 128)])
def test_layernorm_bwd_vs_autograd(ops, dt, R, d):
    torch.manual_seed(32)
    x = (torch.randn(R, d, device='cuda') * 1.5 + 0.3).requires_grad_(True)
    g = torch.randn(d, device='cuda', requires_grad=True)
    b = torch.randn(d, device='cuda', requires_grad=True)
    dy = torch.randn(R, d, device='cuda').to(dt)
    y = torch.nn.functional.layer_norm(x.double(), (d,), g.double(), b.double())
    y.backward(dy.double())
    dx = torch.randn(R, d, device='cuda')
    dx0 = dx.clone()
    dg, db = ops.layernorm_bwd(x.detach(), g.detach(), dy, dx)
    tol = 1e-4 if dt == torch.float32 else 1e-2
    assert rel_err(dx - dx0, x.grad) < tol and rel_err(dg, g.grad) < tol and rel_err(db, b.grad) < tol
    dx2 = dx0.clone()
    assert ops.layernorm_bwd(x.detach(), None, dy, dx2) == (None, None)
    assert rel_err(dx2 - dx0, dy.float()) < 1e-6


def test_small_backward_kernels(ops):
    torch.manual_seed(33)
    for dt in (torch.float32, torch.bfloat16):
        a = torch.randn(70, 45, device='cuda').to(dt)
        assert torch.equal(ops.transpose(a), a.t().contiguous())
        pre = torch.randn(1000, device='cuda').to(dt)
        dy = torch.randn(1000, device='cuda').to(dt)
        y = torch.empty_like(pre)
        ops.gelu_fwd(pre, y)
        assert rel_err(y.float(), torch.nn.functional.gelu(pre.double())) < (1e-6 if dt == torch.float32 else 1e-2)
        pg = pre.double().requires_grad_(True)
        torch.nn.functional.gelu(pg).backward(dy.double())
        out = torch.empty_like(pre)
        ops.gelu_bwd(pre, dy, out)
        assert rel_err(out.float(), pg.grad) < (1e-5 if dt == torch.float32 else 1e-2)
        x = torch.randn(333, 70, device='cuda').to(dt)
        assert rel_err(ops.colsum(x, scale=0.5), 0.5 * x.double().sum(0)) < 1e-5
        xl = torch.randn(5000, 130, device='cuda').to(dt)            # two-level path
        assert rel_err(ops.colsum(xl), xl.double().sum(0)) < 1e-5
        assert torch.equal(ops.colsum(xl), ops.colsum(xl))
    logits = torch.randn(19, 1032, device='cuda') * 3
    tgt = torch.randint(0, 1025, (19,), dtype=torch.int32, device='cuda')
    dl = torch.zeros(19, 1032, device='cuda')
    rows = ops.cross_entropy(logits, tgt, 1025, dlogits=dl, scale=1 / 19)
    lg = logits[:, :1025].double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lg, tgt.long(), reduction='none')
    ref.mean().backward()
    assert rel_err(rows, ref) < 1e-5 and rel_err(dl[:, :1025], lg.grad) < 1e-5 and float(dl[:, 1025:].abs().max()) == 0.0
    ids = torch.randint(0, 50, (3, 9, 4), dtype=torch.int32, device='cuda')
    dx = torch.randn(3 * 11, 64, device='cuda')
    gt = torch.zeros(4, 50, 64, device='cuda')
    ops.embed_bwd(ids, dx, gt, t_split=3, nq_a=4, nq_b=2, rows_per_batch=11, row_offset=2)
    ref = torch.zeros(4, 50, 64, device='cuda', dtype=torch.float64)
    for b in range(3):
        for t in range(9):
            for j in range(4 if t < 3 else 2):
                ref[j, ids[b, t, j]] += dx[b * 11 + 2 + t].double()
    assert rel_err(gt, ref) < 1e-5


def test_ar_training_gradients_vs_reference_golden(tmp_path):
    """fp32 mode against the gradients of the EXECUTED reference (tests/golden/ar_tiny_grads.npz: autograd through the
    reference's own ValleAR.training_step on the ragged batch of ar_tiny.npz)."""
    import os

    import numpy as np
    valle2_b200.set_precision('fp32')
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    g, gg = np.load(os.path.join(gdir, 'ar_tiny.npz')), np.load(os.path.join(gdir, 'ar_tiny_grads.npz'))
    oc = synth.tiny_config('LayerNorm')
    model, _ = build('ValleAR', oc, tmp_path, 0)
    batch = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('tf_') and k not in ('tf_logits', 'tf_loss')}
    loss = model.training_step(batch)
    loss.backward()
    assert abs(loss.item() - float(gg['loss'])) < 1e-4
    for name, p in model.named_parameters():
        gr = p.grad.detach().cpu().double().flatten()
        ref_stats, ref_head = gg['stats.' + name], gg['head.' + name]
        scale = ref_stats[2] + 1e-12
        assert abs(float(gr.norm()) - ref_stats[0]) / (ref_stats[0] + 1e-12) < 1e-4, name
        assert np.abs(gr[:16].numpy() - ref_head).max() / scale < 1e-4, name


def test_collated_batch_feeds_training_step(tmp_path):
    """N2 -> a13: the reference's own collated batch (tests/golden/collate.npz, produced by the executed upstream
    ValleARCollate) goes through ValleAR.training_step; loss equals the oracle's teacher-forced loss on the same dict."""
    import os
    import numpy as np
    from valle.collate import ValleARCollate
    valle2_b200.set_precision('fp32')
    oc = synth.tiny_config('LayerNorm')
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'collate.npz'))
    items = [{'codes': torch.from_numpy(z[f'item{i}_codes']), 'tokens': torch.from_numpy(z[f'item{i}_tokens'])} for i in range(3)]
    batch = ValleARCollate(oc)(items)
    model, sd = build('ValleAR', oc, tmp_path, 3)
    model.train()
    zero_pe_dropout(model)          # parity against the dropout-free oracle (SURVEY K-3)
    ref = vo.ar_teacher_forced(sd, oc, batch['tokens'], batch['codes'], batch['tokens_lens'], batch['codes_lens'], batch['target'])[1]
    loss = model.training_step(batch)
    assert abs(loss.item() - ref.item()) < 1e-4


@pytest.mark.parametrize('mode', ['none', 'prefix'])
def test_forward_lse_feeds_attention_backward(ops, mode):
    """The tensor-core forward's log-sum-exp (vb_attention_prefill_tc lse output) against a float64 softmax of the same
    bf16 q, k, and attention_bwd with that lse against attention_bwd recomputing it: same gradients to bf16 round-off."""
    torch.manual_seed(31)
    B, S, H, Dh = 2, 200, 4, 64
    d = H * Dh
    qkv = (torch.randn(B * S, 3 * d, device='cuda') * 0.7).bfloat16()
    o = torch.empty(B * S, d, device='cuda', dtype=torch.bfloat16)
    lse = torch.full((B, H, S), float('nan'), device='cuda')
    x_lens = torch.tensor([37, 90], device='cuda', dtype=torch.int32) if mode == 'prefix' else None
    kv_lens = torch.tensor([200, 161], device='cuda', dtype=torch.int32)
    mm = ops.MASK_PREFIX_LM if mode == 'prefix' else ops.MASK_NONE
    ops.attention_packed(qkv, o, B, S, H, mask_mode=mm, x_lens=x_lens, kv_lens=kv_lens, use_tc=True, lse=lse)
    q, k = qkv.view(B, S, 3, H, Dh)[:, :, 0].double(), qkv.view(B, S, 3, H, Dh)[:, :, 1].double()
    s = torch.einsum('bihd,bjhd->bhij', q, k) / math.sqrt(Dh)
    i, j = torch.arange(S, device='cuda')[:, None], torch.arange(S, device='cuda')[None, :]
    for b in range(B):
        ok = j < int(kv_lens[b])
        if mode == 'prefix':
            xl = int(x_lens[b])
            ok = ok & ((j < xl) | ((i >= xl) & (j <= i)))
        s[b].masked_fill_(~ok[None], float('-inf'))
    ref = torch.logsumexp(s, dim=-1)
    assert not torch.isnan(lse).any()
    assert (lse.double() - ref).abs().max().item() < 2e-3
    do = (torch.randn(B * S, d, device='cuda') * 0.5).bfloat16()
    g_saved, g_recomp = torch.zeros_like(qkv), torch.zeros_like(qkv)
    ops.attention_bwd(qkv, o, do, g_saved, B, S, H, mask_mode=mm, x_lens=x_lens, kv_lens=kv_lens, lse=lse)
    ops.attention_bwd(qkv, o, do, g_recomp, B, S, H, mask_mode=mm, x_lens=x_lens, kv_lens=kv_lens)
    scale = g_recomp.float().abs().max().item()
    assert (g_saved.float() - g_recomp.float()).abs().max().item() < 2e-2 * scale


def test_train_model_driver_end_to_end(tmp_path):
    """N3: the reference's training entry point (valle/train_model.py:13-44, CLI repaired per A-11) on synthetic items of the
    wire format: collate -> training_step -> backward -> clip -> AdamW + scheduler, gradient accumulation; the loss falls."""
    import json
    from valle import train_model as tm
    cfg = {'num_layers': 2, 'd_model': 256, 'n_heads': 4, 'dim_feedforward': 1024, 'norm': 'LayerNorm', 'dropout': 0.0,
           'batch_size': 4, 'grad_accum': 2, 'max_steps': 6, 'lr': 3e-3, 'lr_warmup': 100, 'log_every_n_steps': 2, 'seed': 7,
           'ckpt_path': str(tmp_path / 'c'), 'log_path': str(tmp_path / 'l')}
    fp = tmp_path / 'cfg.json'
    fp.write_text(json.dumps(cfg))
    valle2_b200.set_precision('bf16')
    lines = []
    losses = tm.train(fp, 'ValleAR', synthetic=8, log=lines.append)
    assert len(losses) == 6 and all(math.isfinite(v) for v in losses)
    assert losses[-1] < losses[0] - 0.5, losses          # 8 items seen repeatedly: the model starts to memorise them
    assert any('ms/step' in ln and 'this step' in ln for ln in lines)
    with pytest.raises(RuntimeError, match='out of scope'):
        tm.train(fp, 'ValleAR')
    # the CLI parses the reference's flags
    tm.main(['-c', str(fp), '-m', 'ValleAR', '--synthetic', '8', '--max-steps', '1'])


def test_dropout_kernel_statistics_and_determinism(ops):
    """vb_dropout: keep fraction 1 - p, survivors scaled by 1 / (1 - p), the mask is a pure function of (seed, site, index)
    (the backward pass recomputes it), different sites / seeds give different masks, both dtypes agree on the mask."""
    n = 1 << 20
    for p in (0.1, 0.5):
        x = torch.ones(n, device='cuda')
        ops.dropout_(x, p, 42, 7)
        keep = (x != 0).float().mean().item()
        assert abs(keep - (1 - p)) < 4 * math.sqrt(p * (1 - p) / n) + 1e-4
        assert torch.allclose(x[x != 0], torch.full((1,), 1 / (1 - p), device='cuda'))
        y = torch.ones(n, device='cuda')
        ops.dropout_(y, p, 42, 7)
        assert torch.equal(x, y)
        z = torch.ones(n, device='cuda', dtype=torch.bfloat16)
        ops.dropout_(z, p, 42, 7)
        assert torch.equal(z != 0, x != 0)
        for seed, site in ((43, 7), (42, 8)):
            w = torch.ones(n, device='cuda')
            ops.dropout_(w, p, seed, site)
            agree = ((w != 0) == (x != 0)).float().mean().item()
            assert abs(agree - ((1 - p) ** 2 + p ** 2)) < 0.01          # independent masks
        r = torch.randn(n, device='cuda')
        t = torch.randn(n, device='cuda')
        want = r + t * x
        ops.dropout_add_(r, t, p, 42, 7)
        assert torch.allclose(r, want, atol=1e-6)
    assert torch.equal(ops.dropout_(torch.ones(8, device='cuda'), 0.0, 1, 1), torch.ones(8, device='cuda'))


class _MaskMul(torch.nn.Module):
    def __init__(self, mask):
        super().__init__()
        self.mask = mask

    def forward(self, x):
        return x * self.mask


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-4), ('bf16', 3e-2)])
def test_ar_training_step_with_dropout_equals_reference_with_the_same_masks(tmp_path, precision, tol):
    """model.train() with config.dropout = 0.1 and the PositionalEncoding's hard-wired p = 0.1 (SURVEY K-3): torch's RNG
    stream cannot be reproduced, so the EXECUTED reference gets this step's masks (generated by the same kernel,
    train.dropout_masks) in place of its five nn.Dropout sites; loss and every parameter gradient must then agree."""
    from oracle import ref_shims
    if not ref_shims.reference_available():
        pytest.skip('executed reference not installed (python -m oracle.build_ref)')
    from valle2_b200 import train
    from valle2_b200.config import ConfigValle
    from valle2_b200.models import ValleAR
    valle2_b200.set_precision(precision)
    oc = synth.tiny_config('LayerNorm')
    fields = {k: getattr(oc, k) for k in type(oc).__dataclass_fields__}
    cfg = ConfigValle(dropout=0.1, ckpt_path=tmp_path / 'c', log_path=tmp_path / 'l', **fields)
    sd = synth.synth_state_dict(synth.ar_state_shapes(oc), 3)
    model = ValleAR(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    B, Tx, Ty = 3, 9, 21
    batch = _ar_batch(oc, B, Tx, Ty, 5)
    plan = train.dropout_plan(model, seed=1234)
    assert plan == {'p': 0.1, 'pe_p': 0.1, 'seed': 1234}
    loss = model.training_step(batch, dropout_seed=1234)
    loss.backward()
    loss_eval_mode = None
    S = Tx + Ty
    masks = {k: v.cpu() for k, v in train.dropout_masks(plan, oc.num_layers, B * S, oc.d_model, oc.dim_feedforward, 'cuda').items()}
    assert 0.8 < (masks['pe'] != 0).float().mean().item() < 0.97
    valle = ref_shims.import_reference()
    try:
        rcfg = valle.config.ConfigValle(dropout=0.1, ckpt_path=tmp_path / 'rc', log_path=tmp_path / 'rl', **fields)
        ref = valle.models.ValleAR(rcfg)
        ref.load_state_dict(sd, strict=True)
        ref.eval()
        with torch.no_grad():
            loss_eval_mode = float(ref.training_step(batch))        # nothing dropped
        ref.train()
        pe = masks['pe'].view(B, S, -1)
        # the reference's PositionalEncoding drops in its (T, B, C) layout (modules.py:76-78)
        ref.tokens_position_emb.dropout = _MaskMul(pe[:, :Tx].transpose(0, 1))
        ref.audio_position_emb.dropout = _MaskMul(pe[:, Tx:].transpose(0, 1))
        for li, layer in enumerate(ref.transformer.layers):
            layer.dropout1 = _MaskMul(masks[f'{li}.attn'].view(B, S, -1))
            layer.ffn.dropout = _MaskMul(masks[f'{li}.ffn_inner'].view(B, S, -1))
            layer.dropout2 = _MaskMul(masks[f'{li}.ffn'].view(B, S, -1))
        ref_loss = ref.training_step(batch)
        ref_loss.backward()
        ref_grads = {n: p.grad.clone() for n, p in ref.named_parameters()}
        ref_loss = float(ref_loss.detach())
    finally:
        ref_shims.release_reference()
    assert abs(float(ref_loss) - loss_eval_mode) > 1e-3                 # the masks do change the step
    assert abs(loss.item() - float(ref_loss)) < (1e-4 if precision == 'fp32' else 5e-2)
    for name, p in model.named_parameters():
        scale = ref_grads[name].abs().max().item() + 1e-12
        err = (p.grad.cpu().double() - ref_grads[name].double()).abs().max().item() / scale
        assert err < tol, (name, err)
    # eval mode drops nothing; a second training step draws new masks
    model.eval()
    assert train.dropout_plan(model) is None
    model.train()
    a, b = train.dropout_plan(model), train.dropout_plan(model)
    assert a['seed'] != b['seed']


@pytest.mark.skipif(__import__('os').environ.get('VALLE_B200_TEST_DEVICE_ASSERT', '0') != '1',
                    reason='provokes a CUDA device-side assert in a child process on purpose; opt in with VALLE_B200_TEST_DEVICE_ASSERT=1 '
                           '(ran green twice on a B200 this round: it was among the 600 passes of the two full runs before it became opt-in)')
def test_out_of_range_training_ids_fail_like_the_reference(tmp_path):
    """An id outside the embedding table / class range makes the reference's nn.Embedding / F.cross_entropy fail (a device-side
    assert on CUDA, modules.py:34, valle_ar.py:85); csrc/train.cu clamps, so train.py asserts first.  A device-side assert
    poisons the CUDA context, hence the child process: the good batch must pass, the bad one must not."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f'''
import sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, "tests")!r})
from oracle import synth
from test_gpu_models import build
import pathlib
oc = synth.tiny_config('LayerNorm')
model, sd = build('ValleAR', oc, pathlib.Path({str(tmp_path)!r}), 3)
model.train()
g = torch.Generator().manual_seed(0)
batch = {{'tokens': torch.randint(0, 256, (2, 6), generator=g), 'tokens_lens': torch.tensor([6, 4]),
         'codes': torch.randint(0, 1024, (2, 9), generator=g), 'codes_lens': torch.tensor([9, 7]),
         'target': torch.randint(0, 1025, (2, 9), generator=g)}}
loss = model.training_step(batch); loss.backward(); torch.cuda.synchronize()
print('GOOD_OK', flush=True)
batch['target'][0, 3] = 1025 if sys.argv[1] == 'target' else batch['target'][0, 3]
if sys.argv[1] == 'codes':
    batch['codes'][1, 2] = 5000
try:
    loss = model.training_step(batch); loss.backward(); torch.cuda.synchronize()
except Exception as e:
    print('BAD_RAISED', type(e).__name__, flush=True)
    sys.exit(0)
print('BAD_PASSED', flush=True)
'''
    for which in ('target', 'codes'):
        r = subprocess.run([sys.executable, '-c', code, which], capture_output=True, text=True, timeout=300)
        assert 'GOOD_OK' in r.stdout, (r.stdout[-500:], r.stderr[-1500:])
        assert 'BAD_PASSED' not in r.stdout, which
        assert 'BAD_RAISED' in r.stdout or r.returncode != 0, (which, r.stdout[-500:], r.stderr[-500:])
