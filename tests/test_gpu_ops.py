"""Kernel-level parity on a B200: every C-ABI entry point against a plain torch / oracle computation of the
same op on the same seeded inputs.  Tolerances: fp32 kernels 1e-5 (relative to the output scale), bf16 tensor-core
kernels 1e-2 (bf16 operands, fp32 accumulate)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import valle_oracle as vo  # noqa: E402


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


@pytest.fixture(scope='module')
def ops():
    from valle2_b200 import ops as _ops
    info = _ops.device_info()
    assert info['cc'][0] == 10
    return _ops


def test_embed_sum_pe(ops):
    torch.manual_seed(0)
    B, T, Q, V, d = 3, 7, 4, 50, 64
    ids = torch.randint(0, V, (B, T, Q), dtype=torch.int32, device='cuda')
    tables = torch.randn(Q, V, d, device='cuda')
    pe = vo.sinusoidal_pe(100, d).cuda()
    out = torch.zeros(B, T + 2, d, device='cuda')
    ops.embed_sum_pe(ids, tables, pe, out.view(-1, d), t_split=3, nq_a=Q, nq_b=2, pos_offset=5,
                     out_rows_per_batch=T + 2, out_row_offset=2)
    ref = torch.zeros(B, T, d, device='cuda')
    for t in range(T):
        nq = Q if t < 3 else 2
        for j in range(nq):
            ref[:, t] = ref[:, t] + tables[j][ids[:, t, j].long()]
        ref[:, t] = ref[:, t] + pe[5 + t]
    assert torch.equal(out[:, 2:], ref)
    assert (out[:, :2] == 0).all()
    pos_b = torch.tensor([0, 10, 20], dtype=torch.int32, device='cuda')
    out2 = torch.zeros(B, 1, d, device='cuda')
    ops.embed_sum_pe(ids[:, :1, :1].contiguous(), tables[:1].contiguous(), pe, out2.view(-1, d), pos_b=pos_b)
    ref2 = tables[0][ids[:, 0, 0].long()] + pe[pos_b.long()]
    assert torch.equal(out2[:, 0], ref2)


@pytest.mark.parametrize('d', [1024, 512, 256, 64, 96])
@pytest.mark.parametrize('ydt', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('R,ns', [(37, 3), (1500, 5), (3, 16), (32, 8), (1, 4), (32, 6)])      # ns >= 4, R <= 1024, d in {256,512,1024}: 4-CTA cluster kernel
def test_residual_layernorm(ops, d, ydt, R, ns):
    torch.manual_seed(1)
    x = torch.randn(R, d, device='cuda') * 2 + 0.5
    g, b = torch.randn(d, device='cuda'), torch.randn(d, device='cuda')
    part = torch.randn(ns, R, d, device='cuda')
    bias = torch.randn(d, device='cuda')
    y = torch.empty(R, d, device='cuda', dtype=ydt)
    x1 = x.clone()
    ops.residual_layernorm(x1, g, b, y, eps=1e-5)
    ref = vo.layer_norm(x.cpu(), g.cpu(), b.cpu())
    tol = 1e-5 if ydt == torch.float32 else 8e-3
    assert rel_err(y.float(), ref) < tol
    assert torch.equal(x1, x)
    x2 = x.clone()
    ops.residual_layernorm(x2, g, b, y, part=part, n_part=ns, part_stride=R * d, bias=bias)
    xr = x.cpu() + (bias.cpu() + part.cpu().sum(0))
    assert rel_err(x2, xr) < 2e-6
    assert rel_err(y.float(), vo.layer_norm(xr, g.cpu(), b.cpu())) < tol
    # residual update only (y = None) and slices + plain cast (the last decode LayerNorm call of a step, K-2)
    x3 = x.clone()
    ops.residual_layernorm(x3, None, None, None, part=part, n_part=ns, part_stride=R * d, bias=bias)
    assert rel_err(x3, xr) < 2e-6
    x4 = x.clone()
    ops.residual_layernorm(x4, None, None, y, part=part, n_part=ns, part_stride=R * d, bias=bias)
    assert rel_err(x4, xr) < 2e-6
    assert torch.equal(y, x4) if ydt == torch.float32 else rel_err(y.float(), x4) < 4e-3     # y = cast(updated x)
    # cast-only mode
    ops.residual_layernorm(x2, None, None, y)
    assert rel_err(y.float(), x2) < (1e-7 if ydt == torch.float32 else 4e-3)


def test_reduce_bias_act(ops):
    torch.manual_seed(2)
    R, N, ns = 5, 1024, 4
    part = torch.randn(ns, R, N, device='cuda')
    bias = torch.randn(N, device='cuda')
    y = torch.empty(R, N, device='cuda', dtype=torch.bfloat16)
    ops.reduce_bias_act(part, ns, R * N, bias, True, y)
    ref = vo.gelu_erf((part.sum(0) + bias).cpu())
    assert rel_err(y.float(), ref) < 8e-3
    y32 = torch.empty(R, N, device='cuda')
    ops.reduce_bias_act(part, ns, R * N, bias, False, y32)
    assert rel_err(y32, (part[0] + part[1] + part[2] + part[3] + bias)) < 1e-6


EPIS = ['none', 'bias', 'gelu', 'residual']


def _ref_linear(x, w, bias, res, epi):
    y = x.double() @ w.double().t()
    if epi != 'none':
        y = y + bias.double()
    if epi == 'gelu':
        y = 0.5 * y * (1 + torch.erf(y / math.sqrt(2)))
    if epi == 'residual':
        y = y + res.double()
    return y


@pytest.mark.parametrize('M,N,K', [(37, 100, 64), (128, 256, 1024), (5, 1025, 256), (200, 192, 60)])
@pytest.mark.parametrize('epi', EPIS)
def test_linear_fp32(ops, M, N, K, epi):
    torch.manual_seed(3)
    x, w = torch.randn(M, K, device='cuda'), torch.randn(N, K, device='cuda') / math.sqrt(K)
    bias, res = torch.randn(N, device='cuda'), torch.randn(M, N, device='cuda')
    y = ops.linear(x, w, None if epi == 'none' else bias, gelu=(epi == 'gelu'),
                   residual=res if epi == 'residual' else None)
    assert rel_err(y, _ref_linear(x, w, bias, res, epi)) < 1e-5


@pytest.mark.parametrize('M,N,K', [(128, 256, 64), (256, 512, 128), (300, 1024, 1024), (57, 192, 64),
                                   (1000, 3072, 1024), (129, 100, 256), (4096, 1024, 4096), (64, 1025, 1024)])
@pytest.mark.parametrize('epi', EPIS)
@pytest.mark.parametrize('ydt', [torch.bfloat16, torch.float32])
def test_linear_bf16_tc(ops, M, N, K, epi, ydt):
    torch.manual_seed(4)
    x = torch.randn(M, K, device='cuda').bfloat16()
    w = (torch.randn(N, K, device='cuda') / math.sqrt(K)).bfloat16()
    bias, res = torch.randn(N, device='cuda'), torch.randn(M, N, device='cuda')
    y = ops.linear(x, w, None if epi == 'none' else bias, gelu=(epi == 'gelu'),
                   residual=res if epi == 'residual' else None, out_dtype=ydt)
    torch.cuda.synchronize()
    err = rel_err(y.float(), _ref_linear(x, w, bias, res, epi))
    assert err < (1e-2 if ydt == torch.bfloat16 else 1e-4), err


def test_linear_bf16_inplace_residual(ops):
    torch.manual_seed(5)
    M, N, K = 384, 1024, 1024
    x = torch.randn(M, K, device='cuda').bfloat16()
    w = (torch.randn(N, K, device='cuda') / math.sqrt(K)).bfloat16()
    bias, res = torch.randn(N, device='cuda'), torch.randn(M, N, device='cuda')
    ref = _ref_linear(x, w, bias, res, 'residual')
    ops.linear(x, w, bias, residual=res, out=res)
    assert rel_err(res, ref) < 1e-4


@pytest.mark.parametrize('M', [1, 5, 16, 32, 33, 64, 200, 256])
@pytest.mark.parametrize('N,K', [(1025, 1024), (3072, 1024), (1024, 4096), (256, 256), (768, 64)])
def test_linear_decode(ops, M, N, K):
    torch.manual_seed(6)
    x = torch.randn(M, K, device='cuda').bfloat16()
    w = (torch.randn(N, K, device='cuda') / math.sqrt(K)).bfloat16()
    ns_max = 32
    part = torch.full((ns_max, M, N), float('nan'), device='cuda')
    ns = ops.linear_decode(x, w, part, M * N, ns_max, flags=(M % 2))     # odd M also exercises the late PDL trigger
    assert ns == ops.linear_decode_splits(N, K, ns_max) and 1 <= ns <= ns_max
    torch.cuda.synchronize()
    y = part[:ns].sum(0)
    assert not torch.isnan(y).any()
    assert rel_err(y, x.double() @ w.double().t()) < 1e-4


@pytest.mark.parametrize('M', [1, 5, 16, 17, 32])
@pytest.mark.parametrize('N,K,epi,ydt,split', [
    (3072, 1024, 'none', torch.float32, 1), (3072, 1024, 'none', torch.float32, 2), (1024, 1024, 'residual', torch.float32, 1),
    (4096, 1024, 'gelu', torch.bfloat16, 1), (1024, 4096, 'none', torch.float32, 1), (1025, 1024, 'none', torch.float32, 1),
    (768, 256, 'none', torch.float32, 1), (256, 256, 'residual', torch.float32, 1), (1024, 256, 'gelu', torch.bfloat16, 1),
    (256, 1024, 'bias', torch.bfloat16, 1), (1025, 256, 'none', torch.float32, 1), (520, 512, 'bias', torch.float32, 1),
    (8192, 1024, 'gelu', torch.bfloat16, 1), (104, 768, 'none', torch.float32, 1)])
def test_linear_decode_rows(ops, M, N, K, epi, ydt, split):
    """csrc/gemm_decode_mma.cu against a float64 matmul of the same bf16 operands: every epilogue, ragged N (1025 = the AR
    logits), split-K slices (K = 4096 -> 4 slices; want_split = 2), M across both m16 tile counts; late PDL trigger on odd M."""
    torch.manual_seed(6)
    x = torch.randn(M, K, device='cuda').bfloat16()
    w = (torch.randn(N, K, device='cuda') / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device='cuda') if epi != 'none' else None
    ref = x.double() @ w.double().t()
    ns = ops.linear_decode_rows_splits(K, split, M)
    assert ns == {768: 3}.get(K, max(split, K // 1024, 1))
    if ns > 1:
        y = torch.full((ns, M, N), float('nan'), device='cuda')
        got_ns = ops.linear_decode_rows(x, w, y, want_split=split, flags=(M % 2))
        assert got_ns == ns
        torch.cuda.synchronize()
        out = y.sum(0)
    else:
        res = torch.randn(M, N, device='cuda')
        y = res.clone() if epi == 'residual' else torch.full((M, N), float('nan'), device='cuda').to(ydt)
        ops.linear_decode_rows(x, w, y, bias=bias, gelu=(epi == 'gelu'), residual=(epi == 'residual'), flags=(M % 2))
        torch.cuda.synchronize()
        out = y
        if epi != 'none':
            ref = ref + bias.double()
        if epi == 'gelu':
            ref = torch.nn.functional.gelu(ref)
        if epi == 'residual':
            ref = ref + res.double()
    assert not torch.isnan(out.float()).any()
    assert rel_err(out, ref) < (1e-4 if ydt == torch.float32 else 6e-3)


@pytest.mark.parametrize('M', [1, 3, 8])
@pytest.mark.parametrize('N,K,epi,ydt,norm', [
    (3072, 1024, 'none', torch.float32, True), (4096, 1024, 'gelu', torch.bfloat16, True), (1025, 1024, 'none', torch.float32, False),
    (768, 256, 'none', torch.float32, True), (1024, 256, 'gelu', torch.bfloat16, True), (1025, 256, 'none', torch.float32, False),
    (512, 512, 'bias', torch.float32, True), (256, 1024, 'residual', torch.float32, True)])
def test_linear_decode_rows_layernorm_on_load(ops, M, N, K, epi, ydt, norm):
    """vb_linear_decode_rows_ln: y = epilogue(LN(x) @ w.T) against the oracle's layer_norm (modules.py:271/276) followed by a
    float64 matmul of the bf16-rounded rows; rows with a large common offset check the two-level (Chan) variance."""
    torch.manual_seed(8)
    x = torch.randn(M, K, device='cuda') * 2 + 3.0
    x[0] += 50.0
    w = (torch.randn(N, K, device='cuda') / math.sqrt(K)).bfloat16()
    gamma, beta = (torch.randn(K, device='cuda'), torch.randn(K, device='cuda')) if norm else (None, None)
    bias = torch.randn(N, device='cuda') if epi != 'none' else None
    hn = vo.layer_norm(x.cpu(), gamma.cpu(), beta.cpu()) if norm else x.cpu()
    ref = hn.bfloat16().double() @ w.cpu().double().t()
    res = torch.randn(M, N, device='cuda')
    y = res.clone() if epi == 'residual' else torch.full((M, N), float('nan'), device='cuda').to(ydt)
    ops.linear_decode_rows_ln(x, w, y, gamma=gamma, beta=beta, bias=bias, gelu=(epi == 'gelu'), residual=(epi == 'residual'),
                              flags=(M % 2))
    torch.cuda.synchronize()
    if epi != 'none':
        ref = ref + bias.cpu().double()
    if epi == 'gelu':
        ref = torch.nn.functional.gelu(ref)
    if epi == 'residual':
        ref = ref + res.cpu().double()
    assert not torch.isnan(y.float()).any()
    # LN output is rounded to bf16 before the MMA on both sides; a 1-ulp flip of a few elements is the remaining noise
    assert rel_err(y.float().cpu(), ref) < (2e-3 if ydt == torch.float32 else 6e-3)


@pytest.mark.parametrize('M', [1, 8])
def test_linear_decode_rows_whole_k_4096(ops, M):
    """want_split=0: FFN2 (K = 4096) inside one CTA with the residual epilogue (modules.py:278), small batches only."""
    torch.manual_seed(9)
    N, K = 1024, 4096
    assert ops.linear_decode_rows_splits(K, 0, M) == 1 and ops.linear_decode_rows_splits(K, 0, 32) == 4
    x = torch.randn(M, K, device='cuda').bfloat16()
    w = (torch.randn(N, K, device='cuda') / math.sqrt(K)).bfloat16()
    bias, res = torch.randn(N, device='cuda'), torch.randn(M, N, device='cuda')
    y = res.clone()
    assert ops.linear_decode_rows(x, w, y, bias=bias, residual=True, want_split=0) == 1
    torch.cuda.synchronize()
    assert rel_err(y, x.double() @ w.double().t() + bias.double() + res.double()) < 1e-4


def test_linear_decode_rows_is_deterministic_and_rejects_bad_shapes(ops):
    torch.manual_seed(7)
    x = torch.randn(32, 1024, device='cuda').bfloat16()
    w = (torch.randn(3072, 1024, device='cuda') / 32).bfloat16()
    a, b = torch.empty(32, 3072, device='cuda'), torch.empty(32, 3072, device='cuda')
    ops.linear_decode_rows(x, w, a)
    ops.linear_decode_rows(x, w, b)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError, match='M = 33'):
        ops.linear_decode_rows(torch.zeros(33, 1024, device='cuda').bfloat16(), w, torch.empty(33, 3072, device='cuda'))
    assert ops.linear_decode_rows_splits(1000) == 0
    with pytest.raises(RuntimeError, match='M = 9'):
        ops.linear_decode_rows_ln(torch.zeros(9, 1024, device='cuda'), w, torch.empty(9, 3072, device='cuda'))


@pytest.mark.parametrize('B', [1, 7, 16, 32, 33, 64])
@pytest.mark.parametrize('d,F', [(1024, 4096), (256, 1024)])
def test_decode_chain_matches_separate_kernels(ops, B, d, F):
    """The persistent chain kernel (out-proj -> LN -> FFN1 -> GELU -> FFN2 -> LN -> QKV, modules.py:271-278) against the
    same sequence launched as separate kernels: split-K slices bit-identical, LN / GELU rows to fp32 round-off; run
    three times back to back to exercise the self-resetting grid-barrier counter."""
    torch.manual_seed(11)
    dev = 'cuda'
    bf = torch.bfloat16
    o = torch.randn(B, d, device=dev).to(bf)
    x0 = torch.randn(B, d, device=dev)
    wo, w1 = (torch.randn(d, d, device=dev) / math.sqrt(d)).to(bf), (torch.randn(F, d, device=dev) / math.sqrt(d)).to(bf)
    w2, wq = (torch.randn(d, F, device=dev) / math.sqrt(F)).to(bf), (torch.randn(3 * d, d, device=dev) / math.sqrt(d)).to(bf)
    bo, b1, b2 = torch.randn(d, device=dev), torch.randn(F, device=dev), torch.randn(d, device=dev)
    g2, be2, g1, be1 = (torch.randn(d, device=dev) for _ in range(4))
    ns = {k: ops.linear_decode_splits(n, kk, 32) for k, (n, kk) in
          {'qkv': (3 * d, d), 'o': (d, d), 'f1': (F, d), 'f2': (d, F)}.items()}

    def bufs():
        return {'x': x0.clone(), 'h': torch.zeros(B, d, device=dev, dtype=bf), 'h2': torch.zeros(B, d, device=dev, dtype=bf),
                'f': torch.zeros(B, F, device=dev, dtype=bf),
                'p_o': torch.full((ns['o'], B, d), float('nan'), device=dev),
                'p_f1': torch.full((ns['f1'], B, F), float('nan'), device=dev),
                'p_f2': torch.full((ns['f2'], B, d), float('nan'), device=dev),
                'p_qkv': torch.full((ns['qkv'], B, 3 * d), float('nan'), device=dev)}

    r = bufs()
    ops.linear_decode(o, wo, r['p_o'], B * d, 32)
    ops.residual_layernorm(r['x'], g2, be2, r['h'], part=r['p_o'], n_part=ns['o'], part_stride=B * d, bias=bo)
    ops.linear_decode(r['h'], w1, r['p_f1'], B * F, 32)
    ops.reduce_bias_act(r['p_f1'], ns['f1'], B * F, b1, True, r['f'])
    ops.linear_decode(r['f'], w2, r['p_f2'], B * d, 32)
    ops.residual_layernorm(r['x'], g1, be1, r['h2'], part=r['p_f2'], n_part=ns['f2'], part_stride=B * d, bias=b2)
    ops.linear_decode(r['h2'], wq, r['p_qkv'], B * 3 * d, 32)
    torch.cuda.synchronize()

    gbar = torch.zeros(64, device=dev, dtype=torch.int32)
    for rep in range(3):
        c = bufs()
        ops.decode_chain([
            ops.chain_gemm(o, wo, c['p_o'], B * d),
            ops.chain_ln(c['x'], g2, be2, c['h'], part=c['p_o'], n_part=ns['o'], part_stride=B * d, bias=bo),
            ops.chain_gemm(c['h'], w1, c['p_f1'], B * F),
            ops.chain_act(c['p_f1'], ns['f1'], B * F, b1, c['f']),
            ops.chain_gemm(c['f'], w2, c['p_f2'], B * d),
            ops.chain_ln(c['x'], g1, be1, c['h2'], part=c['p_f2'], n_part=ns['f2'], part_stride=B * d, bias=b2),
            ops.chain_gemm(c['h2'], wq, c['p_qkv'], B * 3 * d)], B, gbar)
        torch.cuda.synchronize()
        assert int(gbar[0].item()) == 0, 'grid-barrier counter not reset'
        assert rel_err(c['p_o'][:ns['o']].sum(0), r['p_o'][:ns['o']].sum(0)) < 1e-5, 'out-proj slices differ'
        # downstream stages see LN / GELU rows that may differ by fp32 round-off before the bf16 rounding
        assert rel_err(c['x'], r['x']) < 2e-3      # after FFN2: bf16 roundings of h / f may flip by one ulp
        for k in ('h', 'f', 'h2'):
            assert rel_err(c[k].float(), r[k].float()) < 1e-2, k
        for k, n in (('p_f1', 'f1'), ('p_f2', 'f2'), ('p_qkv', 'qkv')):
            assert not torch.isnan(c[k][:ns[n]]).any()
            assert rel_err(c[k][:ns[n]].sum(0), r[k][:ns[n]].sum(0)) < 2e-2, k
    # and against fp64 math end to end
    xr = x0.double() + o.double() @ wo.double().t() + bo.double()
    assert rel_err(c['p_o'].sum(0) + bo + x0, xr) < 1e-4


def test_decode_chain_plain_cast_and_first_layer(ops):
    """LN phase variants: n_part = 0 (first layer: normalise x as is) and gamma = None (plain bf16 cast)."""
    torch.manual_seed(12)
    B, d, N = 5, 1024, 1025
    dev = 'cuda'
    x = torch.randn(B, d, device=dev)
    g, be = torch.randn(d, device=dev), torch.randn(d, device=dev)
    w = (torch.randn(N, d, device=dev) / math.sqrt(d)).bfloat16()
    ns = ops.linear_decode_splits(N, d, 32)
    gbar = torch.zeros(64, device=dev, dtype=torch.int32)
    for gamma, beta in ((g, be), (None, None)):
        h = torch.zeros(B, d, device=dev, dtype=torch.bfloat16)
        part = torch.full((ns, B, N), float('nan'), device=dev)
        xc = x.clone()
        ops.decode_chain([ops.chain_ln(xc, gamma, beta, h), ops.chain_gemm(h, w, part, B * N)], B, gbar)
        torch.cuda.synchronize()
        assert torch.equal(xc, x)
        ref_h = torch.nn.functional.layer_norm(x, (d,), g, be) if gamma is not None else x
        assert rel_err(h.float(), ref_h) < 1e-2
        assert rel_err(part.sum(0), h.double() @ w.double().t()) < 1e-4


@pytest.mark.parametrize('B', [1, 7, 32, 33, 64])
@pytest.mark.parametrize('N,K,cluster', [(3072, 1024, 0), (3072, 1024, 4), (1024, 1024, 0), (1024, 1024, 8), (4096, 1024, 0),
                                       (1024, 4096, 0), (1024, 4096, 8), (1025, 1024, 0), (256, 256, 2), (200, 64, 1),
                                       (1024, 2048, 5), (96, 1536, 3), (384, 512, 1)])
def test_linear_decode_fused_bf16(ops, B, N, K, cluster):
    """Cluster split-K decode GEMM (partials reduced through DSMEM) against fp64 math, all epilogues."""
    torch.manual_seed(21)
    a = torch.randn(B, K, device='cuda').bfloat16()
    w = (torch.randn(N, K, device='cuda') / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device='cuda')
    ref = a.double() @ w.double().t()
    y = torch.full((B, N), float('nan'), device='cuda')
    ops.linear_decode_fused(a, w, y, cluster_k=cluster, flags=B % 2)
    assert rel_err(y, ref) < 1e-4
    y2 = torch.full((B, N), float('nan'), device='cuda')
    ops.linear_decode_fused(a, w, y2, cluster_k=cluster)
    assert torch.equal(y, y2), 'not deterministic'
    yb = torch.zeros(B, N, device='cuda', dtype=torch.bfloat16)
    ops.linear_decode_fused(a, w, yb, bias=bias, gelu=True, cluster_k=cluster)
    assert rel_err(yb.float(), torch.nn.functional.gelu(ref + bias.double())) < 1e-2
    x = torch.randn(B, N, device='cuda')
    x0 = x.clone()
    ops.linear_decode_fused(a, w, x, bias=bias, residual=True, cluster_k=cluster)
    assert rel_err(x, x0.double() + ref + bias.double()) < 1e-4


@pytest.mark.parametrize('B', [1, 5, 32, 40, 64])
@pytest.mark.parametrize('d,N,cluster', [(1024, 4096, 0), (1024, 3072, 0), (1024, 3072, 4), (256, 768, 0), (1024, 1025, 0),
                                       (512, 512, 1), (1024, 512, 16), (512, 256, 3)])
def test_linear_decode_fused_layernorm_on_load(ops, B, d, N, cluster):
    """A = LayerNorm(x) computed inside the GEMM from the fp32 residual rows, row statistics combined across the cluster
    (modules.py:271/276), and the plain-cast variant (no final norm before valle_ar.py:158)."""
    torch.manual_seed(22)
    x = torch.randn(B, d, device='cuda') * 2 + 0.5
    g, be = torch.randn(d, device='cuda'), torch.randn(d, device='cuda')
    w = (torch.randn(N, d, device='cuda') / math.sqrt(d)).bfloat16()
    bias = torch.randn(N, device='cuda')
    h = torch.nn.functional.layer_norm(x, (d,), g, be).bfloat16()          # the operand the kernel builds in smem
    y = torch.full((B, N), float('nan'), device='cuda')
    ops.linear_decode_fused(x, w, y, gamma=g, beta=be, cluster_k=cluster)
    assert rel_err(y, h.double() @ w.double().t()) < 3e-3                    # bf16 roundings of LN(x) may flip by an ulp
    f = torch.zeros(B, N, device='cuda', dtype=torch.bfloat16)
    ops.linear_decode_fused(x, w, f, bias=bias, gelu=True, gamma=g, beta=be, cluster_k=cluster)
    assert rel_err(f.float(), torch.nn.functional.gelu(h.double() @ w.double().t() + bias.double())) < 1e-2
    yc = torch.full((B, N), float('nan'), device='cuda')
    ops.linear_decode_fused(x, w, yc, cluster_k=cluster)                     # plain cast
    assert rel_err(yc, x.bfloat16().double() @ w.double().t()) < 1e-4


def _pool_swizzle(x):
    """bf16 KV pool rows (64 tokens x 64 dims): 16-byte chunk c of token t is stored at chunk c ^ (t & 7)
    (include/valle_b200.h, 'Paged KV pool layout').  The permutation is its own inverse."""
    *lead, T, D = x.shape
    assert T == 64 and D == 64
    t = torch.arange(64, device=x.device)[:, None]
    c = torch.arange(8, device=x.device)[None, :]
    src = ((c ^ (t & 7)) * 8)[:, :, None] + torch.arange(8, device=x.device)[None, None, :]      # (64, 8, 8)
    return torch.gather(x, -1, src.reshape(64, 64).expand(*lead, 64, 64))


def _dense_attention(q, k, v, allowed):
    s = (q.double() @ k.double().transpose(-1, -2)) / math.sqrt(q.shape[-1])
    if allowed is not None:
        s = s.masked_fill(~allowed, float('-inf'))
    return torch.softmax(s, -1) @ v.double()


@pytest.mark.parametrize('dt', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('Dh', [64, 16])
def test_attention_masks(ops, dt, Dh):
    torch.manual_seed(7)
    B, H, S = 3, 4, 45
    d = H * Dh
    qkv = torch.randn(B, S, 3, H, Dh, device='cuda').to(dt)
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    out = torch.empty(B, S, d, device='cuda', dtype=dt)
    tol = 1e-5 if dt == torch.float32 else 1e-2
    # no mask
    ops.attention(q, k, v, out)
    ref = _dense_attention(q.float(), k.float(), v.float(), None).permute(0, 2, 1, 3).reshape(B, S, d)
    assert rel_err(out.float(), ref) < tol
    # prefix-LM + key padding, evaluated in-kernel from lengths
    x_lens = torch.tensor([10, 10, 10], dtype=torch.int32, device='cuda')
    kv_lens = torch.tensor([45, 30, 41], dtype=torch.int32, device='cuda')
    ops.attention(q, k, v, out, mask_mode=ops.MASK_PREFIX_LM, x_lens=x_lens, kv_lens=kv_lens)
    m = vo.build_attn_mask(10, S - 10).cuda()[None, None].expand(B, H, S, S).clone()
    pad = torch.arange(S, device='cuda')[None, :] >= kv_lens[:, None]
    m = m | pad[:, None, None, :]
    ref = _dense_attention(q.float(), k.float(), v.float(), ~m).permute(0, 2, 1, 3).reshape(B, S, d)
    assert rel_err(out.float(), ref) < tol
    # the same mask, materialised (module-level path)
    ops.attention(q, k, v, out, mask_mode=ops.MASK_EXPLICIT, mask=m.to(torch.uint8).contiguous())
    assert rel_err(out.float(), ref) < tol


@pytest.mark.parametrize('B,S,H', [(2, 300, 4), (3, 128, 2), (2, 77, 16), (2, 900, 3), (1, 1, 1), (4, 129, 2)])
@pytest.mark.parametrize('mode', ['none', 'prefix', 'ragged'])
def test_attention_prefill_tc(ops, B, S, H, mode):
    """tcgen05 flash attention vs a float64 dense softmax on the same bf16 qkv."""
    torch.manual_seed(11)
    Dh = 64
    d = H * Dh
    qkv = (torch.randn(B * S, 3 * d, device='cuda') * 1.5).bfloat16()
    out = torch.full((B * S, d), float('nan'), device='cuda', dtype=torch.bfloat16)
    x_len = max(1, S // 3)
    xl = torch.full((B,), x_len, dtype=torch.int32, device='cuda')
    if mode == 'ragged':
        kl = torch.tensor([max(1, S - 17 * b) for b in range(B)], dtype=torch.int32, device='cuda')
    else:
        kl = torch.full((B,), S, dtype=torch.int32, device='cuda')
    mm = ops.MASK_NONE if mode == 'none' else ops.MASK_PREFIX_LM
    ops.attention_packed(qkv, out, B, S, H, mask_mode=mm, x_lens=xl, kv_lens=kl, use_tc=True)
    torch.cuda.synchronize()
    v5 = qkv.view(B, S, 3, H, Dh).float()
    q, k, v = (v5[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    allowed = torch.ones(B, H, S, S, dtype=torch.bool, device='cuda')
    if mode != 'none':
        allowed &= ~vo.build_attn_mask(x_len, S - x_len).cuda()[None, None]
    allowed &= (torch.arange(S, device='cuda')[None, :] < kl[:, None])[:, None, None, :]
    ref = _dense_attention(q, k, v, allowed).permute(0, 2, 1, 3).reshape(B, S, d)
    got = out.view(B, S, d).float()
    for b in range(B):        # query rows beyond kv_len are padding: not compared
        n = int(kl[b])
        assert torch.isfinite(got[b, :n]).all()
        assert rel_err(got[b, :n], ref[b, :n]) < 1.5e-2, (b, rel_err(got[b, :n], ref[b, :n]))


@pytest.mark.parametrize('dt', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('n_tsplit', [1, 4])
@pytest.mark.parametrize('flags', [0, 2, 4])
def test_attn_decode_paged(ops, dt, n_tsplit, flags):
    torch.manual_seed(8)
    B, H, Dh, max_pages = 5, 4, 64, 6
    d = H * Dh
    seq = torch.tensor([1, 63, 64, 200, 319], dtype=torch.int32, device='cuda')
    n_pages = B * max_pages
    perm = torch.randperm(n_pages, device='cuda').to(torch.int32)
    bt = perm.view(B, max_pages).contiguous()
    dense_k = torch.randn(B, H, max_pages * 64, Dh, device='cuda').to(dt)
    dense_v = torch.randn(B, H, max_pages * 64, Dh, device='cuda').to(dt)
    pool = torch.zeros(n_pages, 2, H, 64, Dh, device='cuda', dtype=dt)
    for b in range(B):
        for p in range(max_pages):
            pool[bt[b, p], 0] = dense_k[b, :, p * 64:(p + 1) * 64]
            pool[bt[b, p], 1] = dense_v[b, :, p * 64:(p + 1) * 64]
    if dt == torch.bfloat16:
        pool = _pool_swizzle(pool)
    n_part = 3
    part = torch.randn(n_part, B, 3 * d, device='cuda')
    out = torch.empty(B, d, device='cuda', dtype=dt)
    ws = torch.zeros(ops.attn_decode_ws_bytes(B, H, n_tsplit) // 4 + 16, dtype=torch.int32, device='cuda')
    for rep in range(2):   # second launch checks that the split counters reset themselves
        pool_run = pool.clone()
        ops.attn_decode_paged(part, n_part, B * 3 * d, pool_run, bt, seq, out, B, H, Dh, n_tsplit, ws, flags)
        torch.cuda.synchronize()
        qkv = (part[0] + part[1] + part[2]).view(B, 3, H, Dh)
        for b in range(B):
            n = int(seq[b])
            knew, vnew = qkv[b, 1].to(dt), qkv[b, 2].to(dt)
            K = torch.cat([dense_k[b, :, :n], knew[:, None]], 1).float()
            V = torch.cat([dense_v[b, :, :n], vnew[:, None]], 1).float()
            ref = _dense_attention(qkv[b, 0][:, None].float(), K, V, None)[:, 0].reshape(d)
            assert rel_err(out[b].float(), ref) < (1e-5 if dt == torch.float32 else 1e-2), (b, rep)
            page, slot = int(bt[b, n // 64]), n % 64
            got = _pool_swizzle(pool_run[page]) if dt == torch.bfloat16 else pool_run[page]
            assert torch.equal(got[0, :, slot], knew) and torch.equal(got[1, :, slot], vnew)


def test_kv_scatter(ops):
    torch.manual_seed(9)
    B, S, H, Dh = 2, 70, 4, 64
    d = H * Dh
    qkv = torch.randn(B * S, 3 * d, device='cuda').bfloat16()
    bt = torch.tensor([[3, 0], [1, 2]], dtype=torch.int32, device='cuda')
    lens = torch.tensor([70, 65], dtype=torch.int32, device='cuda')
    pool = torch.zeros(4, 2, H, 64, Dh, device='cuda', dtype=torch.bfloat16)
    ops.kv_scatter_paged(qkv, pool, bt, lens, B, S, H, Dh)
    pool = _pool_swizzle(pool)          # back to plain (token, dim) order
    v = qkv.view(B, S, 3, H, Dh)
    for b in range(B):
        for s in range(S):
            page, slot = int(bt[b, s // 64]), s % 64
            if s < int(lens[b]):
                assert torch.equal(pool[page, 0, :, slot], v[b, s, 1]) and torch.equal(pool[page, 1, :, slot], v[b, s, 2])
            else:
                assert (pool[page, :, :, slot] == 0).all()


@pytest.mark.parametrize('top_k,top_p,temp', [(1, 1.0, 1.0), (50, 1.0, 1.0), (10, 0.8, 0.7), (0, 0.5, 1.0), (0, 1.0, 1.3),
                                              (5, 0.3, 1.0), (2000, 0.9, 1.0)])
def test_sample_matches_oracle(ops, top_k, top_p, temp):
    torch.manual_seed(10)
    R, V, ns = 12, 1025, 2
    part = torch.randn(ns, R, V, device='cuda') * 2
    logits = (part[0] + part[1]).cpu()
    tok = torch.empty(R, dtype=torch.int32, device='cuda')
    lp = torch.empty(R, device='cuda')
    for trial in range(4):
        u = torch.rand(R, generator=torch.Generator().manual_seed(trial))
        ops.sample(part, ns, R * V, V, R, V, temperature=temp, top_k=top_k, top_p=top_p, out_tok=tok, out_logprob=lp,
                   uniforms=u.cuda())
        s_ref, lp_ref = vo.topk_sampling(logits, top_k, top_p, temp, uniforms=u)
        filt = vo.top_k_top_p_filter(logits / temp, top_k, top_p)
        probs = torch.softmax(filt, -1)
        t = tok.cpu().long()
        # every draw lands in the oracle's surviving set ...
        assert (probs.gather(1, t[:, None]) > 0).all()
        # ... and equals the oracle's inverse-CDF draw unless u sits within rounding of a CDF edge
        cdf = probs.double().cumsum(-1)
        same = t == s_ref[:, 0]
        edge = (cdf - u.double()[:, None]).abs().min(-1).values < 1e-5
        assert (same | edge).all(), (t, s_ref[:, 0])
        assert rel_err(lp.cpu()[same], lp_ref[same]) < 1e-5 or same.sum() == 0


def test_sample_greedy_ties_and_hash_rng(ops):
    V = 1025
    logits = torch.zeros(3, V, device='cuda')
    logits[0, 7] = 5.0
    logits[1, [3, 900]] = 2.0
    logits[2, :] = -1.0
    tok = torch.empty(3, dtype=torch.int32, device='cuda')
    lp = torch.empty(3, device='cuda')
    ops.sample(logits, 1, 0, V, 3, V, temperature=1.0, top_k=1, top_p=1.0, out_tok=tok, out_logprob=lp)
    assert tok.tolist() == [7, 3, 0]
    assert (lp.cpu() - torch.tensor([0.0, -math.log(2.0), -math.log(V)])).abs().max() < 1e-5
    # in-kernel RNG: deterministic per (seed, step, row), different across steps, roughly uniform
    R = 4096
    lg = torch.zeros(R, 16, device='cuda')
    t1 = torch.empty(R, dtype=torch.int32, device='cuda')
    t2 = torch.empty(R, dtype=torch.int32, device='cuda')
    step = torch.zeros(1, dtype=torch.int32, device='cuda')
    ops.sample(lg, 1, 0, 16, R, 16, temperature=1.0, top_k=0, top_p=1.0, out_tok=t1, seed=1234, step_ptr=step)
    ops.sample(lg, 1, 0, 16, R, 16, temperature=1.0, top_k=0, top_p=1.0, out_tok=t2, seed=1234, step_ptr=step)
    assert torch.equal(t1, t2)
    step.fill_(1)
    ops.sample(lg, 1, 0, 16, R, 16, temperature=1.0, top_k=0, top_p=1.0, out_tok=t2, seed=1234, step_ptr=step)
    assert not torch.equal(t1, t2)
    hist = torch.bincount(t1.long(), minlength=16).float() / R
    assert (hist - 1 / 16).abs().max() < 0.03


def test_ar_bookkeeping(ops):
    B, eos = 4, 1024
    sample = torch.tensor([5, 1024, 7, 9], dtype=torch.int32, device='cuda')
    logprob = torch.tensor([-1.0, -2.0, -3.0, -4.0], device='cuda')
    last = torch.tensor([1, 2, 1024, 3], dtype=torch.int32, device='cuda')
    slp = torch.zeros(B, device='cuda')
    codes = torch.zeros(B, 8, dtype=torch.int32, device='cuda')
    seq = torch.tensor([10, 11, 12, 13], dtype=torch.int32, device='cuda')
    pos = torch.tensor([3, 3, 3, 3], dtype=torch.int32, device='cuda')
    state = torch.tensor([2, -1], dtype=torch.int32, device='cuda')
    ops.ar_bookkeeping(sample, logprob, last, slp, codes, seq, pos, state, eos)
    assert slp.tolist() == [-1.0, -2.0, 0.0, -4.0]
    assert codes[:, 2].tolist() == [5, 1024, 1024, 9] and last.tolist() == [5, 1024, 1024, 9]
    assert seq.tolist() == [11, 12, 13, 14] and pos.tolist() == [4, 4, 4, 4] and state.tolist() == [3, -1]
    sample.fill_(1024)
    ops.ar_bookkeeping(sample, logprob, last, slp, codes, seq, pos, state, eos)
    assert state.tolist() == [4, 3]
    ops.ar_bookkeeping(sample, logprob, last, slp, codes, seq, pos, state, eos)
    assert state.tolist() == [5, 3]          # the first all-EOS step is kept


@pytest.mark.parametrize('x_t,w_t', [(False, True), (True, True), (True, False)])
@pytest.mark.parametrize('M,N,K', [(304, 1024, 512), (1024, 4096, 21616), (1040, 1024, 21616), (3072, 1024, 700), (136, 256, 64)])
def test_linear_transposed_operands(ops, x_t, w_t, M, N, K):
    """vb_linear_t: MN-major MMA operands read straight from the row-major matrices the backward pass has (dgrad: W (N,K) as
    transposed w; wgrad: dy (R,N) and x (R,K) as transposed x and w, contraction over R = 21 616 rows, ragged last k block)."""
    torch.manual_seed(12)
    K8 = (K + 7) // 8 * 8
    X = (torch.randn(M, K, device='cuda') / math.sqrt(K)).bfloat16()
    W = torch.randn(N, K, device='cuda').bfloat16()
    x = X.t().contiguous() if x_t else X
    w = W.t().contiguous() if w_t else W
    if not x_t and K % 8:
        pytest.skip('K-major operand needs K % 8 == 0')
    if not w_t and K % 8:
        pytest.skip('K-major operand needs K % 8 == 0')
    y = ops.linear_t(x, w, x_t=x_t, w_t=w_t, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_err(y, X.double() @ W.double().t()) < 1e-4
