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
                                   (1000, 3072, 1024), (129, 100, 256), (4096, 1024, 4096), (64, 1025, 1024),
                                   # M >= 1024: CTA pairs (256 x 256 tiles): ragged M, ragged N, N with an empty second half
                                   (1300, 1025, 512), (2048, 384, 256), (1024, 3072, 1024)])
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


@pytest.mark.parametrize('M', [1, 5, 16, 32, 33, 64, 128, 129, 192, 193, 200, 256, 300, 512])
@pytest.mark.parametrize('N,K', [(1025, 1024), (3072, 1024), (1024, 4096), (256, 256), (768, 64)])
def test_linear_decode(ops, M, N, K):
    torch.manual_seed(6)
    x = torch.randn(M, K, device='cuda').bfloat16()
    w = (torch.randn(N, K, device='cuda') / math.sqrt(K)).bfloat16()
    ns_max = 32
    part = torch.full((ns_max, M, N), float('nan'), device='cuda')
    ns = ops.linear_decode(x, w, part, M * N, ns_max, flags=(M % 2))     # odd M also exercises the late PDL trigger
    assert ns == ops.linear_decode_splits(N, K, ns_max, M) and 1 <= ns <= ns_max       # above 128 rows: batch tiles, fewer slices
    if M <= 128:
        assert ns == ops.linear_decode_splits(N, K, ns_max)
    else:
        assert ns * -(-N // 128) * -(-M // 128) <= ops.device_info()['sm_count'] or ns == 1          # one wave
    torch.cuda.synchronize()
    y = part[:ns].sum(0)
    assert not torch.isnan(y).any()
    assert rel_err(y, x.double() @ w.double().t()) < 1e-4


@pytest.mark.parametrize('M', [1, 5, 8, 16])
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
    assert ops.linear_decode_rows_splits(K, 0, M) == 1 and ops.linear_decode_rows_splits(K, 0, 16) == 4
    x = torch.randn(M, K, device='cuda').bfloat16()
    w = (torch.randn(N, K, device='cuda') / math.sqrt(K)).bfloat16()
    bias, res = torch.randn(N, device='cuda'), torch.randn(M, N, device='cuda')
    y = res.clone()
    assert ops.linear_decode_rows(x, w, y, bias=bias, residual=True, want_split=0) == 1
    torch.cuda.synchronize()
    assert rel_err(y, x.double() @ w.double().t() + bias.double() + res.double()) < 1e-4


def test_linear_decode_rows_is_deterministic_and_rejects_bad_shapes(ops):
    torch.manual_seed(7)
    x = torch.randn(16, 1024, device='cuda').bfloat16()
    w = (torch.randn(3072, 1024, device='cuda') / 32).bfloat16()
    a, b = torch.empty(16, 3072, device='cuda'), torch.empty(16, 3072, device='cuda')
    ops.linear_decode_rows(x, w, a)
    ops.linear_decode_rows(x, w, b)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError, match='M = 17'):
        ops.linear_decode_rows(torch.zeros(17, 1024, device='cuda').bfloat16(), w, torch.empty(17, 3072, device='cuda'))
    assert ops.linear_decode_rows_splits(1000) == 0
    with pytest.raises(RuntimeError, match='M = 9'):
        ops.linear_decode_rows_ln(torch.zeros(9, 1024, device='cuda'), w, torch.empty(9, 3072, device='cuda'))


def _dg_buffers(ops, M, N, K):
    plan = ops.decode_gemm_plan(M, N, K)
    ws = torch.full((plan['ws_bytes'] // 4,), float('nan'), device='cuda')
    ctr = torch.zeros(256, dtype=torch.int32, device='cuda')
    return plan, ws, ctr


@pytest.mark.parametrize('M', [1, 9, 16, 20, 32, 33, 64, 100, 128, 256])
@pytest.mark.parametrize('N,K', [(3072, 1024), (1024, 1024), (4096, 1024), (1024, 4096), (1025, 1024), (768, 256), (256, 1024)])
def test_decode_gemm_modes(ops, M, N, K):
    """vb_decode_gemm (tcgen05 swap-AB, in-kernel split-K reduction), PLAIN and RESIDUAL epilogues, against float64
    references computed from the same bf16 operands; repeated launches on the same counters (monotonic arrival counters)
    and bit-identical results between launches (deterministic reduction order)."""
    torch.manual_seed(M * 7 + N + K)
    x = torch.randn(M, K, device='cuda').bfloat16()
    w = (torch.randn(N, K, device='cuda') / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device='cuda')
    plan, ws, ctr = _dg_buffers(ops, M, N, K)
    ref = x.double() @ w.double().t()
    y = torch.full((M, N), float('nan'), device='cuda')
    ops.decode_gemm(x, w, ops.DG_PLAIN, ws=ws, counters=ctr, y32=y)
    y2 = torch.full((M, N), float('nan'), device='cuda')
    ops.decode_gemm(x, w, ops.DG_PLAIN, ws=ws, counters=ctr, y32=y2, flags=ops.FLAG_LATE_TRIGGER)
    torch.cuda.synchronize()
    assert rel_err(y, ref) < 2e-5 and torch.equal(y, y2)
    ops.decode_gemm(x, w, ops.DG_PLAIN, ws=ws, counters=ctr, y32=y, bias=bias)
    assert rel_err(y, ref + bias.double()) < 2e-5
    # the two exchange paths (thread-block cluster / DSMEM where it applies, L2 buffer + counters) sum in the same order
    yg = torch.full((M, N), float('nan'), device='cuda')
    ops.decode_gemm(x, w, ops.DG_PLAIN, ws=ws, counters=ctr, y32=yg, bias=bias, flags=ops.FLAG_DG_GLOBAL)
    torch.cuda.synchronize()
    assert torch.equal(y, yg)
    # RESIDUAL: x += acc + bias, bf16 copy, row statistics chunks
    xres0 = torch.randn(M, N, device='cuda')
    xres = xres0.clone()
    xb = torch.zeros(M, N, device='cuda', dtype=torch.bfloat16)
    chunks = plan['tiles']
    stats = torch.full((M, chunks, 2), float('nan'), device='cuda')
    ops.decode_gemm(x, w, ops.DG_RESIDUAL, ws=ws, counters=ctr, bias=bias, xres=xres, y16=xb, stats_out=stats.view(-1))
    torch.cuda.synchronize()
    want = xres0.double() + ref + bias.double()
    assert rel_err(xres, want) < 2e-5
    assert torch.equal(xb, xres.bfloat16())
    assert torch.isfinite(stats).all()
    assert rel_err(stats[:, :, 0].sum(1), xres.double().sum(1)) < 1e-4
    assert rel_err(stats[:, :, 1].sum(1), (xres.double() ** 2).sum(1)) < 1e-4


@pytest.mark.parametrize('M', [1, 12, 32, 48, 256])
@pytest.mark.parametrize('d,N,n_chunks', [(1024, 3072, 1), (1024, 4096, 8), (1024, 3072, 8), (256, 768, 2), (1024, 1025, 16)])
def test_decode_gemm_folded_layernorm(ops, M, d, N, n_chunks):
    """LN / LN_GELU modes: LayerNorm(x; gamma, beta) . W^T (+ bias, erf-GELU) through the algebraic fold (scaled weights,
    column sums, folded bias, (sum, sum of squares) partials of the raw rows) against the plain formula in float64."""
    torch.manual_seed(M + d + N + n_chunks)
    x32 = torch.randn(M, d, device='cuda') * 1.7 + 0.3
    gamma, beta = 1 + 0.2 * torch.randn(d, device='cuda'), 0.1 * torch.randn(d, device='cuda')
    w32 = torch.randn(N, d, device='cuda') / math.sqrt(d)
    bias = torch.randn(N, device='cuda') * 0.1
    ws_w = (w32 * gamma[None]).bfloat16().contiguous()
    colsum = ws_w.float().sum(1).contiguous()
    bfold = ((w32 * beta[None]).sum(1) + bias).contiguous()
    xb = x32.bfloat16()
    # statistics of the fp32 rows split into n_chunks column chunks, as a producing RESIDUAL launch writes them
    xc = x32.view(M, n_chunks, d // n_chunks)
    stats = torch.stack([xc.sum(-1), (xc * xc).sum(-1)], -1).contiguous()
    plan, ws, ctr = _dg_buffers(ops, M, N, d)
    ln = torch.nn.functional.layer_norm(x32.double(), (d,), gamma.double(), beta.double(), 1e-5)
    ref = ln @ w32.double().t() + bias.double()
    y = torch.full((M, N), float('nan'), device='cuda')
    ops.decode_gemm(xb, ws_w, ops.DG_LN, ws=ws, counters=ctr, bias=bfold, colsum=colsum, stats_in=stats.view(-1),
                    n_chunks_in=n_chunks, eps=1e-5, y32=y)
    torch.cuda.synchronize()
    assert rel_err(y, ref) < 1e-2
    f = torch.zeros(M, N, device='cuda', dtype=torch.bfloat16)
    ops.decode_gemm(xb, ws_w, ops.DG_LN_GELU, ws=ws, counters=ctr, bias=bfold, colsum=colsum, stats_in=stats.view(-1),
                    n_chunks_in=n_chunks, eps=1e-5, y16=f)
    torch.cuda.synchronize()
    assert rel_err(f.float(), torch.nn.functional.gelu(ref)) < 1e-2
    fg = torch.zeros(M, N, device='cuda', dtype=torch.bfloat16)
    ops.decode_gemm(xb, ws_w, ops.DG_LN_GELU, ws=ws, counters=ctr, bias=bfold, colsum=colsum, stats_in=stats.view(-1),
                    n_chunks_in=n_chunks, eps=1e-5, y16=fg, flags=ops.FLAG_DG_GLOBAL)
    torch.cuda.synchronize()
    assert torch.equal(f, fg)


def test_decode_gemm_layer_chain_matches_reference(ops):
    """out-proj -> FFN1 -> FFN2 -> QKV of one decoder layer on the fused launches (modules.py:271-278), full-size shapes,
    B = 32, inside one PDL chain, against the float64 formulas; also replayed from a CUDA graph (counters keep counting)."""
    torch.manual_seed(5)
    B, d, F = 32, 1024, 4096
    mk = lambda n, k: (torch.randn(n, k, device='cuda') / math.sqrt(k))
    wo32, w1_32, w2_32, wq32 = mk(d, d), mk(F, d), mk(d, F), mk(3 * d, d)
    bo, b1, b2 = (torch.randn(n, device='cuda') * 0.1 for n in (d, F, d))
    g1, be1, g2, be2 = (1 + 0.1 * torch.randn(d, device='cuda'), 0.1 * torch.randn(d, device='cuda'),
                        1 + 0.1 * torch.randn(d, device='cuda'), 0.1 * torch.randn(d, device='cuda'))
    o = torch.randn(B, d, device='cuda').bfloat16()
    x0 = torch.randn(B, d, device='cuda')
    wo, w2 = wo32.bfloat16(), w2_32.bfloat16()
    w1s, wqs = (w1_32 * g2[None]).bfloat16().contiguous(), (wq32 * g1[None]).bfloat16().contiguous()
    c1, cq = w1s.float().sum(1).contiguous(), wqs.float().sum(1).contiguous()
    bf1, bfq = ((w1_32 * be2[None]).sum(1) + b1).contiguous(), (wq32 * be1[None]).sum(1).contiguous()
    plans = {k: ops.decode_gemm_plan(B, n, kk) for k, (n, kk) in {'o': (d, d), 'f1': (F, d), 'f2': (d, F), 'qkv': (3 * d, d)}.items()}
    ws = torch.zeros(max(p['ws_bytes'] for p in plans.values()) // 4, device='cuda')
    ctr = torch.zeros(4, 256, dtype=torch.int32, device='cuda')
    ch_o, ch_f2 = plans['o']['tiles'], plans['f2']['tiles']
    stats = torch.zeros(B * max(ch_o, ch_f2) * 2, device='cuda')
    x, xb = x0.clone(), torch.zeros(B, d, device='cuda', dtype=torch.bfloat16)
    f = torch.zeros(B, F, device='cuda', dtype=torch.bfloat16)
    qkv = torch.zeros(B, 3 * d, device='cuda')

    def chain():
        ops.decode_gemm(o, wo, ops.DG_RESIDUAL, ws=ws, counters=ctr[0], bias=bo, xres=x, y16=xb, stats_out=stats)
        ops.decode_gemm(xb, w1s, ops.DG_LN_GELU, ws=ws, counters=ctr[1], bias=bf1, colsum=c1, stats_in=stats, n_chunks_in=ch_o, y16=f)
        ops.decode_gemm(f, w2, ops.DG_RESIDUAL, ws=ws, counters=ctr[2], bias=b2, xres=x, y16=xb, stats_out=stats)
        ops.decode_gemm(xb, wqs, ops.DG_LN, ws=ws, counters=ctr[3], bias=bfq, colsum=cq, stats_in=stats, n_chunks_in=ch_f2, y32=qkv)

    # float64 reference of the same chain (bf16 roundings only where the kernels round: o, weights, f)
    D = torch.double
    x1 = x0.to(D) + o.to(D) @ wo.to(D).t() + bo.to(D)
    h2 = torch.nn.functional.layer_norm(x1, (d,), g2.to(D), be2.to(D), 1e-5)
    f_ref = torch.nn.functional.gelu(h2 @ w1_32.to(D).t() + b1.to(D))
    x2 = x1 + f_ref.bfloat16().to(D) @ w2.to(D).t() + b2.to(D)
    q_ref = torch.nn.functional.layer_norm(x2, (d,), g1.to(D), be1.to(D), 1e-5) @ wq32.to(D).t()
    chain()
    torch.cuda.synchronize()
    assert rel_err(f.float(), f_ref) < 1e-2 and rel_err(x, x2) < 1e-2 and rel_err(qkv, q_ref) < 1e-2
    first = (x.clone(), qkv.clone(), f.clone())
    g = torch.cuda.CUDAGraph()
    x.copy_(x0)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        chain()
    for _ in range(3):
        x.copy_(x0)
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(x, first[0]) and torch.equal(qkv, first[1]) and torch.equal(f, first[2])


@pytest.mark.parametrize('top_k,top_p,temp', [(1, 1.0, 1.0), (50, 1.0, 1.0), (10, 0.8, 0.7)])
def test_ar_step_tail_matches_separate_kernels(ops, top_k, top_p, temp):
    """vb_ar_step_tail == vb_sample + vb_ar_bookkeeping + vb_embed_sum_pe (bit for bit), incl. finished rows and the stop
    detection, over several steps; the row statistics are those of the embedded row."""
    torch.manual_seed(3)
    B, V, d, eos, max_new = 6, 1025, 256, 1024, 5
    table = torch.randn(V + 1, d, device='cuda')
    pe = torch.randn(64, d, device='cuda')
    mk_state = lambda: dict(last=torch.tensor([5, eos, 7, 8, 9, 10], dtype=torch.int32, device='cuda'),
                            slp=torch.zeros(B, device='cuda'), codes=torch.zeros(B, max_new, dtype=torch.int32, device='cuda'),
                            seq=torch.full((B,), 11, dtype=torch.int32, device='cuda'),
                            pos=torch.arange(B, dtype=torch.int32, device='cuda'))
    a, b = mk_state(), mk_state()
    st_a = torch.tensor([0, -1, 0, 0], dtype=torch.int32, device='cuda')
    st_b = torch.tensor([0, -1], dtype=torch.int32, device='cuda')
    seed = torch.tensor([1234], dtype=torch.int64, device='cuda')
    x, xb, stats = torch.zeros(B, d, device='cuda'), torch.zeros(B, d, device='cuda', dtype=torch.bfloat16), torch.zeros(2 * B, device='cuda')
    x_ref = torch.zeros(B, d, device='cuda')
    tok, lp = torch.zeros(B, dtype=torch.int32, device='cuda'), torch.zeros(B, device='cuda')
    for step in range(max_new):
        logits = torch.randn(B, V, device='cuda') * 3
        if step == 3:
            logits[:, eos] += 100.0                              # every running row draws EOS -> stop_step = 3
        ops.ar_step_tail(logits, V, temperature=temp, top_k=top_k, top_p=top_p, uniforms=None, seed=seed, row_offset=4,
                         last=a['last'], sum_logprobs=a['slp'], codes_out=a['codes'], seq_lens=a['seq'], audio_pos=a['pos'],
                         state=st_a, eos=eos, table=table, pe=pe, x=x, xb=xb, stats=stats)
        ops.sample(logits, 1, 0, V, B, V, temperature=temp, top_k=top_k, top_p=top_p, out_tok=tok, out_logprob=lp,
                   seed=1234, step_ptr=st_b, row_offset=4)
        ops.ar_bookkeeping(tok, lp, b['last'], b['slp'], b['codes'], b['seq'], b['pos'], st_b, eos)
        ops.embed_sum_pe(b['last'].view(B, 1, 1), table[None].contiguous(), pe, x_ref, pos_b=b['pos'])
        torch.cuda.synchronize()
        for k in a:
            assert torch.equal(a[k], b[k]), (k, step)
        assert st_a.tolist()[:2] == st_b.tolist() and st_a.tolist()[2:] == [0, 0]
        assert torch.equal(x, x_ref) and torch.equal(xb, x_ref.bfloat16())
        s2 = stats.view(B, 2)
        assert rel_err(s2[:, 0], x_ref.double().sum(1)) < 1e-5 and rel_err(s2[:, 1], (x_ref.double() ** 2).sum(1)) < 1e-5
    assert st_a.tolist()[1] == 3


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


@pytest.mark.parametrize('B,S,H', [(2, 300, 4), (3, 128, 2), (2, 77, 16), (2, 900, 3), (1, 1, 1), (4, 129, 2), (2, 1125, 2), (1, 64, 1),
                                   (1, 65, 1), (1, 192, 1)])
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
        assert rel_err(got[b, :n], ref[b, :n]) < 1e-2, (b, rel_err(got[b, :n], ref[b, :n]))


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


@pytest.mark.parametrize('M,N,K', [(1024, 1024, 256), (1300, 1024, 1024), (2500, 1000, 512), (1025, 300, 64), (4096, 1024, 1024)])
def test_linear_argmax_equals_logits_then_greedy_pick(ops, M, N, K):
    """vb_linear_argmax (logits GEMM + greedy pick in one kernel, valle_nar.py:157-160 with argmax sampling) against the
    unfused pair it replaces -- vb_linear (fp32 logits) followed by vb_sample(top_k=1) -- and against torch.argmax over the
    same logits: integer work, bit-exact.  Ragged M / N (partial last tiles), dense and strided destinations; the key
    scratch is left zero."""
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device='cuda').bfloat16()
    w = (torch.randn(N, K, device='cuda') / math.sqrt(K)).bfloat16()
    logits = ops.linear(x, w, out_dtype=torch.float32)
    want = torch.empty(M, device='cuda', dtype=torch.int32)
    ops.sample(logits, 1, 0, N, M, N, temperature=1.0, top_k=1, top_p=1.0, out_tok=want)
    assert torch.equal(want.long(), logits.argmax(-1))
    keys = torch.zeros(M, device='cuda', dtype=torch.int64)
    got = torch.full((M,), -1, device='cuda', dtype=torch.int32)
    ops.linear_argmax(x, w, keys, got)
    assert torch.equal(got, want)
    assert int(keys.abs().max()) == 0
    # second call on the same scratch, strided destination: column 3 of a (B, T0 + T, 8) code tensor, rows from T0 on
    Bb = 4 if M % 4 == 0 else 1
    T, T0, Q = M // Bb, 5, 8
    ids = torch.full((Bb, T0 + T, Q), -7, device='cuda', dtype=torch.int32)
    ops.linear_argmax(x, w, keys, ids[:, T0:, 3], rows_per_batch=T, batch_stride=(T0 + T) * Q, row_stride=Q)
    assert torch.equal(ids[:, T0:, 3].reshape(-1), want)
    ids[:, T0:, 3] = -7
    assert int((ids != -7).sum()) == 0          # nothing else was touched


def test_linear_argmax_ties_and_degenerate_rows(ops):
    """Ties go to the lowest column (vb_sample's greedy rule, torch.argmax's too): duplicated weight rows give bit-identical
    logits in several columns, also across the 256-column tiles that meet in the atomic.  All-zero activations make every
    logit of a row equal (0.0): token 0."""
    torch.manual_seed(5)
    M, N, K = 1100, 1024, 128
    x = torch.randn(M, K, device='cuda').bfloat16()
    w = (torch.randn(N, K, device='cuda') * 0.1).bfloat16()
    big = (torch.randn(K, device='cuda')).bfloat16()
    x[:, :] = x * 0.01
    x[::3] = big * 0.5                         # rows whose best column is one of the duplicates below
    for c in (7, 300, 301, 700, 1023):         # the same weight row in five columns spread over the four tiles
        w[c] = big
    x[10] = 0
    logits = ops.linear(x, w, out_dtype=torch.float32)
    assert torch.equal(logits[0, 7], logits[0, 1023])
    keys = torch.zeros(M, device='cuda', dtype=torch.int64)
    got = torch.empty(M, device='cuda', dtype=torch.int32)
    ops.linear_argmax(x, w, keys, got)
    want = torch.empty(M, device='cuda', dtype=torch.int32)
    ops.sample(logits, 1, 0, N, M, N, temperature=1.0, top_k=1, top_p=1.0, out_tok=want)
    assert torch.equal(got, want)
    assert int(got[0]) == 7 and int(got[10]) == 0
    from valle2_b200 import _lib
    with pytest.raises(_lib.VBError):          # below the CTA-pair GEMM's shapes: refused, never a silent other path
        ops.linear_argmax(x[:512], w, keys, got)


@pytest.mark.parametrize('ydt', [torch.bfloat16, torch.float32])
@pytest.mark.parametrize('d', [256, 512, 1024])
@pytest.mark.parametrize('affine', [True, False])
def test_embed_sum_pe_norm_equals_embed_then_layernorm(ops, ydt, d, affine):
    """vb_embed_sum_pe_norm (embedding sum + PE + the first (Ada)LayerNorm in one kernel; valle_nar.py:140-152 into
    modules.py:271) against the two kernels it replaces: same residual rows and same normalised rows, bit for bit (same order
    of operations), for both segments of a stage's input (text rows; prompt rows with all codebooks + target rows with the
    first n) written at their row offsets of one (B, S, d) buffer.  (More than 1024 rows: vb_residual_layernorm then runs its
    warp-per-row kernel, whose lane layout the fused kernel shares; at decode-shape row counts it runs a CTA-per-row kernel that
    sums in another order, equal only to fp32 rounding -- second half of the test.)"""
    torch.manual_seed(d)
    B, Tx, Tc, T, Q, V = 3, 5, 7, 400, 8, 50
    S = Tx + Tc + T
    tok = torch.randint(0, 256, (B, Tx, 1), dtype=torch.int32, device='cuda')
    ids = torch.randint(0, V, (B, Tc + T, Q), dtype=torch.int32, device='cuda')
    tok_table = torch.randn(1, 256, d, device='cuda')
    tables = torch.randn(Q, V, d, device='cuda')
    pe = torch.randn(512, d, device='cuda')
    gamma = torch.randn(d, device='cuda') if affine else None
    beta = torch.randn(d, device='cuda') if affine else None
    x0 = torch.full((B * S, d), 7.0, device='cuda')
    ops.embed_sum_pe(tok, tok_table, pe, x0, out_rows_per_batch=S, out_row_offset=0)
    ops.embed_sum_pe(ids, tables, pe, x0, t_split=Tc, nq_a=Q, nq_b=3, out_rows_per_batch=S, out_row_offset=Tx)
    y0 = torch.empty(B * S, d, device='cuda', dtype=ydt)
    ops.residual_layernorm(x0, gamma, beta, y0, eps=1e-5)
    x1 = torch.full((B * S, d), -3.0, device='cuda')
    y1 = torch.zeros(B * S, d, device='cuda', dtype=ydt)
    ops.embed_sum_pe(tok, tok_table, pe, x1, out_rows_per_batch=S, out_row_offset=0, norm_y=y1, gamma=gamma, beta=beta, eps=1e-5)
    ops.embed_sum_pe(ids, tables, pe, x1, t_split=Tc, nq_a=Q, nq_b=3, out_rows_per_batch=S, out_row_offset=Tx, norm_y=y1,
                     gamma=gamma, beta=beta, eps=1e-5)
    assert torch.equal(x0, x1)
    assert torch.equal(y0, y1)
    # and against torch on the same rows
    ref = torch.nn.functional.layer_norm(x0, (d,), gamma, beta, 1e-5) if affine else x0
    assert rel_err(y1.float(), ref) < (1e-2 if ydt == torch.bfloat16 else 1e-5)
    # a decode-shape row count (CTA-per-row LayerNorm on the unfused side): same rows, normalised rows equal to rounding
    Bs, Ts = 2, 9
    xs0, xs1 = torch.empty(Bs * Ts, d, device='cuda'), torch.empty(Bs * Ts, d, device='cuda')
    ys0, ys1 = torch.empty(Bs * Ts, d, device='cuda', dtype=ydt), torch.empty(Bs * Ts, d, device='cuda', dtype=ydt)
    ops.embed_sum_pe(ids[:Bs, :Ts].contiguous(), tables, pe, xs0, nq_b=5)
    ops.residual_layernorm(xs0, gamma, beta, ys0, eps=1e-5)
    ops.embed_sum_pe(ids[:Bs, :Ts].contiguous(), tables, pe, xs1, nq_b=5, norm_y=ys1, gamma=gamma, beta=beta, eps=1e-5)
    assert torch.equal(xs0, xs1)
    assert rel_err(ys1.float(), ys0.float()) < (8e-3 if ydt == torch.bfloat16 else 2e-6)


def test_linear_categorical_draws_from_the_softmax(ops):
    """vb_linear_categorical (valle_nar.py:160, `Categorical(logits / temperature).sample()`, fused into the logits GEMM by the
    Gumbel-max trick): 65 536 rows with IDENTICAL activations are 65 536 independent draws from one known distribution.
    Empirical frequencies against softmax(logits / T): the 20 likeliest classes within 5 sigma, total variation small,
    chi-square per degree of freedom near 1 (would expose correlated noise across rows or columns).  Also: a call is a pure
    function of (seed, step); another step or seed gives other draws; T -> 0 degenerates to the arg-max."""
    torch.manual_seed(3)
    M, N, K, T = 65536, 1024, 64, 0.8
    x1 = torch.randn(1, K, device='cuda').bfloat16()
    x = x1.expand(M, K).contiguous()
    w = (torch.randn(N, K, device='cuda') * 0.3).bfloat16()
    logits = ops.linear(x[:1024], w, out_dtype=torch.float32)[0].double()
    p = torch.softmax(logits / T, -1)
    keys = torch.zeros(M, device='cuda', dtype=torch.int64)

    def draw(seed, step, temperature=T):
        out = torch.full((M,), -1, device='cuda', dtype=torch.int32)
        ops.linear_argmax(x, w, keys, out, temperature=temperature, seed=seed, step=step)
        assert int(keys.abs().max()) == 0
        assert int(out.min()) >= 0 and int(out.max()) < N
        return out.long()

    a = draw(11, 3)
    f = torch.bincount(a, minlength=N).double() / M
    top = p.topk(20).indices
    sigma = (p[top] * (1 - p[top]) / M).sqrt()
    assert ((f[top] - p[top]).abs() < 5 * sigma + 1e-4).all(), (f[top], p[top])
    assert 0.5 * (f - p).abs().sum().item() < 0.06
    big = p * M >= 20
    chi2 = (((f[big] - p[big]) * M) ** 2 / (p[big] * M)).sum().item() / int(big.sum())
    assert 0.6 < chi2 < 1.4, chi2
    # rows are independent of each other: neighbouring rows agree about as often as two independent draws would
    agree = (a[1:] == a[:-1]).double().mean().item()
    assert abs(agree - (p * p).sum().item()) < 5 * math.sqrt((p * p).sum().item() / M) + 1e-3
    assert torch.equal(a, draw(11, 3))                      # pure function of its arguments
    assert (a != draw(11, 4)).double().mean().item() > 0.5  # another stage: other draws
    assert (a != draw(12, 3)).double().mean().item() > 0.5  # another seed: other draws
    cold = draw(11, 3, temperature=1e-4)                    # T -> 0: the arg-max (noise / logit-gap ~ 1e-3)
    assert (cold == logits.argmax().item()).double().mean().item() > 0.999
