"""Parity AT THE BENCHMARKED OPERATING POINT (BASELINE configs[1] / configs[2] shapes), not only at CPU-sized shapes:

  (i)   paged decode attention at B=32, H=16, ctx 750 and 1126 (12-18 pages, 512 CTAs), permuted page table, one split and
        several (thread-block-cluster merge and global ticket merge), against a float64 dense softmax -- 1e-2;
  (ii)  the full-size AR decoder (12 layers, d=1024, 16 heads), text 150 + prompt 226, batch 2 and 32: 64 CUDA-graph
        replayed decode steps starting at context 700, every step's logits against the CPU oracle's teacher-forced logits
        over the same tokens (valle_ar.py:141-171 vs :43-90, the cached == uncached invariant) -- 1e-2 relative, and the
        greedy tokens wherever the oracle's top-1/top-2 margin is outside that tolerance;
  (iii) the full-size NAR decoder at S = 150 + 225 + 525 = 900, batch 2: stage logits against the oracle driven with the
        same codes (valle_nar.py:107-165 repaired) -- bf16 1e-2, fp32 validation mode 2e-5 + token-exact.

The oracle side takes ~1-2 minutes of host time in total (B=32 x 764 positions x 12 layers in fp32 on the CPU).
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import valle2_b200  # noqa: E402
from oracle import synth  # noqa: E402
from oracle import valle_oracle as vo  # noqa: E402
from test_gpu_models import build, oracle_nar_stage_logits, rel_err  # noqa: E402
from test_gpu_ops import _dense_attention, _pool_swizzle  # noqa: E402

TOL = 1e-2          # north star: logits within 1e-2 relative in bf16 compute / fp32 accumulate


@pytest.fixture(autouse=True)
def _restore_precision():
    prev = valle2_b200.get_precision()
    yield
    valle2_b200.set_precision(prev)


@pytest.fixture(scope='module')
def ops():
    from valle2_b200 import ops as _ops
    return _ops


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('ctx', [750, 1126])
@pytest.mark.parametrize('n_tsplit,flags', [(1, 0), (1, 2), (2, 0), (2, 8), (4, 2), (4, 8 | 2)])
def test_attn_decode_paged_bench_shape(ops, ctx, n_tsplit, flags):
    """B=32, H=16 -- the launch shape of the bench (grid (n_tsplit, 16, 32)); ragged contexts around ctx, pages scattered
    over the pool by a random permutation.  flags: 2 = KV streamed before the dependency wait, 8 = ticket merge."""
    torch.manual_seed(21)
    B, H, Dh = 32, 16, 64
    d = H * Dh
    max_pages = (ctx + 64) // 64 + 1
    seq = torch.randint(ctx - 70, ctx + 1, (B,), dtype=torch.int32)
    seq[0], seq[1] = ctx, ctx - 64 * (ctx // 64)           # one full-length row, one that ends inside its first page
    seq = seq.cuda()
    n_pages = B * max_pages
    bt = torch.randperm(n_pages, device='cuda').to(torch.int32).view(B, max_pages).contiguous()
    dense = torch.randn(2, B, H, max_pages * 64, Dh, device='cuda').bfloat16()
    pool = torch.zeros(n_pages, 2, H, 64, Dh, device='cuda', dtype=torch.bfloat16)
    idx = bt.long().view(-1)
    paged = dense.view(2, B, H, max_pages, 64, Dh).permute(1, 3, 0, 2, 4, 5).reshape(n_pages, 2, H, 64, Dh)
    pool[idx] = paged
    pool = _pool_swizzle(pool)
    n_part = 6                                              # the QKV GEMM's split-K slices at d=1024
    part = torch.randn(n_part, B, 3 * d, device='cuda') * 0.5
    out = torch.empty(B, d, device='cuda', dtype=torch.bfloat16)
    ws = torch.zeros(ops.attn_decode_ws_bytes(B, H, n_tsplit) // 4 + 16, dtype=torch.int32, device='cuda')
    qkv = part[0].clone()
    for s in range(1, n_part):
        qkv += part[s]
    qkv = qkv.view(B, 3, H, Dh)
    for rep in range(2):                                    # second launch: the ticket counters reset themselves
        pool_run = pool.clone()
        ops.attn_decode_paged(part, n_part, B * 3 * d, pool_run, bt, seq, out, B, H, Dh, n_tsplit, ws, flags)
        torch.cuda.synchronize()
        worst = 0.0
        for b in range(B):
            n = int(seq[b])
            knew, vnew = qkv[b, 1].bfloat16(), qkv[b, 2].bfloat16()
            K = torch.cat([dense[0, b, :, :n], knew[:, None]], 1).float()
            V = torch.cat([dense[1, b, :, :n], vnew[:, None]], 1).float()
            ref = _dense_attention(qkv[b, 0][:, None].float(), K, V, None)[:, 0].reshape(d)
            worst = max(worst, rel_err(out[b].float(), ref))
            page, slot = int(bt[b, n // 64]), n % 64
            got = _pool_swizzle(pool_run[page])
            assert torch.equal(got[0, :, slot], knew) and torch.equal(got[1, :, slot], vnew), (b, rep)
        assert worst < TOL, (worst, rep)


# ---------------------------------------------------------------------------------------------------------------------
def _oracle_tf_logits(sd, oc, tok, codes_full, rows_per_chunk=8):
    """Teacher-forced oracle logits (B, Ty, V) in row chunks (bounds the (B,H,S,S) attention temporaries on the host)."""
    B, Ty = codes_full.shape
    out = []
    for b0 in range(0, B, rows_per_chunk):
        sl = slice(b0, min(B, b0 + rows_per_chunk))
        n = sl.stop - sl.start
        lg, _ = vo.ar_teacher_forced(sd, oc, tok[sl], codes_full[sl], torch.full((n,), tok.shape[1]), torch.full((n,), Ty))
        out.append(lg)
    return torch.cat(out, 0)


@pytest.mark.parametrize('B', [2, 32])
def test_ar_decode_at_bench_context_vs_oracle(tmp_path, B):
    valle2_b200.set_precision('bf16')
    oc = synth.large_config('LayerNorm', max_audio_len=512)
    model, sd = build('ValleAR', oc, tmp_path, 5)
    Tx, P, ctx0, checked = 150, 226, 700, 64
    warm = ctx0 - (Tx + P)                                  # decode steps before the first checked one (context 376 -> 700)
    g = torch.Generator().manual_seed(31)
    tok = torch.randint(0, 256, (B, Tx), generator=g)
    cod = torch.cat([torch.full((B, 1), oc.bos_token), torch.randint(0, 1024, (B, P - 1), generator=g)], 1)
    eng = model._engine()
    samp = {'temperature': 1.0, 'top_k': 1, 'top_p': 1.0, 'seed': 0}
    st = eng.prefill(tok.cuda(), cod.cuda(), max_new=warm + checked + 2)
    eng.first_token(samp, None, -1)
    eng.decode_step(samp, None, -1)                         # eager launch (module load, function attributes)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        eng.decode_step(samp, None, -1)
    for _ in range(warm - 1):
        graph.replay()
    assert int(st['seq_lens'].min().item()) == ctx0
    got = []
    for _ in range(checked):
        graph.replay()
        got.append(eng.step_logits())
    torch.cuda.synchronize()
    n_gen = 1 + warm + checked
    gen = st['codes_out'][:, :n_gen].long().cpu()
    assert int(st['state'][0, 0].item()) == n_gen
    codes_full = torch.cat([cod, gen[:, :n_gen - 1]], 1)    # teacher forcing over the tokens the GPU actually drew
    ref = _oracle_tf_logits(sd, oc, tok, codes_full)        # (B, P + n_gen - 1, V+1)
    worst, n_clear, n_tok = 0.0, 0, 0
    for j in range(checked):
        k = warm + 1 + j                                    # index of the generated token these logits select
        r = ref[:, P - 1 + k]
        e = rel_err(got[j].cpu(), r)
        worst = max(worst, e)
        assert e < TOL, (j, e)
        top2 = r.topk(2, dim=-1).values
        clear = (top2[:, 0] - top2[:, 1]) > 2 * TOL * r.abs().max()
        n_clear += int(clear.sum())
        n_tok += B
        assert (gen[clear, k] == r.argmax(-1)[clear]).all(), j
    assert n_clear > 0.5 * n_tok, (n_clear, n_tok)          # the token check is not vacuous
    print(f'B={B}: worst relative logit error over {checked} steps at ctx {ctx0}..{ctx0 + checked}: {worst:.2e}; '
          f'{n_clear}/{n_tok} greedy tokens with a clear oracle margin, all equal')


# ---------------------------------------------------------------------------------------------------------------------
def test_nar_full_size_s900_vs_oracle(tmp_path):
    oc = synth.large_config('AdaptiveLayerNorm')
    model, sd = build('ValleNAR', oc, tmp_path, 6)
    g = torch.Generator().manual_seed(41)
    B, Tp, Tt, Tc, T = 2, 50, 100, 225, 525                 # S = 150 + 225 + 525 = 900 (BASELINE configs[2])
    pt, tt = torch.randint(0, 256, (B, Tp), generator=g), torch.randint(0, 256, (B, Tt), generator=g)
    pc, fl = torch.randint(0, 1024, (B, Tc, 8), generator=g), torch.randint(0, 1024, (B, T), generator=g)
    # bf16 (the product): every compared stage's logits against the oracle fed with the GPU's own earlier codebooks
    valle2_b200.set_precision('bf16')
    eng = model._engine()
    codes, trace = eng.generate(pt.cuda(), pc.cuda(), tt.cuda(), fl.cuda(), greedy=True, return_logits=True)
    codes = codes.cpu()
    assert codes.shape == (B, T, 8) and torch.equal(codes[:, :, 0], fl)
    for b, stages in ((0, (1, 4, 7)), (1, (2,))):
        for n in stages:
            ref = oracle_nar_stage_logits(sd, oc, torch.cat([pt[b], tt[b]]), pc[b], codes[b], n)
            e = rel_err(trace[n - 1][b].cpu(), ref)
            assert e < TOL, (b, n, e)
            top2 = ref.topk(2, dim=-1).values
            clear = (top2[:, 0] - top2[:, 1]) > 2 * TOL * ref.abs().max()
            assert clear.sum() > 0.5 * T and (codes[b, clear, n] == ref.argmax(-1)[clear]).all(), (b, n)
    # fp32 validation mode: utterance 0, all seven stages token-exact and 2e-5 on the logits
    valle2_b200.set_precision('fp32')
    eng = model._engine()
    codes32, trace32 = eng.generate(pt[:1].cuda(), pc[:1].cuda(), tt[:1].cuda(), fl[:1].cuda(), greedy=True, return_logits=True)
    ref_codes, ref_trace = vo.nar_generate(sd, oc, pt[0], pc[0], tt[0], fl[0], return_trace=True)
    assert torch.equal(codes32[0].cpu(), ref_codes)
    for n in range(7):
        assert rel_err(trace32[n][0].cpu(), ref_trace[n]) < 2e-5, n
