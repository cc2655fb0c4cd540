#!/usr/bin/env python
"""Round-1 form of the benchmark, kept as a measurement tool: AR decode only (BASELINE configs[1]) at any --batch, used by
tools/ab_r02.sh for A/B runs of the decode step under the VALLE_B200_* switches.  The driver-facing bench is /bench.py.

Benchmark of the Valle2 hot path on B200 (contract: see the task statement / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--no-extras]

Workload (BASELINE.json configs[1]): default VALL-E AR decoder (12 layers, d=1024, 16 heads, F=4096), KV-cached greedy
decode, batch 32 per GPU, text 150 phonemes, 3 s prompt (225 frames + BOS), 750 generated frames; bf16 weights / KV
with fp32 accumulation.  One "step" = one decode step of the whole batch (B new codec tokens).  When K < 750 the
prompt is lengthened so that the mean context over the K timed steps equals the config's mean (750.5 positions).

Prints ONE JSON line on rank 0.  `value` = tokens/s with everything resident in HBM (CUDA-graph replay of the step);
`e2e` = the same metric through ValleAR.generate_batch from pinned HOST tensors (H2D of the prompt, prefill, K+W
decode steps, D2H of the codes all inside the timed region).  `roofline` is measured live with CUDA events on the
dominant kernel (paged decode attention).  `cpu_baseline` times the oracle port on the host cores (bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TX, P0, N_NEW = 150, 226, 750           # phonemes, BOS + 225 prompt frames, generated frames
MEAN_CTX = TX + P0 + (N_NEW - 1) / 2.0  # 750.5


def peaks() -> dict:
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        p['_source'] = 'measured'
        return p
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, '_source': 'fallback'}


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ''
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}


def large_cfg(norm: str, tmp: str, **kw):
    from valle2_b200.config import ConfigValle
    base = dict(num_layers=12, d_model=1024, n_heads=16, dim_feedforward=4096, norm=norm, dropout=0.0,
                max_audio_len=N_NEW, top_k=1, num_beams=1, ckpt_path=os.path.join(tmp, 'c'), log_path=os.path.join(tmp, 'l'))
    base.update(kw)
    return ConfigValle(**base)


def decode_form(eng, sub) -> str:
    if eng._lean_ok(sub):
        return 'lean: linear_decode_rows_kernel (mma.sync, full K per CTA, LayerNorm on load, fused epilogues), 5 launches per layer'
    if eng._tc_ok(sub):
        return ('tc: decode_gemm_kernel (tcgen05 swap-AB, split-K reduced inside the launch, LayerNorm folded, fused epilogues), '
                '5 launches per layer')
    return 'splitk: gemm_tc_kernel<swap-AB split-K> slices + LayerNorm / GELU-reduce kernels, 8 launches per layer'


def ar_step_bytes(B: int, ctx: float, L=12, d=1024, F=4096, V=1025) -> float:
    """Algorithmic HBM bytes of one decode step: all weights once + K/V of every cached position."""
    w = 2 * (L * (3 * d * d + d * d + 2 * d * F) + V * d)
    kv_tok = 2 * L * d * 2
    return w + B * (ctx + 1) * kv_tok


# ------------------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, local_rank: int):
    import valle2_b200
    from valle2_b200 import ops
    from valle2_b200.models import ValleAR

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    valle2_b200.set_precision('bf16')
    tmp = f'/tmp/valle_bench_{os.getpid()}'
    B, K, W = args.batch, args.steps, args.warmup
    torch.manual_seed(0)
    model = ValleAR(large_cfg('LayerNorm', tmp)).eval().to(dev)
    eng = model._engine()
    g = torch.Generator().manual_seed(100 + rank)
    total_steps = K + W
    # lengthen the prompt when fewer than 750 steps are timed so the mean context matches the config
    extra = max(0, int(round(MEAN_CTX - (K - 1) / 2.0 - W)) - (TX + P0))
    P = P0 + extra
    tokens_h = torch.randint(0, 256, (B, TX), generator=g).pin_memory()
    codes_h = torch.cat([torch.full((B, 1), 1025), torch.randint(0, 1024, (B, P - 1), generator=g)], 1).pin_memory()
    samp = {'temperature': 1.0, 'top_k': 1, 'top_p': 1.0, 'seed': 0}

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident measurement: prefill (untimed), W warm-up steps, K timed graph replays ----
    st = eng.prefill(tokens_h.to(dev), codes_h.to(dev), max_new=total_steps + 2)
    eng.first_token(samp, None, -1)
    eng.decode_step(samp, None, -1)                        # eager warm-up launch (module load, func attributes)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        eng.decode_step(samp, None, -1)
    for _ in range(max(W - 1, 0)):
        graph.replay()
    ctx0 = int(st['seq_lens'][0].item())
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        graph.replay()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
        # the only collective of the path: gather the generated codes once at the end (NCCL over NVLink)
        gathered = [torch.empty_like(st['codes_out']) for _ in range(world)]
        torch.distributed.all_gather(gathered, st['codes_out'])
    mean_ctx = ctx0 + (K - 1) / 2.0
    tok_s = B * world * K / (ms * 1e-3)
    step_bytes = ar_step_bytes(B, mean_ctx)
    pk = peaks()
    launches_per_step = eng.launches_per_step()
    n_sub = len(st['subs'])

    result = {
        'metric': 'ar_decode_tokens_per_s', 'value': tok_s, 'unit': 'tokens/s', 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
        'data': 'synthetic',
        'config': {'workload': f'AR decode, default VALL-E AR decoder 12L d1024 h16 F4096, KV-cached greedy, batch {B}/GPU '
                               f'(BASELINE configs[1]); text {TX}, prompt {P}, ctx {ctx0}->{ctx0 + K} (mean {mean_ctx:.1f})',
                   'batch_per_gpu': B, 'global_batch': B * world, 'parallelism': f'dp{world} (utterance sharding, '
                   'one all-gather of the codes at the end)', 'kv_page': 64,
                   'sub_batches': n_sub, 'decode_gemm': decode_form(eng, st['subs'][0]), 'launches_per_step': launches_per_step,
                   'l2_policy': 'inputs larger than L2: every step streams 304 MB of weights + %.0f MB of KV' %
                                ((step_bytes - ar_step_bytes(0, 0)) / 1e6),
                   'step_hbm_bytes': step_bytes, 'step_hbm_frac_of_measured_peak':
                       step_bytes / (ms / K * 1e-3) / (pk['hbm_gbs'] * 1e9)},
        'clocks': clocks, 'gpu_launches': launches_per_step * K,
    }

    if rank == 0:
        # ---- roofline of the dominant kernel (paged decode attention), CUDA events on the launch stream ----
        H, Dh, d = 16, 64, 1024
        ctx_r = int(round(mean_ctx))
        st['seq_lens'].fill_(ctx_r)
        sb = st['subs'][0]                                   # the launch shape the step really uses (one sub-batch)
        Bs = sb['B']
        torch.cuda.synchronize()
        reps = 5
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        lean, tc = eng._lean_ok(sb), eng._tc_ok(sb)
        if lean:
            qkv_src, qkv_np, qkv_ps = sb['r_qkv'], 1, 0
        elif tc:
            qkv_src, qkv_np, qkv_ps = sb['qkv32'], 1, 0
        else:
            qkv_src, qkv_np, qkv_ps = sb['p_qkv'], sb['ns']['qkv'], Bs * 3 * d

        def attn_all_layers():
            for li in range(12):                             # 12 layers x B x ctx KV = > L2, no re-use between launches
                ops.attn_decode_paged(qkv_src, qkv_np, qkv_ps, st['pools'][li], sb['block_table'], sb['seq_lens'],
                                      sb['o'], Bs, H, Dh, sb['n_tsplit'], sb['attn_ws'], eng.attn_flags)

        def timed_graph(fn):
            """Kernel time without host launch overhead: capture fn once, replay it reps times between two events."""
            fn()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                fn()
            gr.replay()
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(reps):
                gr.replay()
            ev[1].record()
            torch.cuda.synchronize()
            return ev[0].elapsed_time(ev[1]) / reps

        att_ms = timed_graph(attn_all_layers) / 12
        att_bytes = Bs * (ctx_r + 1) * 2 * d * 2           # K and V rows of every cached position, bf16
        traffic = None
        try:                                                 # DRAM bytes per launch from the committed ncu --set full capture
            with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
                tr = json.load(fh)['attn_decode_kernel<bf16>' if eng.attn_flags else 'attn_decode_mma_kernel<bf16>']
            # the capture is at B=32, ctx=750; scale linearly to this launch's algorithmic bytes
            traffic = (tr['dram_bytes_read'] + tr['dram_bytes_write']) * att_bytes / tr['algorithmic_bytes']
        except Exception:
            pass
        result['roofline'] = {'bound': 'hbm', 'kernel': 'attn_decode_kernel<bf16> (SIMT)' if eng.attn_flags else 'attn_decode_mma_kernel<bf16>', 'achieved': att_bytes / (att_ms * 1e-3) / 1e9,
                              'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': att_bytes / (att_ms * 1e-3) / 1e9 / pk['hbm_gbs'],
                              'traffic': traffic, 'traffic_source': 'profiles/traffic.json (ncu --set full, scaled to this launch)',
                              'peak_source': pk['_source'] + ' (MEASURED_PEAKS.json hbm_gbs)',
                              'bytes_per_launch': att_bytes, 'us_per_launch': att_ms * 1e3,
                              'share_of_step': att_ms * 12 * n_sub / (ms / K), 'rows_per_launch': Bs,
                              'launches_per_step': 12 * n_sub}
        # weight-streaming GEMMs of one step (incl. their fused reductions / epilogues), same method
        def gemms_all_layers():
            for L in eng.weights.layers:
                if lean:
                    g1, b1, _ = L['norm1']
                    g2, b2, _ = L['norm2']
                    ops.linear_decode_rows_ln(sb['x'], L['wqkv'], sb['r_qkv'][0], gamma=g1[0], beta=b1[0])
                    ops.linear_decode_rows(sb['o'], L['wo'], sb['x'], bias=L['bo'], residual=True)
                    ops.linear_decode_rows_ln(sb['x'], L['w1'], sb['f'], gamma=g2[0], beta=b2[0], bias=L['b1'], gelu=True)
                    ops.linear_decode_rows(sb['f'], L['w2'], sb['x'], bias=L['b2'], residual=True, want_split=0)
                elif tc:
                    dg = sb['dg']
                    ch_o, ch_f2 = dg['o']['tiles'], dg['f2']['tiles']
                    eng._dg(sb, 'qkv', sb['xb'], L['wqkv_s'], ops.DG_LN, bias=L['b_qkv'], colsum=L['c_qkv'], stats_in=sb['stats'],
                            n_chunks_in=ch_f2, y32=sb['qkv32'])
                    eng._dg(sb, 'o', sb['o'], L['wo'], ops.DG_RESIDUAL, bias=L['bo'], xres=sb['x'], y16=sb['xb'], stats_out=sb['stats'])
                    eng._dg(sb, 'f1', sb['xb'], L['w1_s'], ops.DG_LN_GELU, bias=L['b_1'], colsum=L['c_1'], stats_in=sb['stats'],
                            n_chunks_in=ch_o, y16=sb['f'])
                    eng._dg(sb, 'f2', sb['f'], L['w2'], ops.DG_RESIDUAL, bias=L['b2'], xres=sb['x'], y16=sb['xb'], stats_out=sb['stats'])
                else:
                    ops.linear_decode(sb['h'], L['wqkv'], sb['p_qkv'], Bs * 3 * d, 32)
                    ops.linear_decode(sb['o'], L['wo'], sb['p_o'], Bs * d, 32)
                    ops.linear_decode(sb['h'], L['w1'], sb['p_f1'], Bs * 4096, 32)
                    ops.linear_decode(sb['f'], L['w2'], sb['p_f2'], Bs * d, 32)

        sb['x'].zero_()
        gemm_ms = timed_graph(gemms_all_layers)
        wbytes = 2 * 12 * (3 * d * d + d * d + 2 * d * 4096)
        result['gemm_decode'] = {'ms_per_step': gemm_ms, 'achieved_gbs': wbytes / (gemm_ms * 1e-3) / 1e9,
                                 'frac_of_hbm_peak': wbytes / (gemm_ms * 1e-3) / 1e9 / pk['hbm_gbs'], 'launches': 48,
                                 'kernel': decode_form(eng, sb)}

    # ---- end-to-end through the public API from pinned host tensors --------------------------------
    # One untimed call first (allocates the KV pools and captures the step graph of this request shape -- the engine keeps
    # both for later requests of the same shape), then three timed calls; the median is reported and all three are listed.
    # Every timed call does the full job: H2D prompt, prefill, decode, D2H codes.
    def e2e_once():
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        out, n = model.generate_batch(tokens_h.to(dev, non_blocking=True), codes_h.to(dev, non_blocking=True),
                                      max_new=total_steps, ignore_eos=True)
        out_h = out.to('cpu', non_blocking=False)
        t1.record()
        barrier()
        t_ms = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([t_ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            t_ms = float(t.item())
        return t_ms, out_h

    e2e_once()
    runs = [e2e_once() for _ in range(3)]
    e2e_all = sorted(r[0] for r in runs)
    e2e_ms, out_h = e2e_all[1], runs[0][1]
    result['e2e'] = {'value': B * world * total_steps / (e2e_ms * 1e-3), 'unit': 'tokens/s',
                     'h2d_bytes_per_step': (tokens_h.numel() + codes_h.numel()) * 8 / total_steps,
                     'd2h_bytes_per_step': out_h.numel() * 4 / total_steps,
                     'includes': f'H2D prompt, prefill of {TX + P} positions, {total_steps} decode steps (step graph and KV pools of this '
                                 f'request shape are reused from the untimed first call), D2H codes',
                     'ms_total': e2e_ms, 'ms_all_runs': e2e_all, 'runs': 'median of 3 after one untimed call'}

    if rank == 0 and not args.no_extras:
        result['extras'] = extras(args, dev, tmp)
        result['cpu_baseline'] = cpu_baseline(model, B, budget_s=12.0)
    return result


def extras(args, dev, tmp):
    """Secondary numbers of BASELINE.json's metric: AR decode at batch 1, NAR frames/s (config 3, reduced batch if needed)."""
    import valle2_b200
    from valle2_b200.models import ValleAR, ValleNAR
    out = {}
    pk = peaks()
    torch.manual_seed(0)
    model = ValleAR(large_cfg('LayerNorm', tmp)).eval().to(dev)
    eng = model._engine()
    g = torch.Generator().manual_seed(7)
    samp = {'temperature': 1.0, 'top_k': 1, 'top_p': 1.0, 'seed': 0}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # batch 1 (BASELINE configs[1]) over the full 750 frames; batch 128 / 64 = one GPU's share of configs[3] (256 utterances
    # over 2 / 4 GPUs), 150 steps around the same mean context
    for Bx, K1, extra in ((1, N_NEW - 8, 0), (64, 150, 300), (128, 150, 300)):
        tokens = torch.randint(0, 256, (Bx, TX), generator=g).to(dev)
        codes = torch.cat([torch.full((Bx, 1), 1025), torch.randint(0, 1024, (Bx, P0 + extra - 1), generator=g)], 1).to(dev)
        st = eng.prefill(tokens, codes, max_new=K1 + 10)
        eng.first_token(samp, None, -1)
        eng.decode_step(samp, None, -1)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            eng.decode_step(samp, None, -1)
        for _ in range(4):
            graph.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(K1):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        mean_ctx = TX + P0 + extra + 6 + (K1 - 1) / 2
        out[f'ar_decode_b{Bx}'] = {'tokens_per_s': Bx * K1 / (ms * 1e-3), 'us_per_step': ms * 1e3 / K1, 'mean_ctx': mean_ctx,
                                   'launches_per_step': eng.launches_per_step(),
                                   'hbm_frac_of_measured_peak': ar_step_bytes(Bx, mean_ctx) / (ms / K1 * 1e-3) / (pk['hbm_gbs'] * 1e9)}
        del st, graph
    del model, eng
    torch.cuda.empty_cache()
    # NAR config 3: 7 stages, S = 150 + 225 + 525 = 900, batch 64
    try:
        torch.manual_seed(1)
        nar = ValleNAR(large_cfg('AdaptiveLayerNorm', tmp)).eval().to(dev)
        Bn, Tc, Tt = args.nar_batch, 225, 525
        pt = torch.randint(0, 256, (Bn, 50), generator=g).to(dev)
        tt = torch.randint(0, 256, (Bn, 100), generator=g).to(dev)
        pc = torch.randint(0, 1024, (Bn, Tc, 8), generator=g).to(dev)
        fl = torch.randint(0, 1024, (Bn, Tt), generator=g).to(dev)
        use_tc = bool(int(os.environ.get('VALLE_B200_TC_ATTN', '1')))
        nar.generate_batch(pt[:2], pc[:2], tt[:2], fl[:2], use_tc_attention=use_tc)     # warm-up: kernels, func attributes
        nar.generate_batch(pt, pc, tt, fl, use_tc_attention=use_tc)                     # warm-up: workspaces of this shape
        nar_ms = []
        for _ in range(3):              # median of three full runs (the first call of a shape pays cudaMalloc for ~2 GB)
            torch.cuda.synchronize()
            e0.record()
            nar.generate_batch(pt, pc, tt, fl, use_tc_attention=use_tc)
            e1.record()
            torch.cuda.synchronize()
            nar_ms.append(e0.elapsed_time(e1))
        ms = sorted(nar_ms)[1]
        S, d, F, L = 900, 1024, 4096, 12
        flops = 7 * Bn * (L * (S * 2 * (4 * d * d + 2 * d * F) + 4 * S * S * d) + 2 * Tt * d * 1024)
        out['nar'] = {'batch': Bn, 'stage_frames_per_s': Bn * Tt * 7 / (ms * 1e-3), 'utterance_frames_per_s': Bn * Tt / (ms * 1e-3),
                      'ms_total': ms, 'ms_all_runs': nar_ms, 'tflops': flops / (ms * 1e-3) / 1e12,
                      'frac_of_bf16_sustained_peak': flops / (ms * 1e-3) / 1e12 / pk['bf16_tflops_sustained'],
                      'attention': 'tcgen05' if use_tc else 'simt'}
    except Exception as e:  # report, do not hide
        out['nar'] = {'error': repr(e)}
    # Full TTS hand-off (BASELINE configs[3], one GPU's share of 256 utterances over 8 GPUs = 32): AR prefill + 750 decode
    # steps + 7 NAR stages through valle2_b200.tts.synthesize_batch, from host tensors to host code matrices
    try:
        from valle2_b200.tts import synthesize_batch
        torch.manual_seed(0)
        ar_t = ValleAR(large_cfg('LayerNorm', tmp)).eval().to(dev)
        Bt_ = 32
        ptk = torch.randint(0, 256, (Bt_, 50), generator=g)
        ttk = torch.randint(0, 256, (Bt_, 100), generator=g)
        pcd = torch.randint(0, 1024, (Bt_, 225, 8), generator=g)
        for it in range(3):             # first call allocates / captures for this shape, then two timed calls (the second is reported)
            torch.cuda.synchronize()
            e0.record()
            res = synthesize_batch(ar_t, nar, ptk.to(dev), pcd.to(dev), ttk.to(dev), max_new=N_NEW, ignore_eos=True)
            res_h = [r.cpu() for r in res]
            e1.record()
            torch.cuda.synchronize()
            ms_t = e0.elapsed_time(e1)
        out['tts'] = {'utterances': Bt_, 'frames_per_utterance': int(res_h[0].shape[0]), 'codebooks': int(res_h[0].shape[1]),
                      'ms_total': ms_t, 'utterances_per_s': Bt_ / (ms_t * 1e-3),
                      'audio_seconds_per_s': Bt_ * res_h[0].shape[0] / 75.0 / (ms_t * 1e-3),
                      'note': 'AR (prefill 376 + 750 decode steps, greedy, EOS ignored) + NAR stages 2..8, batch 32 = one GPU of the '
                              '8-GPU sharding of configs[3]; codes out on the host'}
        del ar_t
    except Exception as ex:  # noqa: BLE001
        out['tts'] = {'error': repr(ex)[:300]}
    # Training step (BASELINE config 5, per-GPU share): teacher-forced AR, 16 clips of 15 s (Ty = 1126, Tx = 225), bf16
    # operands / fp32 accumulation, forward + backward on the CUDA stack (valle2_b200/train.py)
    try:
        del nar
        torch.cuda.empty_cache()
        torch.manual_seed(2)
        ar = ValleAR(large_cfg('LayerNorm', tmp)).train().to(dev)
        Bt, Txt, Tyt = args.train_batch, 225, 1126
        batch = {'tokens': torch.randint(0, 256, (Bt, Txt), generator=g), 'tokens_lens': torch.full((Bt,), Txt),
                 'codes': torch.randint(0, 1024, (Bt, Tyt), generator=g), 'codes_lens': torch.full((Bt,), Tyt),
                 'target': torch.randint(0, 1025, (Bt, Tyt), generator=g)}
        for it in range(6):             # three warm-up steps (allocator growth, first-use kernel loads), three timed
            if it == 3:
                torch.cuda.synchronize()
                e0.record()
            for p_ in ar.parameters():
                p_.grad = None
            loss = ar.training_step(batch)
            loss.backward()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        S, d, F, L = Txt + Tyt, 1024, 4096, 12
        fwd = Bt * (L * (S * 2 * (4 * d * d + 2 * d * F) + 4 * S * S * d) + 2 * Tyt * d * 1025)
        out['train_step'] = {'batch': Bt, 'seq': S, 'ms_per_step': ms, 'clips_per_s': Bt / (ms * 1e-3), 'loss': float(loss.detach()),
                             'model_tflops': 3 * fwd / (ms * 1e-3) / 1e12,
                             'frac_of_bf16_sustained_peak': 3 * fwd / (ms * 1e-3) / 1e12 / pk['bf16_tflops_sustained'],
                             'note': 'forward + backward, no optimizer step; flops = 3 x dense forward (no causal discount)',
                             'peak_mem_gb': torch.cuda.max_memory_allocated() / 1e9}
        # the NAR half of configs[4]: same clips, all 8 codebooks, stage drawn by training_step (valle_nar.py:76)
        del ar
        torch.cuda.empty_cache()
        torch.manual_seed(3)
        nar_t = ValleNAR(large_cfg('AdaptiveLayerNorm', tmp)).train().to(dev)
        nbatch = {'tokens': batch['tokens'], 'tokens_lens': batch['tokens_lens'],
                  'codes': torch.randint(0, 1024, (Bt, Tyt - 1, 8), generator=g), 'codes_lens': torch.full((Bt,), Tyt - 1)}
        for it in range(6):
            if it == 3:
                torch.cuda.synchronize()
                e0.record()
            for p_ in nar_t.parameters():
                p_.grad = None
            nloss = nar_t.training_step(nbatch)
            nloss.backward()
        e1.record()
        torch.cuda.synchronize()
        nms = e0.elapsed_time(e1) / 3
        Sn = Txt + Tyt - 1
        nfwd = Bt * (L * (Sn * 2 * (4 * d * d + 2 * d * F) + 4 * Sn * Sn * d))
        out['train_step_nar'] = {'batch': Bt, 'seq': Sn, 'ms_per_step': nms, 'clips_per_s': Bt / (nms * 1e-3), 'loss': float(nloss.detach()),
                                 'model_tflops': 3 * nfwd / (nms * 1e-3) / 1e12,
                                 'frac_of_bf16_sustained_peak': 3 * nfwd / (nms * 1e-3) / 1e12 / pk['bf16_tflops_sustained'],
                                 'note': 'ValleNAR.training_step forward + backward (full attention, AdaLN), stage drawn per step'}
        del nar_t
    except Exception as e:  # report, do not hide
        out['train_step'] = {'error': repr(e)}
    return out


# ------------------------------------------------------------------------------------------------------------
def _cpu_decode_rate(sd, oc, B: int, n_steps: int, ctx: int, warm: int = 1):
    """Oracle port (reference algorithm incl. its torch.cat cache growth) on the host cores: tokens/s of n_steps
    KV-cached decode steps at batch B starting from a synthetic cache of `ctx` positions."""
    from oracle import valle_oracle as vo
    d, H, Dh, L = oc.d_model, oc.n_heads, oc.d_model // oc.n_heads, oc.num_layers
    g = torch.Generator().manual_seed(0)
    kv = tuple((torch.randn(B, H, ctx, Dh, generator=g), torch.randn(B, H, ctx, Dh, generator=g)) for _ in range(L))
    x = torch.randn(B, 1, d, generator=g)
    times = []
    for s in range(warm + n_steps):
        t0 = time.perf_counter()
        h, kv = vo.transformer(x, sd, oc, kv_cache=kv, use_cache=True)
        logits = (h @ sd['proj.weight'].t())[:, -1]
        tok, _ = vo.topk_sampling(logits, 1, 1.0, 1.0)
        x = sd['audio_emb.word_embeddings.weight'][tok] + 0.0
        dt = time.perf_counter() - t0
        if s >= warm:
            times.append(dt)
    return B * len(times) / sum(times), sum(times)


def cpu_baseline(model, B: int, budget_s: float):
    from oracle.valle_oracle import OracleConfig
    torch.set_num_threads(os.cpu_count() or 1)
    oc = OracleConfig.from_any(model.config)
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    rate1, t1 = _cpu_decode_rate(sd, oc, B, 1, int(MEAN_CTX), warm=1)
    n = max(2, min(160, int(budget_s / max(t1, 1e-3))))          # bounded sample: ~budget_s seconds of CPU work
    ctx = int(MEAN_CTX - n / 2)                                   # centred on the GPU workload's mean context
    rate, total = _cpu_decode_rate(sd, oc, B, n, ctx, warm=0)
    return {'value': rate, 'unit': 'tokens/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'oracle port (fp32, torch CPU ops, reference algorithm incl. torch.cat KV growth): {n} decode steps at batch '
                      f'{B}, ctx {ctx}->{ctx + n}, {total:.1f} s of CPU work; host has {os.cpu_count()} logical cores'}


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Python/PyTorch and its
    tree does not exist on the GPU box, so this times the oracle port (same algorithm, same ops) on all host threads."""
    if rank != 0:
        return None
    from oracle import synth
    torch.set_num_threads(os.cpu_count() or 1)
    oc = synth.large_config('LayerNorm')
    sd = synth.synth_state_dict(synth.ar_state_shapes(oc), 0)
    K, W = args.steps, args.warmup
    ctx = int(round(MEAN_CTX - (K - 1) / 2.0)) if K < N_NEW else TX + P0
    B = args.batch
    rate, t = _cpu_decode_rate(sd, oc, B, 1, ctx, warm=1)
    while B > 1 and t * (K + W) > 150.0:       # bounded sample: shrink the batch until K steps fit in a few minutes
        B = max(1, B // 4)
        rate, t = _cpu_decode_rate(sd, oc, B, 1, ctx, warm=0)
    k_eff = K if t * (K + W) <= 200.0 else max(1, int(200.0 / t) - W)
    rate, total = _cpu_decode_rate(sd, oc, B, k_eff, ctx, warm=W)
    sample = (f'oracle port on {torch.get_num_threads()} host threads: {k_eff} timed decode steps (of K={K}) at batch {B}, '
              f'ctx {ctx}->{ctx + k_eff}, fp32')
    return {'impl': 'reference', 'metric': 'ar_decode_tokens_per_s', 'value': rate, 'unit': 'tokens/s', 'n_gpus': world,
            'steps': K, 'warmup': W, 'ms_per_step': total / k_eff * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'AR decode, default VALL-E AR decoder 12L d1024 h16 F4096, KV-cached greedy (BASELINE configs[1]); '
                                   'CPU sample: ' + sample, 'batch_per_gpu': B},
            'cpu_baseline': {'value': rate, 'unit': 'tokens/s', 'cores': torch.get_num_threads(), 'kind': 'port', 'sample': sample},
            'e2e': {'value': rate, 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=N_NEW)
    ap.add_argument('--warmup', type=int, default=8)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--nar-batch', type=int, default=64)
    ap.add_argument('--train-batch', type=int, default=16)
    ap.add_argument('--no-extras', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        res = run_reference(args, rank, world)
        if res is not None:
            print(json.dumps(res), flush=True)
        return 0
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'      # keep stdout to the single JSON line
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    res = run_ours(args, rank, world, local_rank)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
