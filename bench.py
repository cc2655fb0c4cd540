#!/usr/bin/env python
"""Benchmark of the Valle2 hot path on B200 (contract: task statement / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--utterances G] [--no-extras]

Workload = BASELINE.json configs[3], the one configuration that is defined at 1, 2, 4 and 8 GPUs: full TTS inference for
G = 256 utterances (text 150 phonemes = 50 prompt + 100 target, 3 s prompt = 225 frames x 8 codebooks, 750 generated frames =
10 s): AR prefill + KV-cached greedy decode of the first codebook + NAR stages 2..8, bf16 weights / KV with fp32
accumulation.  The utterances are SHARDED over the N ranks (strong scaling: G is fixed), every rank decodes its shard
independently and the codes are all-gathered with NCCL at the end -- inside the timed region.  One "step" = one whole job
through `valle2_b200.parallel.generate_sharded` + `valle2_b200.tts.synthesize_batch`.

Prints ONE JSON line on rank 0:
  value     codec frames/s (1 frame = 1 AR token + 7 NAR tokens) with the inputs resident in HBM, CUDA events between
            barriers, max over ranks;  utterances/s = value / 750
  e2e       the same job from pinned HOST tensors: H2D of the shard, job, gather, D2H of the gathered codes
  roofline  the dominant kernel (paged decode attention) at this run's per-GPU batch, timed live with CUDA events
  extras    ar_decode_weak: BASELINE configs[1] (AR decode, batch 32 PER GPU, 750 frames; tokens/s, weak scaling, with the
            step's HBM roofline fraction); on one GPU also batch 1, NAR configs[2], the training steps of configs[4] and the
            torch-eager reference on the same GPU
  cpu_baseline  the EXECUTED reference (oracle/_ref, unmodified) on the host cores, bounded sample (N = 1 only)
  config    the WORKLOAD only (built by workload_config(): both arms print the identical object)
  run       what this arm did with it: launches per decode step, decode GEMM form, sharded_equals_single_gpu, derived rates
`--impl reference` times that executed reference alone (rank 0), same metric / unit / config.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TP, TT, TC, N_NEW = 50, 100, 225, 750   # prompt phonemes, target phonemes, prompt frames, generated frames
TX, P0 = TP + TT, TC + 1                # text length, BOS + prompt
MEAN_CTX = TX + P0 + (N_NEW - 1) / 2.0  # 750.5
Q = 8


def workload_string(G: int) -> str:
    """Identical in both arms (the driver compares it)."""
    return (f'full TTS inference (BASELINE configs[3]): {G} utterances, text {TX} phonemes ({TP} prompt + {TT} target), '
            f'{TC}-frame prompt x {Q} codebooks, {N_NEW} generated frames: default VALL-E AR decoder (12L d1024 h16 F4096) prefill + '
            f'KV-cached greedy decode + NAR stages 2-{Q}; utterances sharded over the GPUs, codes all-gathered at the end')


def workload_config(G: int, world: int) -> dict:
    """`config` of the JSON line: the WORKLOAD only, built by one function so that both arms (ours / --impl reference) print the
    same object; what an arm did with it (launch counts, kernels, samples, derived rates) goes to the line's `run` object."""
    n_local = (G + world - 1) // world
    return {'workload': workload_string(G), 'utterances': G, 'utterances_per_gpu': n_local,
            'parallelism': f'dp{world}: utterance sharding (round robin), one all-gather of the int32 codes + lengths at the end of '
                           'every job, inside the timed region',
            'l2_policy': 'inputs larger than L2: every decode step streams 304 MB of weights + %.0f MB of KV per GPU' %
                         ((ar_step_bytes(n_local, MEAN_CTX) - ar_step_bytes(0, 0)) / 1e6)}


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
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), f'--query-gpu={self.QUERY}',
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
        return ('tc: decode_gemm_kernel (tcgen05 swap-AB, split-K exchanged through DSMEM inside the launch, LayerNorm folded, fused '
                'epilogues), 5 launches per layer')
    return 'splitk: gemm_tc_kernel<swap-AB split-K> slices + LayerNorm / GELU-reduce kernels, 8 launches per layer'


def ar_step_bytes(B: int, ctx: float, L=12, d=1024, F=4096, V=1025) -> float:
    """Algorithmic HBM bytes of one decode step: all weights once + K/V of every cached position."""
    w = 2 * (L * (3 * d * d + d * d + 2 * d * F) + V * d)
    kv_tok = 2 * L * d * 2
    return w + B * (ctx + 1) * kv_tok


def timed_graph(fn, reps: int = 5) -> float:
    """ms per fn(): captured once, replayed reps times between two events (no host launch overhead)."""
    fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ------------------------------------------------------------------------------------------------------------
@torch.inference_mode()
def attention_roofline(eng, ops, ctx: int, pk: dict, step_ms: float | None) -> dict:
    """The dominant kernel at the launch shape the engine really uses (first sub-batch): algorithmic bytes per launch over
    the CUDA-event time of 12 launches on 12 distinct layer pools (> L2, no re-use between launches)."""
    st = eng._state
    sb = st['subs'][0]
    Bs, H, Dh, d = sb['B'], eng.H, eng.Dh, eng.d
    keep = st['seq_lens'].clone()
    st['seq_lens'].fill_(ctx)
    if eng._lean_ok(sb):
        src, n_p, p_s = sb['r_qkv'], 1, 0
    elif eng._tc_ok(sb):
        src, n_p, p_s = sb['qkv32'], 1, 0
    else:
        src, n_p, p_s = sb['p_qkv'], sb['ns']['qkv'], Bs * 3 * d

    def attn_all_layers():
        for li in range(len(eng.weights.layers)):
            ops.attn_decode_paged(src, n_p, p_s, st['pools'][li], sb['block_table'], sb['seq_lens'], sb['o'], Bs, H, Dh,
                                  sb['n_tsplit'], sb['attn_ws'], eng.attn_flags)

    L = len(eng.weights.layers)
    att_ms = timed_graph(attn_all_layers) / L
    st['seq_lens'].copy_(keep)
    att_bytes = Bs * (ctx + 1) * 2 * d * 2               # K and V rows of every cached position, bf16
    traffic = None
    try:                                                 # DRAM bytes per launch from the committed ncu --set full capture
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            tr = json.load(fh)['attn_decode_mma_kernel<bf16>']
        traffic = (tr['dram_bytes_read'] + tr['dram_bytes_write']) * att_bytes / tr['algorithmic_bytes']
    except Exception:
        pass
    out = {'bound': 'hbm', 'kernel': 'attn_decode_mma_kernel<bf16>', 'achieved': att_bytes / (att_ms * 1e-3) / 1e9,
           'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': att_bytes / (att_ms * 1e-3) / 1e9 / pk['hbm_gbs'], 'traffic': traffic,
           'traffic_source': 'profiles/traffic.json (ncu --set full at B=32, ctx=750, scaled to this launch)',
           'peak_source': pk['_source'] + ' (MEASURED_PEAKS.json hbm_gbs)', 'bytes_per_launch': att_bytes,
           'us_per_launch': att_ms * 1e3, 'rows_per_launch': Bs, 'ctx': ctx, 'launches_per_step': L * len(st['subs'])}
    if step_ms:
        out['share_of_decode_step'] = att_ms * L * len(st['subs']) / step_ms
    return out


@torch.inference_mode()
def ar_decode_steady(model, B: int, K: int, extra_prompt: int, dev, gen_seed: int, world: int = 1, warm: int = 8):
    """BASELINE configs[1]: K CUDA-graph-replayed decode steps at batch B (device-resident), prompt lengthened by
    extra_prompt so that the mean context over the K steps is the config's 750.5 when K < 750.  All ranks take part
    (barrier + max over ranks).  Returns (tokens/s of B x world sequences, ms per step, mean ctx, engine)."""
    eng = model._engine()
    g = torch.Generator().manual_seed(gen_seed)
    tokens = torch.randint(0, 256, (B, TX), generator=g).to(dev)
    codes = torch.cat([torch.full((B, 1), 1025), torch.randint(0, 1024, (B, P0 + extra_prompt - 1), generator=g)], 1).to(dev)
    samp = {'temperature': 1.0, 'top_k': 1, 'top_p': 1.0, 'seed': 0}
    st = eng.prefill(tokens, codes, max_new=K + warm + 4)
    eng.first_token(samp, None, -1)
    eng.decode_step(samp, None, -1)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        eng.decode_step(samp, None, -1)
    for _ in range(warm):
        graph.replay()
    ctx0 = int(st['seq_lens'][0].item())
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        graph.replay()
    e1.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    return B * world * K / (ms * 1e-3), ms / K, ctx0 + (K - 1) / 2.0, eng


# ------------------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, local_rank: int):
    import valle2_b200
    from valle2_b200 import _lib, ops, parallel
    from valle2_b200.models import ValleAR, ValleNAR
    from valle2_b200.tts import synthesize_batch

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    valle2_b200.set_precision('bf16')
    tmp = f'/tmp/valle_bench_{os.getpid()}'
    G, K, W = args.utterances, args.steps, args.warmup
    pk = peaks()
    torch.manual_seed(0)
    ar = ValleAR(large_cfg('LayerNorm', tmp)).eval().to(dev)
    torch.manual_seed(1)
    nar = ValleNAR(large_cfg('AdaptiveLayerNorm', tmp)).eval().to(dev)
    g = torch.Generator().manual_seed(100)
    host = {'pt': torch.randint(0, 256, (G, TP), generator=g).pin_memory(),
            'tt': torch.randint(0, 256, (G, TT), generator=g).pin_memory(),
            'pc': torch.randint(0, 1024, (G, TC, Q), generator=g).pin_memory()}
    on_dev = {k: v.to(dev) for k, v in host.items()}

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def decode_fn(src):
        def fn(idx):
            """This rank's shard -> (rows (n, 750 * 8) int32, lens): the public hand-off API, AR then NAR."""
            sel = torch.as_tensor(idx, dtype=torch.long)
            if src is on_dev:
                sel = sel.to(dev)
                pt, tt, pc = (src[k][sel] for k in ('pt', 'tt', 'pc'))
            else:           # pinned host tensors: gather the shard on the host, one H2D copy per tensor
                pt, tt, pc = (src[k][sel].pin_memory().to(dev, non_blocking=True) for k in ('pt', 'tt', 'pc'))
            mats = synthesize_batch(ar, nar, pt, pc, tt, max_new=N_NEW, ignore_eos=True, nar_chunk=args.nar_chunk)
            rows = torch.stack(mats).reshape(len(idx), -1).to(torch.int32)
            return rows, torch.full((len(idx),), rows.shape[1], device=dev, dtype=torch.int32)
        return fn

    def job(src):
        rows, lens = parallel.generate_sharded(decode_fn(src), G, pad_value=-1)
        return rows

    # ---- device-resident measurement: W warm-up jobs, K timed jobs, gather inside --------------------------------
    for _ in range(W):
        out = job(on_dev)
    eng = ar._engine()
    _lib.LAUNCHES = 0
    replays0 = eng.replays
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        out = job(on_dev)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = _lib.LAUNCHES + (eng.replays - replays0) * eng.launches_per_step()
    if world > 1:
        t = torch.tensor([ms, float(launches)], device=dev)
        torch.distributed.all_reduce(t[:1], op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(t[1:], op=torch.distributed.ReduceOp.SUM)
        ms, launches = float(t[0].item()), int(t[1].item())
    assert out.shape == (G, N_NEW * Q) and int(out.min()) >= 0
    n_local = len(parallel.shard_indices(G, rank, world))
    frames_s = G * N_NEW * K / (ms * 1e-3)
    phases = dict(getattr(eng, 'last_phases', {}))

    # ---- sharded == single: rank 0 decodes another rank's shard alone and compares with what the gather delivered ----
    equal = None
    if world > 1:
        other = parallel.shard_indices(G, 1, world)
        if rank == 0:
            rows1, _ = decode_fn(on_dev)(other)
            equal = bool(torch.equal(rows1, out[torch.as_tensor(other, device=dev)]))
            assert equal, 'gathered codes of rank 1 differ from a single-GPU decode of the same utterances'
        barrier()

    # ---- end to end from pinned host tensors (H2D of the shard, job, gather, D2H of the gathered codes) ----------
    ke = max(1, min(K, args.e2e_steps))
    job(host)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(ke):
        out_h = job(host).to('cpu')
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = float(t.item())
    assert torch.equal(out_h, out.cpu()), 'host-fed job differs from the device-resident one'

    st = eng._state
    sb = st['subs'][0]
    result = {
        'metric': 'tts_codec_frames_per_s', 'value': frames_s, 'unit': 'frames/s', 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'bf16',
        'data': 'synthetic',
        'config': workload_config(G, world),
        'run': {'utterances_per_s': frames_s / N_NEW, 'audio_seconds_per_s': frames_s / 75.0, 'kv_page': 64,
                'decode_gemm': decode_form(eng, sb), 'launches_per_decode_step': eng.launches_per_step(),
                'nar_chunk': args.nar_chunk, 'sharded_equals_single_gpu': equal},
        'clocks': clocks, 'gpu_launches': launches,
        'e2e': {'value': G * N_NEW * ke / (e2e_ms * 1e-3), 'unit': 'frames/s',
                'h2d_bytes_per_step': sum(v.numel() for v in host.values()) * 8,
                'd2h_bytes_per_step': out_h.numel() * 4, 'steps': ke, 'ms_per_step': e2e_ms / ke,
                'includes': 'per job: H2D of every rank\'s shard from pinned host tensors, AR prefill, 750 decode steps, NAR stages, '
                            'all-gather, D2H of the gathered codes (step graph and KV pools are reused between jobs)',
                'phases_ms_rank0_last_job': phases},
    }
    if rank == 0:
        result['roofline'] = attention_roofline(eng, ops, int(round(MEAN_CTX)), pk, None)
        if phases.get('ar_prefill_ms'):
            # AR prompt prefill of this rank's shard (tensor-core bound): 2 x params x tokens + attention, SURVEY 8d (120.5 GFLOP / sequence)
            S0, d_, F_, L_ = TX + P0, 1024, 4096, 12
            fl = n_local * (L_ * (S0 * 2 * (4 * d_ * d_ + 2 * d_ * F_) + 4 * S0 * S0 * d_))
            result['e2e']['ar_prefill'] = {'sequences': n_local, 'positions': S0, 'ms': phases['ar_prefill_ms'],
                                           'tflops': fl / (phases['ar_prefill_ms'] * 1e-3) / 1e12,
                                           'frac_of_bf16_sustained_peak': fl / (phases['ar_prefill_ms'] * 1e-3) / 1e12 / pk['bf16_tflops_sustained'],
                                           'note': 'prefix-LM mask not discounted; includes KV scatter into the page pool and the first token'}

    # ---- BASELINE configs[1] on every rank: AR decode, batch 32 per GPU (weak scaling) --------------------------
    if not args.no_extras:
        extras = {}
        Kw = 300
        tok_s, step_ms, mctx, eng32 = ar_decode_steady(ar, 32, Kw, int(round(MEAN_CTX - (Kw - 1) / 2.0 - 8)) - (TX + P0), dev, 7, world)
        if rank == 0:
            bytes32 = ar_step_bytes(32, mctx)
            roof32 = attention_roofline(eng32, ops, int(round(mctx)), pk, step_ms)
            extras['ar_decode_weak'] = {
                'workload': 'BASELINE configs[1]: AR decode, batch 32 per GPU, KV-cached greedy, CUDA-graph steps, device-resident',
                'tokens_per_s': tok_s, 'ms_per_step': step_ms, 'mean_ctx': mctx, 'batch_per_gpu': 32, 'scaling': 'weak',
                'step_hbm_bytes': bytes32, 'step_hbm_frac_of_measured_peak': bytes32 / (step_ms * 1e-3) / (pk['hbm_gbs'] * 1e9),
                'launches_per_step': eng32.launches_per_step(), 'decode_gemm': decode_form(eng32, eng32._state['subs'][0]),
                'roofline_attention': roof32}
        if world > 1:
            del nar
            td = train_step_dp(args, dev, tmp, world, pk)
            if rank == 0:
                extras['train_step_dp'] = td
        if rank == 0 and world == 1:
            extras.update(extras_single_gpu(args, dev, tmp, ar, nar, pk))
            result['cpu_baseline'] = cpu_baseline(budget_s=args.cpu_budget)
        if rank == 0:
            result['extras'] = extras
    return result


def train_step_dp(args, dev, tmp, world: int, pk: dict) -> dict:
    """BASELINE configs[4] across the ranks: teacher-forced AR step, 16 clips of 15 s PER GPU, forward + backward with the
    gradient all-reduce overlapped with the backward pass (parallel.GradReducer), and the same step without any exchange."""
    from valle2_b200 import parallel
    from valle2_b200.models import ValleAR
    torch.cuda.empty_cache()
    torch.manual_seed(2)
    model = ValleAR(large_cfg('LayerNorm', tmp)).train().to(dev)
    g = torch.Generator().manual_seed(11 + int(os.environ.get('RANK', '0')))
    Bt, Txt, Tyt = args.train_batch, 225, 1126
    batch = {'tokens': torch.randint(0, 256, (Bt, Txt), generator=g), 'tokens_lens': torch.full((Bt,), Txt),
             'codes': torch.randint(0, 1024, (Bt, Tyt), generator=g), 'codes_lens': torch.full((Bt,), Tyt),
             'target': torch.randint(0, 1025, (Bt, Tyt), generator=g)}
    red = parallel.GradReducer(model)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    # each variant twice, alternating (the first timed loop of a process also pays allocator growth); the smaller time counts
    for label, reducer in (('local_ms_per_step', None), ('overlapped_allreduce_ms_per_step', red)) * 2:
        for it in range(6):
            if it == 3:
                torch.distributed.barrier()
                torch.cuda.synchronize()
                e0.record()
            for p_ in model.parameters():
                p_.grad = None
            with parallel.reducing(reducer):
                loss = model.training_step(batch)
            loss.backward()
        e1.record()
        torch.distributed.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 3], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        out[label] = min(out.get(label, 1e30), float(t.item()))
    n_grad = sum(p.numel() for p in model.parameters())
    out.update({'workload': 'BASELINE configs[4] per-GPU share: ValleAR.training_step, 16 clips x (225 + 1126) positions per GPU, forward + '
                            'backward, no optimizer step', 'global_batch': Bt * world, 'gradient_bytes_fp32': n_grad * 4,
                'clips_per_s': Bt * world / (out['overlapped_allreduce_ms_per_step'] * 1e-3),
                'slowdown_vs_local': out['overlapped_allreduce_ms_per_step'] / out['local_ms_per_step'],
                'exchange': 'one all_reduce (NCCL AVG) per transformer layer + one for the remaining parameters, enqueued as the backward '
                            'pass finishes each layer'})
    del model, red
    torch.cuda.empty_cache()
    return out


def extras_single_gpu(args, dev, tmp, ar, nar, pk):
    """One GPU only: AR decode at batch 1 (configs[1]), NAR configs[2], training steps of configs[4], torch-eager reference."""
    from valle2_b200.models import ValleAR, ValleNAR
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.Generator().manual_seed(7)
    K1 = N_NEW - 8
    tok_s, step_ms, mctx, eng1 = ar_decode_steady(ar, 1, K1, 0, dev, 9)
    b1 = ar_step_bytes(1, mctx)
    out['ar_decode_b1'] = {'workload': 'BASELINE configs[1]: AR decode, batch 1', 'tokens_per_s': tok_s, 'us_per_step': step_ms * 1e3,
                           'mean_ctx': mctx, 'launches_per_step': eng1.launches_per_step(),
                           'step_hbm_frac_of_measured_peak': b1 / (step_ms * 1e-3) / (pk['hbm_gbs'] * 1e9),
                           'roofline': {'bound': 'hbm', 'achieved': b1 / (step_ms * 1e-3) / 1e9, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                                        'frac': b1 / (step_ms * 1e-3) / (pk['hbm_gbs'] * 1e9), 'bytes_per_step': b1,
                                        'note': 'whole step (all weights + KV once), not one kernel'}}
    # NAR configs[2]: 7 stages, S = 150 + 225 + 525 = 900, batch 64
    try:
        Bn, Tt_ = args.nar_batch, 525
        pt = torch.randint(0, 256, (Bn, TP), generator=g).to(dev)
        tt = torch.randint(0, 256, (Bn, TT), generator=g).to(dev)
        pc = torch.randint(0, 1024, (Bn, TC, Q), generator=g).to(dev)
        fl = torch.randint(0, 1024, (Bn, Tt_), generator=g).to(dev)
        nar.generate_batch(pt[:2], pc[:2], tt[:2], fl[:2])
        nar.generate_batch(pt, pc, tt, fl)
        nar_ms = []
        for _ in range(3):
            torch.cuda.synchronize()
            e0.record()
            nar.generate_batch(pt, pc, tt, fl)
            e1.record()
            torch.cuda.synchronize()
            nar_ms.append(e0.elapsed_time(e1))
        ms = sorted(nar_ms)[1]
        S, d, F, L = 900, 1024, 4096, 12
        flops = 7 * Bn * (L * (S * 2 * (4 * d * d + 2 * d * F) + 4 * S * S * d) + 2 * Tt_ * d * 1024)
        out['nar'] = {'workload': 'BASELINE configs[2]', 'batch': Bn, 'stage_frames_per_s': Bn * Tt_ * 7 / (ms * 1e-3),
                      'utterance_frames_per_s': Bn * Tt_ / (ms * 1e-3), 'ms_total': ms, 'ms_all_runs': nar_ms,
                      'tflops': flops / (ms * 1e-3) / 1e12,
                      'frac_of_bf16_sustained_peak': flops / (ms * 1e-3) / 1e12 / pk['bf16_tflops_sustained']}
    except Exception as e:  # report, do not hide
        out['nar'] = {'error': repr(e)[:300]}
    # torch-eager reference on the same GPU ("the real bar on the box")
    try:
        from oracle import ref_runner
        if ref_runner.available():
            torch.cuda.empty_cache()
            prec = torch.get_float32_matmul_precision()
            r = ref_runner.tts_rate('cuda', 32, 40, 4, warmup=4)
            torch.set_float32_matmul_precision(prec)
            out['gpu_eager_reference'] = {k: r[k] for k in ('frames_per_s', 'ar_tokens_per_s', 'nar_stage_frames_per_s', 'prefill_s',
                                                            'step_s_mean', 'nar_stage_s', 'sample')}
            out['gpu_eager_reference']['note'] = ('unmodified reference modules, torch eager on this GPU, fp32 parameters with the TF32 matmul '
                                                  'mode the reference switches on (valle/utils.py:11)')
        else:
            out['gpu_eager_reference'] = {'unavailable': 'oracle/_ref not installed (python -m oracle.build_ref in the authoring container)'}
    except Exception as e:  # noqa: BLE001
        out['gpu_eager_reference'] = {'error': repr(e)[:300]}
    # Training steps (BASELINE configs[4], per-GPU share): teacher-forced AR and NAR, 16 clips of 15 s, forward + backward
    try:
        torch.cuda.empty_cache()
        torch.manual_seed(2)
        ar_t = ValleAR(large_cfg('LayerNorm', tmp)).train().to(dev)
        Bt, Txt, Tyt = args.train_batch, 225, 1126
        batch = {'tokens': torch.randint(0, 256, (Bt, Txt), generator=g), 'tokens_lens': torch.full((Bt,), Txt),
                 'codes': torch.randint(0, 1024, (Bt, Tyt), generator=g), 'codes_lens': torch.full((Bt,), Tyt),
                 'target': torch.randint(0, 1025, (Bt, Tyt), generator=g)}
        def timed_steps(model_, batch_, n_warm=3, n_timed=5):
            """Per-step CUDA-event times after warm-up (allocator growth, first-use kernel loads); the median is reported, all are listed."""
            import random as _random
            _random.seed(3)             # ValleNAR.training_step draws its stage with random.randint
            evs, loss_ = [], None
            for it in range(n_warm + n_timed):
                for p_ in model_.parameters():
                    p_.grad = None
                if it == n_warm:
                    torch.cuda.synchronize()
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
                loss_ = model_.training_step(batch_)
                loss_.backward()
                b_.record()
                if it >= n_warm:
                    evs.append((a_, b_))
            torch.cuda.synchronize()
            times = sorted(a_.elapsed_time(b_) for a_, b_ in evs)
            return times[len(times) // 2], times, loss_

        ms, ms_all, loss = timed_steps(ar_t, batch)
        S, d, F, L = Txt + Tyt, 1024, 4096, 12
        fwd = Bt * (L * (S * 2 * (4 * d * d + 2 * d * F) + 4 * S * S * d) + 2 * Tyt * d * 1025)
        out['train_step'] = {'batch': Bt, 'seq': S, 'ms_per_step': ms, 'clips_per_s': Bt / (ms * 1e-3), 'loss': float(loss.detach()),
                             'model_tflops': 3 * fwd / (ms * 1e-3) / 1e12,
                             'frac_of_bf16_sustained_peak': 3 * fwd / (ms * 1e-3) / 1e12 / pk['bf16_tflops_sustained'],
                             'ms_all_steps': ms_all,
                             'note': 'forward + backward, no optimizer step; median of 5 steps; flops = 3 x dense forward (no causal discount)',
                             'peak_mem_gb': torch.cuda.max_memory_allocated() / 1e9}
        del ar_t
        torch.cuda.empty_cache()
        torch.manual_seed(3)
        nar_t = ValleNAR(large_cfg('AdaptiveLayerNorm', tmp)).train().to(dev)
        nbatch = {'tokens': batch['tokens'], 'tokens_lens': batch['tokens_lens'],
                  'codes': torch.randint(0, 1024, (Bt, Tyt - 1, 8), generator=g), 'codes_lens': torch.full((Bt,), Tyt - 1)}
        nms, nms_all, nloss = timed_steps(nar_t, nbatch)
        Sn = Txt + Tyt - 1
        nfwd = Bt * (L * (Sn * 2 * (4 * d * d + 2 * d * F) + 4 * Sn * Sn * d))
        out['train_step_nar'] = {'batch': Bt, 'seq': Sn, 'ms_per_step': nms, 'clips_per_s': Bt / (nms * 1e-3), 'loss': float(nloss.detach()),
                                 'model_tflops': 3 * nfwd / (nms * 1e-3) / 1e12,
                                 'frac_of_bf16_sustained_peak': 3 * nfwd / (nms * 1e-3) / 1e12 / pk['bf16_tflops_sustained'],
                                 'ms_all_steps': nms_all,
                                 'note': 'ValleNAR.training_step forward + backward (full attention, AdaLN), stage drawn per step; median of 5 steps'}
        del nar_t
    except Exception as e:  # report, do not hide
        out['train_step'] = {'error': repr(e)[:300]}
    return out


# ------------------------------------------------------------------------------------------------------------
def _reference_rate(n_steps: int, warmup: int, B_ar: int, B_nar: int) -> tuple[dict, str]:
    """Executed reference (oracle/_ref) on the host cores; falls back to the oracle port when it is not installed."""
    from oracle import ref_runner
    torch.set_num_threads(os.cpu_count() or 1)
    if ref_runner.available():
        return ref_runner.tts_rate('cpu', B_ar, n_steps, B_nar, warmup=warmup), 'reference'
    return _port_rate(n_steps, B_ar), 'port'


def _port_rate(n_steps: int, B_ar: int) -> dict:
    """Oracle port (oracle/valle_oracle.py: the reference's algorithm restated) -- only used when oracle/_ref is absent."""
    from oracle import synth
    from oracle import valle_oracle as vo
    oc = synth.large_config('LayerNorm', num_beams=B_ar, max_audio_len=n_steps + 1)
    sd = synth.synth_state_dict(synth.ar_state_shapes(oc), 0)
    g = torch.Generator().manual_seed(100)
    pt, tt = torch.randint(0, 256, (TP,), generator=g), torch.randint(0, 256, (TT,), generator=g)
    pc = torch.randint(0, 1024, (TC, Q), generator=g)
    t0 = time.perf_counter()
    vo.ar_generate(sd, oc, pt, pc, tt, max_steps=1)
    t_pre = time.perf_counter() - t0
    t0 = time.perf_counter()
    vo.ar_generate(sd, oc, pt, pc, tt, max_steps=n_steps + 1)
    t_step = max(time.perf_counter() - t0 - t_pre, 1e-6) / n_steps
    ocn = synth.large_config('AdaptiveLayerNorm')
    sdn = synth.synth_state_dict(synth.nar_state_shapes(ocn), 1)
    fl = torch.randint(0, 1024, (N_NEW,), generator=g)
    t0 = time.perf_counter()
    vo.nar_generate(sdn, ocn, pt, pc, tt, fl)
    t_stage = (time.perf_counter() - t0) / 7
    per_utt = t_pre / B_ar + N_NEW * t_step / B_ar + 7 * t_stage
    return {'frames_per_s': N_NEW / per_utt, 'ar_tokens_per_s': B_ar / t_step, 'nar_stage_frames_per_s': N_NEW / t_stage,
            'prefill_s': t_pre, 'step_s_mean': t_step, 'timed_steps': n_steps, 'nar_stage_s': t_stage, 'B_ar': B_ar, 'B_nar': 1,
            'sample': f'oracle PORT (oracle/_ref not installed): ar_generate at num_beams={B_ar}, {n_steps} cached steps; nar_generate at batch 1'}


def cpu_baseline(budget_s: float):
    """The executed reference on this box's host cores, bounded sample (~budget_s seconds of CPU work)."""
    n_steps = max(4, min(24, int(budget_s / 3 / 0.2)))      # ~0.2 s per step at num_beams=32; prefill + NAR stage take the rest
    r, kind = _reference_rate(n_steps, 2, 32, 2)
    return {'value': r['frames_per_s'], 'unit': 'frames/s', 'cores': torch.get_num_threads(), 'kind': kind, 'sample': r['sample'],
            'ar_tokens_per_s': r['ar_tokens_per_s'], 'nar_stage_frames_per_s': r['nar_stage_frames_per_s'],
            'host_logical_cores': os.cpu_count()}


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return None
    K, W = args.steps, args.warmup
    # bounded sample: one prefill, then W + K cached AR steps (capped so that the arm ends within a few minutes), one NAR stage
    n = max(2, min(K + W, 400))
    r, kind = _reference_rate(n, min(W, n - 1), 32, 2)
    rate = r['frames_per_s']
    return {'impl': 'reference', 'metric': 'tts_codec_frames_per_s', 'value': rate, 'unit': 'frames/s', 'n_gpus': world,
            'steps': K, 'warmup': W, 'ms_per_step': args.utterances * N_NEW / rate * 1e3, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.utterances, world),
            'run': {'utterances_per_s': rate / N_NEW, 'audio_seconds_per_s': rate / 75.0, 'sample': r['sample']},
            'cpu_baseline': {'value': rate, 'unit': 'frames/s', 'cores': torch.get_num_threads(), 'kind': kind, 'sample': r['sample'],
                             'ar_tokens_per_s': r['ar_tokens_per_s'], 'nar_stage_frames_per_s': r['nar_stage_frames_per_s']},
            'e2e': {'value': rate, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--utterances', type=int, default=256)
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--nar-chunk', type=int, default=64)
    ap.add_argument('--nar-batch', type=int, default=64)
    ap.add_argument('--train-batch', type=int, default=16)
    ap.add_argument('--cpu-budget', type=float, default=15.0)
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
