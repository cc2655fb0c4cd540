"""TEST / BASELINE INFRASTRUCTURE ONLY -- time the EXECUTED reference (unmodified `valle` package from /root/reference or
oracle/_ref, imported through oracle/ref_shims.py) on the workload of bench.py.  Used by `bench.py --impl reference`, by its
`cpu_baseline` leg (host cores) and by `extras.gpu_eager_reference` (the same torch-eager modules on the B200).

AR (valle/models/valle_ar.py:92-180, unmodified, through its public `ValleAR.generate`): one utterance with
`num_beams = B` -- the reference has no other way to batch; its beams are independent greedy/sampled replicas,
so B beams cost exactly what B different utterances would.  Per-step times come from a forward-pre-hook on
`model.transformer` (hooks observe, they do not change the code that runs): call 0 is the prefill over text + prompt, call
i >= 1 the KV-cached step i (embedding of ALL positions, `torch.cat` cache growth, top-k filter and `multinomial` included).
NAR (valle_nar.py:107-165): `ValleNAR.generate` raises upstream, so one stage is timed through the
reference's own sub-modules (embeddings, PositionalEncoding, Transformer with AdaptiveLayerNorm, proj_layers) driven by the
repaired stage loop -- the loop of oracle/make_golden.py:golden_nar, batched.
"""
from __future__ import annotations

import os
import tempfile
import time

import torch

from . import ref_shims, synth
from .valle_oracle import OracleConfig

TP, TT, TC, N_NEW = 50, 100, 225, 750       # prompt phonemes, target phonemes, prompt frames, generated frames


def available() -> bool:
    return ref_shims.reference_available()


def _cfg(valle, oc: OracleConfig, tmp: str):
    kw = {k: getattr(oc, k) for k in OracleConfig.__dataclass_fields__}
    return valle.config.ConfigValle(dropout=0.0, ckpt_path=os.path.join(tmp, 'c'), log_path=os.path.join(tmp, 'l'), **kw)


def _sync(device):
    if torch.device(device).type == 'cuda':
        torch.cuda.synchronize()


def time_ar(device: str, B: int, n_steps: int, seed: int = 0) -> dict:
    """`ValleAR.generate` of the executed reference: prefill + n_steps KV-cached greedy steps at num_beams = B."""
    valle = ref_shims.import_reference()
    try:
        tmp = tempfile.mkdtemp(prefix='valle_ref_run_')
        oc = synth.large_config('LayerNorm', num_beams=B, max_audio_len=n_steps + 1, top_k=1)
        torch.manual_seed(seed)
        model = valle.models.ValleAR(_cfg(valle, oc, tmp)).eval().to(device)
        g = torch.Generator().manual_seed(100 + seed)
        pt, tt = torch.randint(0, 256, (TP,), generator=g).to(device), torch.randint(0, 256, (TT,), generator=g).to(device)
        pc = torch.randint(0, 1024, (TC, 8), generator=g).to(device)
        stamps = []

        def hook(_m, _a, _k=None):
            _sync(device)
            stamps.append(time.perf_counter())

        h = model.transformer.register_forward_pre_hook(hook)
        with torch.no_grad():
            model.generate(pt, pc, tt)
        _sync(device)
        t_end = time.perf_counter()
        h.remove()
    finally:
        ref_shims.release_reference()
    stamps.append(t_end)
    d = [stamps[i + 1] - stamps[i] for i in range(len(stamps) - 1)]
    steps = d[1:]                                # d[0] = prefill (+ first sample)
    assert len(steps) >= 1, 'the reference stopped before the first cached step'
    return {'prefill_s': d[0], 'step_s': steps, 'n_steps': len(steps), 'B': B, 'ctx0': TP + TT + TC + 1}


def time_nar_stage(device: str, B: int, seed: int = 1) -> float:
    """Seconds for ONE stage (the first: target embeddings of codebook 1) of the repaired NAR loop on the reference's own
    modules at batch B, S = 150 + 225 + 750."""
    valle = ref_shims.import_reference()
    try:
        tmp = tempfile.mkdtemp(prefix='valle_ref_run_')
        oc = synth.large_config('AdaptiveLayerNorm')
        torch.manual_seed(seed)
        model = valle.models.ValleNAR(_cfg(valle, oc, tmp)).eval().to(device)
        g = torch.Generator().manual_seed(200 + seed)
        tokens = torch.randint(0, 256, (B, TP + TT), generator=g).to(device)
        pc = torch.randint(0, 1024, (B, TC, 8), generator=g).to(device)
        fl = torch.randint(0, 1024, (B, N_NEW), generator=g).to(device)
        with torch.no_grad():
            _sync(device)
            t0 = time.perf_counter()
            emb_prompt = model.codes_embs[0](pc[..., 0])
            for j in range(1, 8):
                emb_prompt = emb_prompt + model.codes_embs[j](pc[..., j])
            x_tok = model.tokens_position_emb(model.tokens_emb(tokens))
            emb_out = model.codes_embs[0](fl)
            codes = model.audio_position_emb(torch.cat([emb_prompt, emb_out], dim=1))
            hdn, _ = model.transformer(torch.cat([x_tok, codes], dim=1), embedding=model.stage_embs[0].weight)
            logits = model.proj_layers[0](hdn[:, TP + TT + TC:])
            _ = logits.argmax(-1)
            _sync(device)
            dt = time.perf_counter() - t0
    finally:
        ref_shims.release_reference()
    return dt


def tts_rate(device: str, B_ar: int, n_steps: int, B_nar: int, warmup: int = 0) -> dict:
    """Frames/s of the full TTS job (AR prefill + 750 cached steps + 7 NAR stages) extrapolated from a bounded sample of the
    executed reference: prefill once, n_steps cached steps (the first `warmup` dropped), one NAR stage."""
    ar = time_ar(device, B_ar, n_steps)
    steps = ar['step_s'][warmup:] or ar['step_s']
    t_step = sum(steps) / len(steps)
    t_stage = time_nar_stage(device, B_nar)
    per_utt = ar['prefill_s'] / B_ar + N_NEW * t_step / B_ar + 7 * t_stage / B_nar
    return {'frames_per_s': N_NEW / per_utt, 'ar_tokens_per_s': B_ar / t_step, 'nar_stage_frames_per_s': B_nar * N_NEW / t_stage,
            'prefill_s': ar['prefill_s'], 'step_s_mean': t_step, 'timed_steps': len(steps), 'nar_stage_s': t_stage,
            'B_ar': B_ar, 'B_nar': B_nar,
            'sample': (f'executed reference ({device}): ValleAR.generate at num_beams={B_ar} (prefill of {ar["ctx0"]} positions + '
                       f'{len(steps)} timed KV-cached steps from ctx {ar["ctx0"]}), one NAR stage at batch {B_nar}, S={TP + TT + TC + N_NEW}; '
                       f'utterance time = prefill/{B_ar} + 750 x step/{B_ar} + 7 x stage/{B_nar}')}
