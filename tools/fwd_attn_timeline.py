"""Per-key-block cycle timeline of the tcgen05 forward attention (one row thread per CTA), NAR shape B=64 S=900 H=16.
    python tools/fwd_attn_timeline.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from valle2_b200 import _lib, ops  # noqa: E402

B, S, H, d = 64, 900, 16, 1024
lib = _lib.load()
qkv = torch.randn(B * S, 3 * d, device='cuda').bfloat16()
o = torch.empty(B * S, d, device='cuda', dtype=torch.bfloat16)
n_cta = ((S + 127) // 128) * H * B
dbg = torch.zeros(n_cta, 32, 8, device='cuda', dtype=torch.int64)
for _ in range(2):
    ops.attention_packed(qkv, o, B, S, H, mask_mode=ops.MASK_NONE, x_lens=None, kv_lens=None, use_tc=True)
_lib.check(lib.vb_attention_prefill_set_debug(dbg.data_ptr()), 'dbg')
ops.attention_packed(qkv, o, B, S, H, mask_mode=ops.MASK_NONE, x_lens=None, kv_lens=None, use_tc=True)
_lib.check(lib.vb_attention_prefill_set_debug(None), 'dbg')
torch.cuda.synchronize()
t = dbg.cpu().numpy().astype(np.float64)
nb = (S + 63) // 64
thread = int(os.environ.get('VALLE_B200_FWD_DBG_THREAD', '64'))
hh = (thread // 32 - 2) // 4            # which half stamped: 0 = even key blocks, 1 = odd
extra = t[:, 31]
t = t[:, :nb]
mine = np.arange(hh, nb, 2)             # the stamping thread's blocks
keep = (t[:, mine, 3] > 0).all(axis=1) & (t[:, :, 7] > 0).all(axis=1)
full = t[keep]
extra = extra[keep]
row = full[:, mine]                      # [cta][own block][slot]
print(f'{len(full)} CTAs with {nb} key blocks; row thread {thread} (half {hh}: blocks {hh}, {hh + 2}, ...); cycles, median over CTAs and its blocks 2..:')
blk = row[:, 1:]
prev = row[:, :-1, 3]
print('  previous P written -> S ready                          %6.0f' % np.median(blk[:, :, 0] - prev))
print('  S ready -> maximum known AND partner exponentials done %6.0f' % np.median(blk[:, :, 1] - blk[:, :, 0]))
print('  -> exp2 / pack done                                    %6.0f' % np.median(blk[:, :, 2] - blk[:, :, 1]))
print('  -> P written (tile free + stores + fence + arrive)     %6.0f' % np.median(blk[:, :, 3] - blk[:, :, 2]))
print('  period of this half (two key blocks)                   %6.0f' % np.median(blk[:, :, 3] - prev))
print(' MMA warp (block t):')
pw = full[:, mine, 3]
print('  P_t written (this row thread) -> P_t seen by the MMA warp %6.0f' % np.median(full[:, mine, 6][:, 1:] - pw[:, 1:]))
print('  P_t seen -> P V_t issued + committed                      %6.0f' % np.median(full[:, 2:, 7] - full[:, 2:, 6]))
print('  P V_{t-1} issued -> Q K_{t+2} operands + buffer ready     %6.0f' % np.median(full[:, 3:, 4] - full[:, :-3, 7]))
print('  -> Q K_{t+2} issued + committed                           %6.0f' % np.median(full[:, 3:, 5] - full[:, 3:, 4]))
print('  MMA warp period per block (P V_t issued -> P V_{t+1})     %6.0f' % np.median(full[:, 2:, 7] - full[:, 1:-1, 7]))
print('  Q K_t issued -> S_t first read by its row thread          %6.0f' % np.median(row[:, 1:, 0] - full[:, mine, 5][:, 1:]))
print('  row thread finished block t-2 -> Q K_t issued             %6.0f' % np.median(full[:, mine, 5][:, 1:] - row[:, :-1, 3]))
print('  Q K_{t-1} issued -> Q K_t issued                          %6.0f' % np.median(full[:, 2:, 5] - full[:, 1:-1, 5]))
print('  CTA lifetime (first S ready -> last P written) median %.0f cycles' % np.median(row[:, -1, 3] - row[:, 0, 0]))
print('  row-thread entry -> first S ready   %6.0f' % np.median(row[:, 0, 0] - extra[:, 0]))
print('  last P written -> last PV done      %6.0f' % np.median(extra[:, 1] - row[:, -1, 3]))
print('  last PV done -> output stored       %6.0f' % np.median(extra[:, 2] - extra[:, 1]))
print('  row-thread entry -> output stored   %6.0f cycles (%.1f us at 1.965 GHz)' % (np.median(extra[:, 2] - extra[:, 0]), np.median(extra[:, 2] - extra[:, 0]) / 1965))
g = dbg.cpu().numpy()[:, 30].astype(np.int64)
start, smid, end = g[:, 0], g[:, 1], g[:, 2]
ok = start > 0
t_all0, t_all1 = start[ok].min(), end[ok].max()
print('  kernel span by %%globaltimer: %.1f us; per-CTA entry -> output stored median %.1f us' % ((t_all1 - t_all0) / 1e3, np.median((end - start)[ok]) / 1e3))
busy = []
for sm in np.unique(smid[ok]):
    m = ok & (smid == sm)
    busy.append(((end - start)[m].sum() / 1e3, m.sum()))
busy = np.array(busy)
print('  per SM: CTAs %.1f, sum of CTA lifetimes %.1f us = %.2f concurrent CTAs over the kernel span' % (busy[:, 1].mean(), busy[:, 0].mean(), busy[:, 0].mean() / ((t_all1 - t_all0) / 1e3)))
