"""Per-phase timing of the persistent decode chain kernel (csrc/decode_chain.cu) from in-kernel %globaltimer stamps.
    python tools/chain_timing.py [B]"""
import ctypes as C
import math
import sys

import torch

sys.path.insert(0, '.')
from valle2_b200 import _lib, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
d, F = 1024, 4096
dev, bf = 'cuda', torch.bfloat16
torch.manual_seed(0)
o = torch.randn(B, d, device=dev).to(bf)
x = torch.randn(B, d, device=dev)
wo, w1 = (torch.randn(d, d, device=dev) / 32).to(bf), (torch.randn(F, d, device=dev) / 32).to(bf)
w2, wq = (torch.randn(d, F, device=dev) / 64).to(bf), (torch.randn(3 * d, d, device=dev) / 32).to(bf)
bo, b1, b2 = torch.randn(d, device=dev), torch.randn(F, device=dev), torch.randn(d, device=dev)
g2, be2, g1, be1 = (torch.randn(d, device=dev) for _ in range(4))
ns = {k: ops.linear_decode_splits(n, kk, 32) for k, (n, kk) in {'qkv': (3 * d, d), 'o': (d, d), 'f1': (F, d), 'f2': (d, F)}.items()}
print('splits', ns)
h, h2, f = torch.zeros(B, d, device=dev, dtype=bf), torch.zeros(B, d, device=dev, dtype=bf), torch.zeros(B, F, device=dev, dtype=bf)
p_o, p_f1 = torch.zeros(ns['o'], B, d, device=dev), torch.zeros(ns['f1'], B, F, device=dev)
p_f2, p_q = torch.zeros(ns['f2'], B, d, device=dev), torch.zeros(ns['qkv'], B, 3 * d, device=dev)
gbar = torch.zeros(64, device=dev, dtype=torch.int32)
sm = ops.device_info()['sm_count']
dbg = torch.zeros(sm, 8, 4, device=dev, dtype=torch.int64)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def phases():
    return [ops.chain_gemm(o, wo, p_o, B * d),
            ops.chain_ln(x, g2, be2, h, part=p_o, n_part=ns['o'], part_stride=B * d, bias=bo),
            ops.chain_gemm(h, w1, p_f1, B * F),
            ops.chain_act(p_f1, ns['f1'], B * F, b1, f),
            ops.chain_gemm(f, w2, p_f2, B * d),
            ops.chain_ln(x, g1, be1, h2, part=p_f2, n_part=ns['f2'], part_stride=B * d, bias=b2),
            ops.chain_gemm(h2, wq, p_q, B * 3 * d)]


names = ['Wo', 'LN2', 'W1', 'GELU', 'W2', 'LN1', 'QKV']
lib = _lib.load()
for flush_l2 in (False, True):
    rows = []
    for it in range(12):
        if flush_l2:
            flush.fill_(it)
        _lib.check(lib.vb_decode_chain_set_debug(dbg.data_ptr()), 'dbg')
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.decode_chain(phases(), B, gbar)
        e1.record()
        torch.cuda.synchronize()
        t = dbg.cpu().numpy().astype('float64')          # [cta][phase][start, work done, arrived, released]
        t0 = t[:, 0, 0].min()
        rows.append((e0.elapsed_time(e1) * 1e3, t, t0))
    _lib.check(lib.vb_decode_chain_set_debug(None), 'dbg')
    us, t, t0 = rows[-1]
    print(f'--- flush_l2={flush_l2}: kernel {us:.1f} us (event), last-phase end - first start = {(t[:, 6, 1].max() - t0) / 1e3:.1f} us')
    for pi, nm in enumerate(names):
        st, wd, ar, rl = (t[:, pi, k] for k in range(4))
        line = f'{nm:5s} start {(st.min() - t0) / 1e3:6.2f}..{(st.max() - t0) / 1e3:6.2f}  work(min/med/max) {((wd - st).min()) / 1e3:5.2f}/{(sorted(wd - st)[len(st) // 2]) / 1e3:5.2f}/{((wd - st).max()) / 1e3:5.2f}'
        if pi < 6:
            line += f'  last arrive {(ar.max() - t0) / 1e3:6.2f}  release {(rl.min() - t0) / 1e3:6.2f}..{(rl.max() - t0) / 1e3:6.2f}  barrier cost {(rl.max() - ar.max()) / 1e3:5.2f}'
        print(line)
