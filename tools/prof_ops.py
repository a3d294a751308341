"""Small driver for ncu captures of single ops (dev tool).

    python tools/prof_ops.py linear_ln|linear_qkv|attention [reps]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scattennet_b200 import functional as F_, _lib as L, synth
from scattennet_b200.functional import Act

which = sys.argv[1] if len(sys.argv) > 1 else "linear_ln"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mode = sys.argv[3] if len(sys.argv) > 3 else "fp16x3"
prec = F_.get_precision(mode)
dev = "cuda"
M = 1600
g = torch.Generator().manual_seed(0)

def lin(n, k):
    l = torch.nn.Linear(k, n); synth.load_synth_(l, 1); return l.to(dev)

if which.startswith("linear"):
    if which == "linear_ln":
        N, K, ep = 256, 256, F_.make_epilogue(residual_mode=L.RES_BEFORE_LN, layer_norm=True)
    elif which == "linear_fc2":
        N, K, ep = 256, 768, F_.make_epilogue(residual_mode=L.RES_BEFORE_LN, layer_norm=True)
    else:
        N, K, ep = 768, 256, F_.make_epilogue(scale_cols=256, scale=0.25)
    xs = [Act(torch.randn(M, K, generator=g).to(dev)).with_planes(prec) for _ in range(3)]
    packs = [F_.PackedLinear([lin(N, K)], None, None) for _ in range(3)]
    res = [torch.randn(M, N, generator=g).to(dev) for _ in range(3)]
    lns = [torch.nn.LayerNorm(N).to(dev) for _ in range(3)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(reps + 3):
        if i == 3: ev0.record()
        F_.linear(prec, xs, packs, ep, residuals=res, lns=lns, out_planes=(which != "linear_qkv"))
    ev1.record(); torch.cuda.synchronize()
    print(which, mode, "avg us per launch", 1e3 * ev0.elapsed_time(ev1) / reps)
elif which == "attention_fa":
    B, T, H, D = 8, 200, 16, 256
    qkv = [F_.split_planes(torch.randn(B * T, 3 * D, generator=g).to(dev), prec) for _ in range(3)]
    km = torch.ones(B, T, dtype=torch.uint8, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(reps + 3):
        if i == 3: ev0.record()
        F_.stream_attention_planes(prec, [(t, 0) for t in qkv], [(t, D) for t in qkv], [(t, 2 * D) for t in qkv], B, T, T, H, 0, key_mask=km)
    ev1.record(); torch.cuda.synchronize()
    print(which, "avg us per launch", 1e3 * ev0.elapsed_time(ev1) / reps)
else:
    B, T, H, D = 8, 200, 16, 256
    qkv = [torch.randn(B * T, 3 * D, generator=g).to(dev) for _ in range(3)]
    km = torch.ones(B, T, dtype=torch.uint8, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(reps + 3):
        if i == 3: ev0.record()
        F_.stream_attention(prec, [t[:, :D] for t in qkv], [t[:, D:2*D] for t in qkv], [t[:, 2*D:] for t in qkv], B, T, T, H, 0, key_mask=km)
    ev1.record(); torch.cuda.synchronize()
    print(which, "avg us per launch", 1e3 * ev0.elapsed_time(ev1) / reps)
