"""Time the BiLSTM alignment head and its recurrent kernel (CUDA events, after warm-up)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scattennet_b200 as S
from scattennet_b200 import functional as F_
from scattennet_b200.alignment_module import alignment_forward
from scattennet_b200 import synth

dev = "cuda"
prec = F_.get_precision("fp16x3")
m = S.AlignmentModule(cls_num=1120, input_size=1024, hidden_size=1024).eval()
synth.load_synth_(m, seed=1)
m = m.to(dev)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for B, T in ((8, 50), (8, 200), (16, 50), (64, 50), (256, 50)):
    gates = torch.randn(B * T, 4096, device=dev)
    w_hh = torch.randn(2, 2048, 512, device=dev) * 0.03
    us = timed(lambda: F_.lstm_bidir(prec, gates, w_hh, B, T, 512, out_f32=False))
    x = F_.Act.from_f32(torch.randn(B * T, 1024, device=dev)).with_planes(prec)
    with torch.no_grad():
        us_head = timed(lambda: alignment_forward(prec, m, x, B, T, clamp=50.0))
    print(f"B={B} T={T}: lstm_bidir {us:.1f} us ({us / T:.2f} us/step), alignment head {us_head:.1f} us")
