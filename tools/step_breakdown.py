"""Per-kernel time of one eager encoder step at a given batch (dev tool): python tools/step_breakdown.py [B] [T] [precision]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scattennet_b200 import MSCAEncoder, synth, functional as F_
from scattennet_b200.config import model_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
mode = sys.argv[3] if len(sys.argv) > 3 else "fp16x3"
cfg = model_config("phoenix-2014t")
m = MSCAEncoder(cfg, 1120, precision=mode, use_graph=False).eval()
synth.load_synth_(m, 0)
m = m.cuda()
kp, mask = synth.synth_batch(B, T, seed=1)
kp, mask = kp.cuda(), mask.cuda()
with torch.no_grad():
    for _ in range(2):
        m(kp, mask)
    torch.cuda.synchronize()
    with F_.profile_ops() as prof:
        m(kp, mask)
    agg = prof.summary()
tot = sum(a["ms"] for a in agg.values())
print(f"B={B} T={T} {mode}: sum of event-bracketed launches {tot:.3f} ms")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{k:46s} x{a['calls']:3d} {a['ms']*1e3:9.1f} us {100*a['ms']/tot:5.1f}%  {a['flops']/a['ms']/1e9 if a['ms'] else 0:7.1f} TFLOP/s {a['bytes']/a['ms']/1e6 if a['ms'] else 0:7.0f} GB/s")
