"""Per-kernel time of one encoder step at a given batch (eager pass with CUDA-event brackets per C-ABI call)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scattennet_b200 import MSCAEncoder, synth, functional as F_
from scattennet_b200.config import model_config
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
prec = sys.argv[3] if len(sys.argv) > 3 else "fp16x3"
cfg = model_config("phoenix-2014t")
m = MSCAEncoder(cfg, 1120, precision=prec).eval(); synth.load_synth_(m, 0); m = m.cuda()
kp, mask = synth.synth_batch(B, T, seed=1); kp, mask = kp.cuda(), mask.cuda()
with torch.no_grad():
    for _ in range(2): m(kp, mask)
    torch.cuda.synchronize()
    with F_.profile_ops() as prof:
        for _ in range(3): m(kp, mask)
    agg = prof.summary()
tot = sum(v["ms"] for v in agg.values())
print(f"B={B} T={T} {prec}: sum of kernel brackets {tot/3:.3f} ms/step")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:32s} {v['calls']//3:3d} calls  {v['ms']/3:8.3f} ms  {100*v['ms']/tot:5.1f}%  {v['flops']/v['ms']/1e9 if v['ms'] else 0:8.1f} TFLOP/s  {v['bytes']/v['ms']/1e6 if v['ms'] else 0:8.0f} GB/s")
