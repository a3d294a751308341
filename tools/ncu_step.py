"""One eager encoder step (with the BiLSTM alignment head) between cudaProfilerStart/Stop, for
`ncu --set full --profile-from-start off` (dev tool; the numbers it prints under ncu are not bench values)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scattennet_b200 import MSCAEncoder, synth
from scattennet_b200.config import model_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
cfg = model_config("phoenix-2014t")
m = MSCAEncoder(cfg, 1120, precision="fp16x3", alignment=True).eval()
synth.load_synth_(m, 0)
m = m.cuda()
kp, mask = synth.synth_batch(B, T, seed=1)
kp, mask = kp.cuda(), mask.cuda()
with torch.no_grad():
    for _ in range(2):
        m(kp, mask)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = m(kp, mask)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok", {k: tuple(v.shape) for k, v in out.items()})
