"""Ablation timing of the tensor-core front end (dev tool): build with
   python -c "from scattennet_b200 import build as B; B.build(extra_flags=('-DSCATT_FT_ABLATE=1',), out='scattennet_b200/libscatt_b.so')"
and run under gpurun; SCATT_FT_DBG bits: 1 no position loads, 2 no gather loads, 4 no stores."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SCATT_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scattennet_b200", "libscatt_b.so")
import torch
from scattennet_b200 import MSCAEncoder, synth, functional as F_
from scattennet_b200.config import model_config
from scattennet_b200.keypoint_module import frontend_forward

T = 200
cfg = model_config("phoenix-2014t")
prec = F_.get_precision("fp16x3")
model = MSCAEncoder(cfg, 1120, precision="fp16x3").eval()
synth.load_synth_(model, 0)
model = model.cuda()
mods = [model.body_encoder, model.left_encoder, model.right_encoder]
idx = model._joint_idx(torch.device("cuda"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B in (8, 256):
    kp, _ = synth.synth_batch(B, T, seed=1)
    kp = kp.cuda()
    for dbg in (0, 1, 2, 4, 3, 7):
        os.environ["SCATT_FT_DBG"] = str(dbg)
        ts = []
        for _ in range(6):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            frontend_forward(prec, mods, kp, idx, B, T)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        print(f"B={B} ablate={dbg}: {sorted(ts)[len(ts)//2]:.1f} us", flush=True)
