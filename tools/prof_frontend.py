"""Runs the K1 front end alone at a given batch (dev tool for `ncu -k regex:frontend`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scattennet_b200 import MSCAEncoder, synth, functional as F_
from scattennet_b200.config import model_config
from scattennet_b200.keypoint_module import frontend_forward

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = 200
cfg = model_config("phoenix-2014t")
prec = F_.get_precision("fp16x3")
model = MSCAEncoder(cfg, 1120, precision="fp16x3").eval()
synth.load_synth_(model, 0)
model = model.cuda()
mods = [model.body_encoder, model.left_encoder, model.right_encoder]
kp, _ = synth.synth_batch(B, T, seed=1)
kp = kp.cuda()
idx = model._joint_idx(torch.device("cuda"))
for _ in range(4):
    out = frontend_forward(prec, mods, kp, idx, B, T)
torch.cuda.synchronize()
print("ok", out[0][0].planes.shape)
