"""Small encoder forward for compute-sanitizer (one tool per gpurun call):

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

Eager launches (no CUDA graph), B = 2, T = 24 with a ragged mask; the fused layer tail is forced on so that the
cluster / DSMEM / TMEM-operand kernel is covered, and both attention kernels run (SCATT_ATTN_PERSIST selects)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scattennet_b200 import MSCAEncoder, synth
from scattennet_b200 import functional as F_
from scattennet_b200.config import VOCAB_STUB, model_config

F_.FUSED_BLOCK_MIN_TILES = 0
cfg = model_config("phoenix-2014t")
model = MSCAEncoder(cfg, VOCAB_STUB, precision=os.environ.get("SCATT_PRECISION", "fp16x3")).eval()
synth.load_synth_(model, 0)
model = model.cuda()
kp, mask = synth.synth_batch(2, 24, seed=1, lengths=[24, 17])
with torch.no_grad():
    out = model(kp.cuda(), mask.cuda())
    torch.cuda.synchronize()
print("sanitize_smoke ok:", {k: tuple(v.shape) for k, v in out.items()}, "finite:", all(bool(torch.isfinite(v).all()) for v in out.values()))
