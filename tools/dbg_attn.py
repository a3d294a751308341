import sys, os
print("start", flush=True)
import torch
print("torch imported", flush=True)
sys.path.insert(0, "/root/repo")
from scattennet_b200 import functional as F_, synth
B, T, kind = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
allpad = len(sys.argv) > 4 and sys.argv[4] == "1"
prec = F_.get_precision("fp16x3")
g = torch.Generator().manual_seed(7)
qkv = torch.randn(B * T, 768, generator=g).cuda()
planes = F_.split_planes(qkv, prec)
lengths = synth.parity_lengths(B, T)
mask = (torch.arange(T)[None] < torch.tensor(lengths)[:, None]).long()
if allpad and B >= 3:
    mask[2] = 0
print("planes ready", flush=True)
act = F_.stream_attention_planes(prec, [(planes, 0)], [(planes, 256)], [(planes, 512)], B, T, T, 16, kind, key_mask=F_.key_mask_u8(mask.cuda()))[0]
torch.cuda.synchronize()
print("ok", B, T, kind, allpad, float(act.planes[0].float().abs().max()))
