"""Cycle breakdown of CTA 0 of the CTC beam-search kernel (dev tool): search warp phases and one producer warp."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scattennet_b200 import functional as F_, _lib as L
dev = "cuda"
buf = torch.zeros(16, dtype=torch.int64, device=dev)
lib = L.load()
B, T, V = 8, 50, 1120
for beam in (5, 16):
    for name, scale in (("flat", 0.05), ("normal", 1.0)):
        x = (torch.randn(B, T, V, generator=torch.Generator().manual_seed(1)) * scale).to(dev)
        lens = torch.full((B,), T, dtype=torch.int32, device=dev)
        for _ in range(3): F_.ctc_beam_decode(x, lens, beam)
        torch.cuda.synchronize()
        L.check(lib.scatt_debug_set_trace(buf.data_ptr()), "on"); F_.ctc_beam_decode(x, lens, beam); torch.cuda.synchronize()
        L.check(lib.scatt_debug_set_trace(None), "off")
        t = buf.cpu().tolist(); buf.zero_()
        n = max(t[6], 1)
        print(f"--- beam {beam}, {name} logits: search loop {t[0]} cyc = {t[0] / n:.0f} per frame: wait {t[1] / n:.0f}, advance {t[2] / n:.0f}, "
              f"score {t[3] / n:.0f}, select {t[4] / n:.0f}, rebuild {t[5] / n:.0f}; producer warp 1: {t[8] / max(t[9], 1):.0f} cyc per frame ({t[9]} frames)")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): F_.ctc_beam_decode(x, lens, beam)
        e1.record(); torch.cuda.synchronize()
        print(f"    kernel {1e3 * e0.elapsed_time(e1) / 10:.1f} us")
