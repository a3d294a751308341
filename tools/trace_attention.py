"""Phase timeline of CTA 0 of the tcgen05 attention kernel (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scattennet_b200 import functional as F_, _lib as L
dev = "cuda"
buf = torch.zeros(16, dtype=torch.int64, device=dev)
lib = L.load()
NAMES = ["start", "setup done", "TMA issued", "Q/K landed", "S ready", "row max done", "P written", "O ready", "row stored", "all done"]
B, T, H, D = (int(sys.argv[1]) if len(sys.argv) > 1 else 8), 200, 16, 256
g = torch.Generator().manual_seed(0)
qkv = [torch.randn(B * T, 3 * D, generator=g).to(dev) for _ in range(3)]
km = torch.ones(B, T, dtype=torch.uint8, device=dev)
for mode in ("fp16x3", "fp16x1"):
    prec = F_.get_precision(mode)
    for kind in (0, 1):
        pl = [F_.split_planes(t, prec) for t in qkv]
        run = lambda: F_.stream_attention_planes(prec, [(t, 0) for t in pl], [(t, D) for t in pl], [(t, 2 * D) for t in pl], B, T, T, H, kind, key_mask=km)
        for _ in range(3): run()
        torch.cuda.synchronize()
        L.check(lib.scatt_debug_set_trace(buf.data_ptr()), "on"); run(); torch.cuda.synchronize(); L.check(lib.scatt_debug_set_trace(None), "off")
        t = buf.cpu().tolist(); buf.zero_()
        print(f"--- {mode} kind={kind}")
        for i, nm in enumerate(NAMES):
            if t[i]: print(f"  {nm:18s} +{t[i]-t[0]:8d} cyc")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            for _ in range(20): run()
        gph.replay(); torch.cuda.synchronize()
        e0.record(); gph.replay(); e1.record(); torch.cuda.synchronize()
        print(f"  graph-replayed launch: {1e3 * e0.elapsed_time(e1) / 20:.1f} us")
