"""Phase timeline of CTA 0 (first tile) of an attn_block launch + CUDA-event time per launch (dev tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
TRACE_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scattennet_b200", "libscatt_trace.so")
if "--build" in sys.argv:  # on the CPU box: the instrumented library travels with the snapshot
    from scattennet_b200 import build as B

    print(B.build(extra_flags=("-DSCATT_BLOCK_TRACE=1",), out=TRACE_LIB))
    sys.exit(0)
os.environ["SCATT_LIB"] = TRACE_LIB
import torch

from scattennet_b200 import _lib as L
from scattennet_b200 import functional as F_
from scattennet_b200 import synth
from scattennet_b200.functional import Act

dev = "cuda"
buf = torch.zeros(160, dtype=torch.int64, device=dev)
lib = L.load()
NAMES = {0: "setup done", 15: "P: first Wo tiles issued",
         10: "M: ctx kb0 landed", 11: "M: ctx kb1", 12: "M: ctx kb2", 13: "M: ctx kb3", 16: "P: ctx issued", 2: "M: out_proj issued",
         5: "E: out_proj complete", 6: "E: h written", 3: "M: h seen", 17: "P: all tiles issued", 4: "M: all MMAs issued",
         7: "E: fc2 complete", 8: "E: tile done", 18: "E: LN1 pass 1 done", 19: "E: LN1 stats combined",
         58: "E: staged (cluster)", 59: "E: landing free seen", 60: "E: peer partial added", 61: "E: LN2 stats combined", 62: "E: LN2 stores issued"}
for j in range(8):
    NAMES[20 + j] = f"M: fc1({j}) issue starts"
    NAMES[30 + j] = f"M: g({j}) seen"
    NAMES[40 + j] = f"E: fc1({j}) complete"
    NAMES[50 + j] = f"E: g({j}) written"


def lin(n, k):
    l = torch.nn.Linear(k, n)
    synth.load_synth_(l, 1)
    return l.to(dev)


D, Fh = 256, 768
for mode in [a for a in sys.argv[1:] if not a.startswith("--")] or ("fp16x3", "fp16x1"):
    prec = F_.get_precision(mode)
    for M, G in ((1600, 3),) if "--small" in sys.argv else ((1600, 3), (51200, 3)):
        g = torch.Generator().manual_seed(0)
        ctx = [Act(torch.randn(M, D, generator=g).to(dev)).with_planes(prec) for _ in range(G)]
        xs = [Act(torch.randn(M, D, generator=g).to(dev)).with_planes(prec) for _ in range(G)]
        pk = lambda n, k: [F_.PackedLinear([lin(n, k)], None, None) for _ in range(G)]
        po, p1, p2 = pk(D, D), pk(Fh, D), pk(D, Fh)
        ln1 = [torch.nn.LayerNorm(D).to(dev) for _ in range(G)]
        ln2 = [torch.nn.LayerNorm(D).to(dev) for _ in range(G)]
        run = lambda: F_.attn_block(prec, ctx, xs, po, ln1, p1, p2, ln2)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = 100.0 * e0.elapsed_time(e1)
        flops = 2.0 * G * M * (D * D + 2 * D * Fh)
        L.check(lib.scatt_debug_set_trace(buf.data_ptr()), "trace on")
        run()
        torch.cuda.synchronize()
        L.check(lib.scatt_debug_set_trace(None), "trace off")
        t = buf.cpu().tolist()
        t = [((v - t[0]) & 0xFFFFFFFF) if v else 0 for v in t]  # 32-bit stamps relative to "setup done"
        t[0] = 0
        stamped = lambda i: i == 0 or t[i] != 0
        print(f"--- {mode} M={M} x{G}: {us:.1f} us per launch, {flops / us * 1e-6:.1f} TFLOP/s algorithmic")
        for i in sorted((i for i in NAMES if stamped(i)), key=lambda i: t[i]):
            print(f"  {NAMES[i]:28s} +{t[i]:8d} cyc")
        print("  ring items: issued -> taken (latency)")
        print("  " + " ".join(f"{k}:{t[64 + k]}->{t[112 + k]}({t[112 + k] - t[64 + k]})" for k in range(48) if t[64 + k]))
        buf.zero_()
