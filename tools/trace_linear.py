"""Phase timeline of CTA 0 of a tcgen05 linear launch (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scattennet_b200 import functional as F_, _lib as L, synth
from scattennet_b200.functional import Act

dev = "cuda"
buf = torch.zeros(16, dtype=torch.int64, device=dev)
lib = L.load()
NAMES = ["start", "setup done", "1st TMA issued", "1st stage landed", "last MMA committed", "acc ready (epi)", "LN stats combined", "epilogue done", "all done", "acc pre-init done"]
g = torch.Generator().manual_seed(0)
def lin(n, k):
    l = torch.nn.Linear(k, n); synth.load_synth_(l, 1); return l.to(dev)
for mode in ("fp16x3",):
    prec = F_.get_precision(mode)
    for name, (M, N, K, kw) in {
        "out_proj+res+LN": (1600, 256, 256, dict(residual_mode=L.RES_BEFORE_LN, layer_norm=True)),
        "fc2+res+LN": (1600, 256, 768, dict(residual_mode=L.RES_BEFORE_LN, layer_norm=True)),
        "qkv": (1600, 768, 256, dict(scale_cols=256, scale=0.25)),
        "fc1+gelu": (1600, 768, 256, dict(act_pre=L.ACT_GELU)),
    }.items():
        xs = [Act(torch.randn(M, K, generator=g).to(dev)).with_planes(prec) for _ in range(3)]
        packs = [F_.PackedLinear([lin(N, K)], None, None) for _ in range(3)]
        res = [torch.randn(M, N, generator=g).to(dev) for _ in range(3)]
        lns = [torch.nn.LayerNorm(N).to(dev) for _ in range(3)]
        ep = F_.make_epilogue(**kw)
        for _ in range(3):
            F_.linear(prec, xs, packs, ep, residuals=res, lns=lns)
        torch.cuda.synchronize()
        L.check(lib.scatt_debug_set_trace(buf.data_ptr()), "trace on")
        F_.linear(prec, xs, packs, ep, residuals=res, lns=lns)
        torch.cuda.synchronize()
        L.check(lib.scatt_debug_set_trace(None), "trace off")
        t = buf.cpu().tolist()
        print(f"--- {mode} {name} M={M} N={N} K={K}")
        for i, nm in enumerate(NAMES):
            if t[i]:
                print(f"  {nm:22s} +{t[i]-t[0]:8d} cyc")
        buf.zero_()
