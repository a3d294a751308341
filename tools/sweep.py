"""Batch sweep (BASELINE config 4, one GPU) and isolated kernel sweep (config 5) - prints markdown tables.

    python tools/sweep.py batch   [--precision fp16x3]     # frames/s vs batch at T=200
    python tools/sweep.py kernels                          # attention / linear launches vs roofline
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scattennet_b200 import MSCAEncoder, synth, functional as F_, _lib as L
from scattennet_b200.config import model_config
from scattennet_b200.functional import Act
from oracle.scatt_oracle import encoder_flops_per_frame  # flop model only (no oracle arithmetic here)

PEAK_TF, PEAK_GBS = 1658.0, 6549.4
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    d = json.load(open(p)); PEAK_TF, PEAK_GBS = d["bf16_tflops"], d["hbm_gbs"]
dev = "cuda"
flush = None

def timed(fn, reps):
    global flush
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for e0, e1 in ev:
        flush.zero_(); e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    ts = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
    return ts[len(ts) // 2]

def batch_sweep(args):
    cfg = model_config(args.config, **({"max_position_embeddings": 512} if args.T > 256 else {}))
    print(f"| batch | T | precision | ms/step | frames/s | model TFLOP/s (algorithmic) | % of bf16 peak ({PEAK_TF:.0f}) |")
    print("|---|---|---|---|---|---|---|")
    for prec in args.precision.split(","):
        model = MSCAEncoder(cfg, 1120, precision=prec, use_graph=True).eval()
        synth.load_synth_(model, 0)
        model = model.to(dev)
        for b in [int(x) for x in args.batches.split(",")]:
            kp, mask = synth.synth_batch(b, args.T, seed=1)
            kp, mask = kp.to(dev), mask.to(dev)
            with torch.no_grad():
                ms = timed(lambda: model(kp, mask), 10 if b >= 256 else 30)
            fl = encoder_flops_per_frame(cfg, args.T, 1120) * b * args.T
            print(f"| {b} | {args.T} | {prec} | {ms:.3f} | {b * args.T / ms * 1e3:,.0f} | {fl / ms / 1e9:.1f} | {100 * fl / ms / 1e9 / PEAK_TF:.2f} |", flush=True)
            model._graphs.clear(); torch.cuda.empty_cache()

def graph_time(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us per launch (back to back, L2-warm)

def kernel_sweep(args):
    H, D = 16, 256
    gen = torch.Generator().manual_seed(0)
    print(f"### stream attention, B*3 streams grouped, H=16, hd=16 (us per grouped launch, back-to-back in a CUDA graph)\n")
    print("| T | B | kind | engine | us | TFLOP/s (4BT^2D, causal half) | % bf16 peak | exp/s (1e12) |")
    print("|---|---|---|---|---|---|---|---|")
    for T in (() if args.only == "linear" else tuple(int(v) for v in args.attn_T.split(","))):
        B = max(1, 12800 // T)
        qkv = [torch.randn(B * T, 3 * D, generator=gen).to(dev) for _ in range(3)]
        km = torch.ones(B, T, dtype=torch.uint8, device=dev)
        for kind, kn in ((0, "self"), (1, "causal"), (2, "cross")):
            for mode in ("fp16x3", "fp16x1", "fp32"):
                prec = F_.get_precision(mode)
                if prec.uses_planes and T <= F_.ATTN_PLANES_MAX_T:
                    pl = [F_.split_planes(t, prec) for t in qkv]
                    fn = lambda: F_.stream_attention_planes(prec, [(t, 0) for t in pl], [(t, D) for t in pl], [(t, 2 * D) for t in pl], B, T, T, H, kind, key_mask=km)
                else:
                    fn = lambda: F_.stream_attention(prec, [t[:, :D] for t in qkv], [t[:, D:2*D] for t in qkv], [t[:, 2*D:] for t in qkv], B, T, T, H, kind, key_mask=km)
                us = graph_time(fn)
                fl = 3 * (2.0 * B * T * (T + 1) * D if kind == 1 else 4.0 * B * T * T * D)
                ex = 3 * B * H * (T * (T + 1) / 2 if kind == 1 else T * T)
                eng = "tcgen05 (TMA-fed)" if (prec.uses_planes and T <= F_.ATTN_PLANES_MAX_T) else "cuda-core fp32"
                print(f"| {T} | {B} | {kn} | {mode} {eng} | {us:.1f} | {fl / us / 1e6:.1f} | {100 * fl / us / 1e6 / PEAK_TF:.2f} | {ex / us / 1e6:.2f} |", flush=True)
    if args.only == "attention":
        return
    print(f"\n### linear (tcgen05), 3 streams grouped (us per launch, back-to-back in a CUDA graph)\n")
    print("| M | N | K | epilogue | mode | us | TFLOP/s (2MNK) | % bf16 peak | GB/s (operands+outputs) |")
    print("|---|---|---|---|---|---|---|---|---|")
    def lin(n, k):
        l = torch.nn.Linear(k, n); synth.load_synth_(l, 1); return l.to(dev)
    for M in (1600, 12800, 51200):
        for (N, K, name, kw, of32, opl) in ((768, 256, "qkv (fp32 out)", dict(scale_cols=256, scale=0.25), True, False),
                                            (768, 256, "fc1+gelu (planes out)", dict(act_pre=L.ACT_GELU), False, True),
                                            (256, 256, "out_proj+res+LN", dict(residual_mode=L.RES_BEFORE_LN, layer_norm=True), True, True),
                                            (256, 768, "fc2+res+LN", dict(residual_mode=L.RES_BEFORE_LN, layer_norm=True), True, True)):
            for mode in ("fp16x3", "fp16x1"):
                prec = F_.get_precision(mode)
                xs = [Act(torch.randn(M, K, generator=gen).to(dev)).with_planes(prec) for _ in range(3)]
                packs = [F_.PackedLinear([lin(N, K)], None, None) for _ in range(3)]
                res = [torch.randn(M, N, generator=gen).to(dev) for _ in range(3)]
                lns = [torch.nn.LayerNorm(N).to(dev) for _ in range(3)]
                ep = F_.make_epilogue(**kw)
                fn = lambda: F_.linear(prec, xs, packs, ep, residuals=res if kw.get("residual_mode") else None, lns=lns if kw.get("layer_norm") else None, out_f32=of32, out_planes=opl)
                us = graph_time(fn, reps=10)
                fl = 3 * 2.0 * M * N * K
                planes = 2 if prec.terms >= 2 else 1
                by = 3 * ((M * K + N * K) * 2 * planes + (M * N * 4 if kw.get("residual_mode") else 0) + M * N * (4 * of32 + 4 * opl))
                print(f"| {M} | {N} | {K} | {name} | {mode} | {us:.1f} | {fl / us / 1e6:.1f} | {100 * fl / us / 1e6 / PEAK_TF:.2f} | {by / us / 1e3:.0f} |", flush=True)
                del xs, packs, res

def fusion_sweep(args):
    """K5 fusion attention (single head, D = 1024, no scaling): fp32 CUDA-core kernel vs the tcgen05 kernel, T' = the pooled
    length the encoder hands to the fusion block (T / 4 for phoenix-2014t, T / 2 for phoenix-2014)."""
    D = 1024
    gen = torch.Generator().manual_seed(0)
    scheds = [("", "")] + ([tuple(x.split(":")) for x in args.schedules.split(",")] if args.schedules else [])
    print("### fusion attention softmax(q k^T) v, D = 1024 (us per launch, back-to-back in a CUDA graph)\n")
    print("| T' | B | engine | schedule (ncols:spc) | us | TFLOP/s (4 B T'^2 D) | % bf16 peak |")
    print("|---|---|---|---|---|---|---|")
    for T, B in ((50, 8), (100, 8), (200, 8), (256, 8), (16, 256), (50, 256), (64, 200), (100, 128), (128, 100), (200, 64), (256, 50)):
        q, k, v = (torch.randn(B * T, D, generator=gen).abs().mul_(0.6).to(dev) for _ in range(3))
        fl = 4.0 * B * T * T * D
        p32 = F_.get_precision("fp32")
        us = graph_time(lambda: F_.fusion_attention(p32, q, k, v, B, T))
        print(f"| {T} | {B} | fp32 cuda-core | - | {us:.1f} | {fl / us / 1e6:.1f} | {100 * fl / us / 1e6 / PEAK_TF:.2f} |", flush=True)
        for mode in ("fp16x3", "fp16x1"):
            prec = F_.get_precision(mode)
            pl = [F_.split_planes(t, prec) for t in (q, k, v)]
            for nc, spc in scheds:
                if mode == "fp16x1" and nc:
                    continue
                for kk, vv in (("SCATT_FUSION_NCOLS", nc), ("SCATT_FUSION_SPC", spc)):
                    if vv: os.environ[kk] = vv
                    else: os.environ.pop(kk, None)
                us = graph_time(lambda: F_.fusion_attention_planes(prec, pl[0], pl[1], pl[2], B, T))
                print(f"| {T} | {B} | {mode} tcgen05 | {nc + ':' + spc if nc else 'default'} | {us:.1f} | {fl / us / 1e6:.1f} | {100 * fl / us / 1e6 / PEAK_TF:.2f} |", flush=True)
        os.environ.pop("SCATT_FUSION_NCOLS", None); os.environ.pop("SCATT_FUSION_SPC", None)


def membound_sweep(args):
    """K1 front end, K4 temporal pool, the row-wise LayerNorm tail and the CTC log-softmax front against the
    measured HBM copy bandwidth.  Buffers are larger than the 126 MB L2 at the big sizes; every launch is timed
    alone after an L2 flush (CUDA events, median of 15)."""
    from scattennet_b200.keypoint_module import frontend_forward
    cfg = model_config("phoenix-2014t")
    prec = F_.get_precision("fp16x3")
    model = MSCAEncoder(cfg, 1120, precision="fp16x3").eval()
    synth.load_synth_(model, 0)
    model = model.to(dev)
    mods = [model.body_encoder, model.left_encoder, model.right_encoder]
    print(f"### memory-bound kernels vs the measured HBM copy bandwidth ({PEAK_GBS:.0f} GB/s); one launch after an L2 flush\n")
    print("| kernel | shape | algorithmic MB | us | GB/s | % of HBM peak |")
    print("|---|---|---|---|---|---|")
    def row(name, shape, nbytes, fn):
        us = timed(fn, 15) * 1e3
        print(f"| {name} | {shape} | {nbytes / 1e6:.1f} | {us:.1f} | {nbytes / us / 1e3:.0f} | {100 * nbytes / us / 1e3 / PEAK_GBS:.1f} |", flush=True)
    T = 200
    for B in (8, 64, 256, 1024):
        for compact in (False, True):
            kp, _ = synth.synth_batch(B, T, seed=1)
            if compact:
                used, idx = model._compact_idx(torch.device(dev))
                kp = kp.index_select(2, used)
            else:
                idx = model._joint_idx(torch.device(dev))
            kp = kp.to(dev).contiguous()
            nbytes = B * T * (48 * 8 + 6 * 256 * 4)  # 384 B of used joints in, 6 branches x 256 x (hi + lo plane) out
            frontend_forward(prec, mods, kp, idx, B, T)
            sym = L.load().scatt_last_kernel().decode()  # frontend_kernel<R> (CUDA cores) or frontend_tc_kernel (tcgen05, large batches)
            row(sym + (" (compact input)" if compact else ""), f"B={B} T={T} K={kp.shape[2]}", nbytes,
                lambda: frontend_forward(prec, mods, kp, idx, B, T))
            del kp
    for B in (8, 256, 1024):
        for Cc in (256, 512):
            Tt = 200 if Cc == 256 else 100
            xs = [torch.randn(B * Tt, Cc, device=dev) for _ in range(3)]
            nbytes = 3 * B * Tt * Cc * (4 + 0.5 * 8)  # fp32 in; fp32 + planes out for half the rows
            row("pool_pairs_kernel (3 streams)", f"B={B} T={Tt} C={Cc}", nbytes, lambda: F_.pool_pairs_group(prec, xs, B, Tt))
            del xs
    for M, N in ((400, 1024), (51200, 1024), (204800, 512)):
        z = torch.randn(M, N, device=dev)
        ln = torch.nn.LayerNorm(N).to(dev)
        nbytes = M * N * (4 + 8)
        row("rowwise_kernel (LayerNorm + ReLU)", f"M={M} N={N}", nbytes,
            lambda: F_.rowwise(prec, z, F_.make_epilogue(layer_norm=True, act_post=L.ACT_RELU), ln))
        del z
    for B in (8, 256, 1024):
        lg = torch.randn(B, 50, 1120, device=dev)
        row("log_softmax_kernel", f"B={B} T'=50 V=1120", 2 * B * 50 * 1120 * 4, lambda: F_.log_softmax_clamp(lg, time_major=True))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["batch", "kernels", "membound", "fusion"])
    ap.add_argument("--precision", default="fp16x3,fp16x1")
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256,512,1024")
    ap.add_argument("--T", type=int, default=200)
    ap.add_argument("--config", default="phoenix-2014t")
    ap.add_argument("--only", default="", help="kernels: 'linear' skips the attention section, 'attention' the linear one")
    ap.add_argument("--attn-T", default="64,128,192,200,256,320,384,448,512", help="kernels: sequence lengths of the attention section")
    ap.add_argument("--schedules", default="", help="fusion: extra ncols:spc overrides, e.g. 128:1,256:1,256:4")
    a = ap.parse_args()
    {"batch": batch_sweep, "kernels": kernel_sweep, "membound": membound_sweep, "fusion": fusion_sweep}[a.what](a)
