"""Operand-precision feasibility probe (CPU emulation; dev tool, not product).

Re-evaluates the oracle with every contraction's operands rounded as a given
tensor-core mode would see them (fp32 accumulate), and prints max-abs error vs
the fp32 oracle for the C1 parity inputs.  Used to choose the default split
mode of the tcgen05 GEMMs ("numerics").
"""
import json, sys, os, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scatt_oracle as O
from scattennet_b200 import synth
from scattennet_b200.config import model_config
import torch.nn.functional as F

def rnd(x, dt):
    return x.to(dt).to(torch.float32)

def split(x, dt, terms):
    hi = rnd(x, dt)
    if terms == 1:
        return [hi]
    return [hi, rnd(x - hi, dt)]

def mm_mode(a, b, mode):
    """a @ b with operand rounding per mode (a: activations, b: weights^T or second operand)."""
    if mode == "fp32":
        return a @ b
    dt = torch.float16 if mode.startswith("fp16") else torch.bfloat16
    kind = mode.split("x")[1]
    if kind == "1":
        return rnd(a, dt) @ rnd(b, dt)
    ah, al = split(a, dt, 2)
    bh, bl = split(b, dt, 2)
    if kind == "2a":   # activation split only
        return ah @ bh + al @ bh
    if kind == "3":
        return ah @ bh + al @ bh + ah @ bl
    raise ValueError(mode)

def run(mode_lin, mode_att, cfg, sd, kp, mask):
    orig_linear, orig_attention, orig_fusion = O.linear, O.attention, O.coordinates_fusion
    def linear(sd_, p, x):
        w = sd_[p + ".weight"]
        return mm_mode(x, w.t(), mode_lin) + sd_[p + ".bias"]
    def attention(sd_, p, x_q, x_kv, additive, heads, kind):
        b, tq, d = x_q.shape; tk = x_kv.shape[1]; hd = d // heads
        q = linear(sd_, p + ".q_proj", x_q) * (hd ** -0.5)
        k = linear(sd_, p + ".k_proj", x_kv)
        v = linear(sd_, p + ".v_proj", x_kv / 2 if kind == "cross" else x_kv)
        q = q.view(b, tq, heads, hd).transpose(1, 2); k = k.view(b, tk, heads, hd).transpose(1, 2); v = v.view(b, tk, heads, hd).transpose(1, 2)
        s = mm_mode(q, k.transpose(-1, -2), mode_att)
        if kind == "causal":
            s = s.masked_fill(torch.ones(tq, tk, dtype=torch.bool).triu(1)[None, None], float("-inf"))
        s = s + additive
        o = mm_mode(torch.softmax(s, -1), v, mode_att)
        return linear(sd_, p + ".out_proj", o.transpose(1, 2).reshape(b, tq, d))
    def fusion(sd_, p, left, right, body):
        l = F.gelu(linear(sd_, p + ".left_se", left)); r = F.gelu(linear(sd_, p + ".right_se", right)); bd = F.gelu(linear(sd_, p + ".body_se", body))
        a = torch.softmax(mm_mode(r, l.transpose(1, 2), mode_att), -1)
        f = O.layer_norm(sd_, p + ".norm", linear(sd_, p + ".out_proj", mm_mode(a, bd, mode_att)))
        return O.inverted_residual(sd_, p + ".inverted_res", f)
    O.linear, O.attention, O.coordinates_fusion = linear, attention, fusion
    try:
        with torch.no_grad():
            return O.encoder_forward(sd, cfg, kp, mask)
    finally:
        O.linear, O.attention, O.coordinates_fusion = orig_linear, orig_attention, orig_fusion

if __name__ == "__main__":
    style = sys.argv[1] if len(sys.argv) > 1 else "perturbed"
    B, T = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (8, 200)
    torch.set_num_threads(8)
    cfg = model_config("phoenix-2014t")
    shapes = {k: tuple(v) for k, v in json.load(open(os.path.join(ROOT, "tests/golden/state_dict_phoenix-2014t.json"))).items()}
    sd = synth.synth_state_dict(shapes, 0, style)
    kp, mask = synth.synth_batch(B, T, 1, synth.parity_lengths(B, T))
    with torch.no_grad():
        ref = O.encoder_forward(sd, cfg, kp, mask)
    keys = ["body_embed", "left_embed", "right_embed", "fuse_embed", "fuse_coord_gloss_logits"]
    print("style", style, "B,T", B, T, {k: round(float(ref[k].abs().max()), 2) for k in keys})
    for ml, ma in [("fp16x1", "fp16x1"), ("fp16x1", "fp32"), ("fp16x2a", "fp16x2a"), ("fp16x3", "fp16x3"), ("bf16x3", "bf16x3"), ("bf16x1", "bf16x1"), ("fp16x3", "fp16x1"), ("fp16x1", "fp16x3")]:
        t0 = time.time()
        out = run(ml, ma, cfg, sd, kp, mask)
        print(f"lin={ml:8s} att={ma:8s}", {k: f"{float((out[k]-ref[k]).abs().max()):.2e}" for k in keys}, f"{time.time()-t0:.0f}s", flush=True)
