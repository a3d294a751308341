"""GPU parity of ``scatt_attn_block`` - out_proj + residual + LayerNorm -> fc1 + GELU -> fc2 + residual + LayerNorm as
one tcgen05 kernel (reference ``model/keypoint_module.py:62-72,98-107``, ``model/layers.py:103-108``) - against an fp64
torch restatement and against the three-launch path it replaces.  Tolerances per precision mode as in
``test_gpu_ops.py`` (LayerNorm outputs are O(1))."""

import pytest
import torch

from scattennet_b200 import _lib as L
from scattennet_b200 import functional as F_
from scattennet_b200 import synth
from scattennet_b200.functional import Act

pytestmark = pytest.mark.gpu
DEV = "cuda"
MODE_TOL = {"fp16x3": 2e-4, "bf16x3": 1e-3, "fp16x2": 1e-2, "fp16x1": 2e-2}


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def make_layer(D, Fh, seed):
    mods = {"out": torch.nn.Linear(D, D), "fc1": torch.nn.Linear(D, Fh), "fc2": torch.nn.Linear(Fh, D),
            "ln1": torch.nn.LayerNorm(D), "ln2": torch.nn.LayerNorm(D)}
    for i, (k, m) in enumerate(mods.items()):
        synth.load_synth_(m, seed * 10 + i)
        if k.startswith("ln"):
            with torch.no_grad():
                g = torch.Generator().manual_seed(seed * 10 + i)
                m.weight.copy_(1.0 + 0.2 * (torch.rand(D, generator=g) - 0.5))
                m.bias.copy_(0.1 * (torch.rand(D, generator=g) - 0.5))
        else:
            with torch.no_grad():
                m.bias.copy_(0.05 * torch.randn(m.bias.shape, generator=torch.Generator().manual_seed(seed * 10 + i + 5)))
        m.to(DEV)
    return mods


def ref_block(ctx, x, m):
    f = torch.nn.functional
    d = lambda t: t.double()
    h = f.layer_norm(d(x) + d(ctx) @ d(m["out"].weight).t() + d(m["out"].bias), (x.shape[1],), d(m["ln1"].weight), d(m["ln1"].bias), 1e-5)
    z = f.gelu(h @ d(m["fc1"].weight).t() + d(m["fc1"].bias)) @ d(m["fc2"].weight).t() + d(m["fc2"].bias)
    return f.layer_norm(h + z, (x.shape[1],), d(m["ln2"].weight), d(m["ln2"].bias), 1e-5)


def run_block(prec, ctxs, xs, layers, out_f32=True, out_planes=True):
    ctx_a = [Act(c).with_planes(prec) for c in ctxs]
    x_a = [Act(x).with_planes(prec) for x in xs]
    pk = lambda key: [F_.PackedLinear([m[key]], None, None) for m in layers]
    return F_.attn_block(prec, ctx_a, x_a, pk("out"), [m["ln1"] for m in layers], pk("fc1"), pk("fc2"),
                         [m["ln2"] for m in layers], out_f32=out_f32, out_planes=out_planes)


@pytest.fixture(params=[0, 1, 2], ids=["auto", "cta", "cluster2"])
def schedule(request):
    """One CTA per row tile, a 2-CTA cluster per row tile (hidden chunks split, partial sums through DSMEM), or the
    library's own choice."""
    L.check(L.load().scatt_debug_set_block_cluster(request.param), "set_block_cluster")
    yield request.param
    L.check(L.load().scatt_debug_set_block_cluster(0), "set_block_cluster")


@pytest.mark.parametrize("mode", list(MODE_TOL))
@pytest.mark.parametrize("M,Fh,G", [(200, 768, 1), (1600, 768, 3), (333, 256, 2), (77, 128, 1), (129, 384, 1), (128, 1024, 1)])
def test_attn_block_vs_fp64(mode, M, Fh, G, schedule):
    prec = F_.get_precision(mode)
    D = 256
    assert L.load().scatt_attn_block_supported(M, D, Fh)
    layers = [make_layer(D, Fh, 3 + g) for g in range(G)]
    ctxs = [rnd(M, D, seed=10 + g) for g in range(G)]
    xs = [rnd(M, D, seed=20 + g) for g in range(G)]
    outs = run_block(prec, ctxs, xs, layers)
    torch.cuda.synchronize()
    for g in range(G):
        ref = ref_block(ctxs[g], xs[g], layers[g])
        err = float((outs[g].f32.double() - ref).abs().max())
        assert err <= MODE_TOL[mode], (mode, M, Fh, g, err)
        rec = outs[g].planes[0].float() + outs[g].planes[1].float()
        rel = 2.0 ** (-21 if mode.startswith("fp16") else -15)
        assert float((rec - outs[g].f32).abs().max()) <= rel * float(outs[g].f32.abs().max()) + 1e-7


@pytest.mark.parametrize("mode", ["fp16x3", "fp16x1"])
def test_attn_block_persistent_many_tiles(mode, schedule):
    """More row tiles than SMs: every CTA walks several tiles (barrier phases, ring and TMEM reuse across tiles);
    the group index changes inside a CTA's walk (per-column parameters reloaded)."""
    prec = F_.get_precision(mode)
    D, Fh, G = 256, 768, 2
    M = 128 * 170 + 19
    layers = [make_layer(D, Fh, 7 + g) for g in range(G)]
    ctxs = [rnd(M, D, seed=30 + g) for g in range(G)]
    xs = [rnd(M, D, seed=40 + g) for g in range(G)]
    outs = run_block(prec, ctxs, xs, layers, out_f32=True, out_planes=False)
    torch.cuda.synchronize()
    for g in range(G):
        ref = ref_block(ctxs[g], xs[g], layers[g])
        err = float((outs[g].f32.double() - ref).abs().max())
        assert err <= MODE_TOL[mode], (mode, g, err)


def test_attn_block_matches_three_launch_path():
    """Same operands through the unfused launches (scatt_linear x 3): the two paths agree to fp32-rounding level."""
    prec = F_.get_precision("fp16x3")
    D, Fh, M = 256, 768, 1600
    m = make_layer(D, Fh, 11)
    ctx, x = rnd(M, D, seed=50), rnd(M, D, seed=51)
    fused = run_block(prec, [ctx], [x], [m], out_f32=True)[0]
    ep_ln = F_.make_epilogue(residual_mode=F_.L.RES_BEFORE_LN, layer_norm=True)
    xa = Act(x).with_planes(prec)
    h = F_.linear(prec, [Act(ctx).with_planes(prec)], [F_.PackedLinear([m["out"]], None, None)], ep_ln, residuals=[xa], lns=[m["ln1"]])
    f = F_.linear(prec, h, [F_.PackedLinear([m["fc1"]], None, None)], F_.make_epilogue(act_pre=F_.L.ACT_GELU))
    y = F_.linear(prec, f, [F_.PackedLinear([m["fc2"]], None, None)], ep_ln, residuals=h, lns=[m["ln2"]])[0]
    torch.cuda.synchronize()
    assert float((fused.f32 - y.f32).abs().max()) <= 2e-5


def test_attn_block_planes_only_output_and_repeatability():
    prec = F_.get_precision("fp16x3")
    D, Fh, M = 256, 768, 900
    m = make_layer(D, Fh, 13)
    ctx, x = rnd(M, D, seed=60), rnd(M, D, seed=61)
    a = run_block(prec, [ctx], [x], [m], out_f32=False, out_planes=True)[0]
    b = run_block(prec, [ctx], [x], [m], out_f32=True, out_planes=True)[0]
    torch.cuda.synchronize()
    assert a.f32 is None
    assert torch.equal(a.planes, b.planes)  # deterministic: no atomics, fixed accumulation order
    ref = ref_block(ctx, x, m)
    assert float(((a.planes[0].float() + a.planes[1].float()).double() - ref).abs().max()) <= 2e-4


# ---- scatt_attn_out_q: the same kernel in its second mode (causal layer tail + the merge layer's q projection)


def run_out_q(prec, ctxs, xs, layers, scale):
    ctx_a = [Act(c).with_planes(prec) for c in ctxs]
    x_a = [Act(x).with_planes(prec) for x in xs]
    pk = lambda key: [F_.PackedLinear([m[key]], None, None) for m in layers]
    return F_.attn_out_q(prec, ctx_a, x_a, pk("out"), [m["ln1"] for m in layers], pk("fc1"), scale)


@pytest.mark.parametrize("mode", ["fp16x3", "bf16x3", "fp16x1"])
@pytest.mark.parametrize("M,G", [(200, 1), (1600, 3), (333, 2), (77, 1), (129, 1)])
def test_attn_out_q_vs_fp64(mode, M, G, schedule):
    """h = LN(x + ctx Wo^T + bo) and q = (h Wq^T + bq) * scale, both as split planes, against fp64 torch - one CTA and a
    2-CTA cluster per row tile, M tails, several streams."""
    prec = F_.get_precision(mode)
    D, scale = 256, 0.25
    assert L.load().scatt_attn_out_q_supported(M, D, D)
    layers = [make_layer(D, D, 23 + g) for g in range(G)]  # "fc1" plays q_proj (256 -> 256)
    ctxs = [rnd(M, D, seed=70 + g) for g in range(G)]
    xs = [rnd(M, D, seed=80 + g) for g in range(G)]
    hs, qs = run_out_q(prec, ctxs, xs, layers, scale)
    torch.cuda.synchronize()
    f, d = torch.nn.functional, (lambda t: t.double())
    for g in range(G):
        m = layers[g]
        h_ref = f.layer_norm(d(xs[g]) + d(ctxs[g]) @ d(m["out"].weight).t() + d(m["out"].bias), (D,), d(m["ln1"].weight), d(m["ln1"].bias), 1e-5)
        q_ref = (h_ref @ d(m["fc1"].weight).t() + d(m["fc1"].bias)) * scale
        h = hs[g].planes[0].double() + hs[g].planes[1].double()
        q = qs[g].planes[0].double() + qs[g].planes[1].double()
        assert hs[g].f32 is None and qs[g].f32 is None
        assert float((h - h_ref).abs().max()) <= MODE_TOL[mode], (mode, M, g)
        assert float((q - q_ref).abs().max()) <= MODE_TOL[mode], (mode, M, g)


def test_attn_out_q_many_tiles_matches_two_launch_path():
    """More row tiles than SMs (every CTA walks several tiles) and agreement with the launches it replaces
    (scatt_linear with the LayerNorm epilogue, then scatt_linear with the q scaling)."""
    prec = F_.get_precision("fp16x3")
    D, G, scale = 256, 2, 0.25
    M = 128 * 160 + 45
    layers = [make_layer(D, D, 31 + g) for g in range(G)]
    ctxs = [rnd(M, D, seed=90 + g) for g in range(G)]
    xs = [rnd(M, D, seed=95 + g) for g in range(G)]
    hs, qs = run_out_q(prec, ctxs, xs, layers, scale)
    pk = lambda key: [F_.PackedLinear([m[key]], None, None) for m in layers]
    h2 = F_.linear(prec, [Act(c).with_planes(prec) for c in ctxs], pk("out"), F_.make_epilogue(residual_mode=F_.L.RES_BEFORE_LN, layer_norm=True),
                   residuals=[Act(x).with_planes(prec) for x in xs], lns=[m["ln1"] for m in layers])
    q2 = F_.linear(prec, h2, pk("fc1"), F_.make_epilogue(scale_cols=D, scale=scale), out_f32=False)
    torch.cuda.synchronize()
    for g in range(G):
        assert float(((hs[g].planes[0].float() + hs[g].planes[1].float()) - h2[g].f32).abs().max()) <= 2e-5
        assert float(((qs[g].planes[0].float() + qs[g].planes[1].float()) - (q2[g].planes[0].float() + q2[g].planes[1].float())).abs().max()) <= 2e-5
