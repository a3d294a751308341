"""GPU parity of the individual C-ABI ops against the oracle / plain torch on
seeded inputs.  Tolerances: exact for gathers, pooling and the split planes'
reconstruction bound; fp32-order noise for the SIMT engine; per-mode bounds
for the tcgen05 engine (stated next to each test)."""

import math

import pytest
import torch

from oracle import scatt_oracle as O
from scattennet_b200 import _lib as L
from scattennet_b200 import functional as F_
from scattennet_b200 import synth
from scattennet_b200.functional import Act

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


class Lin(torch.nn.Linear):
    pass


def make_linear(n, k, seed):
    lin = torch.nn.Linear(k, n)
    synth.load_synth_(lin, seed)
    return lin.to(DEV)


def ref_chain(x, lin, ep_kw, residual=None, ln=None):
    """torch fp64 restatement of the epilogue chain (checker)."""
    z = x.double() @ lin.weight.double().t() + lin.bias.double()
    sc = ep_kw.get("scale_cols", 0)
    if sc:
        z[:, :sc] = z[:, :sc] * ep_kw["scale"]
    act = {L.ACT_NONE: lambda v: v, L.ACT_GELU: torch.nn.functional.gelu, L.ACT_RELU: torch.relu}
    z = act[ep_kw.get("act_pre", 0)](z)
    if ep_kw.get("residual_mode", 0) == L.RES_BEFORE_LN:
        z = z + residual.double()
    if ep_kw.get("layer_norm", False):
        z = torch.nn.functional.layer_norm(z, (z.shape[1],), ln.weight.double(), ln.bias.double(), 1e-5)
    if ep_kw.get("residual_mode", 0) == L.RES_AFTER_LN:
        z = z + residual.double()
    z = act[ep_kw.get("act_post", 0)](z)
    if ep_kw.get("clamp", 0) > 0:
        z = z.clamp(-ep_kw["clamp"], ep_kw["clamp"])
    return z


EPILOGUES = {
    "bias": dict(),
    "qscale": dict(scale_cols=256, scale=0.25),
    "gelu": dict(act_pre=L.ACT_GELU),
    "res_ln": dict(residual_mode=L.RES_BEFORE_LN, layer_norm=True),
    "ln_relu": dict(layer_norm=True, act_post=L.ACT_RELU),
    "ln_res_relu": dict(layer_norm=True, residual_mode=L.RES_AFTER_LN, act_post=L.ACT_RELU),
    "gelu_res_ln": dict(act_pre=L.ACT_GELU, residual_mode=L.RES_BEFORE_LN, layer_norm=True),
    "clamp": dict(clamp=0.5),
}

# max-abs tolerance on O(1)-magnitude outputs (K <= 1024, |x| ~ 1, Xavier weights)
MODE_TOL = {"fp32": 2e-5, "fp16x3": 5e-5, "bf16x3": 2e-4, "fp16x2": 3e-3, "fp16x1": 4e-3, "bf16x1": 3e-2}


@pytest.mark.parametrize("mode", list(MODE_TOL))
@pytest.mark.parametrize("epi", list(EPILOGUES))
@pytest.mark.parametrize("shape", [(200, 256, 256), (1600, 256, 768), (333, 768, 256), (77, 512, 512), (400, 1024, 512), (50, 1120, 1024)])
def test_linear_epilogues(mode, epi, shape):
    M, N, K = shape
    kw = EPILOGUES[epi]
    if kw.get("layer_norm") and N > 1024:
        pytest.skip("LayerNorm rows are at most 1024 wide on this path")
    prec = F_.get_precision(mode)
    x = rnd(M, K, seed=1)
    lin = make_linear(N, K, 2)
    ln = torch.nn.LayerNorm(N)
    with torch.no_grad():
        ln.weight.copy_(1.0 + 0.2 * (torch.rand(N, generator=torch.Generator().manual_seed(3)) - 0.5))
        ln.bias.copy_(0.1 * (torch.rand(N, generator=torch.Generator().manual_seed(4)) - 0.5))
    ln = ln.to(DEV)
    res = rnd(M, N, seed=4) if kw.get("residual_mode", 0) else None
    pk = F_.PackedLinear([lin], None, None)
    out = F_.linear(prec, [Act(x)], [pk], F_.make_epilogue(**kw), residuals=None if res is None else [res],
                    lns=[ln] if kw.get("layer_norm") else None)[0]
    torch.cuda.synchronize()
    ref = ref_chain(x, lin, kw, res, ln)
    err = float((out.f32.double() - ref).abs().max())
    tol = MODE_TOL[mode] * (4.0 if kw.get("layer_norm") else 1.0) * max(1.0, math.sqrt(K / 256))
    assert err <= tol, (mode, epi, shape, err)
    if prec.uses_planes:  # exported planes reconstruct the fp32 output to 2^-21 relative (fp16) / 2^-15 (bf16)
        rec = out.planes[0].float() + out.planes[1].float()
        rel = 2.0 ** (-21 if mode.startswith("fp16") else -15)
        assert float((rec - out.f32).abs().max()) <= rel * float(out.f32.abs().max()) + 1e-7


@pytest.mark.parametrize("mode", ["fp16x3", "fp16x1"])
@pytest.mark.parametrize("epi", ["res_ln", "ln_res_relu", "bias"])
def test_linear_n256_large_m_single_cta_path(mode, epi):
    """N = 256 switches from 2-CTA clusters (small batch) to one CTA per 128-row tile once the row tiles
    alone fill the GPU (> 74 of them); both LayerNorm implementations must agree with the reference chain."""
    M, N, K = 128 * 80 + 17, 256, 256
    kw = EPILOGUES[epi]
    prec = F_.get_precision(mode)
    x = rnd(M, K, seed=1)
    lin = make_linear(N, K, 2)
    ln = torch.nn.LayerNorm(N).to(DEV)
    res = rnd(M, N, seed=4) if kw.get("residual_mode", 0) else None
    out = F_.linear(prec, [Act(x)], [F_.PackedLinear([lin], None, None)], F_.make_epilogue(**kw),
                    residuals=None if res is None else [res], lns=[ln] if kw.get("layer_norm") else None)[0]
    ref = ref_chain(x, lin, kw, res, ln)
    assert float((out.f32.double() - ref).abs().max()) <= MODE_TOL[mode] * 4


@pytest.mark.parametrize("mode", ["fp16x3", "fp16x1"])
@pytest.mark.parametrize("epi", ["ln_relu", "ln_res_relu", "gelu_res_ln", "res_ln"])
@pytest.mark.parametrize("G,M,N,K", [(3, 800, 512, 256), (3, 400, 512, 512), (1, 400, 1024, 1024), (1, 400, 1024, 512), (1, 2300, 1024, 512),
                                     (3, 128 * 13, 512, 256), (1, 128 * 19 + 3, 1024, 256)])
def test_linear_wide_layernorm_cluster_and_tail(mode, epi, G, M, N, K):
    """Rows of 512 / 1024 columns are normalised inside the GEMM by 4- / 8-CTA clusters (DSMEM statistics
    exchange) while the clusters fit one wave, and by the row-wise tail launch beyond; both against the chain."""
    kw = EPILOGUES[epi]
    prec = F_.get_precision(mode)
    fused = bool(L.load().scatt_linear_ln_fused(M, N, G, prec.engine))
    assert fused == (-(-M // 128) * G * (N // 128) <= 148)
    xs = [rnd(M, K, seed=10 + g) for g in range(G)]
    lins = [make_linear(N, K, 20 + g) for g in range(G)]
    lns = []
    for g in range(G):
        ln = torch.nn.LayerNorm(N)
        with torch.no_grad():
            ln.weight.copy_(1.0 + 0.2 * (torch.rand(N, generator=torch.Generator().manual_seed(30 + g)) - 0.5))
            ln.bias.copy_(0.1 * (torch.rand(N, generator=torch.Generator().manual_seed(40 + g)) - 0.5))
        lns.append(ln.to(DEV))
    res = [rnd(M, N, seed=50 + g) for g in range(G)] if kw.get("residual_mode", 0) else None
    for out_f32 in (True, False):
        outs = F_.linear(prec, [Act(x) for x in xs], [F_.PackedLinear([l], None, None) for l in lins], F_.make_epilogue(**kw),
                         residuals=res, lns=lns, out_f32=out_f32)
        torch.cuda.synchronize()
        for g in range(G):
            ref = ref_chain(xs[g], lins[g], kw, None if res is None else res[g], lns[g])
            got = outs[g].f32.double() if outs[g].f32 is not None else outs[g].planes[0].double() + outs[g].planes[1].double()
            if not out_f32 and fused:
                assert outs[g].f32 is None  # planes only: no fp32 round trip when the LayerNorm is fused
            err = float((got - ref).abs().max())
            assert err <= MODE_TOL[mode] * 4.0 * max(1.0, math.sqrt(K / 256)), (mode, epi, G, M, N, K, out_f32, err)


@pytest.mark.parametrize("mode", ["fp16x3", "fp16x1", "bf16x3"])
@pytest.mark.parametrize("epi", list(EPILOGUES))
@pytest.mark.parametrize("shape", [(400, 1024, 1024), (400, 1024, 3072), (130, 512, 2048), (400, 1024, 1536), (77, 640, 1024)])
def test_linear_split_k(mode, epi, shape):
    """Few row tiles and a long K loop (the fusion block at B = 8): K is cut into slices that run as problem slots of one
    GEMM launch, the row-wise reduce kernel adds the fp32 partial sums in slot order and runs the epilogue."""
    M, N, K = shape
    kw = EPILOGUES[epi]
    prec = F_.get_precision(mode)
    assert L.load().scatt_linear_workspace_bytes(1, M, N, K, prec.engine) >= 3 * M * N * 4  # the split path is the one tested
    x = rnd(M, K, seed=1)
    lin = make_linear(N, K, 2)
    ln = torch.nn.LayerNorm(N)
    with torch.no_grad():
        ln.weight.copy_(1.0 + 0.2 * (torch.rand(N, generator=torch.Generator().manual_seed(3)) - 0.5))
        ln.bias.copy_(0.1 * (torch.rand(N, generator=torch.Generator().manual_seed(4)) - 0.5))
    ln = ln.to(DEV)
    res = rnd(M, N, seed=4) if kw.get("residual_mode", 0) else None
    pk = F_.PackedLinear([lin], None, None)
    for res_as_planes in ((False, True) if res is not None else (False,)):
        r = None if res is None else [Act(None, F_.split_planes(res, prec)) if res_as_planes else res]
        out = F_.linear(prec, [Act(x)], [pk], F_.make_epilogue(**kw), residuals=r, lns=[ln] if kw.get("layer_norm") else None)[0]
        ref = ref_chain(x, lin, kw, res, ln)
        tol = MODE_TOL[mode] * (4.0 if kw.get("layer_norm") else 1.0) * max(1.0, math.sqrt(K / 256))
        assert float((out.f32.double() - ref).abs().max()) <= tol, (mode, epi, shape, res_as_planes)
        rec = out.planes[0].float() + out.planes[1].float()
        rel = 2.0 ** (-21 if mode.startswith("fp16") else -15)
        assert float((rec - out.f32).abs().max()) <= rel * float(out.f32.abs().max()) + 1e-7
    # planes-only outputs (what the encoder asks for)
    if res is None:
        out = F_.linear(prec, [Act(x)], [pk], F_.make_epilogue(**kw), lns=[ln] if kw.get("layer_norm") else None, out_f32=False)[0]
        if out.f32 is None:
            ref = ref_chain(x, lin, kw, res, ln)
            rec = out.planes[0].double() + out.planes[1].double()
            assert float((rec - ref).abs().max()) <= MODE_TOL[mode] * (4.0 if kw.get("layer_norm") else 1.0) * max(1.0, math.sqrt(K / 256)) * 2


@pytest.mark.parametrize("mode", ["fp16x3", "fp16x1", "bf16x3"])
@pytest.mark.parametrize("epi", ["bias", "qscale", "gelu", "clamp"])
@pytest.mark.parametrize("shape", [(12800 + 5, 768, 256), (19000, 1120, 512)])
def test_linear_multi_wave_dual_cta_path(mode, epi, shape):
    """Grids larger than one wave run the 2-CTAs-per-SM kernel (128-wide tiles, 4 epilogue warps)."""
    M, N, K = shape
    kw = EPILOGUES[epi]
    prec = F_.get_precision(mode)
    x = rnd(M, K, seed=1)
    lin = make_linear(N, K, 2)
    res = rnd(M, N, seed=4)
    for with_res in (False, True):  # with a residual the accumulator pre-initialisation runs on 4 warps
        kw2 = dict(kw, residual_mode=L.RES_BEFORE_LN) if with_res and epi == "bias" else kw
        out = F_.linear(prec, [Act(x)], [F_.PackedLinear([lin], None, None)], F_.make_epilogue(**kw2),
                        residuals=[res] if kw2.get("residual_mode") else None)[0]
        ref = ref_chain(x, lin, kw2, res, None)
        assert float((out.f32.double() - ref).abs().max()) <= MODE_TOL[mode] * max(1.0, math.sqrt(K / 256))
        rec = out.planes[0].float() + out.planes[1].float()
        assert float((rec - out.f32).abs().max()) <= 2.0 ** (-21 if mode.startswith("fp16") else -15) * float(out.f32.abs().max()) + 1e-7


@pytest.mark.parametrize("mode", ["fp16x3", "fp16x1"])
@pytest.mark.parametrize("epi", ["bias", "gelu", "qscale"])
@pytest.mark.parametrize("out", ["planes", "f32", "both"])
def test_linear_multi_wave_grouped_outputs(mode, epi, out):
    """Multi-wave grids run the persistent kernel (one CTA per SM walks 128 x 128 tiles, two TMEM accumulators):
    three grouped problems, ragged M and N tails, every output combination (the staging differs)."""
    G, M, N, K = 3, 128 * 21 + 77, 736, 320
    kw = EPILOGUES[epi]
    prec = F_.get_precision(mode)
    xs = [rnd(M, K, seed=60 + g) for g in range(G)]
    lins = [make_linear(N, K, 70 + g) for g in range(G)]
    outs = F_.linear(prec, [Act(x) for x in xs], [F_.PackedLinear([l], None, None) for l in lins], F_.make_epilogue(**kw),
                     out_f32=out != "planes", out_planes=out != "f32")
    torch.cuda.synchronize()
    for g in range(G):
        ref = ref_chain(xs[g], lins[g], kw)
        o = outs[g]
        if out != "planes":
            assert float((o.f32.double() - ref).abs().max()) <= MODE_TOL[mode] * max(1.0, math.sqrt(K / 256)), (mode, epi, out, g)
        if out != "f32":
            rec = o.planes[0].double() + o.planes[1].double()
            assert float((rec - ref).abs().max()) <= MODE_TOL[mode] * max(1.0, math.sqrt(K / 256)) + 1e-6, (mode, epi, out, g)


@pytest.mark.parametrize("mode", ["fp16x3", "bf16x3", "fp16x1"])
@pytest.mark.parametrize("shape", [(128 * 80 + 17, 256, 256), (128 * 75, 256, 768)])
def test_linear_residual_from_planes(mode, shape):
    """A residual stream that lives in split planes only (no fp32 copy): `residual_planes` must give the result of
    the fp32 residual it was split from (large-batch LayerNorm kernel), and the output may be planes only as well."""
    M, N, K = shape
    prec = F_.get_precision(mode)
    x = rnd(M, K, seed=1)
    lin = make_linear(N, K, 2)
    ln = torch.nn.LayerNorm(N).to(DEV)
    res = rnd(M, N, seed=4)
    res_planes = F_.split_planes(res, prec)
    res_rec = res_planes[0].float() + res_planes[1].float()  # what the planes hold (== res to 2^-21 / 2^-15 relative)
    for layer_norm in (True,):
        ep = F_.make_epilogue(residual_mode=L.RES_BEFORE_LN, layer_norm=layer_norm)
        kw = dict(lns=[ln] if layer_norm else None)
        want = F_.linear(prec, [Act(x)], [F_.PackedLinear([lin], None, None)], ep, residuals=[res_rec], **kw)[0]
        got = F_.linear(prec, [Act(x)], [F_.PackedLinear([lin], None, None)], ep, residuals=[Act(None, res_planes)],
                        out_f32=False, **kw)[0]
        assert got.f32 is None and got.planes is not None
        assert torch.equal(got.planes, want.planes), (mode, shape, layer_norm)


def test_linear_residual_from_planes_rejected_when_not_fused():
    """A residual given as planes needs a LayerNorm kernel that stages / folds it; a plain GEMM rejects it loudly."""
    prec = F_.get_precision("fp16x3")
    x, res = rnd(64, 256, seed=1), rnd(64, 768, seed=2)
    with pytest.raises(Exception, match="residual"):
        F_.linear(prec, [Act(x)], [F_.PackedLinear([make_linear(768, 256, 2)], None, None)],
                  F_.make_epilogue(residual_mode=L.RES_BEFORE_LN), residuals=[Act(None, F_.split_planes(res, prec))])


@pytest.mark.parametrize("mode", ["fp16x3", "fp16x1"])
@pytest.mark.parametrize("epi", ["res_ln", "ln_res_relu", "gelu_res_ln"])
@pytest.mark.parametrize("G,M,N,K", [(3, 1600, 256, 256), (3, 1600, 256, 768), (1, 77, 256, 128), (3, 400, 512, 512), (1, 400, 1024, 1024)])
def test_linear_cluster_residual_from_planes(mode, epi, G, M, N, K):
    """Small-batch LayerNorm GEMMs (2- / 4- / 8-CTA clusters) with the residual stream kept in split planes only:
    the planes are staged by TMA and hi + lo is added in the epilogue; planes-only output."""
    kw = EPILOGUES[epi]
    prec = F_.get_precision(mode)
    xs = [rnd(M, K, seed=10 + g) for g in range(G)]
    lins = [make_linear(N, K, 20 + g) for g in range(G)]
    lns = [torch.nn.LayerNorm(N).to(DEV) for _ in range(G)]
    res = [rnd(M, N, seed=50 + g) for g in range(G)]
    res_pl = [F_.split_planes(r, prec) for r in res]
    res_rec = [pl[0].float() + pl[1].float() for pl in res_pl]  # what the kernel adds
    outs = F_.linear(prec, [Act(x) for x in xs], [F_.PackedLinear([l], None, None) for l in lins], F_.make_epilogue(**kw),
                     residuals=[Act(None, pl) for pl in res_pl], lns=lns, out_f32=False)
    torch.cuda.synchronize()
    for g in range(G):
        ref = ref_chain(xs[g], lins[g], kw, res_rec[g], lns[g])
        assert outs[g].f32 is None
        got = outs[g].planes[0].double() + outs[g].planes[1].double()
        assert float((got - ref).abs().max()) <= MODE_TOL[mode] * 4.0 * max(1.0, math.sqrt(K / 256)), (mode, epi, G, M, N, K)


@pytest.mark.parametrize("mode", ["fp32", "fp16x3", "fp16x1"])
def test_linear_grouped_matches_single(mode):
    prec = F_.get_precision(mode)
    xs = [rnd(300, 256, seed=10 + g) for g in range(3)]
    lins = [make_linear(768, 256, 20 + g) for g in range(3)]
    packs = [F_.PackedLinear([l], None, None) for l in lins]
    ep = F_.make_epilogue(act_pre=L.ACT_GELU)
    grouped = F_.linear(prec, [Act(x) for x in xs], packs, ep)
    for g in range(3):
        single = F_.linear(prec, [Act(xs[g])], [packs[g]], ep)[0]
        assert torch.equal(grouped[g].f32, single.f32)


def test_packed_linear_concat_and_scale():
    prec = F_.get_precision("fp32")
    a, b = make_linear(256, 256, 1), make_linear(256, 256, 2)
    pk = F_.PackedLinear([a, b], [1.0, 0.5], None)
    x = rnd(64, 256, seed=3)
    y = F_.linear(prec, [Act(x)], [pk], F_.make_epilogue())[0].f32
    ref = torch.cat([x @ a.weight.t() + a.bias, (x / 2) @ b.weight.t() + b.bias], 1)
    assert float((y - ref).abs().max()) <= 2e-5


@pytest.mark.parametrize("fmt", ["fp16x3", "bf16x3"])
def test_split_planes_bound(fmt):
    prec = F_.get_precision(fmt)
    x = rnd(129, 256, seed=5, scale=3.0)
    p = F_.split_planes(x, prec, scale=0.5)
    rec = p[0].float() + p[1].float()
    rel = 2.0 ** (-21 if fmt.startswith("fp16") else -15)
    assert float((rec - 0.5 * x).abs().max()) <= rel * float(x.abs().max())
    assert torch.equal(p[0], (0.5 * x).to(prec.plane_dtype))


ATT_TOL = {"fp32": 3e-5, "fp16x3": 3e-5, "bf16x3": 3e-4, "fp16x1": 1e-2}


@pytest.mark.parametrize("mode", list(ATT_TOL))
@pytest.mark.parametrize("kind", ["self", "causal", "cross"])
@pytest.mark.parametrize("B,T", [(2, 24), (3, 37), (8, 200), (1, 130), (2, 256), (2, 400)])
def test_stream_attention_key_mask(kind, B, T, mode):
    """fp32 -> CUDA-core kernel; fp16xN / bf16xN -> tcgen05 kernel for T <= 256 (fp32 kernel above)."""
    H, D = 16, 256
    prec = F_.get_precision(mode)
    q, k, v = (rnd(B * T, D, seed=s) for s in (1, 2, 3))
    lengths = synth.parity_lengths(B, T)
    mask = (torch.arange(T)[None] < torch.tensor(lengths)[:, None]).long()
    if B >= 3:
        mask[2] = 0  # an all-padded sequence: rows must come out uniform over the permitted keys
    act = F_.stream_attention(prec, [q], [k], [v], B, T, T, H, {"self": 0, "causal": 1, "cross": 2}[kind],
                              key_mask=F_.key_mask_u8(mask.to(DEV)))[0]
    out = act.f32 if act.f32 is not None else act.planes[0].float() + act.planes[1].float()
    qh, kh, vh = (t.cpu().view(B, T, H, 16).transpose(1, 2) for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2)
    if kind == "causal":
        s = s.masked_fill(torch.ones(T, T, dtype=torch.bool).triu(1)[None, None], float("-inf"))
        s = s + O.causal_additive(mask, T, torch.float32)
    else:
        s = s + O.key_padding_additive(mask, torch.float32)
    ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B * T, D)
    assert torch.isfinite(out).all()
    assert float((out.cpu() - ref).abs().max()) <= ATT_TOL[mode]


@pytest.fixture(params=[2, 1], ids=["one_item_per_cta", "persistent"])
def attn_schedule(request):
    """Both schedules of scatt_attention_planes on every shape (the library picks by item count: >= 4096 items run the
    persistent two-group kernel, which the small test shapes would never reach)."""
    L.check(L.load().scatt_debug_set_attn_persist(request.param), "set_attn_persist")
    yield request.param
    L.check(L.load().scatt_debug_set_attn_persist(0), "set_attn_persist")


@pytest.mark.parametrize("mode", ["fp16x3", "bf16x3", "fp16x1"])
@pytest.mark.parametrize("kind", ["self", "causal", "cross"])
@pytest.mark.parametrize("B,T", [(2, 24), (3, 37), (8, 200), (1, 130), (2, 224), (5, 16), (2, 225), (3, 400), (2, 512), (1, 672), (24, 200), (40, 100), (12, 450), (2, 673), (3, 1000), (1, 1568)])
def test_stream_attention_planes(kind, B, T, mode, attn_schedule):
    """TMA-fed tcgen05 kernel: operands are split planes inside a wider [rows, 768] buffer, like the QKV GEMM writes them."""
    H, D = 16, 256
    prec = F_.get_precision(mode)
    qkv = rnd(B * T, 3 * D, seed=7)
    planes = F_.split_planes(qkv, prec)
    lengths = synth.parity_lengths(B, T)
    mask = (torch.arange(T)[None] < torch.tensor(lengths)[:, None]).long()
    if B >= 3:
        mask[2] = 0
    act = F_.stream_attention_planes(prec, [(planes, 0)], [(planes, D)], [(planes, 2 * D)], B, T, T, H,
                                     {"self": 0, "causal": 1, "cross": 2}[kind], key_mask=F_.key_mask_u8(mask.to(DEV)))[0]
    out = act.planes[0].float() + act.planes[1].float()
    q, k, v = (qkv[:, i * D:(i + 1) * D].cpu().view(B, T, H, 16).transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2)
    if kind == "causal":
        s = s.masked_fill(torch.ones(T, T, dtype=torch.bool).triu(1)[None, None], float("-inf"))
        s = s + O.causal_additive(mask, T, torch.float32)
    else:
        s = s + O.key_padding_additive(mask, torch.float32)
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, D)
    assert torch.isfinite(out).all()
    assert float((out.cpu() - ref).abs().max()) <= ATT_TOL[mode]


def test_stream_attention_additive_mask_and_grouping():
    B, T, H, D = 2, 24, 16, 256
    prec = F_.get_precision("fp16x3")
    qs = [rnd(B * T, 3 * D, seed=s) for s in (1, 2, 3)]
    add = rnd(B, 1, T, T, seed=9)
    outs = F_.stream_attention(prec, [t[:, :D] for t in qs], [t[:, D:2 * D] for t in qs], [t[:, 2 * D:] for t in qs], B, T, T, H, 0,
                               additive=add)
    for g in range(3):
        q, k, v = (qs[g][:, i * D:(i + 1) * D].cpu().view(B, T, H, 16).transpose(1, 2) for i in range(3))
        ref = (torch.softmax(q @ k.transpose(-1, -2) + add.cpu(), -1) @ v).transpose(1, 2).reshape(B * T, D)
        rec = outs[g].planes[0].float() + outs[g].planes[1].float()
        assert float((rec.cpu() - ref).abs().max()) <= 3e-5


@pytest.mark.parametrize("B,T,D", [(2, 7, 1024), (8, 50, 1024), (2, 200, 1024), (3, 9, 512)])
def test_fusion_attention(B, T, D):
    prec = F_.get_precision("fp32")
    q, k, v = (rnd(B * T, D, seed=s, scale=0.6).abs() for s in (1, 2, 3))  # GELU-like positive-heavy operands, large logits
    out = F_.fusion_attention(prec, q, k, v, B, T).f32
    qd, kd, vd = (t.double().cpu().view(B, T, D) for t in (q, k, v))
    ref = (torch.softmax(qd @ kd.transpose(1, 2), -1) @ vd).reshape(B * T, D)
    assert float((out.double().cpu() - ref).abs().max()) <= 2e-4


# logits reach ~235 here (1024 positive products, no scaling): one fp32 ulp of the accumulator is 1.5e-5 and a logit error e
# is a relative error e of the probability - the fp32 CUDA-core kernel above is held to 2e-4 on the same operands
FUSION_TOL = {"fp16x3": 3e-4, "bf16x3": 2e-3, "fp16x1": 0.25}


def _fusion_planes_case(mode, B, T, D, out_f32=False):
    prec = F_.get_precision(mode)
    acts = [F_.Act.from_f32(rnd(B * T, D, seed=s, scale=0.6).abs()).with_planes(prec) for s in (1, 2, 3)]
    q, k, v = (a.planes for a in acts)
    out = F_.fusion_attention_planes(prec, q, k, v, B, T, out_f32=out_f32)
    # the kernel sees exactly hi + lo: compare against the fp64 contraction of the reconstructed operands
    qd, kd, vd = ((p[0].double() + p[1].double()).cpu().view(B, T, D) for p in (q, k, v))
    ref = (torch.softmax(qd @ kd.transpose(1, 2), -1) @ vd).reshape(B * T, D)
    rec = (out.planes[0].double() + out.planes[1].double()).cpu()
    assert torch.isfinite(rec).all()
    err = float((rec - ref).abs().max())
    if out_f32:
        assert float((out.f32.double().cpu() - ref).abs().max()) <= FUSION_TOL[mode]
    return err


@pytest.mark.parametrize("mode", ["fp16x3", "bf16x3", "fp16x1"])
@pytest.mark.parametrize("B,T,D", [(2, 7, 1024), (8, 50, 1024), (2, 200, 1024), (3, 37, 512), (1, 256, 1024), (2, 129, 1024),
                                   (1, 64, 256), (160, 50, 1024), (80, 130, 1024)])
def test_fusion_attention_planes(mode, B, T, D):
    """tcgen05 fusion attention (logits ~ 200, no scaling) against fp64; the last two shapes take the
    all-slices-in-one-CTA schedule."""
    assert F_.fusion_attention_planes_supported(F_.get_precision(mode), T, D)
    assert _fusion_planes_case(mode, B, T, D, out_f32=(B == 2)) <= FUSION_TOL[mode]


@pytest.mark.parametrize("ncols,spc", [(64, 1), (64, 4), (128, 1), (128, 2), (128, 8), (256, 1), (256, 2), (256, 4)])
def test_fusion_attention_planes_schedules(ncols, spc, monkeypatch):
    """Every (slice width, slices per CTA) schedule gives the same result (the launcher reads the override per launch)."""
    monkeypatch.setenv("SCATT_FUSION_NCOLS", str(ncols))
    monkeypatch.setenv("SCATT_FUSION_SPC", str(spc))
    for B, T in ((3, 50), (2, 200), (2, 100)):
        assert _fusion_planes_case("fp16x3", B, T, 1024) <= FUSION_TOL["fp16x3"]


def test_fusion_attention_planes_rejects_long_sequences():
    prec = F_.get_precision("fp16x3")
    assert not F_.fusion_attention_planes_supported(prec, 257, 1024)
    assert not F_.fusion_attention_planes_supported(F_.get_precision("fp32"), 50, 1024)


@pytest.mark.parametrize("B,T,C", [(2, 21, 256), (8, 200, 256), (3, 5, 512), (1, 2, 512)])
def test_pool_pairs_exact(B, T, C):
    prec = F_.get_precision("fp16x3")
    x = rnd(B * T, C, seed=1)
    out = F_.pool_pairs(prec, x, B, T)
    ref = O.max_pool_pairs(x.view(B, T, C)).reshape(-1, C)
    assert torch.equal(out.f32, ref)
    assert torch.equal(out.planes[0], ref.to(torch.float16))
    with pytest.raises(RuntimeError):
        F_.pool_pairs(prec, x[:B], B, 1)


def test_rowwise_layernorm_widths():
    prec = F_.get_precision("fp16x3")
    for n in (256, 512, 1024):
        z, r = rnd(77, n, seed=1, scale=2.0), rnd(77, n, seed=2)
        ln = torch.nn.LayerNorm(n)
        with torch.no_grad():
            ln.weight.copy_(1.0 + 0.2 * (torch.rand(n, generator=torch.Generator().manual_seed(3)) - 0.5))
            ln.bias.copy_(0.1 * (torch.rand(n, generator=torch.Generator().manual_seed(4)) - 0.5))
        ln = ln.to(DEV)
        out = F_.rowwise(prec, z, F_.make_epilogue(layer_norm=True, residual_mode=L.RES_AFTER_LN, act_post=L.ACT_RELU), ln, r)
        ref = torch.relu(torch.nn.functional.layer_norm(z.double(), (n,), ln.weight.double(), ln.bias.double(), 1e-5) + r.double())
        assert float((out.f32.double() - ref).abs().max()) <= 1e-5


def test_l2_prefetch_is_a_pure_hint():
    """scatt_l2_prefetch touches no data: buffers of any 16-byte-multiple size (also > one 16 KB piece, also empty),
    one launch for hundreds of buffers; a misaligned buffer is refused."""
    ts = [torch.arange(n, dtype=torch.float32, device=DEV) for n in (4, 4096, 5000, 70000)] + [torch.empty(0, device=DEV)]
    ts += [torch.full((256,), float(i), device=DEV) for i in range(300)]
    before = [t.clone() for t in ts]
    n0 = L.launch_count()
    F_.l2_prefetch(ts)
    torch.cuda.synchronize()
    assert L.launch_count() - n0 == 1 and L.load().scatt_last_kernel().decode() == "l2_prefetch_kernel"
    assert all(torch.equal(a, b) for a, b in zip(ts, before))
    with pytest.raises(L.ScattError):
        F_.l2_prefetch([torch.zeros(64, device=DEV)[1:]])


@pytest.mark.parametrize("B,T", [(8, 200), (3, 50), (5, 77), (1, 128), (2, 300), (40, 200)])
def test_frontend_tensor_core(B, T, monkeypatch):
    monkeypatch.setenv("SCATT_FRONTEND_TC", "2")  # from 128 frames up (the default hands batches below ~31 x 200 to the CUDA-core kernel)
    """Planes-only output of >= 128 frames runs the mapping on tcgen05 (frontend_tc_kernel): oracle parity within the
    split-plane class (hi + lo carries 22 bits), every tile shape: M tail, sequences wrapping inside a tile, one and
    two k-steps (6 / 21 joints), an unsorted index list, both self_attn_x settings."""
    from scattennet_b200.config import model_config
    from scattennet_b200.keypoint_module import KeypointModule, frontend_forward

    cfg = model_config("phoenix-2014t")
    if T > cfg["max_position_embeddings"]:
        cfg = dict(cfg, max_position_embeddings=512)
    kp, _ = synth.synth_batch(B, T, seed=5)
    mods, idxs = [], []
    for i, part in enumerate(("body", "left", "right")):
        m = KeypointModule(cfg[part + "_idx"], cfg["num_frame"], dict(cfg, self_attn_x=(i != 1))).eval()
        synth.load_synth_(m, 40 + i)
        mods.append(m.to(DEV))
        idxs.append(torch.tensor(cfg[part + "_idx"], dtype=torch.int32, device=DEV))
    idxs[0] = torch.tensor([500, 3, 541, 0, 77, 12], dtype=torch.int32, device=DEV)
    prec = F_.get_precision("fp16x3")
    s, c, _ = frontend_forward(prec, mods, kp.to(DEV), idxs, B, T)
    assert L.load().scatt_last_kernel().decode() == "frontend_tc_kernel"
    sf, cf, _ = frontend_forward(prec, mods, kp.to(DEV), idxs, B, T, want_f32=True)  # the CUDA-core kernel (exact fp32 order)
    assert L.load().scatt_last_kernel().decode().startswith("frontend_kernel")
    for g, m in enumerate(mods):
        region = kp[:, :, idxs[g].cpu().long(), :]
        sd = {"m." + k: v.cpu() for k, v in m.state_dict().items()}
        xe, ye = O.coordinate_mapping(sd, "m.coordinate_mapping", region)
        s_in, c_in = (xe, ye) if m.sca.x_self else (ye, xe)
        s_ref = O.layer_norm(sd, "m.sca.first_self_norm", O.position_embed(sd, "m.sca.self_pos_embed", s_in))
        c_ref = O.layer_norm(sd, "m.sca.first_causal_norm", O.position_embed(sd, "m.sca.causal_pos_embed", c_in))
        for act, ref, f32 in ((s[g], s_ref, sf[g]), (c[g], c_ref, cf[g])):
            assert act.f32 is None
            got = act.planes[0].float() + act.planes[1].float()
            assert float((got.cpu().view(B, T, -1) - ref).abs().max()) <= 2e-5
            assert float((got - f32.f32).abs().max()) <= 2e-5


def test_frontend_gather_exact_and_embeddings():
    from scattennet_b200.config import model_config
    from scattennet_b200.keypoint_module import KeypointModule, frontend_forward

    cfg = model_config("phoenix-2014t")
    B, T = 3, 37
    kp, _ = synth.synth_batch(B, T, seed=2)
    mods, idxs = [], []
    for i, part in enumerate(("body", "left", "right")):
        m = KeypointModule(cfg[part + "_idx"], cfg["num_frame"], dict(cfg, self_attn_x=(i != 1))).eval()
        synth.load_synth_(m, 30 + i)
        mods.append(m.to(DEV))
        idxs.append(torch.tensor(cfg[part + "_idx"], dtype=torch.int32, device=DEV))
    # a non-contiguous, unsorted index list must work too (the API is an arbitrary list)
    idxs[0] = torch.tensor([500, 3, 541, 0, 77, 12], dtype=torch.int32, device=DEV)
    prec = F_.get_precision("fp16x3")
    s, c, gathered = frontend_forward(prec, mods, kp.to(DEV), idxs, B, T, want_gathered=True, want_f32=True)
    s2, c2, _ = frontend_forward(prec, mods, kp.to(DEV), idxs, B, T)  # the hot path's form: split planes only
    for g, m in enumerate(mods):
        idx = idxs[g].cpu().long()
        region = kp[:, :, idx, :]
        assert torch.equal(gathered[g].cpu(), region)  # bit-exact region gather
        sd = {"m." + k: v.cpu() for k, v in m.state_dict().items()}
        xe, ye = O.coordinate_mapping(sd, "m.coordinate_mapping", region)
        s_in, c_in = (xe, ye) if m.sca.x_self else (ye, xe)
        s_ref = O.layer_norm(sd, "m.sca.first_self_norm", O.position_embed(sd, "m.sca.self_pos_embed", s_in))
        c_ref = O.layer_norm(sd, "m.sca.first_causal_norm", O.position_embed(sd, "m.sca.causal_pos_embed", c_in))
        assert float((s[g].f32.cpu().view(B, T, -1) - s_ref).abs().max()) <= 1e-5
        assert float((c[g].f32.cpu().view(B, T, -1) - c_ref).abs().max()) <= 1e-5
        assert s2[g].f32 is None and torch.equal(s2[g].planes, s[g].planes) and torch.equal(c2[g].planes, c[g].planes)
    with pytest.raises(IndexError):
        big, _ = synth.synth_batch(1, 257, seed=2)
        frontend_forward(prec, mods, big.to(DEV), idxs, 1, 257)


def test_stream_attention_planes_automatic_schedule():
    """4160 items (>= 4096): the automatic choice is the persistent kernel; same result as the one-item schedule bit for bit."""
    B, T, H, D = 130, 200, 16, 256
    prec = F_.get_precision("fp16x3")
    planes = F_.split_planes(rnd(B * T, 3 * D, seed=11), prec)
    mask = (torch.arange(T)[None] < torch.tensor(synth.parity_lengths(B, T))[:, None]).long()
    km = F_.key_mask_u8(mask.to(DEV))
    run = lambda: F_.stream_attention_planes(prec, [(planes, 0)], [(planes, D)], [(planes, 2 * D)], B, T, T, H, 1, key_mask=km)[0]
    auto = run()
    assert L.load().scatt_last_kernel().decode().startswith("stream_attention_fa2_kernel")
    L.check(L.load().scatt_debug_set_attn_persist(2), "set_attn_persist")
    try:
        one = run()
        assert L.load().scatt_last_kernel().decode().startswith("stream_attention_fa_kernel")
    finally:
        L.check(L.load().scatt_debug_set_attn_persist(0), "set_attn_persist")
    assert torch.equal(auto.planes, one.planes)
