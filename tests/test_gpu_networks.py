"""GPU parity tests for stage 4 (tcgen05 GEMM in both operand formats, LayerNorm, attention, ViT, MAE) and the
end-to-end Annotator, against torch fp32/fp64 references of the same op, the CPU oracle, and the
reference-generated golden fixtures.

Tolerances (north_star): float probabilities within 1e-3 absolute of the reference; labels exact.
The bf16x3 GEMM itself is held to a much tighter bound: |err| <= 2e-5 * sum_k |a||w| (dropped lo*lo
terms + fp32 accumulation); LayerNorm / attention outputs to 2e-6 relative of their split-bf16 range."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ribca_oracle as orc                                  # checker only
from multiplexed_image_annotator_b200 import engine, ops, synth, weights
from multiplexed_image_annotator_b200.cell_type_annotation import model as bmodel
from multiplexed_image_annotator_b200.cell_type_annotation import markerImputer as bimputer

DEV = "cuda"


def _unsplit(t):
    return t[0].double() + t[1].double()


def _unsplit_f16f8(t):
    """Value an f16f8 A-role operand stands for: fp16 plane + e4m3(low byte of the pair plane) / 2^8."""
    hi = t[0].view(torch.float16).double()
    lo = (t[1].view(torch.uint8).reshape(t[1].shape + (2,))[..., 0].contiguous().view(torch.float8_e4m3fn)).double()
    return hi + lo / 256.0


def _gemm_case(m, n, k, precision, epilogue=ops.EPI_STORE, bias=True, table_period=0, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    a = torch.randn((m, k), generator=g, device=DEV)
    w = torch.randn((n, k), generator=g, device=DEV) * 0.05
    b = torch.randn(n, generator=g, device=DEV) if bias else None
    tab = torch.randn((table_period, n), generator=g, device=DEV) if table_period else None
    if precision == "f16f8":
        return _gemm_case_f16f8(a, w, b, tab, epilogue, g)
    a_s, w_s = ops.split_bf16(a), ops.split_bf16(w)
    ref = _unsplit(a_s) @ _unsplit(w_s).T
    bound = (_unsplit(a_s).abs() @ _unsplit(w_s).abs().T)
    if b is not None:
        ref = ref + b.double()
    if tab is not None:
        ref = ref + tab.double()[torch.arange(m, device=DEV) % table_period]
    out0 = None
    if epilogue == ops.EPI_RESIDUAL:
        out0 = torch.randn((m, n), generator=g, device=DEV)
        ref = ref + out0.double()
        out = ops.gemm(a_s, w_s, b, tab, epilogue, out=out0.clone(), precision=precision)
        got = out.double()
    elif epilogue == ops.EPI_GELU:
        ref = torch.nn.functional.gelu(ref)
        got = _unsplit(ops.gemm(a_s, w_s, b, tab, epilogue, precision=precision))
    else:
        got = ops.gemm(a_s, w_s, b, tab, epilogue, precision=precision).double()
    torch.cuda.synchronize()
    err = (got - ref).abs()
    rel = (err / bound.clamp_min(1e-6)).max().item()
    return err.max().item(), rel, ref.abs().max().item()


def _gemm_case_f16f8(a, w, b, tab, epilogue, g):
    """fp16 main pass + e4m3 correction pass against the exact fp64 product of the fp32 operands."""
    m, n = a.shape[0], w.shape[0]
    t = ops.weight_log2_scale(float(w.abs().max().item()))
    a_s = ops.split_planes(a, ops.FMT_F16F8)
    w_s = ops.split_planes(w, ops.FMT_F16F8, w_role=True, log2_scale=t)
    ref = a.double() @ w.double().T
    bound = a.double().abs() @ w.double().abs().T
    if b is not None:
        ref = ref + b.double()
    if tab is not None:
        ref = ref + tab.double()[torch.arange(m, device=DEV) % tab.shape[0]]
    kw = dict(precision="f16f8", w_log2_scale=t + 8)
    if epilogue == ops.EPI_RESIDUAL:
        out0 = torch.randn((m, n), generator=g, device=DEV)
        ref = ref + out0.double()
        got = ops.gemm(a_s, w_s, b, tab, epilogue, out=out0.clone(), **kw).double()
    elif epilogue == ops.EPI_GELU:
        ref = torch.nn.functional.gelu(ref)
        got = _unsplit_f16f8(ops.gemm(a_s, w_s, b, tab, epilogue, **kw))
    elif epilogue == ops.EPI_STORE_SPLIT:
        got = _unsplit(ops.gemm(a_s, w_s, b, tab, epilogue, **kw))
    else:
        got = ops.gemm(a_s, w_s, b, tab, epilogue, **kw).double()
    torch.cuda.synchronize()
    err = (got - ref).abs()
    rel = (err / bound.clamp_min(1e-6)).max().item()
    return err.max().item(), rel, ref.abs().max().item()


GEMM_SHAPES = [(128, 192, 64), (128, 16, 48), (256, 576, 576), (101 * 5, 1728, 576), (1000, 288, 144), (303, 1600, 512),
               (77, 2304, 576), (4096, 576, 2304), (640, 768, 1600), (129, 144, 112)]


@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
def test_gemm_simt_matches_fp64(m, n, k):
    err, rel, mag = _gemm_case(m, n, k, "simt")
    print(f"simt  M={m} N={n} K={k}: max|err|={err:.3e} rel-to-bound={rel:.3e} |ref|max={mag:.2f}")
    assert rel < 2e-6


@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
def test_gemm_tcgen05_bf16x3(m, n, k):
    err, rel, mag = _gemm_case(m, n, k, "bf16x3")
    print(f"bf16x3 M={m} N={n} K={k}: max|err|={err:.3e} rel-to-bound={rel:.3e} |ref|max={mag:.2f}")
    assert rel < 2e-5


@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
def test_gemm_tcgen05_f16f8(m, n, k):
    """Default precision: the two correction terms are rounded to e4m3 (2^-4 of a 2^-11 term each), so the error
    is bounded by 2^-14 * sum |a||w| in the worst case and is ~1e-5 of it for random operands."""
    err, rel, mag = _gemm_case(m, n, k, "f16f8")
    print(f"f16f8 M={m} N={n} K={k}: max|err|={err:.3e} rel-to-bound={rel:.3e} |ref|max={mag:.2f}")
    assert rel < 2.0 ** -14


@pytest.mark.parametrize("epilogue,period", [(ops.EPI_RESIDUAL, 0), (ops.EPI_GELU, 0), (ops.EPI_STORE, 101), (ops.EPI_STORE_SPLIT, 0)])
def test_gemm_tcgen05_f16f8_epilogues(epilogue, period):
    err, rel, mag = _gemm_case(101 * 7, 576, 288, "f16f8", epilogue, True, period, seed=3)
    print(f"f16f8 epilogue {epilogue} period {period}: max|err|={err:.3e} rel={rel:.3e}")
    assert rel < 2.0 ** -13        # + the output's own quantisation (2^-15 relative for the f16f8 planes)


def _planes_value(t, precision):
    return _unsplit_f16f8(t) if precision == "f16f8" else _unsplit(t)


@pytest.mark.parametrize("precision", ["f16f8", "bf16x3", "simt"])
@pytest.mark.parametrize("m,d,k,resid,period", [(101 * 9, 576, 576, True, 0), (333, 288, 1152, True, 0), (101 * 3, 576, 240, False, 101),
                                                (1000, 144, 144, True, 0), (260, 384, 384, True, 0)])
def test_gemm_ln_producer_epilogues(precision, m, d, k, resid, period):
    """RIBCA_EPI_RESIDUAL_LN / STORE_LN (proj, fc2, patch embedding of the LayerNorm-folded flow): the fp32 rows are
    x_old + (A W^T + b) with the same two roundings as the reduce-add epilogue, the planes are the split of exactly the stored
    values, and the slot sums add up to the row's sum / sum of squares."""
    g = torch.Generator(device=DEV).manual_seed(m + d)
    a = torch.randn((m, k), generator=g, device=DEV)
    w = torch.randn((d, k), generator=g, device=DEV) * 0.05
    b = None if period else torch.randn(d, generator=g, device=DEV)
    tab = torch.randn((period, d), generator=g, device=DEV) if period else None
    x0 = torch.randn((m, d), generator=g, device=DEV) * 3 + 0.5
    fmt = ops.plane_format(precision)
    kw = dict(precision=precision)
    if precision == "f16f8":
        t = ops.weight_log2_scale(float(w.abs().max().item()))
        a_s, w_s = ops.split_planes(a, fmt), ops.split_planes(w, fmt, w_role=True, log2_scale=t)
        kw["w_log2_scale"] = t + 8
    else:
        a_s, w_s = ops.split_bf16(a), ops.split_bf16(w)
    # the plain epilogue of the same GEMM is the reference for the fp32 rows: bit-identical
    plain = ops.gemm(a_s, w_s, b, tab, ops.EPI_RESIDUAL if resid else ops.EPI_STORE, out=x0.clone() if resid else None, **kw)
    x, planes, stats, slots = ops.gemm_ln(a_s, w_s, b, tab, ops.EPI_RESIDUAL_LN if resid else ops.EPI_STORE_LN,
                                          out=x0.clone() if resid else None, **kw)
    torch.cuda.synchronize()
    assert torch.equal(x, plain), f"fp32 rows differ from the plain epilogue: max {(x - plain).abs().max().item():.3e}"
    want_planes = ops.split_planes(x, fmt) if fmt == ops.FMT_F16F8 else ops.split_bf16(x)
    assert torch.equal(planes.view(torch.int16), want_planes.view(torch.int16)), "planes are not the split of the stored rows"
    assert 1 <= slots <= 8 and torch.all(stats[:, slots:] == 0)
    s = stats[:, :slots].double().sum(1)
    xd = x.double()
    assert (s[:, 0] - xd.sum(1)).abs().max().item() < 1e-4 * max(1.0, xd.abs().sum(1).max().item()) * 1e-1
    assert ((s[:, 1] - (xd * xd).sum(1)).abs() / (xd * xd).sum(1)).max().item() < 5e-6
    # bit-reproducible statistics: a second launch with another M (another tile schedule) gives the same bits for the shared rows
    m2 = m - 37
    x2, _, stats2, _ = ops.gemm_ln(a_s[:, :m2].contiguous(), w_s, b, tab, ops.EPI_RESIDUAL_LN if resid else ops.EPI_STORE_LN,
                                   out=x0[:m2].clone() if resid else None, **kw)
    assert torch.equal(x2, x[:m2]) and torch.equal(stats2, stats[:m2])


@pytest.mark.parametrize("precision", ["f16f8", "bf16x3", "simt"])
@pytest.mark.parametrize("epilogue", [ops.EPI_STORE, ops.EPI_GELU, ops.EPI_STORE_SPLIT])
@pytest.mark.parametrize("m,d,n", [(101 * 6, 576, 1728), (515, 288, 1152), (300, 144, 576)])
def test_gemm_ln_consumer_matches_layernorm_then_linear(precision, epilogue, m, d, n):
    """LayerNorm folded into its consumer GEMM (qkv, fc1): raw-x planes + W * diag(gamma) + the statistics written by a producer
    epilogue give LayerNorm(x) W^T + b within the operand format's error bound (fp64 reference of the unfolded formula)."""
    g = torch.Generator(device=DEV).manual_seed(m + n)
    gamma = 1 + 0.1 * torch.randn(d, generator=g, device=DEV)
    beta = 0.1 * torch.randn(d, generator=g, device=DEV)
    w = torch.randn((n, d), generator=g, device=DEV) * 0.05
    b = torch.randn(n, generator=g, device=DEV) * 0.1
    # x with a row-dependent offset and scale, produced by a *_LN epilogue (identity-like GEMM would do; use a real one)
    a0 = torch.randn((m, 64), generator=g, device=DEV)
    w0 = torch.randn((d, 64), generator=g, device=DEV) * 0.3
    x0 = (torch.randn((m, d), generator=g, device=DEV) * (0.5 + torch.rand((m, 1), generator=g, device=DEV) * 3)
          + torch.randn((m, 1), generator=g, device=DEV))
    fmt = ops.plane_format(precision)
    kw = dict(precision=precision)
    wg = w * gamma[None, :]
    if precision == "f16f8":
        t0 = ops.weight_log2_scale(float(w0.abs().max().item()))
        x, xa, stats, slots = ops.gemm_ln(ops.split_planes(a0, fmt), ops.split_planes(w0, fmt, w_role=True, log2_scale=t0), None, None,
                                          ops.EPI_RESIDUAL_LN, out=x0.clone(), precision=precision, w_log2_scale=t0 + 8)
        t = ops.weight_log2_scale(float(wg.abs().max().item()))
        w_s = ops.split_planes(wg, fmt, w_role=True, log2_scale=t)
        kw["w_log2_scale"] = t + 8
    else:
        x, xa, stats, slots = ops.gemm_ln(ops.split_bf16(a0), ops.split_bf16(w0), None, None, ops.EPI_RESIDUAL_LN, out=x0.clone(),
                                          precision=precision)
        w_s = ops.split_bf16(wg)
    c1 = wg.double().sum(1).float()
    c2 = (b.double() + w.double() @ beta.double()).float()
    got = ops.gemm_ln(xa, w_s, c2, None, epilogue, stats_in=stats, c1=c1, slots_in=slots, **kw)
    torch.cuda.synchronize()
    xd = x.double()
    ln = torch.nn.functional.layer_norm(xd, (d,), gamma.double(), beta.double(), 1e-6)
    ref = ln @ w.double().T + b.double()
    rstd = 1.0 / torch.sqrt(xd.var(1, unbiased=False, keepdim=True) + 1e-6)
    bound = rstd * (xd.abs() @ wg.double().abs().T) + 1.0        # the contraction runs on raw x: its error scales with rstd * sum |x||W'|
    if epilogue == ops.EPI_GELU:
        ref = torch.nn.functional.gelu(ref)
        got = _planes_value(got, "f16f8" if precision == "f16f8" else "bf16x3")
    elif epilogue == ops.EPI_STORE_SPLIT:
        got = _unsplit(got)
    else:
        got = got.double()
    rel = ((got - ref).abs() / bound).max().item()
    print(f"ln-fold consumer {precision} epi {epilogue} M={m} D={d} N={n}: max|err|={(got - ref).abs().max().item():.3e} rel-to-bound={rel:.3e}")
    assert rel < {"f16f8": 2.0 ** -13, "bf16x3": 3e-5, "simt": 1e-5}[precision]


def test_f16f8_activation_planes():
    """LayerNorm / attention outputs in the f16f8 A-role format decode to the fp32 value within 2^-15 relative."""
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn((101 * 3, 576), generator=g, device=DEV) * 2 + 0.3
    gam = torch.randn(576, generator=g, device=DEV)
    bet = torch.randn(576, generator=g, device=DEV)
    want = torch.nn.functional.layer_norm(x.double(), (576,), gam.double(), bet.double(), 1e-6)
    got = _unsplit_f16f8(ops.layernorm_split(x, gam, bet, 1e-6, fmt=ops.FMT_F16F8))
    # e4m3 of the scaled remainder: 2^-4 relative in its normal range, 2^-10 absolute below 2^-6 (-> 2^-18 of x)
    assert ((got - want).abs() <= 2.0 ** -15 * want.abs() + 2.0 ** -18 + 3e-6).all()
    v = torch.randn((4096,), generator=g, device=DEV) * 3
    dec = _unsplit_f16f8(ops.split_planes(v, ops.FMT_F16F8))
    assert ((dec - v.double()).abs() <= 2.0 ** -15 * v.double().abs() + 2.0 ** -18).all()
    # the second e4m3 of each pair is the fp16 value rounded to 4 significant bits
    pl = ops.split_planes(v, ops.FMT_F16F8)
    hi8 = pl[1].view(torch.uint8).reshape(-1, 2)[:, 1].contiguous().view(torch.float8_e4m3fn).double()
    assert ((hi8 - v.double()).abs() <= 2.0 ** -4 * v.double().abs() + 2.0 ** -9).all()
    for cells, tokens, heads, hd in ((3, 101, 12, 48), (5, 7, 12, 64)):
        d = heads * hd
        qkv = torch.randn((cells * tokens, 3 * d), generator=g, device=DEV)
        q, k, vv = qkv.view(cells, tokens, 3, heads, hd).permute(2, 0, 3, 1, 4).double().unbind(0)
        want = torch.nn.functional.scaled_dot_product_attention(q, k, vv).transpose(1, 2).reshape(cells * tokens, d)
        got = _unsplit_f16f8(ops.attention(qkv, cells, tokens, heads, fmt=ops.FMT_F16F8))
        assert (got - want).abs().max().item() < 4e-5


@pytest.mark.parametrize("epilogue,period", [(ops.EPI_RESIDUAL, 0), (ops.EPI_GELU, 0), (ops.EPI_STORE, 101)])
def test_gemm_tcgen05_epilogues(epilogue, period):
    err, rel, mag = _gemm_case(101 * 7, 576, 288, "bf16x3", epilogue, True, period, seed=3)
    print(f"epilogue {epilogue} period {period}: max|err|={err:.3e} rel={rel:.3e}")
    assert rel < 3e-5


def test_gemm_tcgen05_bf16x1_is_single_pass():
    err3, rel3, _ = _gemm_case(512, 576, 576, "bf16x3", seed=5)
    err1, rel1, _ = _gemm_case(512, 576, 576, "bf16x1", seed=5)
    print(f"bf16x1 rel={rel1:.3e} vs bf16x3 rel={rel3:.3e}")
    assert rel1 < 1e-2 and rel1 > 20 * rel3         # really a lower-precision pass, not the same kernel path


def test_layernorm_and_attention_vs_torch():
    g = torch.Generator(device=DEV).manual_seed(1)
    for m, d in ((101 * 3, 576), (50, 144), (257, 768), (16, 512)):
        x = torch.randn((m, d), generator=g, device=DEV) * 2 + 0.3
        gam = torch.randn(d, generator=g, device=DEV)
        bet = torch.randn(d, generator=g, device=DEV)
        want = torch.nn.functional.layer_norm(x.double(), (d,), gam.double(), bet.double(), 1e-6)
        got = _unsplit(ops.layernorm_split(x, gam, bet, 1e-6))
        assert (got - want).abs().max().item() < 3e-5 * max(1.0, want.abs().max().item())
    for cells, tokens, heads, hd in ((3, 101, 12, 48), (2, 101, 12, 12), (4, 101, 12, 24), (2, 101, 12, 32), (5, 7, 12, 64), (3, 16, 8, 64)):
        d = heads * hd
        qkv = torch.randn((cells * tokens, 3 * d), generator=g, device=DEV)
        q, k, v = qkv.view(cells, tokens, 3, heads, hd).permute(2, 0, 3, 1, 4).double().unbind(0)
        want = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(cells * tokens, d)
        got = _unsplit(ops.attention(qkv, cells, tokens, heads))
        err = (got - want).abs().max().item()
        print(f"attention cells={cells} tokens={tokens} hd={hd}: max|err|={err:.3e}")
        assert err < 2e-5


@pytest.mark.parametrize("cells,tokens,heads,hd", [(3, 101, 12, 48), (2, 101, 12, 12), (4, 101, 12, 24), (2, 101, 12, 32),
                                                   (3, 101, 12, 64), (2, 40, 4, 16), (1, 112, 2, 64), (5, 97, 3, 48), (300, 101, 12, 48), (2, 112, 12, 48), (7, 100, 5, 48)])
def test_attention_tensor_core_vs_torch(cells, tokens, heads, hd):
    g = torch.Generator(device=DEV).manual_seed(cells * 1000 + tokens + hd)
    m, hdp = cells * tokens, (hd + 15) // 16 * 16
    qkv = torch.randn((m, 3, heads, hd), generator=g, device=DEV) * 1.5
    padded = torch.zeros((m, 3, heads, hdp), device=DEV)
    padded[..., :hd] = qkv
    qs = ops.split_bf16(padded.reshape(m, 3 * heads * hdp))
    exact = _unsplit(qs).reshape(m, 3, heads, hdp)[..., :hd]           # what the kernel really sees
    q, k, v = exact.view(cells, tokens, 3, heads, hd).permute(2, 0, 3, 1, 4).unbind(0)
    want = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(m, heads * hd)
    got = _unsplit(ops.attention_tc(qs, cells, tokens, heads, hd))
    got8 = _unsplit_f16f8(ops.attention_tc(qs, cells, tokens, heads, hd, fmt=ops.FMT_F16F8))
    torch.cuda.synchronize()
    assert (got8 - want).abs().max().item() < 2.0 ** -14 * max(1.0, want.abs().max().item())
    err = (got - want).abs().max().item()
    print(f"attention_tc cells={cells} tokens={tokens} heads={heads} hd={hd}: max|err|={err:.3e} |ref|max={want.abs().max().item():.2f}")
    # the output itself is quantised to split-bf16 (16 mantissa bits): 2^-16 of its magnitude
    assert err < 2.0 ** -16 * max(1.0, want.abs().max().item()) * 1.5


def _vit_pair(panel, seed=1):
    sd = weights.random_vit_state(panel, seed=seed)
    ref = orc.make_vit(panel)
    ref.load_state_dict(sd)
    return sd, ref


@pytest.mark.parametrize("panel", ["immune_base", "nerve_cell"])
def test_vit_reference_golden(golden_dir, panel):
    g = np.load(os.path.join(golden_dir, "vit.npz"))
    sd, _ = _vit_pair(panel)
    engines = {"f16f8": engine.VitEngine(panel, sd, DEV), "bf16x3": engine.VitEngine(panel, sd, DEV, precision="bf16x3")}
    engines["simt"] = engines["bf16x3"]                 # same bf16 {hi, lo} planes on the FP32 pipe
    assert engines["f16f8"].precision == "f16f8"        # the default
    x = torch.from_numpy(g[panel + "_x"]).to(DEV)
    for prec, tol_logit, tol_prob in (("f16f8", 6e-4, 3e-4), ("bf16x3", 2e-4, 1e-4), ("simt", 5e-5, 2e-5)):
        probs, logits = engines[prec].forward(x, return_logits=True, precision=prec)
        dl = np.abs(logits.cpu().numpy() - g[panel + "_logits"]).max()
        dp = np.abs(probs.cpu().numpy() - g[panel + "_probs"]).max()
        print(f"{panel} {prec}: max|dlogit|={dl:.3e} max|dprob|={dp:.3e}")
        assert dl < tol_logit and dp < tol_prob


def test_engine_packs_every_precision_on_demand():
    """One engine serves every precision (the re-evaluation levels of exact.py ask for bf16x3 / fp32 on an f16f8 engine):
    the weights are re-packed lazily per operand format and give the same bits as a dedicated engine."""
    sd, _ = _vit_pair("nerve_cell")
    eng = engine.VitEngine("nerve_cell", sd, DEV)           # f16f8 planes
    x = torch.randn((5, 3, 40, 40), device=DEV)
    for prec in ("bf16x3", "f16f8", "fp32"):
        assert torch.equal(eng.forward(x, precision=prec), engine.VitEngine("nerve_cell", sd, DEV, precision=prec).forward(x))
    with pytest.raises(KeyError):
        eng.forward(x, precision="fp8")


@pytest.mark.parametrize("precision", ["f16f8", "bf16x3", "simt"])
@pytest.mark.parametrize("panel", ["immune_full", "immune_base", "nerve_cell"])
def test_vit_layernorm_fold_masks_agree(panel, precision, monkeypatch):
    """RIBCA_LN_FOLD: 0 = LayerNorm kernels, 1 = norm1 inside the qkv GEMM (default), 3 = norm1 and norm2 inside qkv / fc1.
    The same network in the three schedules: probabilities within the operand format's error of each other and of the
    fp64-accurate fp32 path (the folded contraction runs on raw x and W * diag(gamma) instead of on LayerNorm(x) and W)."""
    sd, _ = _vit_pair(panel, seed=11)
    spec = weights.VIT_SPECS[panel]
    x = torch.rand((150, spec.in_chans, 40, 40), device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
    out = {}
    for mask in (0, 1, 3):
        monkeypatch.setenv("RIBCA_LN_FOLD", str(mask))
        eng = engine.VitEngine(panel, sd, DEV, precision=precision)
        assert eng.ln_fold == mask and eng.desc.ln_folded == mask
        out[mask] = eng.forward(x)
        if mask == 0:
            exact = eng.forward(x, precision="fp32")        # plain fp32 FMA path (never folded)
    tol = {"f16f8": 4e-4, "bf16x3": 1.5e-4, "simt": 5e-5}[precision]
    for mask in (0, 1, 3):
        d = (out[mask] - exact).abs().max().item()
        print(f"{panel} {precision} fold mask {mask}: max|dprob| vs fp32 path {d:.3e}")
        assert d < tol
    assert not torch.equal(out[0], out[1])                  # the folded schedule really is another computation


@pytest.mark.parametrize("n", [512, 777, 1300])
def test_vit_interleaved_halves_are_bit_identical_to_the_serial_schedule(n):
    """ribca_set_interleave: a call of >= 512 cells is split in two halves on two streams (LayerNorm of one half under the
    GEMM of the other).  Every row goes through the same kernels, so the probabilities must carry the same bits; the
    work queued on the caller's stream after the call must see the side stream's results (join)."""
    sd, _ = _vit_pair("immune_base", seed=5)
    eng = engine.VitEngine("immune_base", sd, DEV)
    x = torch.randn((n, 7, 40, 40), device=DEV, generator=torch.Generator(device=DEV).manual_seed(n))
    try:
        ops.set_interleave(False)
        serial, serial_logits = eng.forward(x, return_logits=True)
        ops.set_interleave(True)
        for _ in range(3):
            both, both_logits = eng.forward(x, return_logits=True)
            total = both.sum(1)                                  # consumer on the caller's stream right after the join
            assert torch.equal(both, serial) and torch.equal(both_logits, serial_logits)
            assert torch.allclose(total, torch.ones_like(total), atol=1e-5)
    finally:
        ops.set_interleave(False)


@pytest.mark.parametrize("precision", ["f16f8", "bf16x3"])
@pytest.mark.parametrize("panel", ["immune_base", "immune_extended", "immune_full", "structure", "nerve_cell"])
def test_vit_vs_oracle_on_real_patches(panel, precision):
    spec = weights.VIT_SPECS[panel]
    mask = synth.synth_mask(160, 160, seed=6)
    img = orc.normalize(synth.to_uint16(synth.synth_image(mask, spec.in_chans, seed=6)), 0.3, 99.8)
    patches, _, _ = orc.build_patches(img, mask.numpy(), list(range(spec.in_chans)))
    sd, ref = _vit_pair(panel, seed=3)
    with torch.no_grad():
        mean_logits = ref(torch.from_numpy(patches[:32])).mean(0).numpy()
    sd = weights.calibrate_head(sd, mean_logits, 20.0)
    ref.load_state_dict(sd)
    want = orc.vit_probs(ref, patches)
    eng = engine.VitEngine(panel, sd, DEV, max_cells_per_call=50, precision=precision)      # exercises chunking
    got = eng.forward(torch.from_numpy(patches).to(DEV)).cpu().numpy()
    dp = np.abs(got - want).max()
    flips = int((got.argmax(1) != want.argmax(1)).sum())
    gap = np.sort(want, 1)
    gap = gap[:, -1] - gap[:, -2]
    print(f"{panel} {precision}: cells={len(want)} max|dprob|={dp:.3e} argmax flips={flips} min top-2 gap={gap.min():.3e} "
          f"label histogram={np.bincount(want.argmax(1), minlength=want.shape[1]).tolist()}")
    assert dp < 1e-3
    assert flips <= int((gap < 2 * dp).sum())          # raw forward: a flip is only possible inside the error band
    # ... and with the margin-guarded re-evaluation (exact.py) on top, the merged labels are the oracle's, no allowance
    from multiplexed_image_annotator_b200 import exact
    from multiplexed_image_annotator_b200.cell_type_annotation.model import ALL_TYPES, merge_on_device
    x = torch.from_numpy(patches).to(DEV)
    label, conf, counts, margin, st = exact.refine_labels(
        {panel: torch.from_numpy(got).to(DEV)}, lambda pr, want_margin=True: merge_on_device(pr, 0.3, None, want_margin=want_margin),
        lambda sel, prec: {panel: eng.forward(x[sel], precision=prec)})
    want_lab, want_conf = orc.merge_by_voting({panel: want}, 0.3, None)
    assert [ALL_TYPES[k] for k in label.cpu().tolist()] == want_lab, st.as_dict()
    assert np.abs(conf.cpu().numpy() - np.array([float(c) for c in want_conf], dtype=np.float32)).max() < 1e-3


@pytest.mark.parametrize("precision", ["f16f8", "bf16x3"])
@pytest.mark.parametrize("panel,present", [("immune_base", [0, 1, 2, 3, 4, 6]), ("immune_extended", [0, 1, 2, 4, 5, 6, 7, 9]),
                                           ("immune_full", [0, 1, 2, 3, 4, 6, 7, 8, 9, 11, 13, 14])])
def test_mae_vs_oracle_and_golden(golden_dir, panel, present, precision):
    sd = weights.random_mae_state(panel, seed=1)
    eng = engine.MaeEngine(panel, sd, DEV, precision=precision)
    g = np.load(os.path.join(golden_dir, "mae.npz"))
    if panel + "_x" in g:
        x, want = g[panel + "_x"], g[panel + "_out"]
    else:
        spec = weights.MAE_SPECS[panel]
        gen = torch.Generator().manual_seed(8)
        x = (torch.rand((4, spec.channels, 40, 40), generator=gen) * 2 - 1).numpy()
        x[:, [c for c in range(spec.channels) if c not in present]] = -1
        ref = orc.make_mae(panel)
        ref.load_state_dict(sd)
        want = orc.impute(ref, x, present)
    got = eng.impute(torch.from_numpy(x.copy()).to(DEV), present).cpu().numpy()
    for c in present:
        assert np.array_equal(got[:, c], x[:, c])
    d = np.abs(got - want).max()
    print(f"mae {panel} {precision}: max|d|={d:.3e} range of imputed values [{want.min():.3f}, {want.max():.3f}]")
    assert d < 1e-3


@pytest.mark.parametrize("tag,strict", [("full", True), ("impute", False), ("struct_nerve", True)])
def test_annotator_end_to_end_golden(golden_dir, tag, strict, tmp_path, monkeypatch):
    """Annotator.preprocess() + predict() + export_annotations() against the reference's own run."""
    g = np.load(os.path.join(golden_dir, "e2e.npz"))
    monkeypatch.chdir(tmp_path)
    np.save("img.npy", g[tag + "_img"])
    np.save("mask.npy", g[tag + "_mask"])
    markers = [str(m) for m in g[tag + "_markers"]]
    synth.write_marker_file("markers.txt", markers)
    with open("images.csv", "w") as f:
        f.write("image_path,mask_path\nimg.npy,mask.npy\n")
    for panel in weights.VIT_SPECS:
        key = f"{tag}_meanlogits_{panel}"
        sd = weights.random_vit_state(panel, seed=2)
        if key in g:
            sd = weights.calibrate_head(sd, g[key], 20.0)
        bmodel.register_state(panel, sd)
    bimputer.register_state("immune_base", weights.random_mae_state("immune_base", seed=2))
    ann = bmodel.Annotator("markers.txt", "images.csv", "cuda", "./", "g", strict, True, -1, True, 0.3, 99.8, 0.3, 30, None, n_jobs=0)
    ann.preprocess()
    ann.predict(32)
    ann.export_annotations()
    # stages 1-3: bit-exact patches (the imputed channels within the network tolerance)
    for panel, pt in ann.preprocessor.patches[0].items():
        want = g[f"{tag}_patches_{panel}"]
        got = pt.cpu().numpy()
        if tag == "impute":
            assert np.array_equal(got[:, [0, 1, 2, 3, 4, 6]], want[:, [0, 1, 2, 3, 4, 6]])
            assert np.abs(got - want).max() < 1e-3
        else:
            assert np.array_equal(got, want)
    np.testing.assert_allclose(ann.preprocessor.intensity_full[0], g[tag + "_intensity"], rtol=1e-12, atol=1e-14)
    # stage 4: probabilities within 1e-3 of the reference
    worst = 0.0
    for panel, p in ann.probs[0].items():
        worst = max(worst, float(np.abs(p - g[f"{tag}_probs_{panel}"]).max()))
    print(f"e2e {tag}: max|dprob|={worst:.3e}")
    assert worst < 1e-3
    # stage 5 + result assembly: labels exact, confidences within the probability tolerance, CSV identical
    # wherever the printed 3-decimal confidence is not on a rounding boundary
    assert ann.annotations[0] == g[tag + "_labels"].tolist()
    conf = np.array([float(c) for c in ann.confidence[0]])
    assert np.abs(conf - g[tag + "_conf"]).max() < 1e-3
    assert [str(c) for c in ann.cell_types] == g[tag + "_cell_types"].tolist()
    got_rows = open("results/g_annotation_0.csv").read().strip().split("\n")
    want_rows = str(g[tag + "_csv"]).strip().split("\n")
    assert got_rows[0] == want_rows[0] and len(got_rows) == len(want_rows)
    for a, b in zip(got_rows[1:], want_rows[1:]):
        fa, fb = a.split(","), b.split(",")
        assert fa[:2] == fb[:2] and fa[3:] == fb[3:], (a, b)            # id, type, centroid, region identical
        assert abs(float(fa[2]) - float(fb[2])) <= 1.001e-3
    comp = ann.cell_type_composition(reduction=False)[0]
    assert sum(comp.values()) == len(ann.annotations[0])
    # colourised label map painted on the device: background 0, every pixel of a cell = its type colour
    from PIL import Image
    ann.colorize(from_script=True)
    rgb = np.array(Image.open("results/g_colorized_annotation_0.png"))
    mask = g[tag + "_mask"]
    assert rgb.shape == mask.shape + (3,) and (rgb[mask == 0] == 0).all()
    pos = ann.preprocessor.cell_pos_dict[0]
    for j, key in enumerate(list(pos)[:10]):
        rows, cols = pos[key]
        t = int(np.where(ann.cell_types == ann.annotations[0][j])[0][0])
        assert (rgb[rows, cols] == np.array(ann.colors[t], dtype=np.uint8)).all()
        assert (rows, cols) == tuple(a.tolist() for a in np.nonzero(mask == key))
    # heat-map reduction (reference model.py:718-727): mean intensity of the cells of each predicted type
    ann.generate_heatmap(integrate=False)
    types_h, hm = ann.heatmaps[0]
    for t, row in zip(types_h, hm):
        sel = [k for k in range(len(ann.annotations[0])) if ann.annotations[0][k] == t]
        np.testing.assert_allclose(row, np.mean([ann.preprocessor.intensity_full[0][k] for k in sel], axis=0), rtol=1e-13)
    # spatial statistics: the neighbourhood matrix equals the reference's loop over annotations_all with scikit-learn
    from sklearn.neighbors import NearestNeighbors
    m = ann.neighborhood_analysis(integrate=True, normalize=False)
    rows_all = list(ann.annotations_all[0])
    xy = np.array([[np.mean(r["Column"]), np.mean(r["Row"])] for r in rows_all])
    ty = np.array([r["Cell type"] for r in rows_all])
    idx = NearestNeighbors(n_neighbors=25, algorithm="ball_tree").fit(xy).kneighbors(xy, return_distance=False)
    want_m = np.zeros((len(ann.cell_types),) * 2)
    for j in range(len(xy)):
        for kk in idx[j][1:]:
            want_m[ty[j], ty[kk]] += 1
    assert np.array_equal(m, want_m) and os.path.exists("results/g_integrated_neighborhood.csv")
    with pytest.raises(ValueError):
        ann.tissue_region_analysis(3)            # 201 neighbours of < 201 cells: scikit-learn raises in the reference too
    ann.clear_tmp()
    assert not os.path.exists("tmp")


def _write_scene(tmp_path, name, size, n_markers, seed):
    mask = synth.synth_mask(size, size, grid=18, seed=seed)
    img = synth.to_uint16(synth.synth_image(mask, n_markers, seed=seed))
    np.save(tmp_path / f"{name}_img.npy", img)
    np.save(tmp_path / f"{name}_mask.npy", mask.numpy())
    return mask.numpy()


def test_cli_run_and_gui_api_full_sequence(tmp_path, monkeypatch):
    """main.run / main.batch_run / gui_api with the reference's fixed post-processing sequence (main.py:19-28,
    gui_api.py:19-31): heat map, annotation CSV, tissue regions, neighbourhood matrix, colourised maps, composition."""
    import json
    import main as cli
    from multiplexed_image_annotator_b200.cell_type_annotation import gui_api
    monkeypatch.chdir(tmp_path)
    from multiplexed_image_annotator_b200.cell_type_annotation.markerParse import MarkerParser
    markers = list(MarkerParser(strict=True).panels["immune_full"])
    synth.write_marker_file("markers.txt", markers)
    masks = [_write_scene(tmp_path, f"s{k}", 420, 15, 20 + k) for k in range(2)]
    for panel in weights.VIT_SPECS:
        bmodel.register_state(panel, weights.random_vit_state(panel, seed=4))
    ctc = {t: -1 for t in bmodel.ALL_TYPES}
    inten, names = cli.run("markers.txt", "s0_img.npy", "s0_mask.npy", "cuda", "./", "one", 64, True, True, -1, 3, True, 0.3, 99.8,
                           0.3, 30, ctc, 0)
    n0 = len(np.unique(masks[0])) - 1
    assert n0 > 250 and len(inten) == n0 + 1 and isinstance(names, str)
    rows = open("results/one_annotation_0.csv").read().strip().split("\n")
    assert rows[0] == "Cell Index,Cell Type,Confidence,Row,Column,Tissue Region" and len(rows) == n0 + 1
    nb = open("results/one_integrated_neighborhood.csv").read().strip().split("\n")
    assert nb[0].startswith("cell_type,") and len(nb) >= 2
    for f in ("one_colorized_annotation_0.png", "one_confidence_0.png", "one_tissue_region_0.png"):
        assert os.path.exists(os.path.join("results", f)), f
    assert not os.path.exists("tmp")
    # batch CSV through the CLI entry point, then the GUI JSON entry point (regions are computed before the export there)
    with open("batch.csv", "w") as f:
        f.write("image_path,mask_path\ns0_img.npy,s0_mask.npy\ns1_img.npy,s1_mask.npy\n")
    cli.batch_run("markers.txt", "batch.csv", "cuda", "./", "two", 64, True, True, -1, 0, True, 0.3, 99.8, 0.3, 30, ctc, 0)
    assert os.path.exists("results/two_annotation_1.csv") and not os.path.exists("results/two_tissue_region_0.png")
    os.makedirs("work", exist_ok=True)
    json.dump({"marker_file": "markers.txt", "image_file": "s1_img.npy", "mask_file": "s1_mask.npy", "device": "cuda", "main_dir": "./",
               "batch_size": 64, "strict": True, "infer": True, "min_cells": -1, "n_regions": 2, "normalize": True, "blur": 0.3,
               "upper_limit": 99.8, "confidence": 0.3, "cell_size": 30, "cell_type_confidence": ctc}, open("work/hyperparams.json", "w"))
    inten, names = gui_api.gui_api("work")
    rows = open("results/single_run_annotation_0.csv").read().strip().split("\n")
    assert {r.split(",")[-1] for r in rows[1:]} <= {"Region 0", "Region 1"}
    # the same image as a multi-page TIFF + TIFF mask goes through the streaming decoder (decode / upload / stage 1 overlapped
    # plane by plane) and gives the same CSV as the .npy input
    from PIL import Image
    stack = np.load("s1_img.npy")
    ims = [Image.fromarray(p) for p in stack]
    ims[0].save("s1.tif", save_all=True, append_images=ims[1:])
    Image.fromarray(np.load("s1_mask.npy").astype(np.int32)).save("s1_mask.tif")
    cli.run("markers.txt", "s1.tif", "s1_mask.tif", "cuda", "./", "tif", 64, True, True, -1, 0, True, 0.3, 99.8, 0.3, 30, ctc, 0)
    cli.run("markers.txt", "s1_img.npy", "s1_mask.npy", "cuda", "./", "npy", 64, True, True, -1, 0, True, 0.3, 99.8, 0.3, 30, ctc, 0)
    assert open("results/tif_annotation_0.csv").read() == open("results/npy_annotation_0.csv").read()


def test_c1_reference_example_configuration(golden_dir, tmp_path, monkeypatch):
    """BASELINE configs[0] on the GPU path: example mask (1850 cells) + examples/markers.txt -> vit_m + vit_s, merge branch 2,
    against the unmodified reference's CPU run (fixture c1.npz): labels, confidences, CSV rows, probabilities."""
    g = np.load(os.path.join(golden_dir, "c1.npz"))
    mask = g["mask"]
    markers = [str(m) for m in g["markers"]]
    img = synth.to_uint16(synth.synth_image(torch.from_numpy(mask), len(markers), seed=1))
    assert int(img.astype(np.int64).sum()) == int(g["image_checksum"][0])
    monkeypatch.chdir(tmp_path)
    np.save("img.npy", img); np.save("mask.npy", mask)
    synth.write_marker_file("markers.txt", markers)
    with open("images.csv", "w") as f:
        f.write("image_path,mask_path\nimg.npy,mask.npy\n")
    for panel in weights.VIT_SPECS:
        sd = weights.random_vit_state(panel, seed=2)
        if f"meanlogits_{panel}" in g:
            sd = weights.calibrate_head(sd, g[f"meanlogits_{panel}"], 20.0)
        bmodel.register_state(panel, sd)
    ann = bmodel.Annotator("markers.txt", "images.csv", "cuda", "./", "c1", True, True, -1, True, 0.3, 99.8, 0.3, 30, None, n_jobs=0)
    assert ann.preprocessor.predicted_panels() == ["immune_extended", "structure"]
    ann.preprocess()
    ann.predict(128)
    ann.export_annotations()
    worst = max(float(np.abs(ann.probs[0][p] - g[f"probs_{p}"]).max()) for p in ("immune_extended", "structure"))
    differ = [j for j, (a, b) in enumerate(zip(ann.annotations[0], g["labels"].tolist())) if a != b]
    print(f"C1: 1850 cells, max|dprob|={worst:.3e}, labels differing={len(differ)}")
    print(f"C1 re-evaluation: {ann.refine_stats[0].as_dict()}")
    assert worst < 1e-3
    # exact labels by construction (margin-guarded re-evaluation, exact.py).  What no implementation can reproduce is a
    # decision the REFERENCE's own fp32 arithmetic takes inside its rounding band: cell 508 of this fixture has a top-2 vote
    # gap of 7.0e-7 in the reference's probabilities (fp32 eps = 1.2e-7 per operation, measured fp32-vs-fp64 probability
    # error 2.5e-6).  So a label may differ only where the reference's decision margin is below 1e-5 - not 2 * 1e-3 as before.
    ref_dev = {p: torch.from_numpy(g[f"probs_{p}"]).to(DEV) for p in ("immune_extended", "structure")}
    ref_margin = bmodel.merge_on_device(ref_dev, 0.3, None, want_margin=True)[3].cpu().numpy()
    assert all(ref_margin[j] < 1e-5 for j in differ), [(j, float(ref_margin[j])) for j in differ]
    assert len(differ) <= int((ref_margin < 1e-5).sum()) <= 2
    conf = np.array([float(c) for c in ann.confidence[0]])
    keep = np.ones(len(conf), bool); keep[differ] = False
    assert np.abs(conf - g["conf"])[keep].max() < 1e-3
    assert [str(c) for c in ann.cell_types] == g["cell_types"].tolist()
    got_rows = open("results/c1_annotation_0.csv").read().strip().split("\n")
    want_rows = str(g["csv"]).strip().split("\n")
    assert len(got_rows) == len(want_rows) == 1851
    for j, (a, b) in enumerate(zip(got_rows[1:], want_rows[1:])):
        fa, fb = a.split(","), b.split(",")
        assert fa[0] == fb[0] and fa[3:] == fb[3:], (a, b)            # id, centroid, region identical
        if j not in differ:
            assert fa[1] == fb[1] and abs(float(fa[2]) - float(fb[2])) <= 1.001e-3
