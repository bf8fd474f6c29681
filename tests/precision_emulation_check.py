"""CPU emulation of the tensor-core operand formats on the whole vit_l (research script like tests/multi_gpu_check.py: not collected by pytest, not part of the product).

Quantises the operands of every Linear / attention matmul as the kernels would (bf16x1, fp16x1, bf16x3, fp16 + split weights,
fp16 + split activations, f16f8 = fp16 main pass + e4m3 correction pairs) and reports max |dprob| against plain fp32 on the cells
of a synthetic scene; `fold:<mode>` additionally applies LayerNorm algebraically in the consumer GEMM's epilogue
(LN(x) W^T = rstd (x (g*W)^T - mean c1) + c2 on RAW x operands), the formulation proposed in DESIGN.md "what comes next".
    python tests/precision_emulation_check.py 768 immune_full "bf16x3,bf16x1,fp16x1,f16f8:8/bf16x3,fold:f16f8:8/bf16x3"
Results behind DESIGN.md section 4 (1849 cells): bf16x3 5.6e-5, bf16x1 3.4e-2, fp16x1 3.7e-3, f16f8 1.3e-4, folded f16f8 1.2e-4.
"""
import sys, os, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from multiplexed_image_annotator_b200 import synth, weights
from oracle import ribca_oracle as orc
torch.set_num_threads(8)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 768
panel = sys.argv[2] if len(sys.argv) > 2 else "immune_full"
spec = weights.VIT_SPECS[panel]
mask = synth.synth_mask(S, S, grid=18, seed=2, device="cpu")
img = synth.synth_image(mask, 15, seed=2)
img_u16, mask_i32 = synth.to_uint16(img), mask.numpy()
norm = orc.normalize(img_u16, 0.3, 99.8)
idx = list(range(spec.in_chans))
t0 = time.time()
out = orc.build_patches(norm, mask_i32, idx)
patches = out[0] if isinstance(out, tuple) else out
if isinstance(patches, dict): patches = patches["patches"]
patches = torch.as_tensor(np.asarray(patches), dtype=torch.float32)
print("patches", patches.shape, time.time() - t0, flush=True)
sd = weights.random_vit_state(panel, seed=7)

def q_none(x): return x
def q_bf16(x): return x.bfloat16().float()
def q_fp16(x): return x.half().float()
def split(x, q):
    hi = q(x); lo = q(x - hi); return hi, lo

def make_mm(mode):
    if mode == "fp32": return lambda a, w: a @ w
    if mode == "bf16x1": return lambda a, w: q_bf16(a) @ q_bf16(w)
    if mode == "fp16x1": return lambda a, w: q_fp16(a) @ q_fp16(w)
    if mode == "bf16x3":
        def f(a, w):
            ah, al = split(a, q_bf16); wh, wl = split(w, q_bf16)
            return ah @ wh + (al @ wh + ah @ wl)
        return f
    if mode == "fp16x2w":   # activations split, weights hi only
        def f(a, w):
            ah, al = split(a, q_fp16); wh = q_fp16(w)
            return (ah + al) @ wh
        return f
    if mode == "fp16x2a":   # weights split, activations hi only
        def f(a, w):
            wh, wl = split(w, q_fp16); ah = q_fp16(a)
            return ah @ (wh + wl)
        return f
    if mode == "fp16x1_bf16lo2":  # fp16 hi*hi + cross terms with bf16 lo (K-concat emulation)
        def f(a, w):
            ah, al = split(a, q_fp16); wh, wl = split(w, q_fp16)
            return ah @ wh + (q_bf16(al) @ q_bf16(wh) + q_bf16(ah) @ q_bf16(wl))
        return f
    if mode.startswith("f16f8"):
        sa = int(mode.split(":")[1]) if ":" in mode else 8
        E = torch.float8_e4m3fn
        def q8(x): return x.clamp(-448, 448).to(E).float()
        def f(a, w):
            ah = q_fp16(a); al = a - ah
            wh = q_fp16(w); wl = w - wh
            m = w.abs().max().item()
            t = math.floor(math.log2(128.0 / m))
            main = ah @ (wh * 2.0 ** (sa + t))
            cross = q8(al * 2.0 ** sa) @ q8(wh * 2.0 ** t) + q8(ah) @ q8(wl * 2.0 ** (t + sa))
            return (main + cross) * 2.0 ** -(sa + t)
        return f
    raise ValueError(mode)

class PlanMM:
    """Per-GEMM precision plan: `single` names the GEMMs run on plane 0 only (fp16 x fp16, one tensor pass,
    half the L2 -> SM bytes); every other GEMM stays f16f8.  Names: embed, qkv, proj, fc1, fc2, optionally with a block range
    suffix like fc1@0-5.  mode string: plan:qkv+fc1@6-11"""

    def __init__(self, spec_str):
        self.full, self.one = make_mm("f16f8:8"), make_mm("fp16x1")
        self.rules = []
        for item in filter(None, spec_str.split("+")):
            name, _, rng = item.partition("@")
            lo, hi = (int(v) for v in rng.split("-")) if rng else (0, 99)
            self.rules.append((name, lo, hi))
        self.block, self.name = -1, "embed"

    def at(self, block, name):
        self.block, self.name = block, name
        return self

    def __call__(self, a, w):
        single = any(n == self.name and lo <= self.block <= hi for n, lo, hi in self.rules)
        return (self.one if single else self.full)(a, w)


def forward(x, sd, mm, head_w, head_b, attn_mode=None):
    B = x.shape[0]; D = spec.dim; H = spec.heads; hd = D // H
    p = spec.patch
    g = spec.img // p
    cols = x.reshape(B, spec.in_chans, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B * g * g, -1)
    w = sd["patch_embed.proj.weight"].reshape(D, -1)
    named = isinstance(mm, PlanMM)
    at = (lambda blk, name: mm.at(blk, name)) if named else (lambda blk, name: mm)
    t = at(-1, "embed")(cols, w.t()) + sd["patch_embed.proj.bias"]
    t = t.reshape(B, g * g, D)
    t = torch.cat([sd["cls_token"].expand(B, -1, -1), t], 1) + sd["pos_embed"]
    N = t.shape[1]
    amm = mm if attn_mode is None else attn_mode
    for i in range(spec.depth):
        pre = f"blocks.{i}."
        h = F.layer_norm(t, (D,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-6)
        qkv = at(i, "qkv")(h.reshape(-1, D), sd[pre + "attn.qkv.weight"].t()) + sd[pre + "attn.qkv.bias"]
        qkv = qkv.reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        s = amm(q, k.transpose(-1, -2)) * hd ** -0.5
        pr = torch.softmax(s, -1)
        o = amm(pr, v).transpose(1, 2).reshape(B * N, D)
        t = t + (at(i, "proj")(o, sd[pre + "attn.proj.weight"].t()) + sd[pre + "attn.proj.bias"]).reshape(B, N, D)
        h = F.layer_norm(t, (D,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-6)
        u = F.gelu(at(i, "fc1")(h.reshape(-1, D), sd[pre + "mlp.fc1.weight"].t()) + sd[pre + "mlp.fc1.bias"])
        t = t + (at(i, "fc2")(u, sd[pre + "mlp.fc2.weight"].t()) + sd[pre + "mlp.fc2.bias"]).reshape(B, N, D)
    c = F.layer_norm(t[:, 0], (D,), sd["norm.weight"], sd["norm.bias"], 1e-6)
    return c @ head_w.t() + head_b

def forward_folded(x, sd, mm, head_w, head_b, amm):
    B = x.shape[0]; D = spec.dim; H = spec.heads; hd = D // H
    p = spec.patch; g = spec.img // p
    cols = x.reshape(B, spec.in_chans, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B * g * g, -1)
    w = sd["patch_embed.proj.weight"].reshape(D, -1)
    t = mm(cols, w.t()) + sd["patch_embed.proj.bias"]
    t = t.reshape(B, g * g, D)
    t = torch.cat([sd["cls_token"].expand(B, -1, -1), t], 1) + sd["pos_embed"]
    N = t.shape[1]
    def ln_gemm(xr, gam, bet, W, b):
        # LN(x) W^T + b = rstd * (x (gam*W)^T - mean * c1) + c2
        mean = xr.mean(-1, keepdim=True); var = xr.var(-1, unbiased=False, keepdim=True); rstd = (var + 1e-6).rsqrt()
        Wg = W * gam[None, :]
        c1 = Wg.sum(1); c2 = W @ bet + b
        acc = mm(xr, Wg.t())
        return rstd * (acc - mean * c1[None, :]) + c2[None, :]
    for i in range(spec.depth):
        pre = f"blocks.{i}."
        qkv = ln_gemm(t.reshape(-1, D), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"])
        qkv = qkv.reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        s_ = amm(q, k.transpose(-1, -2)) * hd ** -0.5
        pr = torch.softmax(s_, -1)
        o = amm(pr, v).transpose(1, 2).reshape(B * N, D)
        t = t + (mm(o, sd[pre + "attn.proj.weight"].t()) + sd[pre + "attn.proj.bias"]).reshape(B, N, D)
        u = F.gelu(ln_gemm(t.reshape(-1, D), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]))
        t = t + (mm(u, sd[pre + "mlp.fc2.weight"].t()) + sd[pre + "mlp.fc2.bias"]).reshape(B, N, D)
    c = F.layer_norm(t[:, 0], (D,), sd["norm.weight"], sd["norm.bias"], 1e-6)
    return c @ head_w.t() + head_b


def run(mode, hw, hb, n=None, attn=None):
    if mode.startswith("fold:"):
        mm = make_mm(mode[5:]); amm = make_mm(attn) if attn else mm
        outs = []
        with torch.no_grad():
            for a in range(0, n or len(patches), 128):
                outs.append(forward_folded(patches[a:a + 128], sd, mm, hw, hb, amm))
        return torch.cat(outs)
    mm = PlanMM(mode[5:]) if mode.startswith("plan:") else make_mm(mode)
    amm = make_mm(attn) if attn else None
    outs = []
    with torch.no_grad():
        for a in range(0, n or len(patches), 128):
            outs.append(forward(patches[a:a + 128], sd, mm, hw, hb, amm))
    return torch.cat(outs)

with torch.no_grad():
    lg = run("fp32", sd["head.weight"], sd["head.bias"], 256)
    cal = weights.calibrate_head(sd, lg.mean(0).numpy(), 20.0)
    hw, hb = cal["head.weight"], cal["head.bias"]
    t0 = time.time()
    ref = run("fp32", hw, hb).double()
    pref = torch.softmax(ref, 1)
    print("fp32 time", time.time() - t0, "cells", len(ref), "label hist", torch.bincount(pref.argmax(1)).tolist(), flush=True)
    # cross-check against the oracle model
    model = orc.make_vit(panel); model.load_state_dict(cal)
    po = torch.as_tensor(orc.vit_probs(model, patches[:256].numpy()))
    print("emul-vs-oracle fp32 max dprob", (po - pref[:256]).abs().max().item())
    modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["bf16x3", "bf16x1", "fp16x1", "fp16x2w", "fp16x2a"]
    for m in modes:
        mode, _, attn = m.partition("/")
        lgm = run(mode, hw, hb, attn=attn or None).double()
        pm = torch.softmax(lgm, 1)
        d = (pm - pref).abs().max(1).values
        flips = (pm.argmax(1) != pref.argmax(1)).sum().item()
        print(f"{m:24s} max|dprob| {d.max().item():.3e}  p99.9 {d.quantile(0.999).item():.3e}  mean {d.mean().item():.3e}  max|dlogit| {(lgm - ref).abs().max().item():.3e} flips {flips}", flush=True)
