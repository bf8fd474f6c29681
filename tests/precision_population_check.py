#!/usr/bin/env python
"""Full-population precision check on the bench scene: probabilities of every cell from the f16f8 and bf16x3 engines and
from the reference modules in eager fp32, all against the same modules in fp64 (ground truth).
usage: python tests/precision_population_check.py [size] [panel]   (a GPU script like tests/multi_gpu_check.py, not collected by pytest)"""
import json, os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiplexed_image_annotator_b200 import engine, ops, synth, weights
from multiplexed_image_annotator_b200.cell_type_annotation.model import merge_on_device
from oracle import ribca_oracle as orc
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
panel = sys.argv[2] if len(sys.argv) > 2 else "immune_full"
n_ch = weights.VIT_SPECS[panel].in_chans
index = list(range(n_ch))
mask = synth.synth_mask(S, S, grid=18, seed=2, device=dev)
img = torch.from_numpy(synth.to_uint16(synth.synth_image(mask, n_ch, seed=2))).to(dev)
norm = ops.normalize(img, 0.3, 99.8)
cells = ops.cell_stats(mask)
mn = ops.channel_min(norm)
sd = weights.random_vit_state(panel, seed=7)
engs = {p: engine.VitEngine(panel, sd, dev, precision=p) for p in ("f16f8", "bf16x3")}
(p256,), _, _ = ops.build_patches(norm, mask, mn, cells, [index], 0, 256)
_, logits = engs["f16f8"].forward(p256, return_logits=True)
cal = weights.calibrate_head(sd, logits.mean(0).cpu().numpy(), 20.0)
for e in engs.values():
    e.set_head(cal["head.weight"], cal["head.bias"])
ref32 = orc.make_vit(panel); ref32.load_state_dict(cal); ref32 = ref32.to(dev).eval()
ref64 = orc.make_vit(panel); ref64.load_state_dict(cal); ref64 = ref64.to(dev).double().eval()
out = {k: [] for k in ("f16f8", "bf16x3", "fp32", "fp64")}
with torch.no_grad():
    for a in range(0, cells.n, 4096):
        (pt,), _, _ = ops.build_patches(norm, mask, mn, cells, [index], a, min(4096, cells.n - a))
        for k, e in engs.items():
            out[k].append(e.forward(pt).double())
        out["fp32"].append(torch.cat([torch.softmax(ref32(pt[b:b + 128]), 1) for b in range(0, len(pt), 128)]).double())
        out["fp64"].append(torch.cat([torch.softmax(ref64(pt[b:b + 256].double()), 1) for b in range(0, len(pt), 256)]))
P = {k: torch.cat(v) for k, v in out.items()}
truth = P["fp64"]
lab = {k: merge_on_device({panel: v.float()}, 0.3, None)[0] for k, v in P.items()}
res = {"cells": cells.n, "panel": panel}
for k in ("fp32", "bf16x3", "f16f8"):
    d = (P[k] - truth).abs().max(1).values
    res[k] = {"max": d.max().item(), "p99.99": d.quantile(0.9999).item(), "p99.9": d.quantile(0.999).item(), "p99": d.quantile(0.99).item(),
              "mean": d.mean().item(), "n>1e-3": int((d > 1e-3).sum()), "n>5e-4": int((d > 5e-4).sum()), "n>2e-4": int((d > 2e-4).sum()),
              "labels_differ_vs_fp64": int((lab[k] != lab["fp64"]).sum()), "labels_differ_vs_fp32": int((lab[k] != lab["fp32"]).sum())}
    worst = int(d.argmax())
    res[k]["worst_cell"] = {"index": worst, "truth": [round(v, 5) for v in truth[worst].tolist()], "got": [round(v, 5) for v in P[k][worst].tolist()]}
# top-2 gap statistics: how many cells sit within 2e-4 of a tie (a flip there is inside any fp32-level noise)
g = truth.sort(1, descending=True).values
res["cells_with_top2_gap_below"] = {"1e-4": int(((g[:, 0] - g[:, 1]) < 1e-4).sum()), "1e-3": int(((g[:, 0] - g[:, 1]) < 1e-3).sum())}
print(json.dumps(res, indent=1))
