"""Exact labels by construction (multiplexed_image_annotator_b200/exact.py): the decision margin of stage 5, the plain
fp32 forward on the FP32 pipe, and the margin-guarded re-evaluation through HotPath / Annotator.

The checker is the oracle (the reference's algorithm, cta/model.py:397-406 + 481-636) run in float64: with the labels
refined, every cell must carry the label exact arithmetic gives it - no tolerance on labels."""
import os

import numpy as np
import pytest
import torch

from multiplexed_image_annotator_b200 import engine, exact, ops, synth, weights
from multiplexed_image_annotator_b200.cell_type_annotation.model import ALL_TYPES, OTHERS, merge_on_device
from multiplexed_image_annotator_b200.pipeline import HotPath
from oracle import ribca_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _outcome(label, conf):
    return label.astype(np.int64) + 32 * (conf == -1.0)


def _random_tables(rng, n, k, sharp):
    z = rng.normal(size=(n, k)).astype(np.float32) * sharp
    z -= z.max(1, keepdims=True)
    p = np.exp(z)
    return (p / p.sum(1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("panels,ctc", [
    (["immune_full"], None),
    (["immune_base"], {"B cell": 0.5, "CD4 T cell": 0.2}),
    (["structure"], None),
    (["immune_extended", "structure"], None),
    (["immune_base", "nerve_cell"], {"Nerve cell": 0.6, "CD8 T cell": 0.0}),
    (["structure", "nerve_cell"], None),
])
def test_decision_margin_guards_the_outcome(panels, ctc):
    """Any perturbation of the probabilities smaller than half the margin leaves (label, re-labelled?) unchanged, and
    the margin is tight: it never exceeds the distance to the winner's threshold or to the runner-up's takeover."""
    rng = np.random.default_rng(11)
    n = 4000
    full_ctc = None if ctc is None else {t: ctc.get(t, -1) for t in ALL_TYPES}
    tabs = {p: _random_tables(rng, n, len(weights.VIT_SPECS[p].classes), 1.2) for p in panels}
    dev = {p: torch.from_numpy(t).to(DEV) for p, t in tabs.items()}
    label, conf, counts, margin = merge_on_device(dev, 0.3, full_ctc, want_margin=True)
    l0, c0, k0 = merge_on_device(dev, 0.3, full_ctc)
    assert torch.equal(label, l0) and torch.equal(conf, c0) and torch.equal(counts, k0)      # the margin output changes nothing
    base = _outcome(label.cpu().numpy(), conf.cpu().numpy())
    m = margin.cpu().numpy()
    assert np.all(m >= 0) and np.isfinite(m).mean() > 0.99
    finite = np.where(np.isfinite(m), m, 1.0)
    for trial in range(12):
        pert = {}
        for p, t in tabs.items():
            d = rng.uniform(-1, 1, size=t.shape).astype(np.float64)
            if trial % 3 == 0:
                d = np.sign(d)                                                     # worst case: every entry at the bound
            pert[p] = torch.from_numpy((t.astype(np.float64) + d * (0.249 * finite[:, None])).astype(np.float32)).to(DEV)
        # two-model thresholds use min(others...) of BOTH tables: 0.249 * margin per entry keeps every compared
        # difference (vote - vote, vote - threshold) below margin / 2
        l2, c2, _ = merge_on_device(pert, 0.3, full_ctc)
        assert np.array_equal(_outcome(l2.cpu().numpy(), c2.cpu().numpy()), base), f"trial {trial}"
    # tightness for the one-model branch: margin <= |p_best - thr| for a thresholded winner and <= gap to the runner-up
    # whenever the runner-up would give another outcome
    if len(panels) == 1:
        t = tabs[panels[0]]
        names = weights.VIT_SPECS[panels[0]].classes
        srt = np.sort(t, 1)
        best = t.argmax(1)
        thr = np.array([(ctc or {}).get(names[b], -1) for b in best], dtype=np.float32)
        thr = np.where(thr > 0, thr, np.float32(0.3))
        real = np.array([names[b] != "Others" for b in best])
        assert np.all(m[real] <= np.abs(srt[real, -1] - thr[real]) + 1e-7)
        assert np.all(m >= 0)


@pytest.mark.parametrize("panel", ["immune_base", "immune_extended", "immune_full", "structure", "nerve_cell"])
def test_vit_fp32_path_vs_oracle(panel):
    """precision="fp32": fp32 operands + FFMA accumulation (csrc/stage4_fp32.cu) against the oracle's torch fp32 AND fp64."""
    spec = weights.VIT_SPECS[panel]
    mask = synth.synth_mask(160, 160, seed=6)
    img = orc.normalize(synth.to_uint16(synth.synth_image(mask, spec.in_chans, seed=6)), 0.3, 99.8)
    patches, _, _ = orc.build_patches(img, mask.numpy(), list(range(spec.in_chans)))
    sd = weights.random_vit_state(panel, seed=3)
    ref = orc.make_vit(panel)
    ref.load_state_dict(sd)
    with torch.no_grad():
        mean_logits = ref(torch.from_numpy(patches[:32])).mean(0).numpy()
    sd = weights.calibrate_head(sd, mean_logits, 20.0)
    ref.load_state_dict(sd)
    want32 = orc.vit_probs(ref, patches)
    ref64 = orc.make_vit(panel).double()
    ref64.load_state_dict({k: v.double() for k, v in sd.items()})
    with torch.no_grad():
        want64 = torch.softmax(ref64(torch.from_numpy(patches).double()), 1).numpy()
    eng = engine.VitEngine(panel, sd, DEV, max_cells_per_call=37)          # default packs f16f8; fp32 is packed on demand
    x = torch.from_numpy(patches).to(DEV)
    got = eng.forward(x, precision="fp32").cpu().numpy().astype(np.float64)
    d32, d64 = np.abs(got - want32).max(), np.abs(got - want64).max()
    r64 = np.abs(want32 - want64).max()
    print(f"{panel} fp32 path: max|dprob| vs torch fp32 {d32:.2e}, vs fp64 {d64:.2e} (torch fp32 vs fp64: {r64:.2e})")
    assert d64 < 1e-5 and d32 < 1e-5
    # the other precisions still come out of the same engine (lazy packs), bit-identical to a dedicated engine
    b3 = eng.forward(x, precision="bf16x3")
    b3_ded = engine.VitEngine(panel, sd, DEV, precision="bf16x3", max_cells_per_call=37).forward(x)
    assert torch.equal(b3, b3_ded)


def _scene(size, seed, channels=15):
    mask = synth.synth_mask(size, size, grid=18, seed=seed)
    return synth.to_uint16(synth.synth_image(mask, channels, seed=seed)), mask.numpy()


def _calibrated(panel, patches_probe, seed=7):
    sd = weights.random_vit_state(panel, seed=seed)
    ref = orc.make_vit(panel)
    ref.load_state_dict(sd)
    with torch.no_grad():
        mean_logits = ref(torch.from_numpy(patches_probe)).mean(0).numpy()
    return weights.calibrate_head(sd, mean_logits, 20.0)


def test_refined_labels_equal_float64_oracle():
    """Whole hot path on a 1024^2 scene (~3.2 k cells, immune_full -> vit_l): with the re-evaluation on, every label and
    re-labelled flag equals the float64 evaluation of the reference's algorithm; the re-evaluated cells are few; cells
    that were not re-evaluated keep the fast pass's probabilities bit for bit."""
    panel, index = "immune_full", list(range(15))
    img, mask = _scene(1024, 2)
    norm = orc.normalize(img, 0.3, 99.8)
    patches, _, _ = orc.build_patches(norm, mask, index)
    sd = _calibrated(panel, patches[:64])
    ref64 = orc.make_vit(panel).double()
    ref64.load_state_dict({k: v.double() for k, v in sd.items()})
    with torch.no_grad():
        p64 = torch.cat([torch.softmax(ref64(torch.from_numpy(patches[a:a + 256]).double()), 1) for a in range(0, len(patches), 256)]).numpy()
    want_lab, want_conf = orc.merge_by_voting({panel: p64.astype(np.float32)}, 0.3, None)
    # cells that float32 rounding of the float64 probabilities itself could decide either way are not a test of the GPU path
    eng = engine.VitEngine(panel, sd, DEV, max_cells_per_call=1024)
    fast = HotPath({panel: index}, {panel: eng}, device=DEV, shard_cells=False, chunk_cells=1024, exact_labels=0)
    safe = HotPath({panel: index}, {panel: eng}, device=DEV, shard_cells=False, chunk_cells=1024, exact_labels=2)
    r0 = fast.run(img, mask, keep_probs=True)
    r2 = safe.run(img, mask, keep_probs=True)
    st = r2.refine.as_dict()
    names0, names2 = r0.names(), r2.names()
    d0 = sum(a != b for a, b in zip(names0, want_lab))
    d2 = [j for j, (a, b) in enumerate(zip(names2, want_lab)) if a != b]
    dp0 = float(np.abs(r0.probs[panel].cpu().numpy() - p64).max())
    dp2 = float(np.abs(r2.probs[panel].cpu().numpy() - p64).max())
    print(f"refine: {st}; labels differing from fp64: fast pass {d0}, refined {len(d2)}; max|dprob| fast {dp0:.2e} refined {dp2:.2e}")
    assert len(names2) == len(want_lab) == r2.n_cells
    assert d2 == []
    relab = np.array([c == -1 for c in want_conf])
    assert np.array_equal(r2.confidence.numpy() == -1.0, relab)
    assert dp0 < 1e-3 and dp2 <= dp0
    assert 0 < st["level1_bf16x3_cells"] < 0.1 * r2.n_cells and st["level2_fp32_cells"] <= st["level1_bf16x3_cells"]
    # untouched cells: identical probabilities; touched cells: all have a fast-pass margin below eps1
    same = (r0.probs[panel] == r2.probs[panel]).all(1).cpu().numpy()
    m0 = r0.margin.cpu().numpy()
    assert np.all(m0[~same] < exact.EPS1) and (~same).sum() <= st["level1_bf16x3_cells"]
    assert int(r2.counts.sum()) == r2.n_cells
    assert np.array_equal(np.bincount(r2.label.numpy(), minlength=18), r2.counts.numpy())
    # independent of how the cells are batched: another chunking gives the same labels, confidences and probabilities
    r3 = HotPath({panel: index}, {panel: eng}, device=DEV, shard_cells=False, chunk_cells=700, exact_labels=2).run(img, mask, keep_probs=True)
    assert torch.equal(r3.label, r2.label) and torch.equal(r3.confidence, r2.confidence)
    assert torch.equal(r3.probs[panel], r2.probs[panel])


def test_refined_labels_two_models_and_imputer():
    """Merge branch 2 (immune_base with an imputed marker + structure) through the re-evaluation: labels equal the
    oracle's on the same inputs evaluated in float64 for the classifiers."""
    img, mask = _scene(512, 5, channels=12)
    base_idx, struct_idx = [0, 1, 2, 3, 4, -1, 5], [4, 6, 7, 8, 9, 10, 11]
    norm = orc.normalize(img, 0.3, 99.8)
    pb, _, _ = orc.build_patches(norm, mask, base_idx)
    ps, _, _ = orc.build_patches(norm, mask, struct_idx)
    mae_sd = weights.random_mae_state("immune_base", seed=7)
    mae_ref = orc.make_mae("immune_base")
    mae_ref.load_state_dict(mae_sd)
    present = [0, 1, 2, 3, 4, 6]
    pb_imp = orc.impute(mae_ref, pb.copy(), present)
    sds = {"immune_base": _calibrated("immune_base", pb_imp[:64]), "structure": _calibrated("structure", ps[:64], seed=9)}
    p64 = {}
    for p, x in (("immune_base", pb_imp), ("structure", ps)):
        r = orc.make_vit(p).double()
        r.load_state_dict({k: v.double() for k, v in sds[p].items()})
        with torch.no_grad():
            p64[p] = torch.softmax(r(torch.from_numpy(x).double()), 1).numpy().astype(np.float32)
    want_lab, _ = orc.merge_by_voting(p64, 0.3, None)
    models = {p: engine.VitEngine(p, sds[p], DEV) for p in sds}
    mae = engine.MaeEngine("immune_base", mae_sd, DEV)
    hp = HotPath({"immune_base": base_idx, "structure": struct_idx}, models, {"immune_base": (mae, present)}, device=DEV,
                 shard_cells=False, exact_labels=2)
    res = hp.run(img, mask, keep_probs=True)
    differ = [j for j, (a, b) in enumerate(zip(res.names(), want_lab)) if a != b]
    # the imputed marker itself comes from the MAE at reduced precision (bf16x3 at best): a label may only differ where
    # the oracle's own margin is inside that input error band
    m64 = merge_on_device({p: torch.from_numpy(v).to(DEV) for p, v in p64.items()}, 0.3, None, want_margin=True)[3].cpu().numpy()
    print(f"two models + imputer: {res.refine.as_dict()}, labels differing {len(differ)}, their fp64 margins {[float(m64[j]) for j in differ]}")
    assert all(m64[j] < 2e-4 for j in differ) and len(differ) <= 1


# ------------------------------------------------------------------------------------------------
# batch mode stays inside a device-memory budget (the reference bounds memory by spilling to tmp/*.pt, preprocess.py:132-135)
# ------------------------------------------------------------------------------------------------
_BATCH_SCRIPT = r'''
import hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.environ["RIBCA_REPO"])
limit_gb = float(os.environ.get("LIMIT_GB", "0"))
if limit_gb > 0:
    total = torch.cuda.get_device_properties(0).total_memory
    torch.cuda.set_per_process_memory_fraction(limit_gb * 2**30 / total, 0)
from multiplexed_image_annotator_b200 import synth, weights
from multiplexed_image_annotator_b200.cell_type_annotation import model as bmodel
os.chdir(os.environ["WORK"])
with open("images.csv", "w") as f:
    f.write("image_path,mask_path\n")
    for k in range(16):
        m = synth.synth_mask(512, 512, seed=60 + k)
        np.save(f"img{k}.npy", synth.to_uint16(synth.synth_image(m, 7, seed=60 + k))); np.save(f"mask{k}.npy", m.numpy())
        f.write(f"img{k}.npy,mask{k}.npy\n")
synth.write_marker_file("markers.txt", synth.STRUCTURE_MARKERS)
bmodel.register_state("structure", weights.random_vit_state("structure", seed=4))
ann = bmodel.Annotator("markers.txt", "images.csv", "cuda", "./", "mem", True, True, -1, True, 0.3, 99.8, 0.3, 30, None, n_jobs=0)
ann.load_models()
for m in ann.models.values():
    m.max_cells = 256            # small network workspace: the images and patch caches dominate the footprint
ann.preprocess()
ann.predict(128)
ann.colorize(from_script=True)
h = hashlib.sha256()
for lab, conf in zip(ann.labels_index, ann.confidence):
    h.update(lab.tobytes()); h.update(np.asarray(conf, dtype=np.float32).tobytes())
pre = ann.preprocessor
print("RESULT", h.hexdigest(), sum(pre._resident), sum(p is not None for p in pre.patches), f"{torch.cuda.max_memory_allocated() / 2**30:.2f}")
'''


def test_batch_csv_stays_inside_a_hard_memory_limit(tmp_path):
    """A 16-image batch CSV under torch.cuda.set_per_process_memory_fraction: with the default budgets (everything
    resident: 16 stacks + 16 patch caches) the hard limit is hit; with RIBCA_PATCH_CACHE_BYTES / RIBCA_IMAGE_CACHE_BYTES
    sized for two images the same run passes inside the limit and gives the same labels and confidences as an unlimited run."""
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def run(tag, limit_gb, **env):
        work = tmp_path / tag
        work.mkdir()
        e = dict(os.environ, RIBCA_REPO=repo, WORK=str(work), LIMIT_GB=str(limit_gb), **{k: str(v) for k, v in env.items()})
        return subprocess.run([sys.executable, "-c", _BATCH_SCRIPT], env=e, capture_output=True, text=True, timeout=600)

    free = run("free", 0)
    assert free.returncode == 0, free.stderr[-2000:]
    _, want, n_res, n_cached, peak_free = next(l for l in free.stdout.split("\n") if l.startswith("RESULT")).split()
    assert (n_res, n_cached) == ("16", "16")
    limit = round(float(peak_free) * 0.62, 2)                  # well below what the all-resident run needs
    tight = run("tight", limit)
    assert tight.returncode != 0 and "out of memory" in (tight.stderr + tight.stdout).lower(), "the all-resident run should hit the limit"
    small = run("small", limit, RIBCA_PATCH_CACHE_BYTES=80 << 20, RIBCA_IMAGE_CACHE_BYTES=18 << 20)
    assert small.returncode == 0, small.stderr[-2000:]
    _, got, n_res, n_cached, peak = next(l for l in small.stdout.split("\n") if l.startswith("RESULT")).split()
    print(f"batch residency: unlimited peak {peak_free} GiB (16 resident, 16 cached); limit {limit} GiB -> all-resident run OOMs, "
          f"budgeted run peaks at {peak} GiB with {n_res} resident stacks, {n_cached} cached patch sets")
    assert got == want
    assert int(n_res) <= 2 and int(n_cached) <= 2 and float(peak) <= limit


def test_guard_escalates_when_the_fast_pass_breaks_its_error_premise():
    """exact.refine_labels checks its own premise: if the cells it re-evaluates show a fast-pass error above EPS1 / 2, it warns and
    sends EVERY cell to level 1 (with trained checkpoints the f16f8 planes could saturate silently).  Here the "fast pass"
    is the true probability table + 2e-3 of noise and the "higher precision" returns the true table."""
    rng = np.random.default_rng(5)
    n, panel = 6000, "immune_full"
    true = torch.from_numpy(_random_tables(rng, n, len(weights.VIT_SPECS[panel].classes), 1.5)).to(DEV)
    noise = torch.from_numpy(rng.uniform(-2e-3, 2e-3, size=tuple(true.shape)).astype(np.float32)).to(DEV)
    merge = lambda pr, want_margin=True: merge_on_device(pr, 0.3, None, want_margin=want_margin)
    want_label, want_conf, want_counts = merge_on_device({panel: true}, 0.3, None)
    calls = []

    def forward_cells(sel, prec):
        calls.append((int(sel.numel()), prec))
        return {panel: true[sel]}

    with pytest.warns(RuntimeWarning, match="re-evaluating all"):
        label, conf, counts, margin, st = exact.refine_labels({panel: true + noise}, merge, forward_cells, levels=2)
    assert st.escalated and st.reevaluated[0] == n and st.observed_error[0] > 5e-4
    assert torch.equal(label, want_label) and torch.equal(conf, want_conf) and torch.equal(counts, want_counts)
    assert sum(c for c, p in calls if p == "bf16x3") == n
    # ... and with an honest fast pass (error 1e-4) nothing escalates and only the boundary cells are touched
    calls.clear()
    small = noise * 0.05
    import warnings as _w
    with _w.catch_warnings():
        _w.simplefilter("error")
        label, conf, counts, margin, st = exact.refine_labels({panel: true + small}, merge, forward_cells, levels=2)
    assert not st.escalated and 0 < st.reevaluated[0] < n // 10 and st.observed_error[0] <= 1.1e-4
    assert torch.equal(label, want_label) and torch.equal(counts, want_counts)


def test_hot_path_on_an_image_without_cells_and_with_one_cell():
    """Edge cases of the whole path: an empty mask gives empty labels and zero counts (the reference would write an empty CSV),
    a single cell goes through every stage (chunk of 1, no re-evaluation batch of size 0 issues)."""
    panel = "nerve_cell"
    eng = engine.VitEngine(panel, weights.random_vit_state(panel, seed=2), DEV)
    hp = HotPath({panel: [0, 1, 2]}, {panel: eng}, device=DEV)
    img = (np.random.default_rng(0).random((3, 96, 96)) * 4000).astype(np.uint16)
    mask = np.zeros((96, 96), np.int32)
    res = hp.run(img, mask)
    assert res.n_cells == 0 and res.label.numel() == 0 and res.confidence.numel() == 0 and int(res.counts.sum()) == 0
    mask[40:52, 30:44] = 7                                    # one cell with label 7
    res = hp.run(img, mask)
    assert res.n_cells == 1 and res.label.numel() == 1 and int(res.counts.sum()) == 1
    assert ALL_TYPES[int(res.label[0])] in tuple(weights.VIT_SPECS[panel].classes) + (ALL_TYPES[OTHERS],)
