"""The CPU oracle against the golden fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import lost_oracle as LO
from oracle import pruning_oracle as PO


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _seq(z, prefix):
    out, i = [], 0
    while f"{prefix}{i}" in z:
        out.append(z[f"{prefix}{i}"])
        i += 1
    return out


def test_snip_matches_reference(golden_dir):
    z = _load(golden_dir, "snip_tiny.npz")
    w, g, ref_masks = _seq(z, "w"), _seq(z, "g"), _seq(z, "m")
    masks, thr, scores = PO.snip_pruning(w, [g], 0.9)
    assert thr == float(z["threshold"][0])               # bit-exact: same fp32 value printed by train.py:309
    for m, r in zip(masks, ref_masks):
        assert np.array_equal(m, r.astype(bool))
    assert PO.compute_sparsity_global(w, masks) == pytest.approx(float(z["sparsity"][0]), abs=0)


def test_magnitude_tiefree_bit_exact(golden_dir):
    z = _load(golden_dir, "magnitude_tiefree.npz")
    w = _seq(z, "w")
    masks = None
    for r, amount in enumerate(z["amounts"]):
        masks, info = PO.magnitude_masks(w, masks, float(amount))
        ref = _seq(z, f"m{r}_")
        if info["quota"] == info["n_equal"]:
            for m, rm in zip(masks, ref):
                assert np.array_equal(m, rm.astype(bool)), f"round {r}"
        else:
            assert PO.masks_equal_modulo_ties(masks, ref, w, info["threshold"])
        assert PO.compute_sparsity_global(w, masks) == float(z["sparsity"][r])
        masks = [rm.astype(bool) for rm in ref]          # continue from the reference's state


def test_magnitude_with_planted_ties(golden_dir):
    z = _load(golden_dir, "magnitude_tiny.npz")
    w = _seq(z, "w")
    m1, info1 = PO.magnitude_masks(w, None, 0.5)
    ref1 = _seq(z, "m1_")
    assert info1["n_equal"] > 1                            # the fixture really has ties at the threshold
    assert PO.masks_equal_modulo_ties(m1, ref1, w, info1["threshold"])
    assert sum(int(m.sum()) for m in m1) == sum(int(m.sum()) for m in ref1)    # exactly k pruned
    # round 2 starts from the reference's own round-1 masks
    m2, info2 = PO.magnitude_masks(w, [r.astype(bool) for r in ref1], 0.2)
    ref2 = _seq(z, "m2_")
    assert PO.masks_equal_modulo_ties(m2, ref2, w, info2["threshold"])
    assert PO.compute_sparsity_global(w, ref2) == float(z["sparsity"][1])


def test_tie_policy_is_lowest_index_first():
    w = [np.array([0.5, 0.1, 0.1, 0.1, 0.9], np.float32), np.array([0.1, 0.7, 0.1], np.float32)]
    masks, info = PO.magnitude_masks(w, None, 3)           # integer amount = absolute count
    assert info["n_equal"] == 5 and info["quota"] == 3
    assert masks[0].tolist() == [True, False, False, False, True]
    assert masks[1].tolist() == [True, True, True]


def test_k_rules():
    assert PO.snip_k(25502912, 0.9) == 22952620            # train.py:299, SURVEY §8(a-5)
    assert PO.magnitude_k(0.5, 11678912) == 5839456
    assert PO.magnitude_k(0.5, 5) == 2 and PO.magnitude_k(0.5, 7) == 4     # banker's rounding
    assert PO.snip_threshold(np.arange(10, dtype=np.float32), 1.0) == float("inf")
    assert PO.snip_threshold(np.arange(10, dtype=np.float32), 0.0) == -1


def test_snip_all_zero_scores_prunes_everything():
    w = [np.ones(10, np.float32)]
    g = [np.zeros(10, np.float32)]
    masks, thr, _ = PO.snip_pruning(w, [g], 0.5)
    assert thr == 0.0 and not masks[0].any()               # every tie pruned


def test_masked_sgd_matches_torch():
    import torch
    rng = np.random.default_rng(0)
    w0 = rng.standard_normal(1000).astype(np.float32)
    mask = rng.random(1000) > 0.6
    for nesterov in (False, True):
        p = torch.nn.Parameter(torch.from_numpy(w0.copy()))
        opt = torch.optim.SGD([p], lr=0.1, momentum=0.9, weight_decay=1e-4, nesterov=nesterov)
        w, buf = w0.copy(), None
        for step in range(4):
            g = rng.standard_normal(1000).astype(np.float32)
            p.grad = torch.from_numpy(g * mask)            # MulBackward of the reparametrisation
            opt.step()
            w, buf, weff = PO.masked_sgd_step(w, g, buf, mask, 0.1, 0.9, 0.0, 1e-4, nesterov, first_step=(step == 0))
            np.testing.assert_allclose(w, p.detach().numpy(), rtol=2e-6, atol=1e-7)
            assert np.array_equal(weff != 0, mask & (w != 0))


def test_lost_matches_reference(golden_dir):
    z = _load(golden_dir, "lost_cases.npz")
    meta = json.load(open(os.path.join(golden_dir, "lost_cases.json")))
    for name, m in meta.items():
        feats = z[f"{name}_feats"]
        pred, A, scores, seed = LO.lost(feats, m["dims"], m["scales"], tuple(m["init_image_size"]), m["k_patches"])
        assert seed == m["seed"], name
        assert [int(v) for v in pred] == m["pred"], name
        assert np.array_equal((-scores).astype(np.int32), z[f"{name}_degree"]), name
        np.testing.assert_allclose(A[:4, :6], z[f"{name}_A_probe"], rtol=1e-5, atol=1e-4)


def test_lost_replay_from_degrees_matches_reference(golden_dir):
    """The fp64 replay the GPU tests use as their checker reproduces the reference's seed and box from the reference's
    own degrees, and calls every golden decidable."""
    z = _load(golden_dir, "lost_cases.npz")
    meta = json.load(open(os.path.join(golden_dir, "lost_cases.json")))
    for name, m in meta.items():
        seed, pred, decidable = LO.lost_from_degrees(z[f"{name}_feats"], z[f"{name}_degree"], m["dims"], m["scales"],
                                                     tuple(m["init_image_size"]), m["k_patches"])
        assert seed == m["seed"] and [int(v) for v in pred] == m["pred"], name
        assert decidable or name.startswith("random"), name


def test_lost_background_seed_raises():
    M = -np.ones(12, np.float32)
    with pytest.raises(ValueError, match="background"):
        LO.detect_box(M, 5, [3, 4], (48, 64), [16, 16])


def test_torch_port_matches_goldens(golden_dir):
    """The torch-CPU timing port (bench.py's cpu_baseline / --impl reference legs) reproduces the
    unmodified reference's outputs."""
    import torch
    from oracle import torch_port as TP
    z = _load(golden_dir, "snip_tiny.npz")
    w = [torch.from_numpy(a.copy()) for a in _seq(z, "w")]
    g = [torch.from_numpy(a.copy()) for a in _seq(z, "g")]
    masks, thr = TP.snip_mask_build(w, [g], 0.9)
    assert thr == float(z["threshold"][0])
    for m, r in zip(masks, _seq(z, "m")):
        assert np.array_equal(m.numpy().astype(bool), r.astype(bool))
    assert TP.sparsity_percent(w, masks) == float(z["sparsity"][0])
    # multi-batch extension agrees with the numpy oracle
    g2 = [torch.from_numpy((a * 0.5 + 1e-4).astype(np.float32)) for a in _seq(z, "g")]
    masks2, thr2 = TP.snip_mask_build(w, [g, g2], 0.7)
    exp, ethr, _ = PO.snip_pruning([t.numpy() for t in w], [[t.numpy() for t in g], [t.numpy() for t in g2]], 0.7)
    assert thr2 == ethr
    for m, e in zip(masks2, exp):
        assert np.array_equal(m.numpy().astype(bool), e)
    # magnitude, iterative, tie-free fixture: bit-exact with the reference
    z = _load(golden_dir, "magnitude_tiefree.npz")
    w = [torch.from_numpy(a.copy()) for a in _seq(z, "w")]
    masks = None
    for r, amount in enumerate(z["amounts"]):
        masks, k = TP.magnitude_mask_build(w, masks, float(amount))
        ref = _seq(z, f"m{r}_")
        for m, rm in zip(masks, ref):
            assert np.array_equal(m.numpy().astype(bool), rm.astype(bool)), f"round {r}"
