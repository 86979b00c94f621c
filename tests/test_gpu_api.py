"""The reference-facing Python functions (same names/arguments as train.py) on a CUDA model."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.utils.prune as prune

from oracle import pruning_oracle as PO
from tests.tinynet import TinyNet, load_weights

pytestmark = pytest.mark.gpu

from pruning_for_vision_representation_b200 import pruning as P     # noqa: E402
from pruning_for_vision_representation_b200._lib import B200PruneError   # noqa: E402

DEV = torch.device("cuda:0")


def _seq(z, prefix):
    out, i = [], 0
    while f"{prefix}{i}" in z:
        out.append(z[f"{prefix}{i}"])
        i += 1
    return out


def _masks(mods):
    return [m.weight_mask.detach().cpu().numpy().astype(bool) for m in mods]


def test_magnitude_pruning_dropin(golden_dir):
    z = np.load(os.path.join(golden_dir, "magnitude_tiefree.npz"))
    torch.manual_seed(0)
    model = TinyNet()
    mods = load_weights(model, _seq(z, "w"))
    model.to(DEV)
    assert P.compute_sparsity_global(model) == 0.0
    orig_params = [m.weight for m in mods]
    for r, amount in enumerate(z["amounts"]):
        out = P.magnitude_pruning(model, float(amount))
        assert out is model
        ref = _seq(z, f"m{r}_")
        for got, rm in zip(_masks(mods), ref):
            assert np.array_equal(got, rm.astype(bool)), f"round {r}"
        assert P.compute_sparsity_global(model) == float(z["sparsity"][r])
    # torch.nn.utils.prune post-conditions
    assert prune.is_pruned(model)
    sd = model.state_dict()
    assert "f1.weight_orig" in sd and "f1.weight_mask" in sd and "f1.weight" not in sd
    assert sd["f1.weight_mask"].dtype == torch.float32 and sd["f1.weight_mask"].shape == mods[2].weight_orig.shape
    for m, p in zip(mods, orig_params):
        assert m.weight_orig is p                                # same Parameter object (optimizer / DDP refs stay valid)
    x = torch.randn(4, 3, 16, 16, device=DEV)
    model(x).sum().backward()
    for m in mods:
        mask = m.weight_mask.bool()
        assert torch.equal(m.weight, m.weight_orig * m.weight_mask)
        assert torch.count_nonzero(m.weight_orig.grad[~mask]) == 0   # MulBackward: no gradient where pruned
    prune.remove(mods[2], "weight")
    assert "weight_orig" not in mods[2]._parameters and isinstance(mods[2].weight, nn.Parameter)
    assert torch.count_nonzero(mods[2].weight) == int(_seq(z, "m3_")[2].sum())


def test_magnitude_pruning_errors():
    model = TinyNet().to(DEV)
    with pytest.raises(ValueError):
        P.magnitude_pruning(model, 1.5)                         # prune.py:1256-1290
    with pytest.raises(TypeError):
        P.magnitude_pruning(model, "a lot")
    with pytest.raises(ValueError):
        P.magnitude_pruning(model, 10 ** 9)                     # more than there is to prune
    with pytest.raises(B200PruneError):
        P.magnitude_pruning(TinyNet(), 0.2)                     # CPU model: no fallback


def test_snip_pruning_dropin(golden_dir, capsys):
    z = np.load(os.path.join(golden_dir, "snip_tiny.npz"))
    torch.backends.cudnn.allow_tf32 = False          # fp32 convolutions, like the CPU reference run
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(3)                             # same init (incl. biases) as gen_snip_tiny()
    model = TinyNet()
    mods = load_weights(model, _seq(z, "w"))
    model.to(DEV)
    x, y = torch.from_numpy(z["x"]), torch.from_numpy(z["y"])
    crit = nn.CrossEntropyLoss()
    # gradients as the GPU computes them (cuDNN), for the oracle
    model.zero_grad()
    crit(model(x.to(DEV)), y.to(DEV)).backward()
    g = [m.weight.grad.detach().cpu().numpy().copy() for m in mods]
    w = [m.weight.detach().cpu().numpy().copy() for m in mods]
    P.snip_pruning(model, [(x, y)], DEV, crit, target_sparsity=0.9)
    printed = capsys.readouterr().out
    assert "Applying SNIP pruning with target sparsity 0.9..." in printed and "SNIP threshold:" in printed
    exp, thr, _ = PO.snip_pruning(w, [g], 0.9)
    assert f"SNIP threshold: {thr}" in printed                   # same float, same repr as train.py:309
    got = _masks(mods)
    for a, e in zip(got, exp):
        assert np.array_equal(a, e)
    # against the CPU reference run: identical up to the handful of entries whose score moved
    # across the threshold because cuDNN and the CPU convolution round differently
    ref = _seq(z, "m")
    diff = sum(int((a != r.astype(bool)).sum()) for a, r in zip(got, ref))
    assert diff <= 0.01 * sum(a.size for a in got), diff
    assert abs(P.compute_sparsity_global(model) - float(z["sparsity"][0])) < 0.2


def test_snip_multibatch_and_degenerate():
    torch.manual_seed(5)
    model = TinyNet().to(DEV)
    mods = [m for m in model.modules() if isinstance(m, (nn.Conv2d, nn.Linear))]
    crit = nn.CrossEntropyLoss()
    batches = [(torch.randn(4, 3, 16, 16), torch.randint(0, 10, (4,))) for _ in range(3)]
    w = [m.weight.detach().cpu().numpy().copy() for m in mods]
    grads = []
    for xb, yb in batches:
        model.zero_grad()
        crit(model(xb.to(DEV)), yb.to(DEV)).backward()
        grads.append([m.weight.grad.detach().cpu().numpy().copy() for m in mods])
    P.snip_pruning(model, batches, DEV, crit, target_sparsity=0.6, num_batches=3)
    exp, thr, _ = PO.snip_pruning(w, grads, 0.6)
    for a, e in zip(_masks(mods), exp):
        assert np.array_equal(a, e)
    # target 1.0 -> threshold inf -> everything pruned; 0.0 -> threshold -1 -> nothing more pruned
    m2 = TinyNet().to(DEV)
    P.snip_pruning(m2, batches, DEV, crit, target_sparsity=1.0)
    assert P.compute_sparsity_global(m2) == 100.0
    m3 = TinyNet().to(DEV)
    P.snip_pruning(m3, batches, DEV, crit, target_sparsity=0.0)
    assert P.compute_sparsity_global(m3) == 0.0 and prune.is_pruned(m3)


def test_adopts_torch_prune_checkpoint(golden_dir):
    """A model pruned by torch.nn.utils.prune (reference checkpoint format) continues correctly."""
    z = np.load(os.path.join(golden_dir, "magnitude_tiefree.npz"))
    model = TinyNet()
    mods = load_weights(model, _seq(z, "w"))
    for m, rm in zip(mods, _seq(z, "m0_")):
        prune.custom_from_mask(m, "weight", torch.from_numpy(rm.astype(np.float32)))
    model.to(DEV)
    assert P.compute_sparsity_global(model) == float(z["sparsity"][0])
    P.magnitude_pruning(model, float(z["amounts"][1]))
    for got, rm in zip(_masks(mods), _seq(z, "m1_")):
        assert np.array_equal(got, rm.astype(bool))
