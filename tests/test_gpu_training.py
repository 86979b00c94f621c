"""Masked training step (MaskedSGD over K4), pruned-checkpoint helpers and the CLI surface mirror,
checked against torch.nn.utils.prune + torch.optim.SGD — the reference's own step
(train.py:54-67, torch/optim/sgd.py:343-380, torch/nn/utils/prune.py:53-74)."""
import argparse
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.utils.prune as prune

pytestmark = pytest.mark.gpu

from pruning_for_vision_representation_b200 import pruning as P                      # noqa: E402
from pruning_for_vision_representation_b200.checkpoint import (fp32_masks_from_packed, load_pruned,   # noqa: E402
                                                               packed_mask_state)
from pruning_for_vision_representation_b200.cli import add_pruning_args, run_pruning_schedule   # noqa: E402
from pruning_for_vision_representation_b200.masked_sgd import MaskedSGD             # noqa: E402
from tests.tinynet import TinyNet                                                    # noqa: E402

DEV = torch.device("cuda:0")


def _mods(model):
    return [m for m in model.modules() if isinstance(m, (nn.Conv2d, nn.Linear))]


@pytest.mark.parametrize("nesterov", [False, True])
def test_masked_sgd_matches_torch_prune_plus_sgd(nesterov):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    ours = TinyNet().to(DEV)
    ref = copy.deepcopy(ours)
    P.magnitude_pruning(ours, 0.6)
    for mo, mr in zip(_mods(ours), _mods(ref)):
        prune.custom_from_mask(mr, "weight", mo.weight_mask.clone())
    hp = dict(lr=0.1, momentum=0.9, weight_decay=1e-4, nesterov=nesterov)
    opt_o = MaskedSGD(ours, **hp)
    opt_r = torch.optim.SGD(ref.parameters(), **hp)
    sched = torch.optim.lr_scheduler.StepLR(opt_o, step_size=2, gamma=0.5)            # LR schedulers drive param_groups
    sched_r = torch.optim.lr_scheduler.StepLR(opt_r, step_size=2, gamma=0.5)
    crit = nn.CrossEntropyLoss()
    for step in range(5):
        x = torch.randn(8, 3, 16, 16, device=DEV); y = torch.randint(0, 10, (8,), device=DEV)
        for model, opt in ((ours, opt_o), (ref, opt_r)):
            opt.zero_grad()
            crit(model(x), y).backward()
            opt.step()
        sched.step(); sched_r.step()
        for mo, mr in zip(_mods(ours), _mods(ref)):
            torch.testing.assert_close(mo.weight_orig, mr.weight_orig, rtol=2e-5, atol=2e-6)
            torch.testing.assert_close(mo.bias, mr.bias, rtol=2e-5, atol=2e-6)
            assert torch.equal(mo.weight, mo.weight_orig * mo.weight_mask)                # next forward's weight, emitted by the kernel
            assert torch.count_nonzero(mo.weight[mo.weight_mask == 0]) == 0
    torch.testing.assert_close(ours(x), ref(x), rtol=1e-4, atol=1e-5)
    assert P.compute_sparsity_global(ours) == P.compute_sparsity_global(ref)
    # a further pruning round under the fused optimizer
    P.magnitude_pruning(ours, 0.2)
    opt_o.refresh_after_pruning()
    prune.global_unstructured([(m, "weight") for m in _mods(ref)], pruning_method=prune.L1Unstructured, amount=0.2)
    assert abs(P.compute_sparsity_global(ours) - P.compute_sparsity_global(ref)) < 1e-9
    torch.testing.assert_close(ours(x), ref(x), rtol=1e-4, atol=1e-5)


def test_masked_sgd_bf16_weight_emit():
    torch.manual_seed(1)
    model = TinyNet().to(DEV)
    P.magnitude_pruning(model, 0.5)
    opt = MaskedSGD(model, lr=0.05, momentum=0.9, bf16_weights=True)
    x = torch.randn(4, 3, 16, 16, device=DEV)
    model(x).sum().backward()
    opt.step()
    for m, w16 in zip(_mods(model), opt.weff16):
        assert torch.equal(w16, m.weight.detach().to(torch.bfloat16))


def test_checkpoint_roundtrip_and_packed_masks():
    torch.manual_seed(2)
    model = TinyNet().to(DEV)
    P.magnitude_pruning(model, 0.5)
    P.magnitude_pruning(model, 0.2)
    sparsity = P.compute_sparsity_global(model)
    sd = {("module." + k): v.cpu() for k, v in model.state_dict().items()}               # as saved under DDP
    assert "module.f1.weight_orig" in sd and "module.f1.weight_mask" in sd and "module.f1.weight" not in sd
    # continue on the GPU from the checkpoint
    m2 = load_pruned(TinyNet().to(DEV), sd)
    assert prune.is_pruned(m2) and P.compute_sparsity_global(m2) == sparsity
    opt = MaskedSGD(m2, lr=0.1)
    m2(torch.randn(2, 3, 16, 16, device=DEV)).sum().backward(); opt.step()
    assert P.compute_sparsity_global(m2) == sparsity
    # inference copy on the CPU, masks folded in (evaluate_models.py:391-403)
    m3 = load_pruned(TinyNet(), sd, remove=True)
    assert not prune.is_pruned(m3)
    zeros = sum(int((m.weight == 0).sum()) for m in _mods(m3)); total = sum(m.weight.numel() for m in _mods(m3))
    assert 100.0 * zeros / total == sparsity
    # packed <-> fp32 masks
    st = packed_mask_state(model)
    assert st["words"].numel() * 4 < sum(st["numels"])                                  # 1 bit per parameter (+ chunk padding)
    back = fp32_masks_from_packed(st)
    for name, m in model.named_modules():
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            assert torch.equal(back[f"{name}.weight_mask"].view_as(m.weight_mask), m.weight_mask.cpu())


def test_cli_surface(capsys):
    p = add_pruning_args(argparse.ArgumentParser())
    a = p.parse_args([])
    assert (a.pruning_method, a.target_sparsity, a.pruning_rate, a.pruning_threshold, a.starting_pruning_iteration) == \
        ("magnitude", 0.9, 0.2, 95.0, 0)
    with pytest.raises(SystemExit):
        p.parse_args(["--pruning-method", "random"])
    torch.manual_seed(3)
    model = TinyNet().to(DEV)
    a = p.parse_args(["--pruning-method", "magnitude", "--pruning-rate", "0.5", "--pruning-threshold", "80"])
    final = run_pruning_schedule(model, a, None, DEV, nn.CrossEntropyLoss())
    out = capsys.readouterr().out
    assert "Initial sparsity: 0.00%" in out and "Pruning iteration: 0" in out and "Target Pruning Threshold: 80.0%" in out
    assert 80.0 <= final < 90.0 and out.count("Pruning iteration:") == 3                 # 50 -> 75 -> 87.5 %
    m2 = TinyNet().to(DEV)
    a = p.parse_args(["--pruning-method", "snip", "--target-sparsity", "0.7"])
    batch = [(torch.randn(4, 3, 16, 16), torch.randint(0, 10, (4,)))]
    final = run_pruning_schedule(m2, a, batch, DEV, nn.CrossEntropyLoss())
    assert abs(final - 70.0) < 0.5 and "Sparsity after SNIP pruning:" in capsys.readouterr().out
    a.pruning_method = "other"
    with pytest.raises(ValueError, match="Unsupported pruning method"):
        run_pruning_schedule(m2, a, batch, DEV, nn.CrossEntropyLoss())
