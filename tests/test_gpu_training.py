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


# ---- training-loop integration (SURVEY f-2): clipping, GradScaler, EMA, weight-decay groups, checkpoint / resume -----------
class _NormNet(nn.Module):
    """TinyNet with a BatchNorm and a LayerNorm, so that the norm / bias weight-decay groups are not empty."""

    def __init__(self):
        super().__init__()
        self.c1 = nn.Conv2d(3, 13, 3, padding=1)
        self.bn = nn.BatchNorm2d(13)
        self.c2 = nn.Conv2d(13, 24, 3, padding=1)
        self.pool = nn.AdaptiveAvgPool2d(4)
        self.f1 = nn.Linear(24 * 16, 67)
        self.ln = nn.LayerNorm(67)
        self.f2 = nn.Linear(67, 10)

    def forward(self, x):
        x = torch.relu(self.bn(self.c1(x)))
        x = torch.relu(self.c2(x))
        x = self.pool(x).flatten(1)
        return self.f2(torch.relu(self.ln(self.f1(x))))


def _pruned_pair(seed, amount=0.6):
    torch.manual_seed(seed)
    ours = _NormNet().to(DEV)
    ref = copy.deepcopy(ours)
    P.magnitude_pruning(ours, amount)
    for mo, mr in zip(_mods(ours), _mods(ref)):
        prune.custom_from_mask(mr, "weight", mo.weight_mask.clone())
    return ours, ref


def _ref_groups(model, wd, norm_wd, bias_wd):
    """utils.set_weight_decay (utils.py:405-463) restated for the reference side of the comparison."""
    from pruning_for_vision_representation_b200.masked_sgd import set_weight_decay
    return set_weight_decay(model, wd, norm_weight_decay=norm_wd, custom_keys_weight_decay=[("bias", bias_wd)])


class _RefEMA:
    """utils.py:159-170 (AveragedModel, `decay * avg + (1 - decay) * param`, use_buffers=True) restated over the tensors of
    a torch-pruned model: AveragedModel itself deep-copies the model, which torch refuses for a module whose `weight` is
    the non-leaf `weight_mask * weight_orig` that prune.custom_from_mask installs."""

    def __init__(self, model, decay, device):
        self.decay = decay
        self.items = list(model.named_parameters()) + list(model.named_buffers())
        self.avg = [t.detach().clone() for _, t in self.items]
        self.n_averaged = torch.tensor(0, dtype=torch.long, device=device)

    @torch.no_grad()
    def update_parameters(self, model):
        for a, (_, t) in zip(self.avg, self.items):
            if int(self.n_averaged) == 0:
                a.copy_(t.detach())                                   # AveragedModel.update_parameters, n_averaged == 0
            else:
                a.copy_((self.decay * a + (1 - self.decay) * t.detach()).to(a.dtype))
        self.n_averaged += 1

    def state_dict(self):
        sd = {"n_averaged": self.n_averaged}
        sd.update({"module." + name: a for (name, _), a in zip(self.items, self.avg)})
        return sd


@pytest.mark.parametrize("use_scaler", [False, True])
def test_clip_gradscaler_ema_and_decay_groups_match_torch(use_scaler):
    """train.py:54-73 step for step: (scaled) backward, unscale_, clip_grad_norm_ over the masked gradients, (scaler.)step,
    EMA every 2 steps — ours through MaskedSGD / MaskedEMA, the reference through torch prune + SGD + GradScaler +
    AveragedModel.  One step carries an inf gradient: with a scaler both sides must skip it."""
    from pruning_for_vision_representation_b200.masked_sgd import MaskedEMA
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ours, ref = _pruned_pair(4)
    hp = dict(lr=0.05, momentum=0.9, weight_decay=1e-3, nesterov=False)
    opt_o = MaskedSGD(ours, norm_weight_decay=0.0, custom_keys_weight_decay=[("bias", 1e-5)], **hp)
    opt_r = torch.optim.SGD(_ref_groups(ref, 1e-3, 0.0, 1e-5), **hp)
    # the groups of utils.set_weight_decay: pruned weights ("other", fused) / norm / bias
    assert [g["weight_decay"] for g in opt_o.param_groups] == [1e-3, 0.0, 1e-5] == [g["weight_decay"] for g in opt_r.param_groups]
    ema_o = MaskedEMA(ours, 0.9, opt_o)
    ema_r = _RefEMA(ref, 0.9, DEV)
    sc_o = torch.amp.GradScaler("cuda", init_scale=1024.0) if use_scaler else None
    sc_r = torch.amp.GradScaler("cuda", init_scale=1024.0) if use_scaler else None
    crit = nn.CrossEntropyLoss()
    max_norm = 0.5
    for step in range(6):
        x = torch.randn(8, 3, 16, 16, device=DEV); y = torch.randint(0, 10, (8,), device=DEV)
        poison = use_scaler and step == 3
        norms = []
        for model, opt, sc, fused in ((ours, opt_o, sc_o, True), (ref, opt_r, sc_r, False)):
            opt.zero_grad()
            loss = crit(model(x), y)
            if poison:
                loss = loss * float("inf")
            if sc is not None:
                sc.scale(loss).backward()
                sc.unscale_(opt)
                norms.append(opt.clip_grad_norm_(max_norm) if fused else nn.utils.clip_grad_norm_(model.parameters(), max_norm))
                sc.step(opt); sc.update()
            else:
                loss.backward()
                norms.append(opt.clip_grad_norm_(max_norm) if fused else nn.utils.clip_grad_norm_(model.parameters(), max_norm))
                opt.step()
        if not poison:
            torch.testing.assert_close(norms[0], norms[1], rtol=1e-5, atol=1e-7)          # norm of the MASKED gradients
        if step % 2 == 0:
            ema_o.update_parameters(ours); ema_r.update_parameters(ref)
        for mo, mr in zip(_mods(ours), _mods(ref)):
            torch.testing.assert_close(mo.weight_orig, mr.weight_orig, rtol=3e-5, atol=3e-6)
            torch.testing.assert_close(mo.bias, mr.bias, rtol=3e-5, atol=3e-6)
        torch.testing.assert_close(ours.bn.weight, ref.bn.weight, rtol=3e-5, atol=3e-6)
        torch.testing.assert_close(ours.ln.bias, ref.ln.bias, rtol=3e-5, atol=3e-6)
    if use_scaler:
        assert sc_o.get_scale() == sc_r.get_scale() == 512.0                                 # exactly one skipped step on both sides
    sd_o, sd_r = ema_o.state_dict(), ema_r.state_dict()
    assert set(sd_o) == set(sd_r)
    for key in sd_r:
        torch.testing.assert_close(sd_o[key].float(), sd_r[key].float(), rtol=3e-5, atol=3e-6, msg=key)


def test_masked_sgd_state_dict_resume_matches_torch():
    """train.py:507 checkpoints optimizer.state_dict(); a resumed run must continue with the saved momentum."""
    ours, ref = _pruned_pair(5)
    hp = dict(lr=0.1, momentum=0.9, weight_decay=1e-4)
    opt_o, opt_r = MaskedSGD(ours, **hp), torch.optim.SGD(ref.parameters(), **hp)
    crit = nn.CrossEntropyLoss()
    data = [(torch.randn(8, 3, 16, 16, device=DEV), torch.randint(0, 10, (8,), device=DEV)) for _ in range(5)]

    def run(model, opt, batches):
        for x, y in batches:
            opt.zero_grad(); crit(model(x), y).backward(); opt.step()

    run(ours, opt_o, data[:3]); run(ref, opt_r, data[:3])
    ck_model, ck_opt = copy.deepcopy(ours.state_dict()), copy.deepcopy(opt_o.state_dict())
    assert any("momentum_buffer" in s for s in ck_opt["state"].values()) and ck_opt["inner"] is not None
    # a new process: fresh model from the checkpoint, fresh optimizer, load, continue
    m2 = load_pruned(_NormNet().to(DEV), {k: v.cpu() for k, v in ck_model.items()})
    opt2 = MaskedSGD(m2, **hp)
    assert opt2._first
    opt2.load_state_dict(ck_opt)
    assert not opt2._first
    run(m2, opt2, data[3:]); run(ref, opt_r, data[3:])
    for mo, mr in zip(_mods(m2), _mods(ref)):
        torch.testing.assert_close(mo.weight_orig, mr.weight_orig, rtol=3e-5, atol=3e-6)
        torch.testing.assert_close(mo.bias, mr.bias, rtol=3e-5, atol=3e-6)


def test_remove_then_prune_starts_from_all_weights():
    """prune.remove (main_lost.py:67) drops the reparametrisation: a later magnitude_pruning is a fresh global L1 over all N."""
    torch.manual_seed(6)
    m = TinyNet().to(DEV)
    P.magnitude_pruning(m, 0.5)
    for mod in _mods(m):
        prune.remove(mod, "weight")
    P.magnitude_pruning(m, 0.5)                       # the smallest half of ALL entries = the zeros already there
    assert abs(P.compute_sparsity_global(m) - 50.0) < 0.01


def _ddp_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(7)
        model = TinyNet().to(DEV)
        P.magnitude_pruning(model, 0.5)
        opt = MaskedSGD(model, lr=0.1, momentum=0.9)
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[0], broadcast_buffers=False, find_unused_parameters=True)   # train.py:606
        crit = nn.CrossEntropyLoss()
        g = torch.Generator(device="cpu").manual_seed(100 + rank)                 # every rank its own data
        for _ in range(3):
            x = torch.randn(8, 3, 16, 16, generator=g).to(DEV); y = torch.randint(0, 10, (8,), generator=g).to(DEV)
            opt.zero_grad(); crit(ddp(x), y).backward(); opt.clip_grad_norm_(1.0); opt.step()
        flat = torch.cat([m.weight_orig.detach().reshape(-1) for m in _mods(model)] + [m.bias.detach().reshape(-1) for m in _mods(model)])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        if rank == 0:
            out.put(bool(all(torch.equal(gathered[0], t) for t in gathered)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_under_ddp_stay_identical():
    """Under the reference's DDP wrapper (find_unused_parameters=True) the fused leaves are not module parameters; MaskedSGD
    reduces their gradients itself.  Two ranks (gloo, both on GPU 0) with different data: weight_orig identical after 3 steps."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) is True


def test_train_cli_runs_both_schedules(tmp_path, capsys):
    """python -m ...train with the reference's flags, on synthetic data: SNIP one-shot + training, and two magnitude rounds
    with fp16 AMP, clipping, EMA and a checkpoint that holds the reference's keys."""
    from pruning_for_vision_representation_b200 import train as T
    args = T.get_args_parser().parse_args(["--model", "resnet18", "--pruning-method", "snip", "--target-sparsity", "0.8", "--epochs", "1",
                                           "--batch-size", "16", "--synthetic-samples", "48", "--image-size", "32", "--num-classes", "10",
                                           "--clip-grad-norm", "1.0", "--amp-dtype", "bf16", "--output-dir", str(tmp_path)])
    model = T.main(args)
    out = capsys.readouterr().out
    assert "SNIP threshold:" in out and "Final sparsity after SNIP and training:" in out and "Training completed successfully" in out
    assert abs(P.compute_sparsity_global(model) - 80.0) < 0.5 and isinstance(model._b200p_fused_optimizer, MaskedSGD)
    ck = torch.load(tmp_path / "resnet18_checkpoint_snip_0.8.pth", weights_only=False)
    assert {"model", "optimizer", "lr_scheduler", "epoch", "sparsity"} <= set(ck) and "conv1.weight_orig" in ck["model"] and "conv1.weight_mask" in ck["model"]
    args = T.get_args_parser().parse_args(["--model", "resnet18", "--pruning-method", "magnitude", "--pruning-rate", "0.5", "--epochs", "1",
                                           "--batch-size", "16", "--synthetic-samples", "32", "--image-size", "32", "--num-classes", "10",
                                           "--amp", "--clip-grad-norm", "1.0", "--model-ema", "--model-ema-steps", "1", "--max-rounds", "2",
                                           "--norm-weight-decay", "0.0"])
    model = T.main(args)
    out = capsys.readouterr().out
    assert out.count("Pruning iteration:") == 2 and abs(P.compute_sparsity_global(model) - 75.0) < 0.5
