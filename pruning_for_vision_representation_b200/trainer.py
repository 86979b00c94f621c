"""Training-loop integration of the fused masked step: the reference's `create_optimizer`,
`train_one_epoch` and `train_model_to_completion` (train.py:35-89, 372-392, 434-527) with the same signatures, argument
meaning and control flow, over MaskedSGD (one kernel per step for all pruned weights) instead of forward pre-hook +
MulBackward + foreach SGD.

What changes against the reference, and nothing else:
  * the optimizer is MaskedSGD when the model is pruned (torch.optim.SGD on the same groups when it is not);
  * gradient clipping goes through `optimizer.clip_grad_norm_` (the norm of the MASKED gradients, as in the reference,
    computed by one reduction pass and folded into the step kernel);
  * `--amp` keeps the reference's fp16 autocast + GradScaler (train.py:50,609); `amp_dtype="bf16"` (an extension, SURVEY
    §8d config 4) runs bf16 autocast on the bf16 masked weights the step kernel emits and needs no scaler;
  * EMA is a MaskedEMA (same update rule, same cadence `i % model_ema_steps`, same warm-up reset);
  * wandb logging, the evaluation loop and checkpoint files are the caller's business (out of scope): the
    hooks `log_fn` / `eval_fn` / `save_fn` stand where the reference calls them.
"""
import time

import torch

from .masked_sgd import MaskedEMA, MaskedSGD, set_weight_decay
from .pruning import compute_sparsity_global, prunable_modules


def is_pruned(model):
    return any("weight_orig" in m._parameters for _, m in prunable_modules(model))


def build_param_groups(model, args):
    """train.py:447-459: weight-decay groups from --norm-weight-decay / --bias-weight-decay / --transformer-embedding-decay."""
    custom = []
    if getattr(args, "bias_weight_decay", None) is not None:
        custom.append(("bias", args.bias_weight_decay))
    if getattr(args, "transformer_embedding_decay", None) is not None:
        for key in ["class_token", "position_embedding", "relative_position_bias_table"]:
            custom.append((key, args.transformer_embedding_decay))
    return set_weight_decay(model, args.weight_decay, norm_weight_decay=getattr(args, "norm_weight_decay", None),
                            custom_keys_weight_decay=custom if custom else None)


def create_optimizer(args, parameters, model=None):
    """train.py:372-392.  With a pruned `model` the SGD variants return a MaskedSGD over the same groups; RMSprop / AdamW
    (not part of the fused path) and unpruned models get the torch optimizers the reference builds."""
    opt_name = args.opt.lower()
    if opt_name.startswith("sgd"):
        if model is not None and is_pruned(model) and next(model.parameters()).is_cuda:
            return MaskedSGD(model, lr=args.lr, momentum=args.momentum, weight_decay=args.weight_decay,
                             nesterov="nesterov" in opt_name, param_groups=parameters,
                             bf16_weights=getattr(args, "amp_dtype", None) == "bf16")
        return torch.optim.SGD(parameters, lr=args.lr, momentum=args.momentum, weight_decay=args.weight_decay,
                               nesterov="nesterov" in opt_name)
    if opt_name == "rmsprop":
        return torch.optim.RMSprop(parameters, lr=args.lr, momentum=args.momentum, weight_decay=args.weight_decay, eps=0.0316, alpha=0.9)
    if opt_name == "adamw":
        return torch.optim.AdamW(parameters, lr=args.lr, weight_decay=args.weight_decay)
    raise RuntimeError(f"Invalid optimizer {args.opt}. Only SGD, RMSprop and AdamW are supported.")      # train.py:390


def create_lr_scheduler(args, optimizer):
    """The reference's step / cosine / exponential schedules with linear or constant warm-up (train.py:395-431)."""
    name = getattr(args, "lr_scheduler", "steplr").lower()
    epochs, warm = args.epochs, getattr(args, "lr_warmup_epochs", 0)
    if name == "steplr":
        main = torch.optim.lr_scheduler.StepLR(optimizer, step_size=getattr(args, "lr_step_size", 30), gamma=getattr(args, "lr_gamma", 0.1))
    elif name == "cosineannealinglr":
        main = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=epochs - warm, eta_min=getattr(args, "lr_min", 0.0))
    elif name == "exponentiallr":
        main = torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=getattr(args, "lr_gamma", 0.1))
    else:
        raise RuntimeError(f"Invalid lr scheduler '{name}'. Only StepLR, CosineAnnealingLR and ExponentialLR are supported.")
    if warm > 0:
        method = getattr(args, "lr_warmup_method", "constant")
        decay = getattr(args, "lr_warmup_decay", 0.01)
        if method == "linear":
            w = torch.optim.lr_scheduler.LinearLR(optimizer, start_factor=decay, total_iters=warm)
        elif method == "constant":
            w = torch.optim.lr_scheduler.ConstantLR(optimizer, factor=decay, total_iters=warm)
        else:
            raise RuntimeError(f"Invalid warmup lr method '{method}'. Only linear and constant are supported.")
        return torch.optim.lr_scheduler.SequentialLR(optimizer, schedulers=[w, main], milestones=[warm])
    return main


def _autocast(args, scaler):
    if getattr(args, "amp_dtype", None) == "bf16":
        return torch.autocast("cuda", dtype=torch.bfloat16)
    return torch.autocast("cuda", dtype=torch.float16, enabled=scaler is not None)       # train.py:50


def train_one_epoch(model, criterion, optimizer, data_loader, device, epoch, args, model_ema=None, scaler=None,
                    split="train", global_wandb_step=0, log_fn=None):
    """train.py:35-89.  Returns {"loss", "acc1", "acc5", "img_s"} averaged over the epoch (the reference logs them)."""
    model.train()
    fused = isinstance(optimizer, MaskedSGD)
    clip = getattr(args, "clip_grad_norm", None)
    tot_loss = torch.zeros((), device=device)
    tot_a1 = torch.zeros((), device=device)
    tot_a5 = torch.zeros((), device=device)
    n_img, n_it, t_epoch = 0, 0, time.time()
    for i, (image, target) in enumerate(data_loader):
        image, target = image.to(device), target.to(device)
        with _autocast(args, scaler):
            output = model(image)
            loss = criterion(output, target)
        optimizer.zero_grad()
        if scaler is not None:
            scaler.scale(loss).backward()
            if clip is not None:
                scaler.unscale_(optimizer)                      # the leaves sit in param_groups[0]: unscaled and inf-checked with the rest
                if fused:
                    optimizer.clip_grad_norm_(clip)
                else:
                    torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
            scaler.step(optimizer)
            scaler.update()
        else:
            loss.backward()
            if clip is not None:
                if fused:
                    optimizer.clip_grad_norm_(clip)
                else:
                    torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
            optimizer.step()
        if model_ema and i % getattr(args, "model_ema_steps", 32) == 0:
            model_ema.update_parameters(model)
            if epoch < getattr(args, "lr_warmup_epochs", 0):
                model_ema.n_averaged.fill_(0)                  # train.py:71-73: keep copying during warm-up
        with torch.no_grad():
            k5 = min(5, output.shape[1])
            top = output.topk(k5, dim=1).indices
            hit = top == target.view(-1, 1)
            tot_a1 += hit[:, :1].any(dim=1).float().mean() * 100.0
            tot_a5 += hit.any(dim=1).float().mean() * 100.0
            tot_loss += loss.detach().float()
        n_img += image.shape[0]
        n_it += 1
    dt = max(time.time() - t_epoch, 1e-9)
    out = {"loss": float(tot_loss) / max(n_it, 1), "acc1": float(tot_a1) / max(n_it, 1), "acc5": float(tot_a5) / max(n_it, 1),
           "img_s": n_img / dt}
    if log_fn is not None:
        log_fn({f"{split}/{k}": v for k, v in out.items()}, global_wandb_step)
    return out


def train_model_to_completion(model, data_loader, data_loader_test, criterion, args, device, scaler=None, initial_epoch=0,
                              model_ema=None, global_wandb_step_offset=0, eval_fn=None, save_fn=None, log_fn=None):
    """train.py:434-527: a FRESH optimizer and LR scheduler for this round over the reference's weight-decay groups,
    `args.epochs` epochs, checkpoint dict with the reference's keys handed to `save_fn`.  Returns (model, sparsity)."""
    print("Starting standard training to completion")
    model_without_ddp = model.module if getattr(args, "distributed", False) else model
    parameters = build_param_groups(model_without_ddp, args)
    optimizer = create_optimizer(args=args, parameters=parameters, model=model_without_ddp)
    lr_scheduler = create_lr_scheduler(args=args, optimizer=optimizer)
    if model_ema is True:                                       # build the EMA against THIS round's optimizer
        model_ema = MaskedEMA(model_without_ddp, getattr(args, "model_ema_decay", 0.99998), optimizer) if isinstance(optimizer, MaskedSGD) else None
    sparsity = compute_sparsity_global(model_without_ddp)
    print(f"Starting training with sparsity: {sparsity:.2f}%")
    start = time.time()
    history = []
    for epoch in range(initial_epoch, args.epochs):
        step = epoch + global_wandb_step_offset
        if getattr(args, "distributed", False) and hasattr(args, "train_sampler"):
            args.train_sampler.set_epoch(epoch)
        stats = train_one_epoch(model, criterion, optimizer, data_loader, device, epoch, args, model_ema=model_ema, scaler=scaler,
                                global_wandb_step=step, log_fn=log_fn)
        history.append(stats)
        lr_scheduler.step()
        if eval_fn is not None and data_loader_test is not None:
            eval_fn(model, criterion, data_loader_test, device, step)
        if save_fn is not None:
            checkpoint = {"model": model_without_ddp.state_dict(), "optimizer": optimizer.state_dict(),
                          "lr_scheduler": lr_scheduler.state_dict(), "epoch": epoch, "args": args, "sparsity": sparsity}
            if model_ema:
                checkpoint["model_ema"] = model_ema.state_dict()
            if scaler:
                checkpoint["scaler"] = scaler.state_dict()
            save_fn(checkpoint, epoch)
    print(f"Training time {int(time.time() - start)} s")
    train_model_to_completion.last_history = history
    return model, sparsity
