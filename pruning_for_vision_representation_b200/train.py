"""`python -m pruning_for_vision_representation_b200.train` — the reference's train.py command line on the B200 path.

Same flags, defaults and control flow as the reference's `main` / `get_args_parser` (train.py:530-875) for everything
that touches the pruning hot path: --pruning-method {magnitude,snip}, --target-sparsity, --pruning-rate,
--pruning-threshold, --starting-pruning-iteration, --model, --epochs, --batch-size, --opt, --lr, --momentum, --wd,
--norm-weight-decay, --bias-weight-decay, --transformer-embedding-decay, --lr-scheduler and its knobs, --amp,
--clip-grad-norm, --model-ema / --model-ema-steps / --model-ema-decay, --label-smoothing, --seed, --output-dir.
The dataset pipeline (ImageFolder, transforms, samplers, mixup) and wandb are the reference's own plumbing and out of
scope: the loader here is synthetic (`--synthetic-samples` random images of `--image-size`, fixed seed),
which is all this sandbox can offer (no datasets, no network); `--data-path` is accepted and ignored with a note.

    python -m pruning_for_vision_representation_b200.train --model resnet18 --pruning-method snip --target-sparsity 0.9 \\
        --epochs 1 --batch-size 32 --synthetic-samples 256 --image-size 64
"""
import argparse
import os

import torch
import torch.nn as nn

from .cli import add_pruning_args
from .masked_sgd import MaskedEMA, MaskedSGD
from .pruning import compute_sparsity_global, magnitude_pruning, snip_pruning
from .trainer import train_model_to_completion


class SyntheticImages(torch.utils.data.Dataset):
    def __init__(self, n, image_size, num_classes, seed):
        g = torch.Generator().manual_seed(seed)
        self.x = torch.randn(n, 3, image_size, image_size, generator=g)
        self.y = torch.randint(0, num_classes, (n,), generator=g)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i], self.y[i]


def get_args_parser(add_help=True):
    p = argparse.ArgumentParser(description="PyTorch Classification Training with Pruning (B200 path)", add_help=add_help)
    p.add_argument("--data-path", default=None, type=str, help="dataset path (ignored: synthetic data)")
    p.add_argument("--model", default="resnet18", type=str, help="model name")
    p.add_argument("--device", default="cuda", type=str, help="device (cuda only: the pruning path has no CPU fallback)")
    p.add_argument("-b", "--batch-size", default=32, type=int)
    p.add_argument("--epochs", default=90, type=int, metavar="N")
    p.add_argument("--seed", default=1, type=int)
    p.add_argument("-j", "--workers", default=0, type=int, metavar="N")
    p.add_argument("--opt", default="sgd", type=str)
    p.add_argument("--lr", default=0.1, type=float)
    p.add_argument("--momentum", default=0.9, type=float, metavar="M")
    p.add_argument("--wd", "--weight-decay", default=1e-4, type=float, metavar="W", dest="weight_decay")
    p.add_argument("--norm-weight-decay", default=None, type=float)
    p.add_argument("--bias-weight-decay", default=None, type=float)
    p.add_argument("--transformer-embedding-decay", default=None, type=float)
    p.add_argument("--label-smoothing", default=0.0, type=float)
    p.add_argument("--lr-scheduler", default="steplr", type=str)
    p.add_argument("--lr-warmup-epochs", default=0, type=int)
    p.add_argument("--lr-warmup-method", default="constant", type=str)
    p.add_argument("--lr-warmup-decay", default=0.01, type=float)
    p.add_argument("--lr-step-size", default=30, type=int)
    p.add_argument("--lr-gamma", default=0.1, type=float)
    p.add_argument("--lr-min", default=0.0, type=float)
    p.add_argument("--print-freq", default=10, type=int)
    p.add_argument("--output-dir", default="", type=str)
    p.add_argument("--amp", action="store_true", help="fp16 autocast + GradScaler, as the reference")
    p.add_argument("--amp-dtype", default=None, choices=[None, "bf16"], help="extension: bf16 autocast on the bf16 weights the step kernel emits")
    p.add_argument("--model-ema", action="store_true")
    p.add_argument("--model-ema-steps", type=int, default=32)
    p.add_argument("--model-ema-decay", type=float, default=0.99998)
    p.add_argument("--clip-grad-norm", default=None, type=float)
    p.add_argument("--world-size", default=1, type=int)
    p.add_argument("--synthetic-samples", default=512, type=int, help="size of the synthetic training set")
    p.add_argument("--image-size", default=224, type=int)
    p.add_argument("--num-classes", default=1000, type=int)
    p.add_argument("--max-rounds", default=None, type=int, help="stop the magnitude schedule after this many rounds")
    add_pruning_args(p)
    return p


def _ema_for(model, args, optimizer):
    adjust = args.world_size * args.batch_size * args.model_ema_steps / args.epochs          # train.py:634-637
    alpha = min(1.0, (1.0 - args.model_ema_decay) * adjust)
    return MaskedEMA(model, 1.0 - alpha, optimizer)


def main(args):
    import torchvision
    if args.data_path:
        print(f"note: --data-path {args.data_path} ignored, training on {args.synthetic_samples} synthetic images")
    device = torch.device(args.device)
    args.distributed = False
    torch.manual_seed(args.seed)
    dataset = SyntheticImages(args.synthetic_samples, args.image_size, args.num_classes, args.seed)
    data_loader = torch.utils.data.DataLoader(dataset, batch_size=args.batch_size, shuffle=True, num_workers=args.workers, drop_last=True)
    print("Creating model")
    if "vit" in args.model:                                                                  # train.py:592-593
        model = torchvision.models.vit_b_32(weights=None, num_classes=args.num_classes, image_size=args.image_size)
    else:
        model = torchvision.models.get_model(args.model, weights=None, num_classes=args.num_classes)
    model.to(device)
    criterion = nn.CrossEntropyLoss(label_smoothing=args.label_smoothing)
    scaler = torch.amp.GradScaler("cuda") if args.amp else None
    saves = []

    def save_fn(checkpoint, epoch):
        if args.output_dir:
            os.makedirs(args.output_dir, exist_ok=True)
            path = os.path.join(args.output_dir, f"{args.model}_checkpoint_{args.pruning_method}_{args.target_sparsity}.pth")
            torch.save({k: v for k, v in checkpoint.items() if k != "args"}, path)
            saves.append(path)

    if args.pruning_method == "snip":                                                        # train.py:617-652
        snip_pruning(model=model, data_loader=data_loader, device=device, criterion=criterion, target_sparsity=args.target_sparsity)
        sparsity = compute_sparsity_global(model)
        print(f"Sparsity after SNIP pruning: {sparsity:.2f}%")
        _, final = train_model_to_completion(model, data_loader, None, criterion, args, device, scaler=scaler,
                                             model_ema=True if args.model_ema else None, save_fn=save_fn)
        print(f"Final sparsity after SNIP and training: {final:.2f}%")
    elif args.pruning_method == "magnitude":                                                 # train.py:654-708
        it = args.starting_pruning_iteration
        sparsity = compute_sparsity_global(model)
        print(f"Initial sparsity: {sparsity:.2f}%")
        rounds = 0
        while sparsity < args.pruning_threshold and (args.max_rounds is None or rounds < args.max_rounds):
            print(f"Pruning iteration: {it}")
            _, sparsity = train_model_to_completion(model, data_loader, None, criterion, args, device, scaler=scaler,
                                                    model_ema=True if args.model_ema else None,
                                                    global_wandb_step_offset=args.epochs * it, save_fn=save_fn)
            magnitude_pruning(model=model, prune_amount=args.pruning_rate)
            sparsity = compute_sparsity_global(model)
            print(f"Current Sparsity: {sparsity:.2f}%")
            print(f"Target Pruning Threshold: {args.pruning_threshold}%")
            it += 1
            rounds += 1
    else:
        raise ValueError(f"Unsupported pruning method: {args.pruning_method}. Choose 'snip' or 'magnitude'.")
    print("Training completed successfully")
    return model


if __name__ == "__main__":
    main(get_args_parser().parse_args())
