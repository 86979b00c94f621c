"""The pruning surface of the reference's train.py (flags train.py:733-764, dispatch train.py:617-711)
without the trainer around it: the same flag names, defaults and control flow, calling the CUDA path.

    parser = add_pruning_args(argparse.ArgumentParser())
    run_pruning_schedule(model, args, data_loader, device, criterion, train_fn)

`train_fn(model) -> None` stands for the reference's train_model_to_completion (out of scope here)."""
import argparse

from .pruning import compute_sparsity_global, magnitude_pruning, snip_pruning


def add_pruning_args(parser=None):
    parser = parser or argparse.ArgumentParser(description="B200 pruning path", add_help=True)
    parser.add_argument("--pruning-method", default="magnitude", type=str, choices=["magnitude", "snip"],
                        help="Pruning method to use: magnitude-based (iterative) or SNIP (one-shot)")      # train.py:734-740
    parser.add_argument("--target-sparsity", default=0.9, type=float,
                        help="Target sparsity for SNIP pruning (0.0-1.0)")                                   # train.py:741-746
    parser.add_argument("--pruning-rate", default=0.2, type=float,
                        help="Fraction of the surviving weights pruned per magnitude round")                 # train.py:747-752
    parser.add_argument("--pruning-threshold", default=95.0, type=float,
                        help="Stop iterative pruning when the global sparsity (percent) reaches this")        # train.py:753-758
    parser.add_argument("--starting-pruning-iteration", default=0, type=int,
                        help="Pruning iteration to start counting from")                                      # train.py:759-764
    return parser


def run_pruning_schedule(model, args, data_loader, device, criterion, train_fn=None, max_rounds=None):
    """train.py:617-711.  Returns the final sparsity in percent."""
    train_fn = train_fn or (lambda m: None)
    if args.pruning_method == "snip":
        snip_pruning(model=model, data_loader=data_loader, device=device, criterion=criterion,
                     target_sparsity=args.target_sparsity)
        sparsity = compute_sparsity_global(model)
        print(f"Sparsity after SNIP pruning: {sparsity:.2f}%")
        train_fn(model)
        final = compute_sparsity_global(model)
        print(f"Final sparsity after SNIP and training: {final:.2f}%")
        return final
    if args.pruning_method == "magnitude":
        it = args.starting_pruning_iteration
        sparsity = compute_sparsity_global(model)
        print(f"Initial sparsity: {sparsity:.2f}%")
        rounds = 0
        while sparsity < args.pruning_threshold and (max_rounds is None or rounds < max_rounds):
            print(f"Pruning iteration: {it}")
            train_fn(model)
            magnitude_pruning(model=model, prune_amount=args.pruning_rate)
            sparsity = compute_sparsity_global(model)
            print(f"Current Sparsity: {sparsity:.2f}%")
            print(f"Target Pruning Threshold: {args.pruning_threshold}%")
            it += 1
            rounds += 1
        return sparsity
    raise ValueError(f"Unsupported pruning method: {args.pruning_method}. Choose 'snip' or 'magnitude'.")   # train.py:711
