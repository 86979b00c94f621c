"""B200-native pruning hot path of EIDOSLAB/pruning-for-vision-representation.

Python host mirror of the reference's functions over hand-written sm_100a CUDA kernels reached
through the C-ABI of include/b200prune.h (ctypes).  No CPU fallback.

    from pruning_for_vision_representation_b200 import (
        snip_pruning, magnitude_pruning, compute_sparsity_global,     # train.py:241-369
        lost, patch_scoring, detect_box)                              # object_discovery.py:23-134
"""
from ._lib import B200PruneError  # noqa: F401

_LAZY = {
    "snip_pruning": "pruning", "magnitude_pruning": "pruning", "compute_sparsity_global": "pruning",
    "prunable_modules": "pruning", "export_masks": "pruning",
    "ParamPlan": "plan",
    "MaskedSGD": "masked_sgd",
    "lost": "object_discovery", "patch_scoring": "object_discovery", "detect_box": "object_discovery",
    "lost_batched": "object_discovery",
    "load_pruned": "checkpoint", "packed_mask_state": "checkpoint", "fp32_masks_from_packed": "checkpoint",
    "add_pruning_args": "cli", "run_pruning_schedule": "cli",
    "ShardedMaskBuilder": "distributed",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod = importlib.import_module(f".{_LAZY[name]}", __name__)
        return getattr(mod, name)
    raise AttributeError(name)
