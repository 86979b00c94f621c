"""Pruned-checkpoint helpers.

The reference's checkpoints carry a pruned layer as `<name>.weight_orig` + `<name>.weight_mask`
(fp32 0/1, same shape) and no `<name>.weight` (train.py:504-521); every consumer re-implements
`prune.identity -> load_state_dict -> prune.remove` (evaluate_models.py:391-403, main_lost.py:169-182,
explain.py:231-243).  These helpers do that once, keep the format, and move between the fp32 masks and
the bit-packed mask the kernels use (1/32 of the size)."""
from collections import OrderedDict

import torch
import torch.nn.utils.prune as prune

from . import _lib as L
from .pruning import B200MaskMethod, _get_state, export_masks, prunable_modules


def load_pruned(model, state_dict, remove=False, strict=True):
    """Load a reference-format pruned checkpoint into `model` (any device).

    Strips a DDP `module.` prefix, reparametrises exactly the layers the checkpoint has masks for,
    loads, and either keeps the reparametrisation (remove=False: training continues, masks adopted by
    the CUDA path on the next call) or folds the masks into plain weights (remove=True: inference,
    like evaluate_models.py:401)."""
    sd = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in state_dict.items())
    masked = {k[:-len(".weight_mask")] for k in sd if k.endswith(".weight_mask")}
    mods = dict(model.named_modules())
    for name in masked:
        m = mods[name]
        if not prune.is_pruned(m):
            if next(m.parameters()).is_cuda:
                _identity_b200(m)
            else:
                prune.identity(m, "weight")
    model.load_state_dict(sd, strict=strict)
    if hasattr(model, "_b200p_state"):
        object.__delattr__(model, "_b200p_state")        # masks changed underneath: rebuild lazily
    if remove:
        for name in masked:
            prune.remove(mods[name], "weight")
    return model


def _identity_b200(module):
    """prune.identity with this package's hook object (so MaskedSGD can take the layer over)."""
    orig = module._parameters.pop("weight")
    module.register_parameter("weight_orig", orig)
    module.register_buffer("weight_mask", torch.ones_like(orig))
    method = B200MaskMethod()
    method._tensor_name = "weight"
    module.register_forward_pre_hook(method)
    setattr(module, "weight", method.apply_mask(module))


def packed_mask_state(model):
    """{'words': int32 CPU tensor, 'numels': [...], 'names': [...]} — the kernel-side mask, 1 bit per
    parameter, chunk-major as documented in include/b200prune.h."""
    mask, plan = export_masks(model)
    names = [n for n, m in prunable_modules(model) if "weight_orig" in m._parameters]
    return {"words": mask.cpu(), "numels": list(plan.numels), "names": names, "chunk": L.CHUNK}


def fp32_masks_from_packed(state):
    """Inverse of packed_mask_state for consumers of the reference format: {name + '.weight_mask': fp32}."""
    from .plan import unpack_mask_words
    segs = unpack_mask_words(state["words"].numpy(), state["numels"])
    return {f"{n}.weight_mask": torch.from_numpy(s.astype("float32")) for n, s in zip(state["names"], segs)}
