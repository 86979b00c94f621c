"""MaskedSGD — the masked training step of a pruned model as ONE kernel per step (K4, csrc/sgd.cu).

In the reference a pruned Conv2d/Linear costs, per step and per parameter, a forward pre-hook
(`weight = weight_mask * weight_orig`, torch/nn/utils/prune.py:53-74), its MulBackward, and the
foreach passes of `torch.optim.SGD` (torch/optim/sgd.py:343-380, built at train.py:372-392).
Here the prunable weights share one ParamPlan and one launch does all of it:

    g       = mask ? grad(weight) : 0                 (MulBackward)
    g      += wd * weight_orig;  buf = mu * buf + (1 - damp) * g;  g = nesterov ? g + mu * buf : buf
    weight_orig -= lr * g                             (SGD, same update order as torch)
    weight  = mask ? weight_orig : 0                  (next forward's masked weight, fp32 and/or bf16)

`module.weight` becomes a persistent leaf tensor maintained by the kernel (the pre-hook returns it
instead of re-multiplying), so pruned weights and their gradients never re-densify.  As in the
reference, pruned entries of `weight_orig` keep decaying and keep a momentum buffer.  Every other
parameter (biases, norms, unpruned tensors) is stepped by an inner `torch.optim.SGD` with the same
hyper-parameters; `param_groups[0]` is the fused group, so LR schedulers work unchanged.
"""
import torch
import torch.nn as nn
import torch.nn.utils.prune as prune

from . import _lib as L
from ._lib import B200PruneError
from .pruning import B200MaskMethod, _get_state, _flat, prunable_modules


class MaskedSGD(torch.optim.Optimizer):
    def __init__(self, model, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False,
                 norm_weight_decay=None, bf16_weights=False):
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")       # torch/optim/sgd.py
        modules = [m for _, m in prunable_modules(model) if "weight_orig" in m._parameters]
        if not modules:
            raise B200PruneError("MaskedSGD: the model has no pruned Conv2d/Linear (prune it first)")
        st = _get_state(model, modules)
        if st.mask is None:
            raise B200PruneError("MaskedSGD: no packed mask (prune the model with this package or load a pruned checkpoint)")
        self.state_ref, self.plan, self.modules = st, st.plan, modules
        fused = [m._parameters["weight_orig"] for m in modules]
        fused_ids = {id(p) for p in fused}
        norm_classes = (nn.modules.batchnorm._BatchNorm, nn.LayerNorm, nn.GroupNorm)
        norm_params, other = [], []
        for mod in model.modules():
            for p in mod.parameters(recurse=False):
                if not p.requires_grad or id(p) in fused_ids:
                    continue
                (norm_params if isinstance(mod, norm_classes) and norm_weight_decay is not None else other).append(p)
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
        groups = [{"params": fused}]
        if other:
            groups.append({"params": other})
        if norm_params:
            groups.append({"params": norm_params, "weight_decay": norm_weight_decay})          # utils.set_weight_decay
        super().__init__(groups, defaults)
        self.inner = torch.optim.SGD([{k: v for k, v in g.items()} for g in self.param_groups[1:]], lr=lr, momentum=momentum,
                                     dampening=dampening, weight_decay=weight_decay, nesterov=nesterov) if len(groups) > 1 else None
        # persistent buffers: momentum, effective weights (the leaves autograd differentiates)
        dev = st.device
        self.bufs = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in fused]
        self.weff = [torch.zeros_like(p, memory_format=torch.contiguous_format).requires_grad_(True) for p in fused]
        self.weff16 = [torch.zeros_like(p, dtype=torch.bfloat16, memory_format=torch.contiguous_format) for p in fused] if bf16_weights else None
        self._grad_zero = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in fused]
        self._first = True
        self.plan.bind(L.SLOT_W, [_flat(p.detach()) for p in fused])
        self.plan.bind(L.SLOT_BUF, [_flat(b) for b in self.bufs])
        self.plan.bind(L.SLOT_WEFF, [_flat(w.detach()) for w in self.weff])
        if self.weff16 is not None:
            self.plan.bind(L.SLOT_WEFF16, [_flat(w) for w in self.weff16])
        st._w_ptrs = None
        self.plan.apply_mask(st.mask, L.EMIT_WEFF | (L.SGD_EMIT_WEFF16 if bf16_weights else 0))
        # hand the leaves to the forward pre-hooks
        for m, w in zip(modules, self.weff):
            for hook in m._forward_pre_hooks.values():
                if isinstance(hook, prune.BasePruningMethod) and hook._tensor_name == "weight":
                    if not isinstance(hook, B200MaskMethod):
                        raise B200PruneError("MaskedSGD: module was reparametrised by torch.nn.utils.prune; "
                                             "call magnitude_pruning / snip_pruning of this package (or load_pruned) first")
                    hook.fused_weight = w
            setattr(m, "weight", w)

    def zero_grad(self, set_to_none=True):
        for w in self.weff:
            w.grad = None
        if self.inner is not None:
            self.inner.zero_grad(set_to_none=set_to_none)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g0 = self.param_groups[0]
        grads = []
        for w, z in zip(self.weff, self._grad_zero):
            g = w.grad
            if g is None:
                g = z                                            # torch skips params without grad; wd/momentum still need a pass here
            elif not g.is_contiguous():
                g = g.contiguous()
            grads.append(_flat(g))
        self.plan.bind(L.SLOT_G, grads)
        flags = L.SGD_EMIT_WEFF | (L.SGD_EMIT_WEFF16 if self.weff16 is not None else 0)
        if g0["nesterov"]:
            flags |= L.SGD_NESTEROV
        if self._first:
            flags |= L.SGD_FIRST_STEP
        self.plan.masked_sgd_step(self.state_ref.mask, float(g0["lr"]), float(g0["momentum"]), float(g0["dampening"]),
                                  float(g0["weight_decay"]), flags)
        self._first = False
        if self.inner is not None:
            for gi, go in zip(self.inner.param_groups, self.param_groups[1:]):
                for k in ("lr", "momentum", "dampening", "weight_decay", "nesterov"):
                    gi[k] = go[k]
            self.inner.step()
        return loss

    def refresh_after_pruning(self):
        """Call after another pruning round changed the mask: re-emits the masked weights."""
        self.plan.apply_mask(self.state_ref.mask, L.EMIT_WEFF | (L.SGD_EMIT_WEFF16 if self.weff16 is not None else 0))
