"""MaskedSGD — the masked training step of a pruned model as ONE kernel per step (K4, csrc/sgd.cu).

In the reference a pruned Conv2d/Linear costs, per step and per parameter, a forward pre-hook
(`weight = weight_mask * weight_orig`, torch/nn/utils/prune.py:53-74), its MulBackward, and the foreach passes of
`torch.optim.SGD` (torch/optim/sgd.py:343-380, built at train.py:372-392 over the groups of utils.set_weight_decay,
utils.py:405-463).  Here the prunable weights share one ParamPlan and one launch does all of it:

    g       = mask ? c * grad(weight) : 0             (MulBackward; c = unscale x clip x 1/world, read on the device)
    g      += wd * weight_orig;  buf = mu * buf + (1 - damp) * g;  g = nesterov ? g + mu * buf : buf
    weight_orig -= lr * g                             (SGD, same update order as torch)
    weight  = mask ? weight_orig : 0                  (next forward's masked weight, fp32 and/or bf16)

`module.weight` becomes a persistent LEAF tensor maintained by the kernel (the pre-hook returns it instead of
re-multiplying), so pruned weights and their gradients never re-densify.  As in the reference, pruned entries of
`weight_orig` keep decaying and keep a momentum buffer.

What the rest of the reference's training step (train.py:35-89) sees:
  * `param_groups[0]` holds the leaves (they carry the gradients), the other groups — exactly the ones
    `utils.set_weight_decay` would build: norm / bias / custom keys / other — are stepped by an inner fused
    `torch.optim.SGD`; LR schedulers drive `param_groups` as usual; momentum lives in `self.state`, so
    `state_dict()` / `load_state_dict()` checkpoint and resume like `torch.optim.SGD` (train.py:507);
  * `torch.amp.GradScaler`: `unscale_` and the inf check walk `param_groups`, so they see the leaves; `scaler.step`
    hands `grad_scale` / `found_inf` over (`_step_supports_amp_scaling`) and the kernel applies them on the device;
  * gradient clipping: `optimizer.clip_grad_norm_(max_norm)` — the norm of the MASKED gradients (what
    `weight_orig.grad` holds in the reference) plus the other parameters', one reduction pass, the coefficient folded
    into the step.  `nn.utils.clip_grad_norm_(model.parameters())` cannot see the leaves: do not use it here;
  * DDP: the leaves are not module parameters, DDP does not reduce them; `step()` all-reduces their gradients itself
    (one coalesced NCCL call) when a process group is initialised and folds the 1/world into the same factor.
"""
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.utils.prune as prune

from . import _lib as L
from ._lib import B200PruneError
from .pruning import B200MaskMethod, _get_state, _flat, prunable_modules

_NORM_CLASSES = (nn.modules.batchnorm._BatchNorm, nn.LayerNorm, nn.GroupNorm, nn.modules.instancenorm._InstanceNorm,
                 nn.LocalResponseNorm)


def set_weight_decay(model, weight_decay, norm_weight_decay=None, norm_classes=None, custom_keys_weight_decay=None):
    """Parameter groups with the semantics of the reference's utils.set_weight_decay (utils.py:405-463): "other",
    "norm" (only when norm_weight_decay is given) and one group per custom key, a key matching the bare parameter name
    ("bias") or, when it contains a dot, the dotted path.  Returns [{"params": [...], "weight_decay": wd}, ...]."""
    norm_classes = tuple(norm_classes) if norm_classes else _NORM_CLASSES
    groups = {"other": [], "norm": []}
    decay = {"other": weight_decay, "norm": norm_weight_decay}
    keys = []
    for key, wd in (custom_keys_weight_decay or []):
        groups[key] = []
        decay[key] = wd
        keys.append(key)

    def visit(module, prefix):
        for name, p in module.named_parameters(recurse=False):
            if not p.requires_grad:
                continue
            for key in keys:
                target = f"{prefix}.{name}" if prefix and "." in key else name
                if key == target:
                    groups[key].append(p)
                    break
            else:
                (groups["norm"] if norm_weight_decay is not None and isinstance(module, norm_classes) else groups["other"]).append(p)
        for child_name, child in module.named_children():
            visit(child, f"{prefix}.{child_name}" if prefix else child_name)

    visit(model, "")
    return [{"params": ps, "weight_decay": decay[key]} for key, ps in groups.items() if ps]


class MaskedSGD(torch.optim.Optimizer):
    _step_supports_amp_scaling = True      # GradScaler.step passes grad_scale / found_inf instead of syncing on found_inf

    def __init__(self, model, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False,
                 norm_weight_decay=None, custom_keys_weight_decay=None, bf16_weights=False, param_groups=None,
                 process_group=None, sync_grads=None):
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")       # torch/optim/sgd.py
        modules = [m for _, m in prunable_modules(model) if "weight_orig" in m._parameters]
        if not modules:
            raise B200PruneError("MaskedSGD: the model has no pruned Conv2d/Linear (prune it first)")
        st = _get_state(model, modules)
        if st.mask is None:
            raise B200PruneError("MaskedSGD: no packed mask (prune the model with this package or load a pruned checkpoint)")
        self.state_ref, self.plan, self.modules = st, st.plan, modules
        masters = [m._parameters["weight_orig"] for m in modules]
        master_ids = {id(p): i for i, p in enumerate(masters)}
        # groups as utils.set_weight_decay builds them (or the caller's own), minus the fused master weights
        groups = param_groups if param_groups is not None else set_weight_decay(
            model, weight_decay, norm_weight_decay=norm_weight_decay, custom_keys_weight_decay=custom_keys_weight_decay)
        rest, fused_wd = [], set()
        for g in groups:
            g = dict(g)
            keep = []
            for p in g["params"]:
                if id(p) in master_ids:
                    fused_wd.add(g.get("weight_decay", weight_decay))
                else:
                    keep.append(p)
            if keep:
                g["params"] = keep
                rest.append(g)
        if len(fused_wd) > 1:
            raise B200PruneError("MaskedSGD: the pruned weights must share one weight decay (they sit in different groups)")
        dev = st.device
        # persistent buffers: momentum, effective weights (the leaves autograd differentiates)
        self.bufs = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in masters]
        self.weff = [torch.zeros_like(p, memory_format=torch.contiguous_format).requires_grad_(True) for p in masters]
        self.weff16 = [torch.zeros_like(p, dtype=torch.bfloat16, memory_format=torch.contiguous_format) for p in masters] if bf16_weights else None
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
        fused_group = {"params": self.weff}
        if fused_wd:
            fused_group["weight_decay"] = fused_wd.pop()
        super().__init__([fused_group] + rest, defaults)
        inner_groups = [{k: v for k, v in g.items()} for g in self.param_groups[1:]]
        self.inner = None
        if inner_groups:
            try:
                self.inner = torch.optim.SGD(inner_groups, lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                             nesterov=nesterov, fused=True)
            except (RuntimeError, TypeError, ValueError):
                self.inner = torch.optim.SGD(inner_groups, lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                             nesterov=nesterov)
        self._inner_fused = self.inner is not None and bool(self.inner.defaults.get("fused"))
        self._grad_zero = None
        self._ctl = torch.tensor([1.0, 0.0], dtype=torch.float32, device=dev)
        self._stats = torch.zeros(2, dtype=torch.float64, device=dev)
        self._clip_coef = None
        self.process_group = process_group
        self.sync_grads = sync_grads                            # None: all-reduce the leaves' gradients iff torch.distributed is initialised
        self.ema = None
        self.plan.bind(L.SLOT_W, [_flat(p.detach()) for p in masters])
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
        object.__setattr__(model, "_b200p_fused_optimizer", self)

    # ---- bookkeeping ---------------------------------------------------------------------------------------------------
    @property
    def _first(self):
        """torch's SGD creates the momentum buffer at the first step; ours exists from the start, `state` says which it is."""
        return "momentum_buffer" not in self.state.get(self.weff[0], {})

    def zero_grad(self, set_to_none=True):
        for w in self.weff:
            w.grad = None
        if self.inner is not None:
            self.inner.zero_grad(set_to_none=set_to_none)
        self._clip_coef = None

    def _bind_grads(self):
        grads, missing = [], 0
        for i, w in enumerate(self.weff):
            g = w.grad
            if g is None:
                # torch skips parameters without a gradient; the fused launch covers all tensors, so a missing gradient is
                # a zero one here: weight decay and momentum still act on that tensor (documented deviation)
                if self._grad_zero is None:
                    self._grad_zero = [None] * len(self.weff)
                if self._grad_zero[i] is None:
                    self._grad_zero[i] = torch.zeros_like(w, memory_format=torch.contiguous_format)
                g = self._grad_zero[i]
                missing += 1
            elif not g.is_contiguous():
                g = g.contiguous()
            grads.append(_flat(g))
        if missing == len(self.weff):
            raise B200PruneError("MaskedSGD.step: no pruned weight has a gradient (did backward run through module.weight?)")
        self.plan.bind(L.SLOT_G, grads)
        return grads

    def _world(self):
        if self.sync_grads is False or not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.process_group)

    def _allreduce_leaf_grads(self, grads):
        """What DDP does for module parameters (train.py:604-607): sum over ranks; the 1/world is folded into the step."""
        handles = [dist.all_reduce(g, group=self.process_group, async_op=True) for g in grads]
        for h in handles:
            h.wait()

    # ---- clipping (train.py:57-66) -----------------------------------------------------------------------------------
    @torch.no_grad()
    def clip_grad_norm_(self, max_norm):
        """Global 2-norm clipping over the MASKED gradients of the pruned weights and all other parameters' gradients:
        what `nn.utils.clip_grad_norm_(model.parameters(), max_norm)` computes on the reference's model.  Call it where the
        reference does (after `scaler.unscale_(optimizer)` when a GradScaler is used).  The coefficient is applied to the
        other parameters' gradients in place and to the pruned weights inside the step kernel.  Returns the total norm
        (0-d tensor, no host sync)."""
        grads = self._bind_grads()
        world = self._world()
        if world > 1 and not getattr(self, "_synced", False):
            self._allreduce_leaf_grads(grads)
            self._synced = True
        self.plan.grad_stats(self.state_ref.mask, self._stats)
        sq = self._stats[0] / float(world * world)              # gradients are summed over ranks, the mean is what DDP leaves
        others = [p.grad for g in self.param_groups[1:] for p in g["params"] if p.grad is not None]
        if others:
            norms = torch._foreach_norm(others)
            sq = sq + torch.stack([n.double() for n in norms]).square().sum()
        total = sq.sqrt().float()
        coef = torch.clamp(float(max_norm) / (total + 1e-6), max=1.0)      # torch.nn.utils.clip_grad_norm_
        if others:
            torch._foreach_mul_(others, coef)
        self._clip_coef = coef
        return total

    # ---- the step -------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g0 = self.param_groups[0]
        grads = self._bind_grads()
        world = self._world()
        if world > 1 and not getattr(self, "_synced", False):
            self._allreduce_leaf_grads(grads)
        self._synced = False
        grad_scale = getattr(self, "grad_scale", None)          # set by GradScaler.step when the gradients are still scaled
        found_inf = getattr(self, "found_inf", None)
        ctl = None
        if grad_scale is not None or found_inf is not None or self._clip_coef is not None or world > 1:
            mul = torch.ones((), dtype=torch.float32, device=self._ctl.device)
            if grad_scale is not None:
                mul = mul / grad_scale.to(torch.float32).reshape(())
            if self._clip_coef is not None:
                mul = mul * self._clip_coef
            if world > 1:
                mul = mul / float(world)
            self._ctl[0] = mul
            self._ctl[1] = found_inf.to(torch.float32).reshape(()) if found_inf is not None else 0.0
            ctl = self._ctl
        flags = L.SGD_EMIT_WEFF | (L.SGD_EMIT_WEFF16 if self.weff16 is not None else 0)
        if g0["nesterov"]:
            flags |= L.SGD_NESTEROV
        first = self._first
        if first:
            flags |= L.SGD_FIRST_STEP
        self.plan.masked_sgd_step(self.state_ref.mask, float(g0["lr"]), float(g0["momentum"]), float(g0["dampening"]),
                                  float(g0["weight_decay"]), flags, ctl=ctl)
        if first and g0["momentum"] != 0:
            if found_inf is not None and bool(found_inf.item()):         # a skipped FIRST step leaves no buffer behind (rare; one sync)
                pass
            else:
                for w, b in zip(self.weff, self.bufs):
                    self.state[w]["momentum_buffer"] = b
        self._clip_coef = None
        if self.inner is not None:
            for gi, go in zip(self.inner.param_groups, self.param_groups[1:]):
                for k in ("lr", "momentum", "dampening", "weight_decay", "nesterov"):
                    gi[k] = go[k]
            if self._inner_fused and (grad_scale is not None or found_inf is not None):
                self.inner.grad_scale, self.inner.found_inf = grad_scale, found_inf
                try:
                    self.inner.step()
                finally:
                    del self.inner.grad_scale, self.inner.found_inf
            else:
                if grad_scale is not None:
                    others = [p.grad for g in self.param_groups[1:] for p in g["params"] if p.grad is not None]
                    torch._foreach_div_(others, grad_scale.to(torch.float32))
                if found_inf is None or not bool(found_inf.item()):
                    self.inner.step()
        return loss

    # ---- checkpoint / resume (train.py:504-521 saves optimizer.state_dict()) ------------------------------------------
    def state_dict(self):
        sd = super().state_dict()
        sd["inner"] = self.inner.state_dict() if self.inner is not None else None
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        inner = state_dict.pop("inner", None)
        super().load_state_dict(state_dict)
        # the loaded momentum tensors are new objects: copy them into the buffers the plan is bound to
        for w, b in zip(self.weff, self.bufs):
            st = self.state.get(w, {})
            if "momentum_buffer" in st and st["momentum_buffer"] is not None:
                b.copy_(st["momentum_buffer"])
                st["momentum_buffer"] = b
        if inner is not None and self.inner is not None:
            self.inner.load_state_dict(inner)

    def refresh_after_pruning(self):
        """Call after another pruning round changed the mask: re-emits the masked weights."""
        self.plan.apply_mask(self.state_ref.mask, L.EMIT_WEFF | (L.SGD_EMIT_WEFF16 if self.weff16 is not None else 0))


class MaskedEMA:
    """ExponentialMovingAverage (utils.py:159-170: AveragedModel with `decay * avg + (1 - decay) * param`, use_buffers=True)
    for a model trained with MaskedSGD: the pruned master weights are averaged by one kernel over the whole plan
    (csrc/sgd.cu k_ema_update), every other parameter and buffer by torch foreach ops.  `update_parameters(model)` and
    `n_averaged` behave like the reference's object (train.py:69-73 resets n_averaged during warm-up: the next update is a
    copy); `state_dict()` has AveragedModel's layout (`module.<name>` + `n_averaged`)."""

    def __init__(self, model, decay, optimizer=None):
        opt = optimizer if optimizer is not None else getattr(model, "_b200p_fused_optimizer", None)
        if opt is None:
            raise B200PruneError("MaskedEMA needs the model's MaskedSGD optimizer")
        self.decay, self.opt, self.plan = float(decay), opt, opt.plan
        masters = [m._parameters["weight_orig"] for m in opt.modules]
        self._fused_ids = {id(p) for p in masters}
        self.avg_fused = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in masters]
        self.plan.bind(L.SLOT_EMA, [_flat(a) for a in self.avg_fused])
        self._names, self._src, self._avg = [], [], []
        self._fused_names = []
        for name, t in list(model.named_parameters()) + list(model.named_buffers()):
            if id(t) in self._fused_ids:
                self._fused_names.append(name)
                continue
            self._names.append(name); self._src.append(t); self._avg.append(t.detach().clone())
        self.n_averaged = torch.tensor(0, dtype=torch.long, device=masters[0].device)
        self._n_host = 0

    @torch.no_grad()
    def update_parameters(self, model=None):
        if int(self.n_averaged.item()) == 0 or self._n_host == 0:        # the reference may have reset n_averaged (warm-up)
            self._n_host = 0
        copy = self._n_host == 0
        self.plan.ema_update(self.decay, copy=copy)
        fl = [(a, s) for a, s in zip(self._avg, self._src) if a.is_floating_point()]
        other = [(a, s) for a, s in zip(self._avg, self._src) if not a.is_floating_point()]
        if copy:
            for a, s in zip(self._avg, self._src):
                a.copy_(s.detach())
        else:
            if fl:
                avgs, srcs = [a for a, _ in fl], [s.detach() for _, s in fl]
                torch._foreach_mul_(avgs, self.decay)
                torch._foreach_add_(avgs, srcs, alpha=1.0 - self.decay)
            for a, s in other:                                            # integer buffers (num_batches_tracked): the reference's lambda too
                a.copy_((self.decay * a + (1 - self.decay) * s.detach()).to(a.dtype))
        self.n_averaged += 1
        self._n_host += 1

    def state_dict(self):
        sd = {"n_averaged": self.n_averaged.clone()}
        for name, a in zip(self._fused_names, self.avg_fused):
            sd["module." + name] = a
        for name, a in zip(self._names, self._avg):
            sd["module." + name] = a
        return sd
