"""Host-side mirror of the reference's pruning functions (same names, arguments, prints and
error behaviour) on top of the sm_100a kernels:

    snip_pruning(model, data_loader, device, criterion, target_sparsity=0.9)   train.py:241-319
    magnitude_pruning(model, prune_amount=0.2)                                 train.py:322-344
    compute_sparsity_global(model)                                             train.py:347-369

Post-conditions kept from `torch.nn.utils.prune`: every pruned module has the
Parameter `weight_orig` (the same object as the old `weight`), the fp32 buffer `weight_mask`,
the attribute `weight`, and a `BasePruningMethod` forward pre-hook, so `prune.is_pruned`,
`prune.remove` and the `*.weight_orig` / `*.weight_mask` checkpoint format keep working.  The
kernel-side truth is the bit-packed mask held in `model._b200p_state`.

No CPU fallback: a model that is not on a CUDA device raises B200PruneError.
"""
import torch
import torch.nn as nn
import torch.nn.utils.prune as prune

from . import _lib as L
from ._lib import B200PruneError
from .plan import ParamPlan


def prunable_modules(model):
    """Every nn.Linear / nn.Conv2d in named_modules() order (train.py:263-265, 334-336)."""
    return [(name, m) for name, m in model.named_modules() if isinstance(m, (nn.Linear, nn.Conv2d))]


class B200MaskMethod(prune.BasePruningMethod):
    """Forward pre-hook object standing in for prune.CustomFromMask / PruningContainer.  The mask
    itself is produced globally by the CUDA kernels; this object only re-materialises
    `weight = weight_mask * weight_orig` (torch/nn/utils/prune.py:53-74) so that autograd sees the
    same graph as with the reference.  In fused mode (MaskedSGD) `weight` is a persistent leaf
    maintained by the optimizer kernel and the hook does nothing."""

    PRUNING_TYPE = "unstructured"

    def __init__(self):
        self.fused_weight = None      # set by MaskedSGD

    def compute_mask(self, t, default_mask):
        return default_mask

    def apply_mask(self, module):
        if self.fused_weight is not None:
            return self.fused_weight
        return super().apply_mask(module)


class PruneState:
    """Plan + packed mask of one model's prunable set."""

    def __init__(self, modules, device):
        self.modules = modules
        self.device = device
        self.plan = ParamPlan([m.weight.numel() for m in modules], device)
        self.mask = None               # packed int32 words once something has been pruned
        self.n_alive = self.plan.total
        self._w_ptrs = None
        self._mf_ptrs = None
        self.maskf = None

    def __deepcopy__(self, memo):
        return None                    # copies of the model (EMA, deepcopy) rebuild their own state lazily

    def __reduce__(self):
        return (type(None), ())        # never pickled with the model; checkpoints carry weight_orig / weight_mask

    def weights(self):
        """fp32 master weights: weight_orig if the module is reparametrised, else weight."""
        out = []
        for m in self.modules:
            p = m._parameters.get("weight_orig", None)
            if p is None:
                p = m._parameters["weight"]
            out.append(p)
        return out

    def bind_weights(self):
        ws = [_flat(p.detach()) for p in self.weights()]
        ptrs = [w.data_ptr() for w in ws]
        if ptrs != self._w_ptrs:
            self.plan.bind(L.SLOT_W, ws)
            self._w_ptrs = ptrs

    def ensure_mask_buffers(self):
        """fp32 `weight_mask` tensors (reference checkpoint format), created on first use."""
        if self.maskf is None:
            self.maskf = []
            for m in self.modules:
                buf = m._buffers.get("weight_mask", None)
                if buf is None or not buf.is_cuda or buf.dtype != torch.float32 or not buf.is_contiguous():
                    buf = torch.ones_like(_param_of(m), dtype=torch.float32, memory_format=torch.contiguous_format)
                self.maskf.append(buf)
        ptrs = [b.data_ptr() for b in self.maskf]
        if ptrs != self._mf_ptrs:
            self.plan.bind(L.SLOT_MASKF, [_flat(b) for b in self.maskf])
            self._mf_ptrs = ptrs

    def install(self):
        """Reparametrise every module once (prune.py:76-206 without the per-tensor arithmetic)."""
        for m, buf in zip(self.modules, self.maskf):
            if "weight_orig" not in m._parameters:
                orig = m._parameters.pop("weight")
                m.register_parameter("weight_orig", orig)
                m.register_buffer("weight_mask", buf)
                method = B200MaskMethod()
                method._tensor_name = "weight"
                m.register_forward_pre_hook(method)
                setattr(m, "weight", method.apply_mask(m))
            else:
                if m._buffers.get("weight_mask", None) is not buf:
                    m._buffers["weight_mask"] = buf
                for hook in m._forward_pre_hooks.values():
                    if isinstance(hook, prune.BasePruningMethod) and hook._tensor_name == "weight":
                        setattr(m, "weight", hook.apply_mask(m))
                        break


def _param_of(m):
    p = m._parameters.get("weight_orig", None)
    return p if p is not None else m._parameters["weight"]


def _flat(t):
    if not t.is_contiguous():
        raise B200PruneError("prunable weights / gradients must be contiguous (default memory format)")
    return t.view(-1)


def _get_state(model, modules):
    if not modules:
        return None
    dev = _param_of(modules[0]).device
    if dev.type != "cuda":
        raise B200PruneError("the B200 pruning path needs the model on a CUDA device (no CPU fallback)")
    st = getattr(model, "_b200p_state", None)
    if st is not None and st.mask is not None and any("weight_orig" not in m._parameters or "weight_mask" not in m._buffers for m in st.modules):
        st = None          # prune.remove() (main_lost.py:67, evaluate_models.py:401) took the reparametrisation away: start from all N weights
    if st is not None and st.maskf is not None and any(
            "weight_mask" in m._buffers and m._buffers["weight_mask"] is not buf for m, buf in zip(st.modules, st.maskf)):
        st = None          # torch.nn.utils.prune (or a checkpoint load) replaced the mask buffers underneath: re-adopt
    if st is None or len(st.modules) != len(modules) or any(a is not b for a, b in zip(st.modules, modules)) \
            or st.device != dev:
        st = PruneState(modules, dev)
        # adopt masks installed by torch.nn.utils.prune (e.g. a loaded reference checkpoint)
        if any("weight_mask" in m._buffers for m in modules):
            st.ensure_mask_buffers()          # modules without a mask get an (unregistered) all-ones one
            st.mask = st.plan.new_mask()
            st.plan.mask_pack_from_f32(st.mask)
            _, st.n_alive = st.plan.count_zeros(st.mask, use_weights=False)
        object.__setattr__(model, "_b200p_state", st)
    return st


def magnitude_pruning(model, prune_amount=0.2):
    """Global magnitude pruning (train.py:322-344 -> prune.global_unstructured + L1Unstructured):
    removes the `prune_amount` fraction (or absolute number, if int) of the SURVIVING Conv2d/Linear
    weights with the smallest |w|; ties at the cut are resolved lowest-flat-index-first."""
    modules = [m for _, m in prunable_modules(model)]
    prune._validate_pruning_amount_init(prune_amount)            # prune.py:1256-1290 error behaviour
    if not modules:
        return model
    st = _get_state(model, modules)
    plan = st.plan
    prune._validate_pruning_amount(prune_amount, st.n_alive)     # prune.py:1293-1311
    k = prune._compute_nparams_toprune(prune_amount, st.n_alive) # prune.py:1331-1354 (banker's rounding)
    st.bind_weights()
    st.ensure_mask_buffers()
    new_mask = plan.new_mask()
    if k == 0:                                                   # prune.py:533: mask unchanged
        plan.select_begin(0, L.MODE_EXACT_K)
        plan.emit_masks(L.KEY_ABS_W, L.MODE_EXACT_K, new_mask, st.mask, force=1, outputs=L.EMIT_MASKF)
    else:
        # select + emit-by-patch on the packed mask, then the fp32 `weight_mask` buffers (checkpoint format) are
        # expanded from it: the keys are read once, not twice
        plan.mask_build(L.KEY_ABS_W, k, L.MODE_EXACT_K, new_mask, st.mask)
        plan.mask_unpack_to_f32(new_mask)
    st.mask = new_mask
    st.n_alive -= k
    st.install()
    return model


def snip_pruning(model, data_loader, device, criterion, target_sparsity=0.9, num_batches=1):
    """Single-shot Network Pruning based on Connection Sensitivity (train.py:241-319).

    `num_batches` > 1 is the multi-batch extension of SURVEY §8(c): score = sum_b |w * g_b| in fp32,
    batches added in loader order; 1 reproduces the reference bit for bit."""
    print(f"Applying SNIP pruning with target sparsity {target_sparsity}...")
    modules = [m for _, m in prunable_modules(model) if hasattr(m, "weight")]
    data_iter = iter(data_loader)
    st = None
    stashed = []                         # gradient sets of all mini-batches stay resident (B x N x 4 bytes)
    for b in range(num_batches):
        images, targets = next(data_iter)
        images = images.to(device)
        targets = targets.to(device)
        model.zero_grad()
        outputs = model(images)
        loss = criterion(outputs, targets)
        loss.backward()
        if st is None:
            # modules whose weight received no gradient are left alone (train.py:288 `if grad_key in grads`)
            modules = [m for m in modules if _grad_of(m) is not None]
            if not modules:
                raise RuntimeError("torch.cat(): expected a non-empty list of Tensors")   # train.py:294
            st = _get_state(model, modules)
        grads = [_flat(_grad_of(m)) for m in modules]
        if num_batches > 1:
            for m in modules:            # take the tensors over: the next backward allocates fresh ones, no copy
                _grad_param(m).grad = None
        stashed.append(grads)
    scores = [torch.empty(m.weight.numel(), dtype=torch.float32, device=st.device) for m in modules]
    st.plan.bind(L.SLOT_SCORE, scores)
    # |w| uses the weight the forward saw: the masked one if the module is already pruned
    st.plan.bind(L.SLOT_W, [_flat(m.weight.detach()) for m in modules])
    st._w_ptrs = None
    plan = st.plan
    n = plan.total
    k = int(n * target_sparsity)                                 # train.py:299
    tables = [plan.pointer_table(L.SLOT_G, g) for g in stashed]
    st.ensure_mask_buffers()
    new_mask = plan.new_mask()
    if 0 < k < n:
        # one fused pass: score = sum_b |w * g_b|, added in batch order (bit-identical to per-batch accumulation),
        # classified against the sampled bracket while it is written; dead entries score 0 and count, as in the reference
        if st.mask is None:
            # fresh model (the usual SNIP case): the sweep writes the provisional packed mask, the finish kernel patches it,
            # and the fp32 `weight_mask` buffers (checkpoint format) are expanded from the 0.125 B/param packed mask —
            # the scores are never re-read (a full emit with fp32 outputs would: 8 B/param more)
            plan.snip_mask_build(tables, k, new_mask)
            plan.mask_unpack_to_f32(new_mask)
        else:
            plan.snip_score_select(tables, k)
            plan.emit_masks(L.KEY_SCORE, L.MODE_SNIP_STRICT, new_mask, st.mask, outputs=L.EMIT_MASKF)
        threshold = None
    else:
        plan.score_accumulate_multi(tables, accumulate=False)
        threshold = float("inf") if k >= n else -1               # train.py:300-303
        plan.select_begin(0, L.MODE_SNIP_STRICT)
        plan.emit_masks(L.KEY_SCORE, L.MODE_SNIP_STRICT, new_mask, st.mask, force=3, forced_threshold=float(threshold),
                        outputs=L.EMIT_MASKF)
    if num_batches > 1:
        for m, g in zip(modules, stashed[-1]):
            _grad_param(m).grad = g.view_as(_grad_param(m))      # leave .grad populated like the reference does
    res = plan.result()                                          # the reference's `.item()` sync (train.py:307)
    if threshold is None:
        threshold = float(res["threshold"])
    print(f"SNIP threshold: {threshold}")
    st.mask = new_mask
    st.n_alive = int(res["n_kept"])
    st.install()
    return model


def _grad_param(m):
    """The tensor that carries this module's weight gradient: the fused leaf under MaskedSGD (autograd differentiates
    `module.weight` there, `weight_orig.grad` stays None), else weight_orig / weight."""
    for hook in m._forward_pre_hooks.values():
        if isinstance(hook, B200MaskMethod) and hook.fused_weight is not None:
            return hook.fused_weight
    p = m._parameters.get("weight_orig", None)
    if p is None:
        p = m._parameters.get("weight", None)
    return p


def _grad_of(m):
    p = _grad_param(m)
    return None if p is None else p.grad


def compute_sparsity_global(model):
    """Global sparsity in percent (train.py:347-369): zeros of the effective weight of every
    Conv2d/Linear, one host sync in total instead of one per module."""
    modules = [m for m in model.modules() if isinstance(m, (nn.Conv2d, nn.Linear))]
    if not modules:
        return 0.0
    st = _get_state(model, modules)
    st.bind_weights()
    zeros, _ = st.plan.count_zeros(st.mask, use_weights=True)
    total = st.plan.total
    if total == 0:
        return 0.0
    return 100.0 * zeros / total


def export_masks(model):
    """Packed mask words (int32 CUDA tensor) and the plan, for checkpointing / inspection."""
    st = getattr(model, "_b200p_state", None)
    if st is None or st.mask is None:
        raise B200PruneError("model has not been pruned by this package")
    return st.mask, st.plan
