"""Host-side mirror of the reference's LOST functions (object_discovery.py:23-134) over the sm_100a
kernels of csrc/lost.cu, reached through the C-ABI (include/b200prune.h):

    lost(feats, dims, scales, init_image_size, k_patches=100) -> (pred, A, scores, seed)
    patch_scoring(M, threshold=0.)                            -> (sel, cent)
    detect_box(A, seed, dims, initial_im_size=None, scales=None) -> (pred, pred_feats)
    lost_batched(feats_list | feats[B,N,d], dims, scales, init_image_sizes, k_patches=100)

Same argument meaning, return types and error behaviour as the reference
(`ValueError("The seed is in the background component.")`).  Tie policy: the reference's unstable
`argsort` is pinned to "lowest patch index first among equal degrees" (SURVEY §8c).
No CPU fallback: tensors must live on a CUDA device.
"""
import ctypes

import numpy as np
import torch

from . import _lib as L
from ._lib import B200PruneError, LostImage, check


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _require_cuda_f32(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise B200PruneError(f"{what} must be a CUDA tensor (the B200 LOST path has no CPU fallback)")
    if t.dtype != torch.float32:
        raise B200PruneError(f"{what} must be float32")


def _box_to_pred(box, scales):
    """[xmin, ymin, xmax, ymax] with the number types the reference produces (object_discovery.py:120-128):
    Python ints when the scales are ints, floats otherwise."""
    ints = all(isinstance(s, (int, np.integer)) for s in scales)
    return [int(round(v)) for v in box] if ints else [float(v) for v in box]


_META_DTYPE = np.dtype([("feat_offset", "<i8"), ("a_offset", "<i8"), ("out_offset", "<i8"), ("dim0", "<i4"), ("dim1", "<i4"),
                        ("img_h", "<i4"), ("img_w", "<i4"), ("scale0", "<f4"), ("scale1", "<f4")])
assert _META_DTYPE.itemsize == ctypes.sizeof(LostImage)


def _meta_records(feat_offs, dims, scales, sizes):
    """b200p_lost_image_t records of a batch as one numpy structured array (no per-image Python objects): `dims` and
    `sizes` are one pair / triple for the whole batch or one per image."""
    B = len(feat_offs)
    rec = np.zeros(B, _META_DTYPE)
    dims = np.asarray(dims, np.int64).reshape(-1, 2)
    rec["dim0"], rec["dim1"] = dims[:, 0], dims[:, 1]
    ns = (rec["dim0"].astype(np.int64) * rec["dim1"]) + np.zeros(B, np.int64)
    rec["feat_offset"] = feat_offs
    rec["out_offset"][1:] = np.cumsum(ns)[:-1]
    rec["a_offset"][1:] = np.cumsum(ns * ns)[:-1]
    if sizes is not None:
        sz = np.asarray(sizes, np.int64)
        sz = sz.reshape(-1, sz.shape[-1])
        rec["img_h"], rec["img_w"] = sz[:, -2], sz[:, -1]
    rec["scale0"], rec["scale1"] = float(scales[0]), float(scales[1])
    return rec, ns


class _Split:
    """List-like view of a flat tensor cut into per-image pieces; the pieces are only made when asked for (a batch of
    256 images would otherwise pay for 256 tensor views per call)."""

    def __init__(self, flat, sizes, shapes=None):
        self.flat, self.sizes, self.shapes = flat, sizes, shapes
        self.starts = np.concatenate(([0], np.cumsum(sizes)))

    def __len__(self):
        return len(self.sizes)

    def __getitem__(self, i):
        if i < 0:
            i += len(self.sizes)
        t = self.flat[int(self.starts[i]):int(self.starts[i + 1])]
        return t.view(*self.shapes[i]) if self.shapes is not None else t

    def __iter__(self):
        return (self[i] for i in range(len(self.sizes)))


class _RawBase:
    """Stand-in for the `feats` tensor of a batch whose images live in separate tensors: the lowest key address."""

    def __init__(self, ptr, device, keepalive):
        self.ptr, self.device, self.keepalive = ptr, device, keepalive

    def data_ptr(self):
        return self.ptr


class _Batch:
    """Device buffers of one b200p_lost_batched call."""

    def __init__(self, feats, row_stride, d, metas, dev, want_A, k_patches, gram_impl, ns=None):
        lib = L.require_cuda()
        if gram_impl is None:
            gram_impl = DEFAULT_GRAM_IMPL
        n_images = len(metas)
        if isinstance(metas, np.ndarray):
            arr = ctypes.cast(metas.ctypes.data, ctypes.POINTER(LostImage))
            self._metas = metas
            total_p, total_a = int(ns.sum()), int((ns * ns).sum())
        else:
            arr = (LostImage * n_images)(*metas)
            ns = np.array([m.dim0 * m.dim1 for m in metas], np.int64)
            total_p, total_a = int(ns.sum()), int((ns * ns).sum())
        ws_bytes = ctypes.c_int64()
        check(lib.b200p_lost_workspace_bytes(n_images, total_p, 0 if want_A else total_a, int(d), gram_impl, ctypes.byref(ws_bytes)),
              "lost_workspace_bytes")
        self.workspace = torch.empty(int(ws_bytes.value), dtype=torch.uint8, device=dev)
        self.A = torch.empty(total_a, dtype=torch.float32, device=dev) if want_A else None
        self.degree = torch.empty(total_p, dtype=torch.int32, device=dev)
        self.seed = torch.empty(n_images, dtype=torch.int32, device=dev)
        self.box = torch.empty(n_images, 4, dtype=torch.float32, device=dev)
        self.status = torch.empty(n_images, dtype=torch.int32, device=dev)
        self.ns = ns
        check(lib.b200p_lost_batched(dev.index, feats.data_ptr(), int(row_stride), int(d), arr, n_images, int(k_patches),
                                     self.A.data_ptr() if want_A else None, self.degree.data_ptr(), self.seed.data_ptr(),
                                     self.box.data_ptr(), self.status.data_ptr(), self.workspace.data_ptr(),
                                     int(ws_bytes.value), gram_impl, _stream(dev)), "lost_batched")


def _meta(feat_off, a_off, out_off, dims, scales, init_image_size):
    m = LostImage()
    m.feat_offset, m.a_offset, m.out_offset = int(feat_off), int(a_off), int(out_off)
    m.dim0, m.dim1 = int(dims[0]), int(dims[1])
    if init_image_size is not None:
        m.img_h, m.img_w = int(init_image_size[-2]), int(init_image_size[-1])
    else:
        m.img_h = m.img_w = 0
    m.scale0, m.scale1 = float(scales[0]), float(scales[1])
    return m


# TMA + tcgen05 3xTF32 on CTA pairs, features read in place (falls back to pre-split operands when the
# layout is not TMA-addressable); L.LOST_GRAM_TC2 / L.LOST_GRAM_TC = pre-split pairs / single CTAs,
# L.LOST_GRAM_FFMA = fp32 CUDA cores
DEFAULT_GRAM_IMPL = L.LOST_GRAM_TC2D


def lost(feats, dims, scales, init_image_size, k_patches=100, gram_impl=None):
    """object_discovery.py:23-69.  feats: [1, N, d] fp32 CUDA tensor (any row stride, e.g. the k slice of
    a qkv buffer), dims = [w_featmap, h_featmap] with N = dims[0]*dims[1].
    Returns (pred ndarray[4], A [N,N] tensor, scores [N] tensor = -degree, seed 0-d int64 tensor)."""
    _require_cuda_f32(feats, "feats")
    if feats.dim() != 3 or feats.shape[0] != 1:
        raise B200PruneError("lost(): feats must have shape [1, N, d] (the reference squeezes a batch of one)")
    f = feats[0]
    n, d = f.shape
    if n != int(dims[0]) * int(dims[1]):
        raise RuntimeError(f"shape '[{dims[0]}, {dims[1]}]' is invalid for input of size {n}")   # the reshape at :101
    if f.stride(1) != 1:
        f = f.contiguous()
    dev = f.device
    b = _Batch(f, f.stride(0), d, [_meta(0, 0, 0, dims, scales, init_image_size)], dev, True, k_patches, gram_impl)
    status, seed, box = int(b.status[0].item()), b.seed[0].to(torch.int64), b.box[0].tolist()   # the reference's sync (:104)
    if status != 0:
        raise ValueError("The seed is in the background component.")
    pred = _box_to_pred(box, scales)
    A = b.A.view(n, n)
    scores = -(b.degree.to(torch.float32))
    return np.asarray(pred), A, scores, seed


def lost_batched(feats, dims, scales, init_image_sizes, k_patches=100, return_A=False, gram_impl=None):
    """LOST over a batch in one call.  feats: [B, N, d] tensor (shared dims) or a list of [N_b, d]
    tensors with per-image dims / init_image_sizes.  Returns a dict of CUDA tensors:
    box [B,4] (xmin,ymin,xmax,ymax), seed [B], status [B] (1 = seed in background), degree (list of
    views), and A (list of [N_b,N_b] views) if return_A.  No host sync."""
    if isinstance(feats, torch.Tensor):
        _require_cuda_f32(feats, "feats")
        if feats.dim() != 3:
            raise B200PruneError("lost_batched(): feats must be [B, N, d]")
        if feats.stride(2) != 1 or feats.stride(0) % 4 or feats.stride(1) % 1:
            feats = feats.contiguous()
        B, n, d = feats.shape
        base, row_stride = feats, feats.stride(1)
        feat_offs = np.arange(B, dtype=np.int64) * feats.stride(0)
        dims_l = dims
        sizes_l = init_image_sizes
    else:
        feats = [f if f.dim() == 2 else f.reshape(-1, f.shape[-1]) for f in feats]
        for f in feats:
            _require_cuda_f32(f, "feats")
        d = feats[0].shape[1]
        strides = {f.stride(0) for f in feats}
        if len(strides) == 1 and all(f.stride(1) == 1 and f.shape[1] == d for f in feats):
            # every image is read IN PLACE, wherever its keys live (views of one producer buffer, k slices of separate qkv
            # outputs, ...): the records carry each image's offset from the lowest address; no concatenation, no copy
            row_stride = feats[0].stride(0)
            ptrs = np.array([f.data_ptr() for f in feats], dtype=np.int64)
            lowest = int(ptrs.min())
            base = _RawBase(lowest, feats[0].device, feats)
            feat_offs = (ptrs - lowest) // 4
        else:                                           # mixed row strides: gather once (not the driver's path)
            base = torch.cat([f.reshape(-1, d) for f in feats], dim=0)
            row_stride = d
            feat_offs = np.concatenate(([0], np.cumsum([f.shape[0] for f in feats])[:-1])).astype(np.int64) * d
        B = len(feats)
        dims_l, sizes_l = list(dims), list(init_image_sizes)
    metas, ns = _meta_records(feat_offs, dims_l, scales, sizes_l)
    b = _Batch(base, row_stride, d, metas, base.device, return_A, k_patches, gram_impl, ns)
    out = {"box": b.box, "seed": b.seed, "status": b.status, "degree": _Split(b.degree, ns), "_keepalive": b}
    if return_A:
        out["A"] = _Split(b.A, ns * ns, [(int(n_i), int(n_i)) for n_i in ns])
    return out


def patch_scoring(M, threshold=0.):
    """object_discovery.py:72-90: (sel, cent) with cent = -degree (fp32) and sel = patches by ascending
    degree, lowest index first among equals.  M: [N, N] fp32 CUDA tensor (not modified)."""
    _require_cuda_f32(M, "M")
    if M.dim() != 2 or M.shape[0] != M.shape[1]:
        raise B200PruneError("patch_scoring(): M must be a square matrix")
    if M.stride(1) != 1:
        M = M.contiguous()
    lib = L.require_cuda()
    n = M.shape[0]
    degree = torch.empty(n, dtype=torch.int32, device=M.device)
    sel = torch.empty(n, dtype=torch.int64, device=M.device)
    check(lib.b200p_lost_patch_scoring(M.device.index, M.data_ptr(), n, M.stride(0), float(threshold),
                                       degree.data_ptr(), sel.data_ptr(), _stream(M.device)), "lost_patch_scoring")
    return sel, -(degree.to(torch.float32))


def detect_box(A, seed, dims, initial_im_size=None, scales=None):
    """object_discovery.py:93-134.  A: correlation vector with dims[0]*dims[1] entries (any shape),
    seed: int or 0-d tensor.  Returns (pred [xmin,ymin,xmax,ymax], pred_feats [ymin,xmin,ymax,xmax])."""
    _require_cuda_f32(A, "A")
    lib = L.require_cuda()
    w_featmap, h_featmap = int(dims[0]), int(dims[1])
    M = A.reshape(w_featmap, h_featmap).contiguous()                      # :101 (raises like the reference on a size mismatch)
    dev = M.device
    box = torch.empty(4, dtype=torch.float32, device=dev)
    fbox = torch.empty(4, dtype=torch.int32, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    h, w = (int(initial_im_size[0]), int(initial_im_size[1])) if initial_im_size else (0, 0)
    check(lib.b200p_lost_detect_box(dev.index, M.data_ptr(), w_featmap, h_featmap, int(seed), float(scales[0]), float(scales[1]),
                                    h, w, box.data_ptr(), fbox.data_ptr(), status.data_ptr(), _stream(dev)), "lost_detect_box")
    if int(status.item()) != 0:
        raise ValueError("The seed is in the background component.")       # :110-111
    return _box_to_pred(box.tolist(), scales), [int(v) for v in fbox.tolist()]
