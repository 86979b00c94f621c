"""LOST driver loop around the batched kernels: what the per-image loop of
main_lost_original.py:166-360 does with `lost()`, for whole batches of images.

    keys = keys_from_qkv(qkv)                      # in-place view of the k slice, no copy (:251-263)
    preds = discover(names, keys_list, dims_list, sizes_list, patch_size=16)
    stats = corloc(preds, gt_boxes)                # CorLoc: any IoU >= 0.5 (:335-338)
    save_predictions(preds, folder, stats)         # preds.pkl + results.txt in the reference's format (:346-360)

The feature extractor (DINO ViT, networks.py) and the dataset wrappers are outside the hot path; any
producer of last-layer qkv tensors works."""
import os
import pickle

import numpy as np
import torch

from .object_discovery import lost_batched


def keys_from_qkv(qkv, drop_cls=True):
    """qkv: [B, T, 3*D] output of the last block's qkv Linear.  Returns the patch keys [B, T-1, D] as a
    strided VIEW (k = qkv[..., D:2D]; the reference's reshape/permute/transpose/reshape of
    main_lost_original.py:251-263 lands on exactly these elements); the kernels read it in place."""
    if qkv.dim() != 3 or qkv.shape[-1] % 3:
        raise ValueError("qkv must be [B, T, 3*D]")
    d = qkv.shape[-1] // 3
    k = qkv[:, :, d:2 * d]
    return k[:, 1:, :] if drop_cls else k


def discover(names, keys, dims, init_image_sizes, patch_size=16, k_patches=100, max_batch=512):
    """Runs LOST over many images.  keys: list of [N_i, d] CUDA tensors (or views), dims: list of
    [h_feat, w_feat], init_image_sizes: list of (3, H, W) before padding.  Images are processed in
    batches of `max_batch` (varlen).  Returns {name: np.int64[4] box or None if the seed fell into the
    background component (the reference raises there, object_discovery.py:110-111)}."""
    preds = {}
    scales = [patch_size, patch_size]
    for i in range(0, len(names), max_batch):
        sl = slice(i, i + max_batch)
        # strided views (the k slice of a qkv output) are handed over as they are: the kernels read them in place
        feats = [k if k.dim() == 2 else (k[0] if k.dim() == 3 and k.shape[0] == 1 else k.reshape(-1, k.shape[-1])) for k in keys[sl]]
        feats = [f if f.stride(1) == 1 else f.contiguous() for f in feats]
        out = lost_batched(feats, dims[sl], scales, init_image_sizes[sl], k_patches=k_patches)
        box, status = out["box"].cpu().numpy(), out["status"].cpu().numpy()       # one sync per batch
        for name, b, st in zip(names[sl], box, status):
            preds[name] = None if st else np.asarray([int(round(v)) for v in b], dtype=np.int64)
    return preds


def bbox_iou(box, boxes, eps=1e-7):
    """Plain IoU of one [x1,y1,x2,y2] box against an [n,4] array (the default path of datasets.py:312-364)."""
    box = torch.as_tensor(box, dtype=torch.float64)
    boxes = torch.as_tensor(boxes, dtype=torch.float64).reshape(-1, 4)
    iw = (torch.minimum(box[2], boxes[:, 2]) - torch.maximum(box[0], boxes[:, 0])).clamp(0)
    ih = (torch.minimum(box[3], boxes[:, 3]) - torch.maximum(box[1], boxes[:, 1])).clamp(0)
    inter = iw * ih
    w1, h1 = box[2] - box[0], box[3] - box[1] + eps
    w2, h2 = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1] + eps
    return inter / (w1 * h1 + w2 * h2 - inter + eps)


def corloc(preds, gt_boxes, iou_threshold=0.5):
    """CorLoc over the images that have ground truth: a hit if any GT box has IoU >= 0.5 with the
    prediction (main_lost_original.py:335-338).  Returns (percentage, hits, count)."""
    hits = cnt = 0
    for name, gt in gt_boxes.items():
        if gt is None or len(gt) == 0:
            continue
        cnt += 1
        pred = preds.get(name)
        if pred is not None and bool((bbox_iou(pred, gt) >= iou_threshold).any()):
            hits += 1
    return (100.0 * hits / cnt if cnt else 0.0), hits, cnt


def save_predictions(preds, folder, stats=None):
    """preds.pkl (dict name -> box) and results.txt ('corloc,%.1f,,') as main_lost_original.py:346-360."""
    os.makedirs(folder, exist_ok=True)
    with open(os.path.join(folder, "preds.pkl"), "wb") as f:
        pickle.dump(preds, f)
    if stats is not None:
        with open(os.path.join(folder, "results.txt"), "w") as f:
            f.write("corloc,%.1f,,\n" % stats[0])
