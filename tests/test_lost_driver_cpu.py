"""Host-side logic of the LOST driver (CorLoc, IoU, prediction files, k-slice view): CPU only."""
import os
import pickle

import numpy as np
import torch

from pruning_for_vision_representation_b200 import lost_driver as D


def test_keys_from_qkv_matches_reference_reshape():
    B, T, nh, hd = 2, 7, 3, 4
    Dm = nh * hd
    qkv = torch.randn(B, T, 3 * Dm)
    # main_lost_original.py:251-263
    q, k, v = qkv.reshape(B, T, 3, nh, hd).permute(2, 0, 3, 1, 4)
    ref = k.transpose(1, 2).reshape(B, T, -1)[:, 1:, :]
    got = D.keys_from_qkv(qkv)
    assert torch.equal(got, ref) and got.data_ptr() == qkv[:, 1:, Dm:].data_ptr() and not got.is_contiguous()


def test_iou_and_corloc():
    assert float(D.bbox_iou([0, 0, 10, 10], [[0, 0, 10, 10]])[0]) > 0.999
    assert abs(float(D.bbox_iou([0, 0, 10, 10], [[5, 0, 15, 10]])[0]) - 1 / 3) < 1e-6
    assert float(D.bbox_iou([0, 0, 10, 10], [[20, 20, 30, 30]])[0]) == 0.0
    preds = {"a": np.array([0, 0, 10, 10]), "b": np.array([0, 0, 10, 10]), "c": None, "d": np.array([1, 1, 2, 2])}
    gts = {"a": np.array([[1, 1, 10, 10]]), "b": np.array([[8, 8, 30, 30], [100, 100, 110, 110]]), "c": np.array([[0, 0, 5, 5]]), "d": []}
    pct, hits, cnt = D.corloc(preds, gts)
    assert (hits, cnt) == (1, 3) and abs(pct - 100 / 3) < 1e-9          # 'd' has no ground truth and is skipped


def test_save_predictions(tmp_path):
    preds = {"img1": np.array([1, 2, 3, 4], dtype=np.int64)}
    D.save_predictions(preds, str(tmp_path), (61.94, 1, 1))
    assert pickle.load(open(os.path.join(tmp_path, "preds.pkl"), "rb"))["img1"].tolist() == [1, 2, 3, 4]
    assert open(os.path.join(tmp_path, "results.txt")).read() == "corloc,61.9,,\n"
