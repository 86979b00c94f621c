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


def test_vit_feature_producer_contract():
    from pruning_for_vision_representation_b200.vit_features import ViTFeatures
    torch.manual_seed(0)
    m = ViTFeatures(patch_size=16, dim=48, depth=2, heads=3, img_size=64).eval()
    img = torch.randn(2, 3, 75, 100)                       # not a multiple of 16: padded to 80 x 112
    qkv, (h, w) = m.last_qkv(img)
    assert (h, w) == (5, 7) and qkv.shape == (2, 1 + 35, 3 * 48)
    keys = D.keys_from_qkv(qkv)
    assert keys.shape == (2, 35, 48)
    # the hook contract of main_lost_original.py:222-263: k of the last block, CLS dropped
    captured = {}
    hnd = m.blocks[-1].attn.qkv.register_forward_hook(lambda mod, i, o: captured.setdefault("qkv", o))
    m.last_qkv(img); hnd.remove()
    B, T = 2, 36
    q, k, v = captured["qkv"].reshape(B, T, 3, 3, 16).permute(2, 0, 3, 1, 4)
    assert torch.equal(keys, k.transpose(1, 2).reshape(B, T, -1)[:, 1:, :])
    # native grid: position embeddings are used as they are
    assert m.interpolate_pos(4, 4) is m.pos_embed and m.interpolate_pos(5, 7).shape == (1, 36, 48)


def test_vit_producer_pinned_to_the_reference(golden_dir):
    """SURVEY f-4: the producer's position-embedding interpolation equals the reference's own interpolate_embeddings
    (vision_transformer.py:781-858) and its key view equals the k the reference extracts (main_lost_original.py:251-263):
    fixtures written by tests/golden/make_golden.py from the unmodified reference."""
    from pruning_for_vision_representation_b200.vit_features import ViTFeatures
    z = np.load(os.path.join(golden_dir, "vit_producer.npz"))
    vit = ViTFeatures(patch_size=16, dim=12, depth=1, heads=2, img_size=224)
    with torch.no_grad():
        vit.pos_embed.copy_(torch.from_numpy(z["pos_embed"]))
        for (h, w) in ((13, 17), (30, 30)):
            got = vit.interpolate_pos(h, w)
            torch.testing.assert_close(got, torch.from_numpy(z[f"interp_{h}x{w}"]), rtol=0, atol=0)
        assert vit.interpolate_pos(14, 14) is vit.pos_embed
    k = D.keys_from_qkv(torch.from_numpy(z["qkv"]))
    assert torch.equal(k, torch.from_numpy(z["k_feats"])) and not k.is_contiguous()
