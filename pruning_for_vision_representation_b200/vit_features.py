"""Self-contained DINO-style ViT feature producer for LOST.

The reference takes its patch keys from facebookresearch/dino's ViT (networks.py:23-93, unvendored and
unpinned) through a forward hook on the last block's `attn.qkv` Linear (main_lost_original.py:220-228).
This module is a plain-PyTorch restatement of that architecture (ViT-S/16: D = 384, 6 heads, 12 pre-LN
blocks, MLP ratio 4, class token, learned position embeddings bicubically interpolated to the (H/p, W/p)
grid of the input) that exposes the same tensor: `last_qkv(img) -> [B, T, 3D]`.  It is the producer of
the hot path's input, not part of it: cuDNN/cuBLAS through PyTorch, random init unless weights are loaded
(there is no network here).  `lost_driver.keys_from_qkv` turns its output into the key view LOST reads.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.scale = (dim // heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x, return_qkv=False):
        B, T, D = x.shape
        qkv_flat = self.qkv(x)                                                      # [B, T, 3D] — what the reference hooks
        qkv = qkv_flat.reshape(B, T, 3, self.heads, D // self.heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        out = F.scaled_dot_product_attention(q, k, v)
        out = self.proj(out.transpose(1, 2).reshape(B, T, D))
        return (out, qkv_flat) if return_qkv else out


class _Block(nn.Module):
    def __init__(self, dim, heads, mlp_ratio):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = nn.Sequential(nn.Linear(dim, int(dim * mlp_ratio)), nn.GELU(), nn.Linear(int(dim * mlp_ratio), dim))

    def forward(self, x, return_qkv=False):
        if return_qkv:
            y, qkv = self.attn(self.norm1(x), True)
            x = x + y
            return x + self.mlp(self.norm2(x)), qkv
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class ViTFeatures(nn.Module):
    def __init__(self, patch_size=16, dim=384, depth=12, heads=6, mlp_ratio=4.0, img_size=224):
        super().__init__()
        self.patch_size, self.dim = patch_size, dim
        self.patch_embed = nn.Conv2d(3, dim, patch_size, patch_size)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.grid0 = img_size // patch_size
        self.pos_embed = nn.Parameter(torch.zeros(1, 1 + self.grid0 * self.grid0, dim))
        self.blocks = nn.ModuleList([_Block(dim, heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)

    def interpolate_pos(self, h, w, align_corners=True):
        """Bicubic interpolation of the patch position embeddings to an h x w grid, class token untouched.
        Default = the reference's own `interpolate_embeddings` (vision_transformer.py:781-858: reshape to the square
        grid, `nn.functional.interpolate(..., mode="bicubic", align_corners=True)`, back) — PINNED against it by
        tests/golden/vit_producer.npz.  align_corners=False is facebookresearch/dino's interpolate_pos_encoding, the
        (unvendored, unpinned) producer of main_lost_original.py; weights from that code base want that variant."""
        if h == self.grid0 and w == self.grid0:
            return self.pos_embed
        cls, patch = self.pos_embed[:, :1], self.pos_embed[:, 1:]
        patch = patch.permute(0, 2, 1).reshape(1, self.dim, self.grid0, self.grid0)
        patch = F.interpolate(patch, size=(h, w), mode="bicubic", align_corners=align_corners)
        return torch.cat([cls, patch.reshape(1, self.dim, h * w).permute(0, 2, 1)], dim=1)

    def pad_to_patch(self, img):
        """Zero-pad H and W up to a multiple of the patch size (main_lost_original.py:188-196)."""
        p = self.patch_size
        H, W = img.shape[-2:]
        return F.pad(img, (0, (-W) % p, 0, (-H) % p))

    @torch.no_grad()
    def last_qkv(self, img):
        """img: [B, 3, H, W].  Returns (qkv [B, 1 + h*w, 3D] of the LAST block, (h, w))."""
        img = self.pad_to_patch(img)
        x = self.patch_embed(img)
        h, w = x.shape[-2:]
        x = x.flatten(2).transpose(1, 2)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1) + self.interpolate_pos(h, w)
        for blk in self.blocks[:-1]:
            x = blk(x)
        _, qkv = self.blocks[-1](x, return_qkv=True)
        return qkv, (h, w)


def vit_small_16(**kw):
    return ViTFeatures(patch_size=16, dim=384, depth=12, heads=6, **kw)
