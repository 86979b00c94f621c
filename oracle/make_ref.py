"""oracle/make_ref.py — stage the UNMODIFIED reference for the reference arm (TEST / BASELINE INFRASTRUCTURE).

    python oracle/make_ref.py            # copies the files below from /root/reference into oracle/_ref/

The reference is pure Python; /root/reference does not exist on the GPU box, oracle/_ref/ (git-ignored, shipped with the
snapshot like the built .so) does.  Nothing is edited: the files are byte-for-byte copies, listed with their sha256 in
oracle/_ref/MANIFEST.json.  `load()` imports `train` and `object_discovery` from there — with `skimage` (imported by
datasets.py:20, not installed, not on the path) stubbed and wandb disabled — so that `bench.py --impl reference` and the
`cpu_baseline` leg time `train.snip_pruning` / `train.magnitude_pruning` (train.py:241-344) and `object_discovery.lost`
(object_discovery.py:23-69) themselves, not a port.  Only bench.py's reference legs, tests/ and __graft_entry__.build()
touch this module; the product never does.
"""
import hashlib
import importlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("B200P_REFERENCE", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
# train.py and what it imports at module level (train.py:6-22), object_discovery.py and its `from datasets import bbox_iou`,
# vision_transformer.py for the position-embedding interpolation the ViT producer is pinned against
FILES = ["train.py", "presets.py", "utils.py", "sampler.py", "transforms.py", "object_discovery.py", "datasets.py",
         "vision_transformer.py"]


def make(src=REF_SRC, dst=REF_DST):
    """Copy the reference files (unmodified) and write the manifest.  Returns the manifest dict."""
    if not os.path.isdir(src):
        raise FileNotFoundError(f"{src} is not available (the reference only exists in the build container)")
    os.makedirs(dst, exist_ok=True)
    manifest = {}
    for name in FILES:
        shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
        with open(os.path.join(dst, name), "rb") as f:
            manifest[name] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": manifest}, f, indent=1)
    return manifest


def available(dst=REF_DST):
    return all(os.path.exists(os.path.join(dst, n)) for n in FILES)


def load(dst=REF_DST):
    """(train, object_discovery) modules of the staged reference."""
    if not available(dst):
        raise ImportError(f"{dst} is incomplete: run `python oracle/make_ref.py` in the build container")
    sys.dont_write_bytecode = True
    os.environ.setdefault("WANDB_MODE", "disabled")
    for name in ("skimage", "skimage.io"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if dst not in sys.path:
        sys.path.insert(0, dst)
    return importlib.import_module("train"), importlib.import_module("object_discovery")


if __name__ == "__main__":
    m = make()
    print(f"staged {len(m)} reference files in {REF_DST}")
    t, od = load()
    print("import ok:", t.snip_pruning.__name__, t.magnitude_pruning.__name__, od.lost.__name__)
