"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/b200prune.h declares, the ctypes binding covers all of them, the product fails loudly
without a GPU, and nothing in the product package touches oracle/."""
import ctypes
import os
import re

import pytest
import torch

from pruning_for_vision_representation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200prune.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200p_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/b200prune.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert lib.b200p_version() == _lib.ABI_VERSION


def test_constants_match_header():
    src = open(HEADER).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(B200P_[A-Z0-9_]+)\s+(-?\d+)\b", src)}
    assert defs["B200P_CHUNK"] == _lib.CHUNK and defs["B200P_WORDS_PER_CHUNK"] == _lib.WORDS_PER_CHUNK
    assert (defs["B200P_SLOT_W"], defs["B200P_SLOT_G"], defs["B200P_SLOT_SCORE"], defs["B200P_SLOT_BUF"], defs["B200P_SLOT_WEFF"],
            defs["B200P_SLOT_MASKF"], defs["B200P_SLOT_WEFF16"]) == (_lib.SLOT_W, _lib.SLOT_G, _lib.SLOT_SCORE, _lib.SLOT_BUF,
                                                                    _lib.SLOT_WEFF, _lib.SLOT_MASKF, _lib.SLOT_WEFF16)
    assert (defs["B200P_MODE_SNIP_STRICT"], defs["B200P_MODE_EXACT_K"]) == (_lib.MODE_SNIP_STRICT, _lib.MODE_EXACT_K)
    assert (defs["B200P_SGD_NESTEROV"], defs["B200P_SGD_FIRST_STEP"], defs["B200P_SGD_EMIT_WEFF"], defs["B200P_SGD_EMIT_WEFF16"]) == \
        (_lib.SGD_NESTEROV, _lib.SGD_FIRST_STEP, _lib.SGD_EMIT_WEFF, _lib.SGD_EMIT_WEFF16)
    assert ctypes.sizeof(_lib.SelectResult) == 72 and ctypes.sizeof(_lib.LostImage) == 48
    assert (defs["B200P_SHARD_SAMPLE"], defs["B200P_SHARD_SWEEP"], defs["B200P_SHARD_FINISH"], defs["B200P_SHARD_TIES"], defs["B200P_SHARD_EMIT"],
            defs["B200P_SHARD_PUSH"], defs["B200P_SHARD_ALL"]) == (_lib.SHARD_SAMPLE, _lib.SHARD_SWEEP, _lib.SHARD_FINISH, _lib.SHARD_TIES,
                                                                  _lib.SHARD_EMIT, _lib.SHARD_PUSH, _lib.SHARD_ALL)


@pytest.mark.skipif(torch.cuda.is_available(), reason="no-GPU behaviour")
def test_fails_loudly_without_a_gpu():
    lib = _lib.load()
    assert lib.b200p_device_count() == 0
    with pytest.raises(_lib.B200PruneError, match="no CUDA device"):
        _lib.require_cuda()
    numel = (ctypes.c_int64 * 1)(4096)
    handle = ctypes.c_void_p()
    rc = lib.b200p_plan_create(0, 1, numel, 0, ctypes.byref(handle))
    assert rc == -2 and handle.value is None and _lib.last_error()          # B200P_ECUDA, message set, no plan
    from pruning_for_vision_representation_b200 import pruning
    from tests.tinynet import TinyNet
    with pytest.raises(_lib.B200PruneError):
        pruning.magnitude_pruning(TinyNet(), 0.2)                             # CPU model: no fallback path exists
    from pruning_for_vision_representation_b200 import object_discovery as OD
    with pytest.raises(_lib.B200PruneError):
        OD.lost(torch.zeros(1, 12, 8), [3, 4], [16, 16], (3, 48, 64))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pruning_for_vision_representation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                # nor reaches it any other way: no path into oracle/ (or the staged reference) is ever built or executed
                assert not re.search(r"""["'/]oracle["'/]|oracle\._ref|importlib\.import_module\(\s*["']oracle""", text), f
