"""ctypes binding of include/b200prune.h.

The library is loaded from the in-tree `libb200prune.so` (built by `build.py`).  There is no
CPU fallback: if the library is missing or no CUDA device is visible, the compute entry
points raise `B200PruneError` — they never route through PyTorch ops or the oracle.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200prune.so")

# constants mirrored from include/b200prune.h
CHUNK = 4096
WORDS_PER_CHUNK = 128
SLOT_W, SLOT_G, SLOT_SCORE, SLOT_BUF, SLOT_WEFF, SLOT_MASKF, SLOT_WEFF16, SLOT_EMA = range(8)
NUM_SLOTS = 8
MODE_SNIP_STRICT, MODE_EXACT_K = 0, 1
KEY_ABS_W, KEY_SCORE = 0, 1
EMIT_MASKF, EMIT_WEFF = 1, 2
SGD_NESTEROV, SGD_FIRST_STEP, SGD_EMIT_WEFF, SGD_EMIT_WEFF16 = 1, 2, 4, 8
LOST_GRAM_FFMA, LOST_GRAM_TC, LOST_GRAM_TC2, LOST_GRAM_TC2D = 0, 1, 2, 3
OPT_SELECT_IMPL, OPT_TIME_SWEEP, OPT_REUSE_SAMPLE, OPT_COOP_GRID = 1, 2, 3, 4
SHARD_SAMPLE, SHARD_SWEEP, SHARD_FINISH, SHARD_TIES, SHARD_EMIT, SHARD_PUSH, SHARD_ALL = 1, 2, 4, 8, 16, 32, 63
SELECT_SAMPLED, SELECT_EXACT = 0, 1


class B200PruneError(RuntimeError):
    pass


class SelectResult(ctypes.Structure):
    _fields_ = [
        ("k", ctypes.c_uint64), ("n_valid", ctypes.c_uint64), ("n_less", ctypes.c_uint64),
        ("n_equal", ctypes.c_uint64), ("quota", ctypes.c_uint64), ("n_kept", ctypes.c_uint64),
        ("threshold", ctypes.c_float), ("thr_key", ctypes.c_uint32),
        ("passes_full", ctypes.c_uint32), ("collected", ctypes.c_uint32),
        ("miss", ctypes.c_uint32), ("reserved_", ctypes.c_uint32),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class LostImage(ctypes.Structure):
    _fields_ = [
        ("feat_offset", ctypes.c_int64), ("a_offset", ctypes.c_int64), ("out_offset", ctypes.c_int64),
        ("dim0", ctypes.c_int32), ("dim1", ctypes.c_int32),
        ("img_h", ctypes.c_int32), ("img_w", ctypes.c_int32),
        ("scale0", ctypes.c_float), ("scale1", ctypes.c_float),
    ]


_P = ctypes.c_void_p
_I = ctypes.c_int
_I64 = ctypes.c_int64
_U64 = ctypes.c_uint64
_F = ctypes.c_float

# name -> (restype, argtypes); every symbol include/b200prune.h declares
SIGNATURES = {
    "b200p_last_error": (ctypes.c_char_p, []),
    "b200p_version": (_I, []),
    "b200p_device_count": (_I, []),
    "b200p_plan_create": (_I, [_I, _I, ctypes.POINTER(_I64), _I64, ctypes.POINTER(_P)]),
    "b200p_plan_destroy": (_I, [_P]),
    "b200p_plan_total": (_I64, [_P]),
    "b200p_plan_num_chunks": (_I64, [_P]),
    "b200p_plan_mask_words": (_I64, [_P]),
    "b200p_plan_seg_chunk_start": (_I64, [_P, _I]),
    "b200p_plan_bind": (_I, [_P, _I, ctypes.POINTER(_P), _P]),
    "b200p_ptrtable_create": (_I, [_P, _I, ctypes.POINTER(_P), _P, ctypes.POINTER(_P)]),
    "b200p_ptrtable_destroy": (_I, [_P]),
    "b200p_ptrtables_update": (_I, [ctypes.POINTER(_P), ctypes.POINTER(_P), _I, _I, _P]),
    "b200p_plan_bind_table": (_I, [_P, _I, _P]),
    "b200p_plan_set_option": (_I, [_P, _I, _I64]),
    "b200p_plan_kernel_time_ms": (_I, [_P, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_I64)]),
    "b200p_plan_hist_ptr": (_P, [_P]),
    "b200p_plan_state_ptr": (_P, [_P]),
    "b200p_score_accumulate": (_I, [_P, _I, _I64, _I64, _P]),
    "b200p_score_accumulate_multi": (_I, [_P, ctypes.POINTER(_P), _I, _I, _I64, _I64, _P]),
    "b200p_select_kth": (_I, [_P, _I, _P, _U64, _I, _P]),
    "b200p_select_begin": (_I, [_P, _U64, _I, _I, _P]),
    "b200p_select_hist": (_I, [_P, _I, _I, _P, _I64, _I64, _P]),
    "b200p_select_scan": (_I, [_P, _I, _P]),
    "b200p_select_ties": (_I, [_P, _I, _P, _I64, _I64, _U64, _P]),
    "b200p_select_ties_count": (_I, [_P, _I, _P, _I64, _I64, _P, _P]),
    "b200p_select_ties_scan": (_I, [_P, _I64, _I64, _P, _I, _P]),
    "b200p_sum_parts": (_I, [_I, _P, _P, _I, _I64, _I64, _P]),
    "b200p_plan_chunk_flat_start": (_I64, [_P, _I64]),
    "b200p_select_result": (_I, [_P, ctypes.POINTER(SelectResult), _P]),
    "b200p_emit_masks": (_I, [_P, _I, _I, _I, _F, _P, _P, _I, _I64, _I64, _P]),
    "b200p_mask_build": (_I, [_P, _I, _P, _U64, _I, _P, _P]),
    "b200p_snip_mask_build": (_I, [_P, ctypes.POINTER(_P), _I, _U64, _P, _P]),
    "b200p_snip_mask_build_refresh": (_I, [_P, ctypes.POINTER(_P), ctypes.POINTER(_P), _I, _U64, _P, _P]),
    "b200p_snip_score_select": (_I, [_P, ctypes.POINTER(_P), _I, _U64, _P, _P]),
    "b200p_count_zeros": (_I, [_P, _P, _P, _I, _P]),
    "b200p_mask_pack_from_f32": (_I, [_P, _P, _P]),
    "b200p_mask_unpack_to_f32": (_I, [_P, _P, _P]),
    "b200p_apply_mask": (_I, [_P, _P, _I, _P]),
    "b200p_mask_grads": (_I, [_P, _P, _P]),
    "b200p_masked_sgd_step": (_I, [_P, _P, _F, _F, _F, _F, _I, _P]),
    "b200p_masked_sgd_step_ctl": (_I, [_P, _P, _F, _F, _F, _F, _I, _P, _P]),
    "b200p_grad_stats": (_I, [_P, _P, _P, _P]),
    "b200p_ema_update": (_I, [_P, _F, _I, _P]),
    "b200p_lost_workspace_bytes": (_I, [_I, _I64, _I64, _I, _I, ctypes.POINTER(_I64)]),
    "b200p_lost_batched": (_I, [_I, _P, _I64, _I, ctypes.POINTER(LostImage), _I, _I, _P, _P, _P, _P, _P,
                                _P, _I64, _I, _P]),
    "b200p_lost_last_trace": (_I, [ctypes.POINTER(_U64)]),
    "b200p_lost_finish_trace": (_I, [_I, ctypes.POINTER(_U64)]),
    "b200p_comm_trace": (_I, [_P, ctypes.POINTER(_U64)]),
    "b200p_select_last_trace": (_I, [ctypes.POINTER(_U64)]),
    "b200p_lost_patch_scoring": (_I, [_I, _P, _I, _I64, _F, _P, _P, _P]),
    "b200p_lost_detect_box": (_I, [_I, _P, _I, _I, _I, _F, _F, _I, _I, _P, _P, _P, _P]),
    "b200p_comm_create": (_I, [_I, _I, _I, _I64, _I64, ctypes.POINTER(_P)]),
    "b200p_comm_destroy": (_I, [_P]),
    "b200p_comm_window_bytes": (_I64, [_P]),
    "b200p_comm_window": (_P, [_P]),
    "b200p_comm_mask_ptr": (_P, [_P]),
    "b200p_comm_score_ptr": (_P, [_P]),
    "b200p_comm_score_cap": (_I64, [_P]),
    "b200p_comm_ipc_handle": (_I, [_P, _P]),
    "b200p_comm_connect_ipc": (_I, [_P, _P]),
    "b200p_comm_connect_local": (_I, [_P, ctypes.POINTER(_P)]),
    "b200p_comm_error": (_I, [_P, _P]),
    "b200p_comm_barrier": (_I, [_P, _P]),
    "b200p_comm_mask_allgather": (_I, [_P, _P, _I64, _I64, _P]),
    "b200p_comm_score_push_table": (_I, [_P, _P, ctypes.POINTER(_I64), _P, ctypes.POINTER(_P)]),
    "b200p_sharded_mask_build": (_I, [_P, _P, _I, _P, _U64, _I, _I64, _I64, _I, _P]),
    "b200p_snip_mask_build_host": (_I, [_P, _P, ctypes.POINTER(_P), _I, _U64, _P, ctypes.POINTER(SelectResult)]),
    "b200p_magnitude_mask_build_host": (_I, [_P, _P, _P, _U64, _P, ctypes.POINTER(SelectResult)]),
}

_lib = None
ABI_VERSION = 200            # must equal b200p_version() of the loaded library (struct layouts / signatures above)


def load(build_if_missing=True):
    """Load libb200prune.so (building it in-tree if absent and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise B200PruneError(f"{LIB_PATH} is missing; run `python -m pruning_for_vision_representation_b200.build`")
        _build.build()
    # a library older than the binding is caught by the ABI version below (no mtime games: file times do not survive the
    # trip to the GPU box, and several ranks must never rebuild the same file at once)
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # e.g. libcudart not found
        raise B200PruneError(f"cannot load {LIB_PATH}: {e}") from e
    lib.b200p_version.restype = ctypes.c_int
    if lib.b200p_version() != ABI_VERSION:
        raise B200PruneError(f"{LIB_PATH} has ABI version {lib.b200p_version()}, this binding needs {ABI_VERSION}: rebuild it "
                             "(`python -m pruning_for_vision_representation_b200.build --force`)")
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().b200p_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise B200PruneError(f"{what or 'b200prune call'} failed (code {rc}): {last_error()}")


def require_cuda():
    """Fail loudly when the CUDA path cannot run (no silent fallback)."""
    lib = load()
    if lib.b200p_device_count() <= 0:
        raise B200PruneError("no CUDA device visible: the B200 pruning path has no CPU fallback")
    return lib
