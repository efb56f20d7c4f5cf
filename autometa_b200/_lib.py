"""ctypes binding of libautometa_b200.so (the C ABI in include/autometa_b200.h).

There is no CPU fallback: if the shared library is missing the import fails
loudly, and every device entry point raises when CUDA is unusable.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libautometa_b200.so")

AM_OK, AM_EINVAL, AM_ECUDA, AM_ECAPACITY, AM_EFORMAT = 0, -1, -2, -3, -4
NORM_METHODS = {"am_clr": 0, "clr": 1, "ilr": 2}
GROUP_BASES = 64
TILE_BASES = 2048


class AutometaB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libautometa_b200 error {code}: {msg}")
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
        "`python -c 'import __graft_entry__ as g; g.build()'` (or `python autometa_b200/build.py`). "
        "autometa_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int
_f64 = C.c_double

_SIGS = {
    "am_version": (C.c_int, []),
    "am_last_error": (C.c_char_p, []),
    "am_launch_count": (_i64, []),
    "am_set_reserved_sms": (C.c_int, [_i32]),
    "am_get_reserved_sms": (C.c_int, []),
    "am_kmer_num_columns": (C.c_int, [_i32]),
    "am_kmer_column_lut": (C.c_int, [_i32, _p]),
    "am_kmer_column_names": (C.c_int, [_i32, _p]),
    "am_fasta_index": (C.c_int, [_p, _i64, _p, _p, _p, _p, _i64, _p]),
    "am_fasta_index_mt": (C.c_int, [_p, _i64, _p, _p, _p, _p, _i64, _p, _i32]),
    "am_pack_plan": (C.c_int, [_p, _i64, _p, _p, _p]),
    "am_pack_code_words": (_i64, [_i64]),
    "am_pack_bitmap_words": (_i64, [_i64]),
    "am_pack_host": (C.c_int, [_p, _p, _i64, _p, _i64, _p, _p, _p, _p, _i64, _p, _i32]),
    "am_pack_work_words": (_i64, [_i64]),
    "am_pack_device": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _p, _p, _p, _i64, _p, _p, _p]),
    "am_pack_device_gc": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _p, _p, _p, _i64, _p, _p, _p, _p]),
    "am_gc_finish": (C.c_int, [_p, _i64, _p, _p]),
    "am_tsv_write_f64": (C.c_int, [C.c_char_p, _p, _i64, _p, _p, _i64, _i64, _p, _i32]),
    "am_tsv_write_i32": (C.c_int, [C.c_char_p, _p, _i64, _p, _p, _i64, _i64, _p, _i32]),
    "am_tsv_shape": (C.c_int, [_p, _i64, _p, _p]),
    "am_tsv_parse_f64": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _i32]),
    "am_count_plan": (C.c_int, [_p, _i64, _i32, _i32, _p, _i64, _p, _p, _i64, _p]),
    "am_count_kmers": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _p, _i64, _i64, _i32, _p, _p, _p, _p, _p]),
    "am_normalize_work_bytes": (_i64, [_i32]),
    "am_normalize": (C.c_int, [_p, _i64, _i32, _p, _p, _i32, _i32, _p, _p, _p, _p, _p]),
    "am_normalize_planes": (C.c_int, [_p, _i64, _i32, _p, _i32, _p, _p, _p, _p, _i32, _p, _p, _p]),
    "am_colsum_work_bytes": (_i64, [_i32]),
    "am_colsum": (C.c_int, [_p, _i64, _i32, _p, _p, _p]),
    "am_colstats_work_bytes": (_i64, [_i32]),
    "am_colstats": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _p]),
    "am_gram_i8_work_bytes": (_i64, [_i64, _i32, _i32]),
    "am_gram_i8": (C.c_int, [_p, _i64, _i32, _p, _p, _i32, _p, _p, _p]),
    "am_planes_bytes": (_i64, [_i64, _i32, _i32]),
    "am_plane_scales": (C.c_int, [_p, _p, _f64, _i32, _i32, _p, _p, _p]),
    "am_quantize_planes": (C.c_int, [_p, _i64, _i32, _p, _p, _i32, _p, _p]),
    "am_gram_planes_work_bytes": (_i64, [_i64, _i32, _i32]),
    "am_gram_planes": (C.c_int, [_p, _i64, _i32, _i32, _p, _p, _p, _p]),
    "am_cov_finish": (C.c_int, [_p, _i32, _p, _p, _p, _p, _p]),
    "am_count_flags": (C.c_int, [_p, _i32, _p, _i64, _p, _p]),
    "am_stats_pack": (C.c_int, [_p, _i32, _p, _i64, _p, _p]),
    "am_gram_recenter": (C.c_int, [_p, _i32, _p, _p, _p, _p, _p]),
    "am_table_stats": (C.c_int, [_p, _p, _i64, _i32, _p, _p, _p]),
    "am_axpb_f64": (C.c_int, [_p, _i64, _f64, _f64, _p]),
    "am_zero_async": (C.c_int, [_p, _i64, _p]),
    "am_pca_finish": (C.c_int, [_p, _p, _i32, _p, _p]),
    "am_gram_work_bytes": (_i64, [_i32]),
    "am_gram": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _p]),
    "am_eigh_work_bytes": (_i64, [_i32]),
    "am_eigh": (C.c_int, [_p, _i32, _p, _p, _p, _i32, _f64, _p]),
    "am_eigh_topk_work_bytes": (_i64, [_i32, _i32]),
    "am_eigh_topk": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _p, _p]),
    "am_eigh_topk_signal": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _p, _p, C.c_uint32, _p]),
    "am_stream_wait_geq": (C.c_int, [_p, _p, C.c_uint32]),
    "am_project_planes_work_bytes": (_i64, [_i32, _i32]),
    "am_project_planes": (C.c_int, [_p, _i64, _i32, _i32, _p, _p, _p, _p, _i32, _p, _p, _p]),
    "am_project": (C.c_int, [_p, _i64, _i32, _p, _p, _i32, _p, _p]),
    "am_dbscan_work_bytes": (_i64, [_i64]),
    "am_eps_union": (C.c_int, [_p, _i64, _i32, _f64, _p, _p, _p, _p, _p]),
    "am_labels_from_parent": (C.c_int, [_p, _i64, _p, _p, _p, _p]),
    "am_iota_i32": (C.c_int, [_p, _i64, _p]),
    "am_metrics_work_bytes": (_i64, [_i64, _i32]),
    "am_cluster_metrics": (C.c_int, [_p, _i64, _p, _p, _p, _p, _i64, _i32, _p, _p, _f64, _f64, _f64, _f64,
                                     _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "am_counts_to_columns": (C.c_int, [_p, _i64, _i64, _p, _p, _p, _i32]),
    "am_columns_to_counts": (C.c_int, [_p, _p, _i32, _i64, _i64, _p, _p, _p, _i32]),
    "am_eps_grid": (C.c_int, [_f64, _p, _p, _i32, _p]),
    "am_eps_union_dyn": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _p, _p]),
    "am_sweep_args_size": (_i64, []),
    "am_sweep_ctl_bytes": (_i64, [_i64, _i32]),
    "am_sweep_init": (C.c_int, [_p, _p]),
    "am_sweep_steps": (C.c_int, [_p, _i32, _p, _p]),
    "am_sweep_graph_create": (C.c_int, [_p, _p, _p]),
    "am_sweep_graph_launch": (C.c_int, [_p, _p, _i32, _p, _p]),
    "am_sweep_graph_destroy": (C.c_int, [_p]),
    "am_sweep_advance_host": (C.c_int, [_p, _p, _i32, _i32, _f64]),
}

#: indices into the loop state of the device-resident sweep (include/autometa_b200.h, AM_SWEEP_*)
SWEEP_CTL_WORDS = 32
(SWEEP_EPS, SWEEP_STEP, SWEEP_BEST, SWEEP_DONE, SWEEP_NSTEPS, SWEEP_BEST_STEP, SWEEP_BEST_NC, SWEEP_LAST_NC,
 SWEEP_LAST_MED, SWEEP_TAKE, SWEEP_INV_CELL, SWEEP_EPS2, SWEEP_SPAN) = range(13)
SWEEP_BEST_EPS = 16


class SweepArgs(C.Structure):
    """``am_sweep_args`` (include/autometa_b200.h), field for field."""

    _fields_ = [
        ("d_X", _p), ("n_points", _i64), ("F", C.c_int32), ("n_markers", C.c_int32),
        ("h_min", _f64 * 4), ("h_max", _f64 * 4),
        ("d_parent", _p), ("d_dbscan_work", _p), ("d_labels", _p), ("d_n_clusters", _p),
        ("d_marker_rows", _p), ("d_marker_cols", _p), ("d_marker_vals", _p), ("nnz", _i64),
        ("d_coverage", _p), ("d_gc", _p), ("cutoffs", _f64 * 4),
        ("d_completeness", _p), ("d_purity", _p), ("d_cov_std", _p), ("d_gc_std", _p),
        ("d_present", _p), ("d_single", _p), ("d_summary", _p), ("d_metrics_work", _p), ("d_ctl", _p),
        ("d_best_labels", _p), ("d_best_completeness", _p), ("d_best_purity", _p), ("d_best_cov_std", _p),
        ("d_best_gc_std", _p), ("d_best_present", _p), ("d_best_single", _p),
        ("trace_cap", C.c_int32), ("reserved", C.c_int32),
    ]

EXPORTS = tuple(_SIGS)

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args


if lib.am_sweep_args_size() != C.sizeof(SweepArgs):  # header / binding mismatch: fail loudly
    raise ImportError("am_sweep_args: the library's struct is %d bytes, the binding's %d"
                      % (lib.am_sweep_args_size(), C.sizeof(SweepArgs)))


def check(rc: int) -> None:
    if rc != AM_OK:
        raise AutometaB200Error(rc, lib.am_last_error().decode(errors="replace"))


def np_ptr(a: np.ndarray) -> int:
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def t_ptr(t) -> int:
    """data pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous()
    return t.data_ptr()


def require_cuda(device=None):
    """Resolve the CUDA device for a call; raise if CUDA is unusable (no fallback)."""
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError(
            "autometa_b200 needs a CUDA device (sm_100a); torch.cuda.is_available() is False "
            "and there is no CPU fallback."
        )
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"autometa_b200 runs on CUDA devices only, got {device}")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def stream_ptr(device) -> int:
    """Raw ``cudaStream_t`` of torch's current stream on ``device`` (the C call, not the Stream object: this
    runs ~25 times per pass and the host-side cost of queueing a pass is what bounds small shards)."""
    import torch

    idx = device.index if getattr(device, "index", None) is not None else torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL = _NullCtx()


def on_device(device):
    """``torch.cuda.device(device)`` unless ``device`` is already current (then a no-op context)."""
    import torch

    if device.index is None or torch._C._cuda_getDevice() == device.index:
        return _NULL
    return torch.cuda.device(device)


def launch_count() -> int:
    return int(lib.am_launch_count())
