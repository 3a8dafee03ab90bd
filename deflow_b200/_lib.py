"""ctypes binding of the C-ABI library (include/deflow_b200.h) and its in-tree build recipe.

The library is the product: there is no CPU or PyTorch fallback.  If ``libdeflow_b200.so`` is
missing or cannot be loaded, every entry point raises ``RuntimeError`` -- loudly, by design.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libdeflow_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]

_lib = None


class IndexArgs(C.Structure):
    """dfb_index_args"""
    _fields_ = [("F", C.c_int), ("Nmax", C.c_int), ("pt_stride", C.c_int), ("pil_cap", C.c_longlong),
                ("voxel_size", C.c_float * 3), ("range", C.c_float * 6),
                ("pts", C.c_void_p), ("keys", C.c_void_p), ("bitmap", C.c_void_p), ("word_rank", C.c_void_p),
                ("blk_cnt", C.c_void_p), ("pt_slot", C.c_void_p), ("counts", C.c_void_p), ("pt_xyz", C.c_void_p),
                ("pt_coor", C.c_void_p), ("pt_idx", C.c_void_p), ("pt_offs", C.c_void_p), ("pt_pillar", C.c_void_p),
                ("pil_cnt", C.c_void_p), ("pil_coor", C.c_void_p), ("pil_pix", C.c_void_p),
                ("pil_start", C.c_void_p), ("sorted_pt", C.c_void_p), ("csr_rec", C.c_void_p),
                ("scan_ws", C.c_void_p), ("tickets", C.c_void_p), ("zero_base", C.c_void_p),
                ("zero_bytes", C.c_longlong), ("occ", C.c_void_p)]


class PfnArgs(C.Structure):
    """dfb_pfn_args"""
    _fields_ = [("F", C.c_int), ("H", C.c_int), ("W", C.c_int), ("training", C.c_int),
                ("voxel_size", C.c_float * 3), ("center_off", C.c_float * 3), ("eps", C.c_float),
                ("momentum", C.c_float),
                ("counts", C.c_void_p), ("pt_xyz", C.c_void_p), ("pt_coor", C.c_void_p), ("pt_pillar", C.c_void_p),
                ("pil_cnt", C.c_void_p), ("pil_coor", C.c_void_p), ("pil_pix", C.c_void_p),
                ("pil_start", C.c_void_p), ("sorted_pt", C.c_void_p), ("weight", C.c_void_p),
                ("gamma", C.c_void_p), ("beta", C.c_void_p), ("running_mean", C.c_void_p),
                ("running_var", C.c_void_p), ("pil_mean", C.c_void_p), ("stats", C.c_void_p),
                ("bn_params", C.c_void_p), ("pil_feats", C.c_void_p), ("image", C.c_void_p),
                ("image_bf16", C.c_int), ("pil_cap", C.c_longlong), ("csr_rec", C.c_void_p), ("pt_mask", C.c_void_p),
                ("partials", C.c_void_p), ("pil_hdr", C.c_void_p), ("image_ready_event", C.c_void_p),
                ("range_min", C.c_float * 3), ("phase", C.c_int), ("sync_stats", C.c_void_p), ("sync_counts", C.c_void_p)]


class PfnBwdArgs(C.Structure):
    """dfb_pfn_bwd_args"""
    _fields_ = [("fwd", PfnArgs), ("grad_image", C.c_void_p), ("grad_weight", C.c_void_p),
                ("grad_gamma", C.c_void_p), ("grad_beta", C.c_void_p), ("bwd_stats", C.c_void_p),
                ("grad_accum", C.c_void_p), ("phase", C.c_int), ("sync_bwd_stats", C.c_void_p)]


class ConvArgs(C.Structure):
    """dfb_conv_args"""
    _fields_ = [("mode", C.c_int), ("n", C.c_int), ("H", C.c_int), ("W", C.c_int), ("ksize", C.c_int),
                ("stride", C.c_int), ("n_src", C.c_int), ("x", C.c_void_p * 2), ("cin", C.c_int * 2),
                ("cin_total", C.c_int), ("cin_off", C.c_int), ("cout", C.c_int), ("w", C.c_void_p),
                ("bias", C.c_void_p), ("y", C.c_void_p), ("y_fp32", C.c_int), ("stats", C.c_void_p),
                ("x_lo", C.c_void_p * 2), ("split3", C.c_int), ("stats_sum_only", C.c_int),
                ("y2", C.c_void_p), ("cin2", C.c_int), ("grad_bias", C.c_void_p)]


class PackDesc(C.Structure):
    """dfb_pack_desc"""
    _fields_ = [("w", C.c_void_p), ("w_fwd", C.c_void_p), ("w_dgrad", C.c_void_p), ("first", C.c_longlong),
                ("cout", C.c_int), ("cin", C.c_int), ("ksize", C.c_int), ("pad_", C.c_int)]


class EvalTables(C.Structure):
    """dfb_eval_tables"""
    _fields_ = [("fg_bg", C.c_ubyte * 256), ("meta", C.c_ubyte * 256), ("speed_splits", C.c_double * 52),
                ("dist_splits", C.c_double * 6), ("n_speed", C.c_int), ("n_dist", C.c_int)]


class UnpackDesc(C.Structure):
    """dfb_unpack_desc"""
    _fields_ = [("wacc", C.c_void_p), ("grad", C.c_void_p), ("first", C.c_longlong),
                ("cout", C.c_int), ("cin", C.c_int), ("ksize", C.c_int), ("accumulate", C.c_int)]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile deflow_b200/csrc/*.cu for sm_100a into deflow_b200/lib/libdeflow_b200.so (nvcc
    cross-compiles without a GPU).  Objects are cached by mtime."""
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "deflow_b200.h"))
    newest_hdr = max(os.path.getmtime(h) for h in hdrs)
    jobs, objs = [], []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), newest_hdr):
            jobs.append(["nvcc", *NVCC_FLAGS, "-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose and (r.stdout or r.stderr):
            print(r.stdout + r.stderr, file=sys.stderr)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB_PATH):
        run(["nvcc", "-shared", "-o", LIB_PATH, *objs, "-lcudart"])
    return LIB_PATH


def _declare(lib):
    vp, i32, i64, f32p = C.c_void_p, C.c_int, C.c_longlong, C.POINTER(C.c_float)
    lib.dfb_last_error.restype = C.c_char_p
    lib.dfb_last_error.argtypes = []
    lib.dfb_version.restype = i32
    lib.dfb_launch_count.restype = i64
    sig = {
        "dfb_grid_size": [f32p, f32p, C.POINTER(C.c_int)],
        "dfb_dynamic_voxelize_forward": [vp, i32, i32, f32p, f32p, vp, vp],
        "dfb_scatter_index": [vp, i32, C.POINTER(C.c_int), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
        "dfb_scatter_reduce": [vp, i32, i32, vp, vp, vp, i32, i32, vp, vp],
        "dfb_dynamic_point_to_voxel_backward": [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp],
        "dfb_index_workspace": [i32, i32, f32p, f32p, C.POINTER(i64), C.POINTER(i64)],
        "dfb_pillar_index": [C.POINTER(IndexArgs), vp],
        "dfb_ego_warp": [vp, vp, vp, vp, i32, i32, vp, i64, vp, vp, vp],
        "dfb_pfn_forward": [C.POINTER(PfnArgs), vp],
        "dfb_pfn_backward": [C.POINTER(PfnBwdArgs), vp],
        "dfb_zero_fill": [vp, i64, i32, vp],
        "dfb_clear_rows": [vp, i32, vp, vp, i32, i64, vp],
        "dfb_decoder_gather": [vp, vp, i32, i32, i32, i32, vp, i32, vp, vp, vp, i32, i32, vp],
        "dfb_decoder_gather_backward": [vp, i32, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp, i32, i32, vp],
        "dfb_decoder_gather_backward_rows": [vp, i32, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp, i32, i32, vp, vp],
        "dfb_gather_img_rows_add": [vp, i32, i32, i32, vp, i32, vp, vp, i32, i32, vp],
        "dfb_add_cat2": [vp, vp, vp, vp, i64, i32, vp, vp],
        "dfb_flow_loss": [i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, i32, vp],
        "dfb_conv_pack_weights": [vp, i32, i32, i32, i32, vp, vp, vp],
        "dfb_conv_pack_weights_multi": [vp, i32, i64, i32, vp],
        "dfb_wgrad_unpack_multi": [vp, i32, i64, vp],
        "dfb_split_bf16x2": [vp, i64, vp, vp, vp],
        "dfb_conv2d": [C.POINTER(ConvArgs), vp],
        "dfb_conv2d_wgrad": [C.POINTER(ConvArgs), vp, vp, i32, vp],
        "dfb_bn2d_finalize": [vp, C.c_double, i32, i32, C.c_float, C.c_float, vp, vp, vp, vp, vp, vp],
        "dfb_bn_gelu_apply": [vp, vp, i32, i64, vp, i32, vp],
        "dfb_bn_gelu_backward": [vp, vp, vp, i32, i64, i32, vp, vp, vp, vp, vp, i32, vp],
        "dfb_bn_gelu_backward_phase": [vp, vp, vp, i32, i64, i32, vp, vp, vp, vp, vp, i32, i32, C.c_double, vp],
        "dfb_channel_sum": [vp, i32, i64, vp, vp, i32, vp],
        "dfb_upsample2x": [vp, i32, i32, i32, i32, vp, i32, i32, vp],
        "dfb_offset_encode": [vp, vp, vp, i32, i32, i32, vp, i32, vp],
        "dfb_offset_encode_backward": [vp, vp, i32, i32, vp, vp, vp],
        "dfb_to_bf16_pad": [vp, i32, i32, i32, vp, vp],
        "dfb_gru_rh": [vp, vp, i32, i32, vp, i32, vp],
        "dfb_gru_update": [vp, vp, vp, i32, i32, vp, vp, i32, vp],
        "dfb_gru_bwd1": [vp, vp, vp, vp, i32, i32, vp, vp, vp, i32, vp],
        "dfb_gru_bwd2": [vp, vp, vp, i32, i32, vp, vp, i32, vp],
        "dfb_acc_bf16": [vp, vp, vp, i64, i32, vp],
        "dfb_head_out": [vp, vp, vp, i32, vp, i32, vp],
        "dfb_head_out_backward": [vp, vp, vp, i32, i32, vp, vp, vp, i32, vp],
        "dfb_eval_accumulate": [vp, vp, vp, i32, vp, vp, vp, i64, C.POINTER(EvalTables), vp, vp],
        "dfb_chamfer_forward": [vp, i32, vp, i32, vp, vp, vp, vp, vp, vp],
        "dfb_chamfer_backward": [vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp],
        "dfb_hard_voxelize_assign": [vp, i32, i32, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp],
        "dfb_conv3x3_dgrad_colsum": [vp, i32, i32, i32, i32, i32, vp, vp, i32, i32, i32, vp, vp, vp],
        "dfb_collate_pad": [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
        "dfb_gru_fused_forward": [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp],
        "dfb_gru_fused_backward": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.restype = i32
        fn.argtypes = args
    lib.dfb_index_scan_workspace.restype = i64
    lib.dfb_index_scan_workspace.argtypes = [i32, i64, i64]
    lib.dfb_hard_voxelize_workspace.restype = i64
    lib.dfb_hard_voxelize_workspace.argtypes = [i32]
    lib.dfb_collate_workspace.restype = i64
    lib.dfb_collate_workspace.argtypes = [i32, i32]
    return sig


EXPORTS = ["dfb_last_error", "dfb_version", "dfb_launch_count", "dfb_grid_size", "dfb_dynamic_voxelize_forward",
           "dfb_scatter_index", "dfb_scatter_reduce", "dfb_dynamic_point_to_voxel_backward", "dfb_index_workspace",
           "dfb_index_scan_workspace", "dfb_pillar_index", "dfb_ego_warp", "dfb_pfn_forward", "dfb_pfn_backward", "dfb_zero_fill", "dfb_clear_rows", "dfb_decoder_gather",
           "dfb_decoder_gather_backward", "dfb_flow_loss", "dfb_conv_pack_weights", "dfb_conv_pack_weights_multi", "dfb_wgrad_unpack_multi", "dfb_split_bf16x2", "dfb_conv2d", "dfb_conv2d_wgrad",
           "dfb_bn2d_finalize", "dfb_bn_gelu_apply", "dfb_bn_gelu_backward", "dfb_channel_sum", "dfb_upsample2x",
           "dfb_offset_encode", "dfb_offset_encode_backward", "dfb_to_bf16_pad", "dfb_gru_rh", "dfb_gru_update",
           "dfb_gru_bwd1", "dfb_gru_bwd2", "dfb_acc_bf16", "dfb_head_out", "dfb_head_out_backward",
           "dfb_gru_fused_forward", "dfb_gru_fused_backward", "dfb_collate_workspace", "dfb_collate_pad",
           "dfb_eval_accumulate", "dfb_bn_gelu_backward_phase",
           "dfb_chamfer_forward", "dfb_chamfer_backward", "dfb_hard_voxelize_workspace", "dfb_hard_voxelize_assign",
           "dfb_conv3x3_dgrad_colsum", "dfb_decoder_gather_backward_rows", "dfb_gather_img_rows_add", "dfb_add_cat2"]


def lib():
    """The loaded library; raises RuntimeError if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(deflow_b200 has no CPU / PyTorch fallback)")
        try:
            handle = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise RuntimeError(f"cannot load {LIB_PATH}: {e}") from e
        _declare(handle)
        _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().dfb_last_error().decode()
        raise RuntimeError(f"deflow_b200 {what}: {msg} (code {rc})")


def launch_count() -> int:
    return int(lib().dfb_launch_count())
