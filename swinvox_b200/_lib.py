"""ctypes binding of libswinvox_b200.so (include/swinvox_b200.h).

There is no fallback: if the CUDA library is missing or its struct layout disagrees with this
mirror, importing the product path raises.  Build it with ``python __graft_entry__.py build``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SVX_LIB_PATH: an instrumented build of the same sources (tools/probes/slab_profile.sh); never a different backend
LIB_PATH = os.environ.get("SVX_LIB_PATH") or os.path.join(_HERE, "libswinvox_b200.so")

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_GELU = 0, 1, 2, 3
A_PLAIN, A_GATHER, A_FLAT, A_SLAB3, A_IM2COL = 0, 1, 2, 3, 4
OPERAND_DEFAULT, OPERAND_TF32, OPERAND_BF16 = 0, 1, 2
IO_OUT_BF16, IO_RES_BF16 = 1, 2
DT_IN_BF16, DT_OUT_BF16, DT_BF16 = 1, 2, 3
EPI_STD, EPI_DEC_TAIL, EPI_POOL8, EPI_CONVT8 = 0, 1, 2, 3
POOL_MAX, POOL_AVG = 0, 1

i32, i64, f32, ptr = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", i32), ("N", i32), ("K", i32), ("Kpad", i32), ("Npad", i32), ("block_n", i32), ("a_mode", i32),
        ("A", ptr), ("lda", i64),
        ("in_D", i32), ("in_H", i32), ("in_W", i32), ("in_Cs", i32), ("in_c0", i32), ("Cin", i32),
        ("out_D", i32), ("out_H", i32), ("out_W", i32),
        ("stride_d", i32), ("stride_h", i32), ("stride_w", i32),
        ("ntaps", i32), ("taps", ptr), ("taps_host", ptr), ("valid_D", i32), ("valid_H", i32), ("valid_W", i32),
        ("W", ptr), ("bias", ptr), ("residual", ptr), ("out", ptr),
        ("o_base", i64), ("o_sn", i64), ("o_sd", i64), ("o_sh", i64), ("o_sw", i64),
        ("act", i32), ("act_param", f32), ("res_after_act", i32), ("out_scale", f32),
        ("round_tf32", i32), ("epi_mode", i32), ("epi_aux", ptr), ("epi_out2", ptr),
        ("o2_base", i64), ("o2_sn", i64), ("o2_sd", i64), ("o2_sh", i64), ("o2_sw", i64),
        ("cin_live", i32), ("res_via_mma", i32),
        ("cls_cout", i32), ("acc_scale", f32), ("c_sd", i64), ("c_sh", i64), ("c_sw", i64),
        ("c2_sd", i64), ("c2_sh", i64), ("c2_sw", i64),
        ("range_flag", ptr), ("operand_kind", i32), ("io_flags", i32),
    ]


class Im2colDesc(C.Structure):
    _fields_ = [
        ("inp", ptr), ("out", ptr),
        ("N", i32), ("C", i32), ("D", i32), ("H", i32), ("W", i32),
        ("s_n", i64), ("s_c", i64), ("s_d", i64), ("s_h", i64), ("s_w", i64),
        ("KD", i32), ("KH", i32), ("KW", i32), ("stride", i32), ("pad_d", i32), ("pad_h", i32), ("pad_w", i32),
        ("OD", i32), ("OH", i32), ("OW", i32), ("Kpad", i32), ("round_tf32", i32),
    ]


class PoolDesc(C.Structure):
    _fields_ = [
        ("inp", ptr), ("out", ptr),
        ("N", i32), ("C", i32), ("D", i32), ("H", i32), ("W", i32), ("in_Cs", i32), ("out_Cs", i32),
        ("KD", i32), ("KH", i32), ("KW", i32), ("SD", i32), ("SH", i32), ("SW", i32),
        ("PD", i32), ("PH", i32), ("PW", i32), ("OD", i32), ("OH", i32), ("OW", i32),
        ("mode", i32), ("round_tf32", i32), ("dtype", i32),
    ]


class LnRowsDesc(C.Structure):
    _fields_ = [
        ("inp", ptr), ("out", ptr), ("gamma", ptr), ("beta", ptr),
        ("rows", i32), ("C", i32), ("merge", i32), ("H", i32), ("W", i32), ("eps", f32), ("round_tf32", i32), ("dtype", i32),
    ]


class LnSampleDesc(C.Structure):
    _fields_ = [
        ("inp", ptr), ("out", ptr), ("gamma", ptr), ("beta", ptr),
        ("N", i32), ("L", i32), ("eps", f32), ("round_tf32", i32), ("dtype", i32),
    ]


class WinAttnDesc(C.Structure):
    _fields_ = [
        ("qkv", ptr), ("out", ptr), ("bias", ptr),
        ("N", i32), ("H", i32), ("W", i32), ("C", i32), ("heads", i32), ("shift", i32), ("scale", f32),
        ("round_tf32", i32), ("dtype", i32), ("reserved0", i32), ("range_flag", ptr),
    ]


class DwConvDesc(C.Structure):
    _fields_ = [
        ("inp", ptr), ("out", ptr), ("w", ptr), ("bias", ptr),
        ("N", i32), ("H", i32), ("W", i32), ("C", i32), ("k", i32), ("OH", i32), ("OW", i32), ("round_tf32", i32), ("dtype", i32),
    ]


class ViewAttnDesc(C.Structure):
    _fields_ = [
        ("qkv", ptr), ("out", ptr),
        ("B", i32), ("V", i32), ("P", i32), ("R", i32), ("heads", i32), ("scale", f32), ("round_tf32", i32), ("dtype", i32),
    ]


class BilinearDesc(C.Structure):
    _fields_ = [
        ("inp", ptr), ("skip", ptr), ("out", ptr),
        ("N", i32), ("IH", i32), ("IW", i32), ("OH", i32), ("OW", i32), ("C", i32), ("round_tf32", i32), ("dtype", i32),
    ]


class MergeFuseDesc(C.Structure):
    _fields_ = [("weights", ptr), ("coarse", ptr), ("out", ptr), ("B", i32), ("V", i32), ("P", i32)]


class MetricsDesc(C.Structure):
    _fields_ = [
        ("logits", ptr), ("gt", ptr), ("prob_thresholds", ptr), ("counts", ptr),
        ("B", i32), ("P", i32), ("T", i32), ("reserved0", i32), ("bce_q20", ptr),
    ]


class TransposeDesc(C.Structure):
    _fields_ = [
        ("inp", ptr), ("out", ptr),
        ("N", i32), ("C", i32), ("P", i32), ("Cs", i32), ("to_channels_last", i32), ("round_tf32", i32),
        ("row_w", i32), ("row_pitch", i32), ("row_x0", i32), ("dtype", i32),
    ]


class Conv3to1Desc(C.Structure):
    _fields_ = [("inp", ptr), ("w", ptr), ("bias", ptr), ("out", ptr),
                ("N", i32), ("D", i32), ("H", i32), ("W", i32), ("Cs", i32), ("c0", i32), ("Cin", i32), ("slope", f32)]


class MlpDesc(C.Structure):
    _fields_ = [("x", ptr), ("ldx", i64), ("W1", ptr), ("b1", ptr), ("W2", ptr), ("b2", ptr),
                ("residual", ptr), ("out", ptr), ("ldo", i64),
                ("M", i32), ("C", i32), ("hidden", i32), ("round_tf32", i32),
                ("ln_gamma", ptr), ("ln_beta", ptr), ("ln_eps", f32), ("reserved0", i32)]


class BinvoxDecodeDesc(C.Structure):
    _fields_ = [("payload", ptr), ("offsets", ptr), ("out", ptr), ("status", ptr),
                ("B", i32), ("d0", i32), ("d1", i32), ("d2", i32), ("fix_coords", i32)]


class BinvoxEncodeDesc(C.Structure):
    _fields_ = [("volume", ptr), ("threshold", f32), ("payload", ptr), ("nbytes", ptr),
                ("B", i32), ("d0", i32), ("d1", i32), ("d2", i32), ("axis_xyz", i32)]


class PreprocessDesc(C.Structure):
    _fields_ = [("inp", ptr), ("out", ptr), ("N", i32), ("H", i32), ("W", i32), ("C", i32), ("OH", i32), ("OW", i32),
                ("y0", i32), ("y1", i32), ("x0", i32), ("x1", i32), ("mean", f32 * 3), ("std", f32 * 3), ("bg_norm", f32 * 3),
                ("reserved0", i32), ("windows", ptr), ("bg_norm_n", ptr)]


class ResizeDesc(C.Structure):
    _fields_ = [("inp", ptr), ("out", ptr), ("NC", i32), ("IH", i32), ("IW", i32), ("OH", i32), ("OW", i32), ("reserved0", i32)]


# order must match svx_desc_sizes()
DESC_TYPES = [GemmDesc, Im2colDesc, PoolDesc, LnRowsDesc, LnSampleDesc, WinAttnDesc, DwConvDesc,
              ViewAttnDesc, BilinearDesc, MergeFuseDesc, MetricsDesc, TransposeDesc, Conv3to1Desc, MlpDesc, ResizeDesc]

# op name -> (immediate symbol, plan_add symbol, descriptor type)
OPS = {
    "gemm": ("svx_gemm", "svx_plan_add_gemm", GemmDesc),
    "im2col": ("svx_im2col", "svx_plan_add_im2col", Im2colDesc),
    "pool": ("svx_pool", "svx_plan_add_pool", PoolDesc),
    "layernorm_rows": ("svx_layernorm_rows", "svx_plan_add_layernorm_rows", LnRowsDesc),
    "layernorm_sample": ("svx_layernorm_sample", "svx_plan_add_layernorm_sample", LnSampleDesc),
    "window_attention": ("svx_window_attention", "svx_plan_add_window_attention", WinAttnDesc),
    "dwconv": ("svx_dwconv", "svx_plan_add_dwconv", DwConvDesc),
    "view_attention": ("svx_view_attention", "svx_plan_add_view_attention", ViewAttnDesc),
    "bilinear_add": ("svx_bilinear_add", "svx_plan_add_bilinear_add", BilinearDesc),
    "merger_fuse": ("svx_merger_fuse", "svx_plan_add_merger_fuse", MergeFuseDesc),
    "voxel_metrics": ("svx_voxel_metrics", "svx_plan_add_voxel_metrics", MetricsDesc),
    "transpose": ("svx_transpose", "svx_plan_add_transpose", TransposeDesc),
    "conv3to1": ("svx_conv3to1", "svx_plan_add_conv3to1", Conv3to1Desc),
    "mlp": ("svx_mlp", "svx_plan_add_mlp", MlpDesc),
    "resize_bilinear": ("svx_resize_bilinear", "svx_plan_add_resize_bilinear", ResizeDesc),
}

OTHER_SYMBOLS = ["svx_abi_version", "svx_last_error", "svx_desc_sizes", "svx_device_info", "svx_plan_create",
                 "svx_plan_destroy", "svx_plan_num_ops", "svx_plan_run", "svx_plan_run_range",
                 "svx_plan_time_ops", "svx_plan_num_launches", "svx_plan_set_lane", "svx_plan_add_join"]

ALL_SYMBOLS = OTHER_SYMBOLS + [s for v in OPS.values() for s in v[:2]]
# data-format entry points (svx_io.cu): present in the CUDA library only (the CPU twin of the tests has no copy)
IO_SYMBOLS = {"svx_binvox_decode": BinvoxDecodeDesc, "svx_binvox_encode": BinvoxEncodeDesc,
              "svx_preprocess": PreprocessDesc}


class SvxError(RuntimeError):
    pass


_lib = None


def bind(path):
    """dlopen `path` and declare every prototype; raises if a symbol or a struct size is off."""
    if not os.path.exists(path):
        raise SvxError(
            f"{path} not found: the CUDA extension is not built. Run `python __graft_entry__.py build` "
            "(nvcc, sm_100a). swinvox_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(path)
    for name in ALL_SYMBOLS:
        if not hasattr(lib, name):
            raise SvxError(f"{path} does not export {name}")
    lib.svx_last_error.restype = C.c_char_p
    lib.svx_plan_create.restype = ptr
    lib.svx_plan_destroy.argtypes = [ptr]
    lib.svx_plan_destroy.restype = None
    lib.svx_plan_num_ops.argtypes = [ptr]
    lib.svx_plan_num_launches.argtypes = [ptr]
    lib.svx_plan_run.argtypes = [ptr, ptr, C.c_int]
    lib.svx_plan_run_range.argtypes = [ptr, C.c_int, C.c_int, ptr]
    lib.svx_plan_time_ops.argtypes = [ptr, ptr, C.c_int, C.POINTER(C.c_float)]
    lib.svx_plan_set_lane.argtypes = [ptr, C.c_int]
    lib.svx_plan_add_join.argtypes = [ptr]
    lib.svx_desc_sizes.argtypes = [C.POINTER(i32), C.c_int]
    lib.svx_device_info.argtypes = [C.c_int, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    for name, desc_t in IO_SYMBOLS.items():
        if hasattr(lib, name):
            getattr(lib, name).argtypes = [C.POINTER(desc_t), ptr]
        elif os.path.abspath(path) == os.path.abspath(LIB_PATH):
            raise SvxError(f"{path} does not export {name}")
    for imm, add, desc_t in OPS.values():
        getattr(lib, imm).argtypes = [C.POINTER(desc_t), ptr]
        getattr(lib, add).argtypes = [ptr, C.POINTER(desc_t)]
    sizes = (i32 * len(DESC_TYPES))()
    n = lib.svx_desc_sizes(sizes, len(DESC_TYPES))
    if n != len(DESC_TYPES):
        raise SvxError(f"descriptor count mismatch: library {n}, python {len(DESC_TYPES)}")
    for t, s in zip(DESC_TYPES, sizes):
        if C.sizeof(t) != s:
            raise SvxError(f"struct layout mismatch for {t.__name__}: python {C.sizeof(t)} bytes, library {s}")
    if lib.svx_abi_version() != 2:
        raise SvxError("ABI version mismatch")
    return lib


def get():
    global _lib
    if _lib is None:
        _lib = bind(LIB_PATH)
    return _lib


def check(rc, lib=None):
    if rc != 0:
        lib = lib or get()
        raise SvxError(lib.svx_last_error().decode("utf-8", "replace"))
