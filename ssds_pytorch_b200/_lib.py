"""ctypes binding of libssdsb200.so (the C ABI declared in include/ssdsb200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.
PyTorch is used only as the owner of device memory and streams; the library itself has no
torch dependency and takes raw device pointers.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssdsb200.so")

SSDSB_MAX_LEVELS = 8
ABI_VERSION = 202        # == ssdsb_version(); bumped whenever include/ssdsb200.h changes incompatibly


class Level(C.Structure):
    """mirror of `ssdsb_level` (include/ssdsb200.h)."""
    _fields_ = [("conf", C.c_void_p), ("loc", C.c_void_p), ("anchors", C.c_void_p),
                ("A", C.c_int), ("C", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("stride", C.c_int)]


class LossLevel(C.Structure):
    """mirror of `ssdsb_loss_level` (include/ssdsb200.h)."""
    _fields_ = [("conf", C.c_void_p), ("loc", C.c_void_p), ("anchors", C.c_void_p), ("depth", C.c_void_p),
                ("box_target", C.c_void_p), ("A", C.c_int), ("C", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("stride", C.c_int)]


class ConvDesc(C.Structure):
    """mirror of `ssdsb_conv_desc` (include/ssdsb200.h)."""
    _fields_ = [(n, C.c_int) for n in (
        "N", "H", "W", "Cin", "Cout", "KH", "KW", "stride", "pad", "Ho", "Wo",
        "x_cstride", "out_cstride", "res_cstride", "x_row_pixels", "chunk", "x_kind", "w_rows", "relu", "out_mode", "n_loc",
        "sigmoid")]


class MbconvDesc(C.Structure):
    """mirror of `ssdsb_mbconv_desc` (include/ssdsb200.h)."""
    _fields_ = [(n, C.c_int) for n in (
        "N", "H", "W", "Cin", "hid", "Cout", "stride", "residual", "relu_expand", "relu_dw", "relu_project",
        "w_exp_rows", "w_proj_rows", "out_cstride")]


def _load():
    # (re)build in-tree if the library is missing or older than its sources (no-op when the stamp
    # matches, e.g. on the GPU box where the prebuilt .so travels with the snapshot)
    from . import build as _build
    try:
        _build.build()
    except Exception as e:   # noqa: BLE001
        # never load a library that does not match the sources it sits next to: a stale .so with a changed
        # struct layout or argument list would corrupt memory silently
        why = "is missing" if not os.path.exists(LIB_PATH) else "is STALE (csrc/ or include/ changed since it was built)"
        raise ImportError(
            f"{LIB_PATH} {why} and could not be rebuilt ({e}). There is no CPU/PyTorch fallback for the hot "
            "path; build it where nvcc exists: python -m ssds_pytorch_b200.build") from e
    lib = C.CDLL(LIB_PATH)
    # ABI handshake: version + the size of the one struct passed by pointer with many fields
    lib.ssdsb_abi_info.restype, lib.ssdsb_abi_info.argtypes = C.c_int, [C.POINTER(C.c_int)]
    abi = (C.c_int * 4)()
    lib.ssdsb_abi_info(abi)
    if abi[0] != ABI_VERSION or abi[1] != C.sizeof(ConvDesc) or abi[2] != C.sizeof(Level) or abi[3] != SSDSB_MAX_LEVELS:
        raise ImportError(f"{LIB_PATH}: ABI mismatch (library version {abi[0]}, sizeof(ssdsb_conv_desc) {abi[1]}, "
                          f"sizeof(ssdsb_level) {abi[2]}, max levels {abi[3]}; binding expects {ABI_VERSION}, "
                          f"{C.sizeof(ConvDesc)}, {C.sizeof(Level)}, {SSDSB_MAX_LEVELS})")
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    fp = C.POINTER(C.c_float)
    lp = C.POINTER(Level)
    llp = C.POINTER(LossLevel)
    sig = {
        "ssdsb_version": (i, []),
        "ssdsb_abi_info": (i, [C.POINTER(C.c_int)]),
        "ssdsb_last_error_string": (C.c_char_p, []),
        "ssdsb_generate_anchors": (i, [i, fp, i, fp, i, vp, vp]),
        "ssdsb_anchor_grid": (i, [vp, i, i, i, i, vp, vp]),
        "ssdsb_box2delta": (i, [vp, vp, i, vp, vp]),
        "ssdsb_delta2box": (i, [vp, vp, i, i, i, i, vp, vp]),
        "ssdsb_decode_workspace_bytes": (sz, [lp, i, i, i]),
        "ssdsb_decode": (i, [lp, i, i, f, i, i, vp, vp, vp, vp, vp, sz, vp]),
        "ssdsb_nms_workspace_bytes": (sz, [i, i, i]),
        "ssdsb_nms": (i, [vp, vp, vp, i, i, f, i, i, vp, vp, vp, vp, vp, vp, sz, vp]),
        "ssdsb_match_iou": (i, [vp, i, i, vp, i, i, i, i, i, f, f, f, vp, vp, vp, vp]),
        "ssdsb_multibox_loss_workspace_bytes": (sz, [i, i, i, i, i]),
        "ssdsb_multibox_loss": (i, [vp, vp, vp, i, i, i, i, i, i, vp, vp, sz, vp]),
        "ssdsb_multibox_loss_sum": (i, [vp, vp, i, i, i, i, i, i, vp, vp, vp, sz, vp]),
        "ssdsb_loss_sum_workspace_bytes": (sz, [i, i, i, i]),
        "ssdsb_focal_loss": (i, [vp, vp, i, i, i, i, i, f, f, vp, vp]),
        "ssdsb_focal_loss_sum": (i, [vp, vp, i, i, i, i, i, f, f, vp, vp, vp, sz, vp]),
        "ssdsb_loc_loss": (i, [vp, vp, i, i, i, i, i, f, vp, vp]),
        "ssdsb_loc_loss_sum": (i, [vp, vp, vp, i, i, i, i, i, f, vp, vp, sz, vp]),
        "ssdsb_multibox_loss_sum_backward": (i, [vp, vp, i, i, i, i, i, i, vp, vp, vp, sz, vp]),
        "ssdsb_focal_loss_sum_backward": (i, [vp, vp, i, i, i, i, i, f, f, vp, vp, vp]),
        "ssdsb_loc_loss_sum_backward": (i, [vp, vp, vp, i, i, i, i, i, f, vp, vp, vp]),
        "ssdsb_detection_loss_workspace_bytes": (sz, [llp, i, i]),
        "ssdsb_detection_loss": (i, [llp, i, i, vp, i, f, f, i, i, f, f, i, f, vp, vp, vp, vp, vp, sz, vp]),
        "ssdsb_conv1x1_pair_bf16": (i, [i, i, i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "ssdsb_conv_last_launch": (i, [C.POINTER(C.c_int)]),
        "ssdsb_mbconv_bf16": (i, [C.POINTER(MbconvDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "ssdsb_mbconv_last_launch": (i, [C.POINTER(C.c_int)]),
        "ssdsb_mbconv_profile": (i, [C.POINTER(C.c_ulonglong)]),
        "ssdsb_conv2d_bf16": (i, [C.POINTER(ConvDesc), vp, vp, vp, vp, vp, vp, vp]),
        "ssdsb_pack_image_s2d": (i, [vp, i, i, i, i, f, f, i, i, vp, vp]),
        "ssdsb_maxpool3x3s2_nhwc_bf16": (i, [vp, i, i, i, i, vp, vp]),
        "ssdsb_maxpool5x5s1_nhwc_bf16": (i, [vp, i, i, i, i, i, vp, i, vp]),
        "ssdsb_upsample2x_add_nhwc_bf16": (i, [vp, vp, i, i, i, i, vp]),
        "ssdsb_upsample2x_concat_nhwc_bf16": (i, [vp, vp, i, i, i, i, i, vp, vp]),
        "ssdsb_dwconv3x3_nhwc_bf16": (i, [vp, vp, vp, i, i, i, i, i, i, vp, vp]),
        "ssdsb_bifpn_fuse_nhwc_bf16": (i, [vp, vp, vp, i, f, f, f, i, i, i, i, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype, fn.argtypes = res, args
    return lib, sig


lib, SIGNATURES = _load()

SSDSB_OK, ERR_INVALID, ERR_WORKSPACE, ERR_CUDA, ERR_UNSUPPORTED = 0, -1, -2, -3, -4


def check(rc, what):
    """Map C status codes to the exception types the reference raises."""
    if rc == SSDSB_OK:
        return
    msg = f"{what}: {lib.ssdsb_last_error_string().decode(errors='replace')}"
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f"{msg} (status {rc})")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def dev_f32(t, device=None):
    """contiguous fp32 CUDA tensor (no copy if it already is one)."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    if device is None:
        device = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    if not t.is_cuda and not torch.cuda.is_available():
        raise RuntimeError("ssds_pytorch_b200 needs a CUDA device (B200); no CPU fallback exists")
    return t.to(device=device, dtype=torch.float32).contiguous()


_workspaces = {}


def workspace(nbytes, device):
    """Grow-only scratch buffer owned by the caller side (torch), 256-byte aligned — one per (device,
    stream): work queued on different streams (e.g. the detector's post-processing side stream) never
    shares scratch memory, and a buffer that is replaced by a larger one is only ever recycled by the
    caching allocator on the stream that used it."""
    stream = torch.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0
    key = (device.type, device.index, stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws
