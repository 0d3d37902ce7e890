"""ctypes binding of ``csrc/libseevcn_b200.so`` (C-ABI declared in ``include/seevcn_b200.h``).

Torch is used only for device memory and the current stream: tensors cross the boundary
as raw ``data_ptr()`` integers.  There is no fallback — a missing library or a non-CUDA
tensor raises.
"""
import ctypes
import os
from ctypes import c_int, c_size_t, c_void_p, c_char_p, POINTER

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libseevcn_b200.so")

_lib = None

P = c_void_p
I = c_int


class VcnParams(ctypes.Structure):
    """struct seevcn_vcn_params (include/seevcn_b200.h)."""
    _names = ["pose_enc0", "pose_enc2", "pose_enc4", "pose_fc0", "pose_fc2",
              "enc1_0", "enc1_3", "enc2_0", "enc2_3", "fc0", "fc2", "fc4"]
    _fields_ = [(f"{n}_{s}", c_void_p) for n in _names for s in ("w", "b")] + [
        ("num_coarse", c_int), ("viewer_centred", c_int)]


# name -> (restype, argtypes); every symbol include/seevcn_b200.h declares
SIGNATURES = {
    "seevcn_abi_version": (I, []),
    "seevcn_last_error": (c_char_p, []),
    "seevcn_launch_count": (ctypes.c_ulonglong, []),
    "seevcn_check_device": (I, [I]),
    "seevcn_prof_enable": (I, [I]),
    "seevcn_prof_report": (I, [c_char_p, c_size_t]),
    "seevcn_points_in_boxes": (I, [I, I, I, P, P, P, P]),
    "seevcn_points_in_boxes_dense": (I, [I, I, P, P, P, P]),
    "seevcn_points_in_boxes_dense_trig": (I, [I, I, P, P, P, P, P]),
    "seevcn_crop_workspace_bytes": (c_size_t, [I, I, I]),
    "seevcn_crop_points_in_boxes": (I, [I, I, I, P, P, P, P, P, P, P, c_size_t, P]),
    "seevcn_select_objects": (I, [I, I, P, I, P, P, P, P]),
    "seevcn_resample_gather": (I, [I, I, I, I, P, P, P, P, P, P, P, P, P]),
    "seevcn_resample_gather_rng": (I, [I, I, I, I, ctypes.c_uint, P, P, P, P, P, P, P, P]),
    "seevcn_resample_perm": (ctypes.c_uint, [ctypes.c_uint] * 4),
    "seevcn_furthest_point_sampling": (I, [I, I, I, P, P, P, P]),
    "seevcn_gather_points": (I, [I, I, I, I, P, P, P, P]),
    "seevcn_group_points": (I, [I, I, I, I, I, P, P, P, P]),
    "seevcn_knn": (I, [I, I, I, I, P, P, P, P, P]),
    "seevcn_knn_surface_select_workspace_bytes": (c_size_t, [I, I, I]),
    "seevcn_knn_surface_select": (I, [I, I, I, I, I, P, P, P, P, P, c_size_t, P]),
    "seevcn_largest_cluster": (I, [I, I, I, ctypes.c_double, I, P, P, P, P, P]),
    "seevcn_largest_cluster_periodic": (I, [I, I, I, ctypes.c_double, I, P, P, P, P, P, P]),
    "seevcn_vcn_create": (I, [POINTER(VcnParams), POINTER(c_void_p), P]),
    "seevcn_vcn_destroy": (None, [P]),
    "seevcn_vcn_workspace_bytes": (c_size_t, [P, I, I]),
    "seevcn_vcn_forward": (I, [P, I, I, P, P, P, P, P, P, c_size_t, I, P]),
    "seevcn_linear_bf16_workspace_bytes": (c_size_t, [I, I, I]),
    "seevcn_linear_bf16": (I, [I, I, I, P, P, P, P, I, I, P, P, P, c_size_t, P]),
    "seevcn_mean_vfe": (I, [I, I, I, P, P, P, P]),
    "seevcn_mean_vfe_int": (I, [I, I, I, P, P, P, P]),
    "seevcn_dynamic_voxelize_workspace_bytes": (c_size_t, [I, I, I, POINTER(c_int)]),
    "seevcn_dynamic_voxelize": (I, [I, I, P, POINTER(ctypes.c_float), POINTER(ctypes.c_float), POINTER(c_int),
                                    I, I, I, P, P, P, P, P, c_size_t, P]),
    "seevcn_dynamic_voxelize_frames": (I, [I, I, P, I, I, P, P, POINTER(ctypes.c_float), POINTER(ctypes.c_float), POINTER(c_int),
                                           I, I, P, P, P, P, P, c_size_t, P]),
    "seevcn_dynamic_voxelize_spliced": (I, [I, I, P, P, I, I, P, P, P, POINTER(ctypes.c_float), POINTER(ctypes.c_float),
                                            POINTER(c_int), I, I, P, P, P, P, P, c_size_t, P]),
    "seevcn_splice_workspace_bytes": (c_size_t, [I, I, I]),
    "seevcn_splice": (I, [I, I, P, I, I, P, P, P, ctypes.c_double, P, I, P, P, P, P, I, P, P, c_size_t, P]),
    "seevcn_unique_rows_frames_workspace_bytes": (c_size_t, [I, I, I, I]),
    "seevcn_unique_rows_frames": (I, [I, I, I, P, P, P, I, P, P, P, c_size_t, P]),
    "seevcn_hard_voxelize_workspace_bytes": (c_size_t, [I, I, I]),
    "seevcn_hard_voxelize": (I, [I, I, P, POINTER(ctypes.c_float), POINTER(ctypes.c_float), POINTER(c_int),
                                 I, I, P, P, P, P, P, c_size_t, P]),
    "seevcn_hard_voxelize_frames_workspace_bytes": (c_size_t, [I, I, I, I]),
    "seevcn_hard_voxelize_frames": (I, [I, I, I, P, P, POINTER(ctypes.c_float), POINTER(ctypes.c_float), POINTER(c_int),
                                        I, I, P, P, P, P, P, c_size_t, P]),
    "seevcn_project_points": (I, [I, P, POINTER(ctypes.c_double), POINTER(ctypes.c_double), POINTER(ctypes.c_double), I, I, I,
                                  P, P, P, P]),
    "seevcn_points_in_masks": (I, [I, I, I, I, P, P, P, P, P, P]),
    "seevcn_dbscan_largest": (I, [I, I, P, P, P, I, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                  ctypes.c_double, I, I, P, P, P, P, c_size_t, P]),
    "seevcn_resample_lists": (I, [I, I, I, ctypes.c_uint, P, P, P, P, P, P]),
    "seevcn_mask_points_by_range_workspace_bytes": (c_size_t, [I]),
    "seevcn_mask_points_by_range": (I, [I, I, P, POINTER(ctypes.c_float), P, P, P, c_size_t, P]),
    "seevcn_shuffle_points": (I, [I, I, ctypes.c_uint, P, P, P]),
    "seevcn_shuffle_perm": (ctypes.c_uint, [ctypes.c_uint] * 3),
    "seevcn_chamfer": (I, [I, I, I, P, P, P, P, P]),
    "seevcn_copy_to_pinned": (I, [P, P, c_size_t, P]),
    "seevcn_gather_pack": (I, [I, I, I, I, I, ctypes.c_longlong, P, P, P, P, P, P]),
    "seevcn_gather_broadcast": (I, [P, c_size_t, POINTER(c_void_p), I, P]),
    "seevcn_copy_from_pinned": (I, [P, P, c_size_t, P]),
}


def lib():
    """Load the C-ABI library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().seevcn_last_error()
        raise RuntimeError(f"seevcn_b200 error {rc}: {msg.decode() if msg else ''}")


_checked_devices = set()


def require_cuda(*tensors):
    """All tensors must be contiguous CUDA tensors on the current sm_100 device."""
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("seevcn_b200 ops need CUDA tensors (no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("seevcn_b200 ops need contiguous tensors")
        dev = t.device.index
        if dev not in _checked_devices:
            check(lib().seevcn_check_device(dev))
            _checked_devices.add(dev)


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream():
    """The current torch stream of the current device as a raw ``cudaStream_t``."""
    if _raw_stream is not None:      # no Stream object per call: this sits on every launch path
        return c_void_p(_raw_stream(torch.cuda.current_device()))
    return c_void_p(torch.cuda.current_stream().cuda_stream)


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NO_GUARD = _NoGuard()


def device_guard(dev):
    """``torch.cuda.device(dev)`` only when ``dev`` is not already current (the guard costs ~10 us per wrapper call)."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return _NO_GUARD if idx == torch.cuda.current_device() else torch.cuda.device(idx)


_workspaces = {}


def workspace(dev, nbytes, tag):
    """Grow-only scratch buffer per (device, current stream, tag).  Kernels of successive calls on one stream are
    ordered, so a stage's scratch can be reused by its next call instead of going through the allocator."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    sid = _raw_stream(idx) if _raw_stream is not None else torch.cuda.current_stream(idx).cuda_stream
    key = (idx, sid, tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=torch.device("cuda", idx))
        _workspaces[key] = buf
    return buf


def farray(vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


def iarray(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def prof_enable(on=True):
    """Start/stop the library's CUDA-event timing of its launch groups (``seevcn_prof_enable``)."""
    return lib().seevcn_prof_enable(1 if on else 0)


def prof_report():
    """-> {group name: (launch groups, total ms)} since ``prof_enable(True)``; synchronises the events."""
    buf = ctypes.create_string_buffer(1 << 16)
    check(lib().seevcn_prof_report(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.rsplit(" ", 2)
        out[name] = (int(cnt), float(ms))
    return out
