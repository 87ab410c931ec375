"""ctypes binding of libsuo_b200.so (include/suo_b200.h).  No fallbacks: if the
library is missing or no sm_100 device is present the calls raise."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsuo_b200.so")

SUO_OPT_CONV_BACKEND, SUO_OPT_TF32_PASSES, SUO_OPT_USE_GRAPH, SUO_OPT_CONV_PERSISTENT, SUO_OPT_MULTISTREAM, SUO_OPT_CONV_MATH, SUO_OPT_CONV_FUSE, SUO_OPT_CONV_PAIR, SUO_OPT_PDL, SUO_OPT_CONV_HALO, SUO_OPT_BA_BLOCK_DIAGONAL, SUO_OPT_PNP_MAX_POINTS = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12
SUO_OPT_SLAM_SFM = 13

_lib = None
vp = C.c_void_p


class SuoError(RuntimeError):
    pass


def exported_symbols():
    """Every entry point include/suo_b200.h declares (tests check the .so exports them all)."""
    return ["suo_create", "suo_destroy", "suo_last_error", "suo_set_option", "suo_kernel_launches",
            "suo_load_weights", "suo_forward", "suo_heatmap_reduce", "suo_crop_concat", "suo_conv2d",
            "suo_pnp_batch", "suo_ba_batch", "suo_solve_keypoints", "suo_frames", "suo_profile_network", "suo_check_range",
            "suo_forward_kp_priors", "suo_render_priors", "suo_chi2_inlier_counts", "suo_frames_u8",
            "suo_ba_last_errors", "suo_edge_linearize", "suo_frames_u8_submit", "suo_frames_wait", "suo_record_bytes",
            "suo_pack_records", "suo_allgather_results", "suo_slam_frame", "suo_activation_bytes"]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SuoError(f"{LIB_PATH} not built: run `python __graft_entry__.py` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.suo_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.suo_destroy.argtypes = [vp]
        L.suo_destroy.restype = None
        L.suo_last_error.argtypes = [vp]
        L.suo_last_error.restype = C.c_char_p
        L.suo_set_option.argtypes = [vp, C.c_int, C.c_int]
        L.suo_kernel_launches.argtypes = [vp]
        L.suo_kernel_launches.restype = C.c_longlong
        L.suo_load_weights.argtypes = [vp, vp, C.c_size_t]
        L.suo_forward.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp] + [vp] * 7 + [C.c_int, vp]
        L.suo_forward_kp_priors.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp] + [vp] * 7 + [C.c_int, vp]
        L.suo_render_priors.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]
        L.suo_chi2_inlier_counts.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_double, C.c_double, vp, C.c_int, vp]
        L.suo_heatmap_reduce.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp] + [vp] * 6 + [C.c_int, vp]
        L.suo_crop_concat.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, C.c_int,
                                      C.c_int, vp]
        L.suo_conv2d.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp,
                                 vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp]
        L.suo_pnp_batch.argtypes = [vp, vp, vp, vp, C.c_int, C.c_double, C.c_uint64, vp, vp, vp, C.c_int, vp]
        L.suo_ba_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp,
                                   C.c_int, C.c_double, C.c_double, C.c_int, vp, C.c_int, vp]
        L.suo_solve_keypoints.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_double, C.c_double,
                                          C.c_uint64, C.c_int, vp, vp, vp, vp, C.c_int, vp]
        L.suo_frames.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp, vp, C.c_double,
                                 C.c_double, C.c_uint64, C.c_int] + [vp] * 6 + [C.c_int, vp]
        L.suo_frames_u8.argtypes = L.suo_frames.argtypes
        L.suo_check_range.argtypes = [vp]
        L.suo_ba_last_errors.argtypes = [vp, vp, C.c_int, C.c_int, vp]
        L.suo_edge_linearize.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp]
        L.suo_frames_u8_submit.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp, C.c_double,
                                           C.c_double, C.c_uint64, C.c_int] + [vp] * 6 + [vp, C.c_int, vp]
        L.suo_frames_wait.argtypes = [vp, C.c_int]
        L.suo_activation_bytes.argtypes = [vp, C.POINTER(C.c_size_t)]
        L.suo_activation_bytes.restype = C.c_size_t
        L.suo_slam_frame.argtypes = ([vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int, C.c_int] + [vp] * 7 +
                                     [C.c_double, C.c_double, C.c_double, C.c_int, C.c_uint64] + [vp] * 14 + [vp, C.c_int, C.c_int, vp])
        L.suo_record_bytes.argtypes = [C.c_int]
        L.suo_record_bytes.restype = C.c_size_t
        L.suo_pack_records.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, vp, C.c_int, vp]
        L.suo_allgather_results.argtypes = [vp, vp, vp, C.c_size_t, C.c_int, vp, vp]
        L.suo_profile_network.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), vp]
        _lib = L
    return _lib


def ptr(a):
    """Raw address of a numpy array / torch tensor / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous(), "tensor must be contiguous"
        return a.data_ptr()
    raise TypeError(type(a))


class Context:
    """One suo_ctx (one GPU)."""

    def __init__(self, device: int = 0, max_crops: int = 64, crop_res: int = 256, num_kp: int = 41):
        self._h = vp()
        rc = lib().suo_create(device, max_crops, crop_res, num_kp, C.byref(self._h))
        if rc != 0:
            raise SuoError(f"suo_create failed (rc={rc}): no sm_100 GPU visible as device {device}? "
                           "libsuo_b200 has no CPU path")
        self.device, self.max_crops, self.crop_res, self.num_kp = device, max_crops, crop_res, num_kp

    def close(self):
        if self._h:
            lib().suo_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise SuoError(f"libsuo_b200 rc={rc}: {lib().suo_last_error(self._h).decode()}")

    @property
    def handle(self):
        return self._h

    def set_option(self, opt, val):
        self.check(lib().suo_set_option(self._h, opt, val))

    def kernel_launches(self) -> int:
        return int(lib().suo_kernel_launches(self._h))

    def activation_bytes(self):
        """(allocated, one-allocation-per-tensor) bytes of the network's activation tensors."""
        u = C.c_size_t(0)
        return int(lib().suo_activation_bytes(self._h, C.byref(u))), int(u.value)

    def load_weights(self, blob: bytes):
        buf = np.frombuffer(blob, dtype=np.uint8)
        self.check(lib().suo_load_weights(self._h, buf.ctypes.data, buf.size))
