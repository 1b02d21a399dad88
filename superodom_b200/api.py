"""Python host-side binding of the C ABI (include/superodom_b200.h) via ctypes.

Used by the parity tests and bench.py; the reference's own host language is C++, whose binding is the header-only
shim include/superodom_b200/LidarSlam.hpp.  No torch types cross this boundary: numpy arrays for host buffers,
raw device pointers (ints) for the *_device entry points.  There is no CPU fallback: `Context()` raises if the
shared library is missing or no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SO_LIB_PATH") or os.path.join(_HERE, "libsuperodom_b200.so")      # SO_LIB_PATH: A/B builds (scripts/)
MAX_ICP_ITERS = 32

EXPORTS = [
    "so_create", "so_destroy", "so_last_error", "so_device_available", "so_set_stream",
    "so_map_set_resolution", "so_map_set_origin", "so_map_get_origin", "so_map_shift", "so_map_set_points",
    "so_map_set_edge_points", "so_map_add_surf", "so_map_add_edge", "so_map_add_scan", "so_map_add_registered_scan", "so_map_add_scan_edge", "so_map_counts_5x5", "so_map_download", "so_map_size",
    "so_scan_prefilter", "so_scan_deskew", "so_scan_extract_uniform", "so_register", "so_register_prefiltered", "so_register_injected", "so_set_pose_sink", "so_register_batch", "so_register_batch_edges", "so_register_batch_device", "so_correspond", "so_correspond_edge", "so_evaluate",
    "so_knn", "so_knn_device", "so_sampling_indices", "so_kernel_launches", "so_bytes_copied", "so_build_flags", "so_profile_enable", "so_profile_get",
]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_map_points", C.c_uint32), ("max_scan_points", C.c_uint32),
                ("max_batch", C.c_uint32), ("line_res", C.c_float), ("plane_res", C.c_float)]


class IcpOpts(C.Structure):
    _fields_ = [("max_icp_iters", C.c_int32), ("max_surface_features", C.c_int32), ("lm_max_iterations", C.c_int32),
                ("yaw_ratio", C.c_float), ("skip_map_checks", C.c_int32), ("use_pose_prior", C.c_int32),
                ("visual_confidence_factor", C.c_float), ("prior_uncertainty", C.c_float * 3)]


class IcpResult(C.Structure):
    _fields_ = [("pose", C.c_double * 7), ("pose_opt", C.c_double * 7), ("status", C.c_int32), ("n_iterations", C.c_int32),
                ("iter_n_surf", C.c_int32 * MAX_ICP_ITERS), ("iter_n_edge", C.c_int32 * MAX_ICP_ITERS),
                ("iter_dtrans", C.c_double * MAX_ICP_ITERS), ("iter_drot", C.c_double * MAX_ICP_ITERS),
                ("iter_lm_steps", C.c_int32 * MAX_ICP_ITERS), ("iter_lm_successful", C.c_int32 * MAX_ICP_ITERS),
                ("iter_lm_termination", C.c_int32 * MAX_ICP_ITERS), ("iter_cost", C.c_double * MAX_ICP_ITERS),
                ("hist_obs", C.c_int32 * 9), ("hist_reject_plane", C.c_int32 * 7), ("hist_reject_line", C.c_int32 * 7),
                ("cov", C.c_double * 36),
                ("pos_err", C.c_double), ("pos_dir", C.c_double * 3), ("pos_inv_cond", C.c_double),
                ("ori_err_deg", C.c_double), ("ori_dir", C.c_double * 3), ("ori_inv_cond", C.c_double),
                ("total_translation", C.c_double), ("total_rotation", C.c_double),
                ("translation_from_last", C.c_double), ("rotation_from_last", C.c_double),
                ("map_surf_5x5", C.c_int32), ("map_edge_5x5", C.c_int32), ("scan_surf_num", C.c_int32), ("scan_edge_num", C.c_int32),
                ("pos_in_localmap", C.c_int32 * 3), ("prediction_source", C.c_int32),
                ("time_ms", C.c_double), ("time_total_ms", C.c_double), ("knn_searched", C.c_int32), ("knn_verified", C.c_int32)]


CORR_DTYPE = np.dtype([("n", "<f8", 3), ("d", "<f8"), ("w", "<f8"), ("nn", "<u4", 5), ("nn_d2", "<f4", 5),
                       ("status", "u1"), ("obs", "u1", 3), ("pad_", "u1", 4)], align=True)

EDGE_CORR_DTYPE = np.dtype([("a", "<f8", 3), ("b", "<f8", 3), ("w", "<f8"), ("nn", "<u4", 10), ("selected_mask", "<u4"),
                            ("status", "u1"), ("n_selected", "u1"), ("pad_", "u1", 2)], align=True)

_lib = None


def load_library():
    """dlopen the in-tree shared library and declare signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} missing: run `python -m superodom_b200.build` (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    L.so_create.restype = C.c_void_p
    L.so_create.argtypes = [C.POINTER(Config)]
    L.so_destroy.argtypes = [C.c_void_p]
    L.so_last_error.restype = C.c_char_p
    L.so_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.so_map_set_resolution.argtypes = [C.c_void_p, C.c_float, C.c_float]
    L.so_map_set_origin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.so_map_get_origin.argtypes = [C.c_void_p, C.c_void_p]
    L.so_map_shift.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.so_map_set_points.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
    L.so_map_add_surf.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
    L.so_map_add_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]
    L.so_map_add_registered_scan.argtypes = [C.c_void_p, C.c_void_p]
    L.so_map_add_scan_edge.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]
    L.so_map_set_edge_points.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
    L.so_map_add_edge.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
    L.so_correspond_edge.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    L.so_map_counts_5x5.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.so_map_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.so_map_size.restype = C.c_size_t
    L.so_map_size.argtypes = [C.c_void_p]
    L.so_scan_prefilter.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.so_scan_deskew.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_double, C.c_void_p, C.c_void_p, C.c_size_t,
                                 C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.so_scan_extract_uniform.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_float, C.c_int,
                                          C.c_void_p, C.c_size_t, C.c_void_p]
    L.so_register.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                              C.c_void_p, C.c_void_p, C.c_void_p]
    L.so_register_prefiltered.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.so_register_injected.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.so_set_pose_sink.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    L.so_register_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
    L.so_register_batch_edges.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                          C.c_void_p, C.c_void_p, C.c_void_p]
    L.so_register_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    L.so_correspond.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int32,
                                C.c_void_p, C.c_void_p, C.c_void_p]
    L.so_evaluate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.so_knn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    L.so_knn_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    L.so_sampling_indices.argtypes = [C.c_uint32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
    L.so_kernel_launches.restype = C.c_uint64
    L.so_kernel_launches.argtypes = [C.c_void_p, C.c_int]
    L.so_bytes_copied.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.so_profile_enable.argtypes = [C.c_void_p, C.c_int]
    L.so_profile_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    _lib = L
    return L


def build_flags() -> int:
    return int(load_library().so_build_flags())


def sampling_indices(n: int, max_surface_features: int) -> np.ndarray:
    """Indices a registration capped at max_surface_features processes (calculateSamplingRate + shouldProcessPoint,
    LidarSlam.cpp:346-359).  Host-only."""
    L = load_library()
    out = np.empty(n, np.uint32)
    cnt = C.c_size_t(0)
    if L.so_sampling_indices(n, max_surface_features, out.ctypes.data_as(C.c_void_p), n, C.byref(cnt)) != 0:
        raise SuperOdomError("so_sampling_indices failed")
    return out[:cnt.value].copy()


def device_available() -> bool:
    return bool(load_library().so_device_available())


class SuperOdomError(RuntimeError):
    pass


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One registration context == one LidarSLAM + LocalMap instance on one GPU."""

    def __init__(self, device: int = 0, max_map_points: int = 4 << 20, max_scan_points: int = 262144, max_batch: int = 1,
                 plane_res: float = 0.4, line_res: float = 0.2):
        self.L = load_library()
        cfg = Config(device, max_map_points, max_scan_points, max_batch, line_res, plane_res)
        h = self.L.so_create(C.byref(cfg))
        if not h:
            raise SuperOdomError("so_create failed: " + self.L.so_last_error().decode())
        self.h = C.c_void_p(h)
        self.max_batch = max_batch

    def close(self):
        if getattr(self, "h", None):
            self.L.so_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc < 0:
            raise SuperOdomError(f"{what} failed ({rc}): " + self.L.so_last_error().decode())
        return rc

    # ---- stream / instrumentation
    def set_stream(self, cuda_stream_ptr: int | None):
        self._chk(self.L.so_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)), "so_set_stream")

    def kernel_launches(self, reset: bool = False) -> int:
        return int(self.L.so_kernel_launches(self.h, int(reset)))

    def bytes_copied(self, reset: bool = False):
        a, b = C.c_uint64(), C.c_uint64()
        self._chk(self.L.so_bytes_copied(self.h, C.byref(a), C.byref(b), int(reset)), "so_bytes_copied")
        return a.value, b.value

    def profile_enable(self, on: bool):
        self._chk(self.L.so_profile_enable(self.h, int(on)), "so_profile_enable")

    def profile_get(self, cls: int, reset: bool = False):
        ms = C.c_double()
        n = C.c_uint64()
        self._chk(self.L.so_profile_get(self.h, cls, C.byref(ms), C.byref(n), int(reset)), "so_profile_get")
        return ms.value, n.value

    # ---- map
    def map_set_resolution(self, line_res: float, plane_res: float):
        self._chk(self.L.so_map_set_resolution(self.h, line_res, plane_res), "so_map_set_resolution")

    def map_set_points(self, xyzi: np.ndarray):
        a = np.ascontiguousarray(xyzi, dtype=np.float32)
        assert a.ndim == 2 and a.shape[1] >= 3
        stride = a.shape[1] * 4
        ioff = 12 if a.shape[1] >= 4 else stride
        self._chk(self.L.so_map_set_points(self.h, _p(a), a.shape[0], stride, ioff), "so_map_set_points")

    def map_set_edge_points(self, xyzi: np.ndarray):
        a = np.ascontiguousarray(xyzi, dtype=np.float32)
        stride = a.shape[1] * 4
        self._chk(self.L.so_map_set_edge_points(self.h, _p(a), a.shape[0], stride, 12 if a.shape[1] >= 4 else stride), "so_map_set_edge_points")

    def map_add_edge(self, xyzi: np.ndarray):
        a = np.ascontiguousarray(xyzi, dtype=np.float32)
        stride = a.shape[1] * 4
        self._chk(self.L.so_map_add_edge(self.h, _p(a), a.shape[0], stride, 12 if a.shape[1] >= 4 else stride), "so_map_add_edge")

    def map_add_scan_edge(self, scan_xyzi: np.ndarray, pose7):
        a = np.ascontiguousarray(scan_xyzi, dtype=np.float32)
        pose = np.ascontiguousarray(pose7, dtype=np.float64)
        stride = a.shape[1] * 4
        self._chk(self.L.so_map_add_scan_edge(self.h, _p(a), a.shape[0], stride, 12 if a.shape[1] >= 4 else stride, _p(pose)), "so_map_add_scan_edge")

    def map_add_surf(self, xyzi: np.ndarray):
        a = np.ascontiguousarray(xyzi, dtype=np.float32)
        stride = a.shape[1] * 4
        self._chk(self.L.so_map_add_surf(self.h, _p(a), a.shape[0], stride, 12 if a.shape[1] >= 4 else stride), "so_map_add_surf")

    def map_add_scan(self, scan_xyzi: np.ndarray, pose7):
        a = np.ascontiguousarray(scan_xyzi, dtype=np.float32)
        pose = np.ascontiguousarray(pose7, dtype=np.float64)
        stride = a.shape[1] * 4
        self._chk(self.L.so_map_add_scan(self.h, _p(a), a.shape[0], stride, 12 if a.shape[1] >= 4 else stride, _p(pose)), "so_map_add_scan")

    def map_add_registered_scan(self, pose7):
        """Insert the surf scan last passed to register() (still on the device) at pose7."""
        pose = np.ascontiguousarray(pose7, dtype=np.float64)
        self._chk(self.L.so_map_add_registered_scan(self.h, _p(pose)), "so_map_add_registered_scan")

    def map_set_origin(self, t):
        t = np.ascontiguousarray(t, dtype=np.float64)
        o = np.zeros(3, np.int32)
        self._chk(self.L.so_map_set_origin(self.h, _p(t), _p(o)), "so_map_set_origin")
        return o

    def map_origin(self):
        o = np.zeros(3, np.int32)
        self._chk(self.L.so_map_get_origin(self.h, _p(o)), "so_map_get_origin")
        return o

    def map_shift(self, t):
        t = np.ascontiguousarray(t, dtype=np.float64)
        o = np.zeros(3, np.int32)
        self._chk(self.L.so_map_shift(self.h, _p(t), _p(o)), "so_map_shift")
        return o

    def map_counts_5x5(self, ijk, with_edge: bool = False):
        ijk = np.ascontiguousarray(ijk, dtype=np.int32)
        ne, ns = C.c_int32(), C.c_int32()
        self._chk(self.L.so_map_counts_5x5(self.h, _p(ijk), C.byref(ne), C.byref(ns)), "so_map_counts_5x5")
        return (ne.value, ns.value) if with_edge else ns.value

    def map_size(self) -> int:
        return int(self.L.so_map_size(self.h))

    def map_download(self, mode: int = 0, ijk=None) -> np.ndarray:
        n = C.c_size_t()
        cap = self.map_size()
        out = np.zeros((max(cap, 1), 4), np.float32)
        ij = np.ascontiguousarray(ijk if ijk is not None else [0, 0, 0], dtype=np.int32)
        self._chk(self.L.so_map_download(self.h, mode, _p(ij), _p(out), cap, C.byref(n)), "so_map_download")
        return out[: n.value]

    # ---- scan pre-filter
    def scan_prefilter(self, scan_xyzi: np.ndarray, line_res: float, plane_res: float, auto_voxel_size: bool = True, download: bool = True):
        """-> (filtered float32 [m,4], line_res, plane_res, average_distance); download=False leaves the filtered cloud on the device
        (for register_prefiltered) and returns its point count in place of the array."""
        a = np.ascontiguousarray(scan_xyzi, dtype=np.float32)
        lr, pr = C.c_float(line_res), C.c_float(plane_res)
        out = np.empty((max(len(a), 1), 4), np.float32) if download else None
        n = C.c_size_t()
        avg = C.c_double(0.0)
        stride = a.shape[1] * 4
        self._chk(self.L.so_scan_prefilter(self.h, _p(a), a.shape[0], stride, 12 if a.shape[1] >= 4 else stride, int(auto_voxel_size),
                                           C.byref(lr), C.byref(pr), _p(out) if download else None, len(out) if download else 0, C.byref(n), C.byref(avg)),
                  "so_scan_prefilter")
        return (out[: n.value] if download else int(n.value)), lr.value, pr.value, avg.value

    def scan_deskew(self, points: np.ndarray, time_col: int, lidar_start_time: float, sample_times, sample_poses, imu_only: bool = False, T_i_l=None):
        """points: float32 [n, C] with x,y,z first and the per-point time in column time_col; rewritten IN PLACE.
        -> (start_pose7, n_past_end)"""
        assert points.dtype == np.float32 and points.flags.c_contiguous and points.ndim == 2
        st = np.ascontiguousarray(sample_times, np.float64)
        sp = np.ascontiguousarray(sample_poses, np.float64).reshape(-1, 7)
        til = np.ascontiguousarray(T_i_l if T_i_l is not None else [0, 0, 0, 0, 0, 0, 1], np.float64)
        out = np.zeros(7, np.float64)
        past = C.c_size_t(0)
        self._chk(self.L.so_scan_deskew(self.h, _p(points), points.shape[0], points.shape[1] * 4, time_col * 4, float(lidar_start_time), _p(st), _p(sp),
                                        len(st), int(imu_only), _p(til), _p(out), C.byref(past)), "so_scan_deskew")
        return out, past.value

    def scan_extract_uniform(self, points: np.ndarray, time_col: int, skip_num: int, block_range: float, int_abs: bool = False):
        """-> float32 [m, 4] {x, y, z, intensity = time} (featureExtraction::uniformFeatureExtraction)"""
        a = np.ascontiguousarray(points, dtype=np.float32)
        out = np.zeros((max(len(a), 1), 4), np.float32)
        n = C.c_size_t()
        self._chk(self.L.so_scan_extract_uniform(self.h, _p(a), a.shape[0], a.shape[1] * 4, time_col * 4, int(skip_num), float(block_range), int(int_abs),
                                                 _p(out), len(out), C.byref(n)), "so_scan_extract_uniform")
        return out[: n.value]

    # ---- registration
    @staticmethod
    def _opts(max_icp_iters, max_surface_features=0, lm_max_iterations=4, yaw_ratio=0.0, skip_map_checks=False, pose_prior=None):
        """pose_prior = (visual_confidence_factor, (ux, uy, uz)) enables the SE3AbsolutatePoseFactor rows."""
        o = IcpOpts(max_icp_iters, max_surface_features, lm_max_iterations, yaw_ratio, int(skip_map_checks), 0, 0.0, (C.c_float * 3)(0, 0, 0))
        if pose_prior is not None:
            o.use_pose_prior = 1
            o.visual_confidence_factor = float(pose_prior[0])
            o.prior_uncertainty = (C.c_float * 3)(*[float(v) for v in pose_prior[1]])
        return o

    def register(self, scan_xyzi: np.ndarray, pose7, max_icp_iters: int, max_surface_features: int = 0, edge_xyzi=None, **kw) -> IcpResult:
        s = np.ascontiguousarray(scan_xyzi, dtype=np.float32)
        pose = np.ascontiguousarray(pose7, dtype=np.float64)
        o = self._opts(max_icp_iters, max_surface_features, **kw)
        r = IcpResult()
        stride = s.shape[1] * 4
        e, ne = None, 0
        if edge_xyzi is not None and len(edge_xyzi):
            ea = np.ascontiguousarray(edge_xyzi, dtype=np.float32)
            assert ea.shape[1] == s.shape[1]
            e, ne = _p(ea), ea.shape[0]
        self._chk(self.L.so_register(self.h, _p(s), s.shape[0], e, ne, stride, 12 if s.shape[1] >= 4 else stride,
                                     _p(pose), C.byref(o), C.byref(r)), "so_register")
        return r

    def set_pose_sink(self, d_rows_ptr: int | None, cap_rows: int = 0, first_row: int = 0):
        """Device buffer [cap_rows, 8] float64 that every following registration call appends {pose_opt[7], status + 256 n_iter} rows to."""
        self._chk(self.L.so_set_pose_sink(self.h, C.c_void_p(d_rows_ptr or 0), cap_rows, first_row), "so_set_pose_sink")

    def register_injected(self, scan_xyzi: np.ndarray, pose7, max_icp_iters: int, nn_ids: np.ndarray, max_surface_features: int = 0, **kw) -> IcpResult:
        """so_register with caller-supplied neighbour sets: nn_ids [iters, n, 5] (int64 with -1 = none, or uint32 with 0xFFFFFFFF)."""
        s = np.ascontiguousarray(scan_xyzi, dtype=np.float32)
        pose = np.ascontiguousarray(pose7, dtype=np.float64)
        ids = np.asarray(nn_ids)
        assert ids.ndim == 3 and ids.shape[1:] == (s.shape[0], 5)
        if ids.dtype != np.uint32:
            ids = np.where(ids < 0, 0xFFFFFFFF, ids).astype(np.uint32)
        ids = np.ascontiguousarray(ids)
        o = self._opts(max_icp_iters, max_surface_features, **kw)
        r = IcpResult()
        stride = s.shape[1] * 4
        self._chk(self.L.so_register_injected(self.h, _p(s), s.shape[0], stride, 12 if s.shape[1] >= 4 else stride, _p(pose), C.byref(o),
                                              _p(ids), ids.shape[0], C.byref(r)), "so_register_injected")
        return r

    def correspond_edge(self, edge_xyzi: np.ndarray, pose7):
        e = np.ascontiguousarray(edge_xyzi, dtype=np.float32)
        pose = np.ascontiguousarray(pose7, dtype=np.float64)
        corr = np.zeros(e.shape[0], EDGE_CORR_DTYPE)
        hr = np.zeros(7, np.int32)
        stride = e.shape[1] * 4
        self._chk(self.L.so_correspond_edge(self.h, _p(e), e.shape[0], stride, 12 if e.shape[1] >= 4 else stride, _p(pose), _p(corr), _p(hr)),
                  "so_correspond_edge")
        return corr, hr

    def register_prefiltered(self, pose7, max_icp_iters: int, max_surface_features: int = 0, **kw) -> IcpResult:
        """so_register of the cloud the last scan_prefilter call left on the device."""
        pose = np.ascontiguousarray(pose7, dtype=np.float64)
        o = self._opts(max_icp_iters, max_surface_features, **kw)
        res = IcpResult()
        rc = self.L.so_register_prefiltered(self.h, _p(pose), C.byref(o), C.byref(res))
        if rc < 0:
            self._chk(rc, "so_register_prefiltered")
        return res

    def register_batch(self, scans_xyzi: np.ndarray, n_points, poses, max_icp_iters: int, max_surface_features: int = 0, **kw):
        """scans_xyzi: float32 [sum(n_points), 4] host array (pinned or pageable)."""
        n_points = np.ascontiguousarray(n_points, dtype=np.uint32)
        poses = np.ascontiguousarray(poses, dtype=np.float64)
        ns = len(n_points)
        o = self._opts(max_icp_iters, max_surface_features, **kw)
        res = (IcpResult * ns)()
        assert scans_xyzi.dtype == np.float32 and scans_xyzi.flags.c_contiguous
        stride = scans_xyzi.shape[1] * 4
        self._chk(self.L.so_register_batch(self.h, _p(scans_xyzi), _p(n_points), ns, stride, 12, _p(poses), C.byref(o), res),
                  "so_register_batch")
        return res

    def register_batch_edges(self, scans_xyzi: np.ndarray, n_points, edges_xyzi: np.ndarray, n_edge, poses, max_icp_iters: int,
                             max_surface_features: int = 0, **kw):
        """register_batch with one edge cloud per scan: edges_xyzi float32 [sum(n_edge), 4], clouds back to back."""
        n_points = np.ascontiguousarray(n_points, dtype=np.uint32)
        n_edge = np.ascontiguousarray(n_edge, dtype=np.uint32)
        poses = np.ascontiguousarray(poses, dtype=np.float64)
        ns = len(n_points)
        assert len(n_edge) == ns
        o = self._opts(max_icp_iters, max_surface_features, **kw)
        res = (IcpResult * ns)()
        scans_xyzi = np.ascontiguousarray(scans_xyzi, dtype=np.float32)
        edges_xyzi = np.ascontiguousarray(edges_xyzi, dtype=np.float32).reshape(-1, 4)
        self._chk(self.L.so_register_batch_edges(self.h, _p(scans_xyzi), _p(n_points), _p(edges_xyzi) if len(edges_xyzi) else None, _p(n_edge), ns,
                                                 16, 12, _p(poses), C.byref(o), res), "so_register_batch_edges")
        return res

    def register_batch_device(self, d_scans_ptr: int, n_points, poses, max_icp_iters: int, max_surface_features: int = 0, **kw):
        n_points = np.ascontiguousarray(n_points, dtype=np.uint32)
        poses = np.ascontiguousarray(poses, dtype=np.float64)
        ns = len(n_points)
        o = self._opts(max_icp_iters, max_surface_features, **kw)
        res = (IcpResult * ns)()
        self._chk(self.L.so_register_batch_device(self.h, C.c_void_p(d_scans_ptr), _p(n_points), ns, _p(poses), C.byref(o), res),
                  "so_register_batch_device")
        return res

    def correspond(self, scan_xyzi: np.ndarray, pose7, max_surface_features: int = 0):
        s = np.ascontiguousarray(scan_xyzi, dtype=np.float32)
        pose = np.ascontiguousarray(pose7, dtype=np.float64)
        corr = np.zeros(s.shape[0], CORR_DTYPE)
        ho = np.zeros(9, np.int32)
        hr = np.zeros(7, np.int32)
        stride = s.shape[1] * 4
        self._chk(self.L.so_correspond(self.h, _p(s), s.shape[0], stride, 12 if s.shape[1] >= 4 else stride, _p(pose),
                                       max_surface_features, _p(corr), _p(ho), _p(hr)), "so_correspond")
        return corr, ho, hr

    def evaluate(self, pose7):
        pose = np.ascontiguousarray(pose7, dtype=np.float64)
        H = np.zeros((6, 6))
        g = np.zeros(6)
        cost = C.c_double()
        self._chk(self.L.so_evaluate(self.h, _p(pose), _p(H), _p(g), C.byref(cost)), "so_evaluate")
        return H, g, cost.value

    # ---- k-NN
    def knn(self, q_xyz: np.ndarray, k: int = 5, max_d2: float = 0.0):
        q = np.ascontiguousarray(q_xyz, dtype=np.float32)
        nq = q.shape[0]
        idx = np.empty((nq, k), np.uint32)
        d2 = np.empty((nq, k), np.float32)
        self._chk(self.L.so_knn(self.h, _p(q), nq, q.shape[1] * 4, k, C.c_float(max_d2), _p(idx), _p(d2)), "so_knn")
        return idx, d2

    def knn_device(self, d_q_ptr: int, nq: int, k: int, max_d2: float, d_idx_ptr: int, d_d2_ptr: int):
        self._chk(self.L.so_knn_device(self.h, C.c_void_p(d_q_ptr), nq, k, C.c_float(max_d2), C.c_void_p(d_idx_ptr),
                                       C.c_void_p(d_d2_ptr)), "so_knn_device")
