"""ctypes front-end of the CPU oracle (oracle/vgicp_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (fast_gicp_b200/) never imports this module.

Layouts: points (N,3) float32 C-contiguous; covariances (N,9) float32, each row a column-major 3x3 (the reference's
Eigen::Matrix3f memory image); poses 4x4 float64 given/returned as numpy (4,4) in the usual row/col indexing
(converted to Eigen's column-major memory image at the boundary).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libvgicp_oracle.so")

# gicp_settings.hpp:6,8 numeric enum order
REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS = range(5)
DIRECT27, DIRECT7, DIRECT1, DIRECT_RADIUS = range(4)


def build(force=False):
    src = os.path.join(_HERE, "vgicp_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={k: v for k, v in os.environ.items() if k not in ("CC", "CFLAGS")})
    return _LIB_PATH


class LsqParams(C.Structure):
    _fields_ = [
        ("max_iterations", C.c_int),
        ("rotation_epsilon", C.c_double),
        ("transformation_epsilon", C.c_double),
        ("use_gauss_newton", C.c_int),
        ("lm_max_iterations", C.c_int),
        ("lm_init_lambda_factor", C.c_double),
    ]


class LsqResult(C.Structure):
    _fields_ = [
        ("T", C.c_double * 16),
        ("H", C.c_double * 36),
        ("iterations", C.c_int),
        ("converged", C.c_int),
        ("n_linearize", C.c_int),
        ("n_error", C.c_int),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp, ip, dp, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p
        L.orc_vector3i_hash.restype = C.c_uint64
        L.orc_vector3i_hash.argtypes = [C.c_int] * 3
        L.orc_voxel_coords.argtypes = [fp, C.c_int, C.c_float, ip]
        L.orc_hashes.argtypes = [ip, C.c_int, C.POINTER(C.c_uint64)]
        L.orc_knn_bruteforce.argtypes = [fp, C.c_int, C.c_int, ip, fp]
        L.orc_knn_kdtree.argtypes = [fp, C.c_int, C.c_int, ip]
        L.orc_covariances.argtypes = [fp, C.c_int, C.c_int, ip, fp]
        L.orc_covariances_rbf.argtypes = [fp, C.c_int, C.c_float, C.c_float, fp]
        L.orc_eig3_direct.argtypes = [fp, fp, fp]
        L.orc_regularize.argtypes = [fp, C.c_int, C.c_int]
        L.orc_voxelmap_build.restype = vp
        L.orc_voxelmap_build.argtypes = [fp, fp, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_ndt_voxelmap_build.restype = vp
        L.orc_ndt_voxelmap_build.argtypes = [fp, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_align_ndt.argtypes = [vp, fp, fp, C.c_int, ip, C.c_int, C.POINTER(LsqParams), dp, C.c_int, C.POINTER(LsqResult)]
        L.orc_evaluate_ndt.restype = C.c_double
        L.orc_evaluate_ndt.argtypes = [vp, fp, fp, C.c_int, ip, C.c_int, dp, dp, dp, dp, C.c_int, C.POINTER(C.c_long)]
        L.orc_voxelmap_free.argtypes = [vp]
        L.orc_voxelmap_num_buckets.argtypes = [vp]
        L.orc_voxelmap_num_voxels.argtypes = [vp]
        L.orc_voxelmap_get.argtypes = [vp, ip, ip, ip, fp, fp]
        L.orc_offsets.argtypes = [C.c_int, C.c_double, ip, C.c_int]
        L.orc_transform_points.argtypes = [fp, fp, C.c_int, fp]
        L.orc_find_correspondences.restype = C.c_long
        L.orc_find_correspondences.argtypes = [vp, fp, C.c_int, fp, ip, C.c_int, ip, C.c_long]
        L.orc_compute_derivatives.restype = C.c_double
        L.orc_compute_derivatives.argtypes = [vp, fp, fp, ip, C.c_long, fp, fp, dp, dp, C.c_int]
        L.orc_se3_exp.argtypes = [dp, dp]
        L.orc_ldlt_solve6.argtypes = [dp, dp, dp]
        L.orc_lsq_default_params.argtypes = [C.POINTER(LsqParams)]
        L.orc_align_f32.argtypes = [vp, fp, fp, C.c_int, ip, C.c_int, C.POINTER(LsqParams), dp, C.c_int, C.POINTER(LsqResult)]
        L.orc_evaluate_f32.restype = C.c_double
        L.orc_evaluate_f32.argtypes = [vp, fp, fp, C.c_int, ip, C.c_int, dp, dp, dp, dp, C.c_int, C.POINTER(C.c_long)]
        L.orc64_covariances.argtypes = [fp, C.c_int, C.c_int, C.c_int, dp, C.c_int]
        L.orc64_align.argtypes = [fp, dp, C.c_int, fp, dp, C.c_int, C.c_double, ip, C.c_int, C.POINTER(LsqParams), dp, C.c_int, C.POINTER(LsqResult)]
        L.orc64_align_gicp.argtypes = [fp, dp, C.c_int, fp, dp, C.c_int, C.c_double, C.POINTER(LsqParams), dp, C.c_int, C.POINTER(LsqResult)]
        L.orc_num_threads.restype = C.c_int
        L.orc_remove_near_origin.argtypes = [fp, C.c_int, fp]
        L.orc_approximate_voxel_grid.argtypes = [fp, C.c_int, C.c_float, C.c_int, fp]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _pose_in(T):
    """(4,4) numpy -> column-major 16 doubles (Eigen::Isometry3d::data())."""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float64).T).reshape(16)


def _pose_out(buf):
    return np.array(buf, dtype=np.float64).reshape(4, 4).T.copy()


def default_params(**kw):
    p = LsqParams()
    lib().orc_lsq_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def num_threads():
    return lib().orc_num_threads()


def vector3i_hash(x, y, z):
    return int(lib().orc_vector3i_hash(int(x), int(y), int(z)))


def voxel_coords(pts, res):
    pts = _f32(pts)
    out = np.empty((len(pts), 3), dtype=np.int32)
    lib().orc_voxel_coords(_p(pts, C.c_float), len(pts), C.c_float(res), _p(out, C.c_int))
    return out


def hashes(coords):
    coords = _i32(coords)
    out = np.empty(len(coords), dtype=np.uint64)
    lib().orc_hashes(_p(coords, C.c_int), len(coords), _p(out, C.c_uint64))
    return out


def knn(pts, k, method="kdtree", return_d2=False):
    pts = _f32(pts)
    idx = np.empty((len(pts), k), dtype=np.int32)
    if method == "bruteforce":
        d2 = np.empty((len(pts), k), dtype=np.float32)
        rc = lib().orc_knn_bruteforce(_p(pts, C.c_float), len(pts), k, _p(idx, C.c_int), _p(d2, C.c_float))
        if rc:
            raise ValueError("knn: bad k")
        return (idx, d2) if return_d2 else idx
    rc = lib().orc_knn_kdtree(_p(pts, C.c_float), len(pts), k, _p(idx, C.c_int))
    if rc:
        raise ValueError("knn: bad k")
    return idx


def covariances(pts, nbr):
    pts, nbr = _f32(pts), _i32(nbr)
    out = np.empty((len(pts), 9), dtype=np.float32)
    lib().orc_covariances(_p(pts, C.c_float), len(pts), nbr.shape[1], _p(nbr, C.c_int), _p(out, C.c_float))
    return out


def covariances_rbf(pts, kernel_width=0.5, max_dist=3.0):
    """covariance_estimation_rbf.cu:59-151 (NearestNeighborMethod::GPU_RBF_KERNEL); defaults = fast_vgicp_cuda_impl.hpp:31."""
    pts = _f32(pts)
    out = np.empty((len(pts), 9), dtype=np.float32)
    lib().orc_covariances_rbf(_p(pts, C.c_float), len(pts), C.c_float(kernel_width), C.c_float(max_dist), _p(out, C.c_float))
    return out


def eig3_direct(cov9):
    cov9 = _f32(cov9).reshape(9)
    ev = np.empty(3, dtype=np.float32)
    V = np.empty(9, dtype=np.float32)
    lib().orc_eig3_direct(_p(cov9, C.c_float), _p(ev, C.c_float), _p(V, C.c_float))
    return ev, V.reshape(3, 3).T.copy()


def regularize(cov9, method=REG_PLANE):
    out = _f32(cov9).copy()
    lib().orc_regularize(_p(out, C.c_float), len(out), int(method))
    return out


def estimate_covariances(pts, k=20, method=REG_PLANE, knn_method="kdtree", kernel_width=0.5, max_dist=3.0):
    if knn_method == "rbf":  # calculate_*_covariances_rbf (fast_vgicp_cuda.cu:205-219): RBF-weighted covariance + the same regulariser
        return regularize(covariances_rbf(pts, kernel_width, max_dist), method)
    nbr = knn(pts, k, knn_method)
    return regularize(covariances(pts, nbr), method)


def offsets(method=DIRECT1, radius=-1.0):
    n = lib().orc_offsets(int(method), float(radius), None, 0)
    if n < 0:
        raise ValueError("bad neighbour search method")
    out = np.empty((n, 3), dtype=np.int32)
    lib().orc_offsets(int(method), float(radius), _p(out, C.c_int), n)
    return out


class VoxelMap:
    """GaussianVoxelMap (gaussian_voxelmap.cu) built by the oracle."""

    def __init__(self, pts, cov9, res=1.0, init_buckets=8192, max_scan=10, accum_double=False):
        """cov9=None builds the NDT map (points only, MIN_EIG-regularised voxel covariances)."""
        self.pts = _f32(pts)
        L = lib()
        if cov9 is None:
            self.cov = None
            self._h = L.orc_ndt_voxelmap_build(_p(self.pts, C.c_float), len(self.pts), C.c_float(res), init_buckets, max_scan, int(accum_double))
        else:
            self.cov = _f32(cov9)
            self._h = L.orc_voxelmap_build(_p(self.pts, C.c_float), _p(self.cov, C.c_float), len(self.pts), C.c_float(res), init_buckets, max_scan, int(accum_double))
        self.res = float(np.float32(res))
        self.num_buckets = L.orc_voxelmap_num_buckets(self._h)
        self.num_voxels = L.orc_voxelmap_num_voxels(self._h)
        B, V = self.num_buckets, self.num_voxels
        self.bucket_coord = np.empty((B, 3), dtype=np.int32)
        self.bucket_id = np.empty(B, dtype=np.int32)
        self.vox_n = np.empty(V, dtype=np.int32)
        self.vox_mean = np.empty((V, 3), dtype=np.float32)
        self.vox_cov = np.empty((V, 9), dtype=np.float32)
        L.orc_voxelmap_get(self._h, _p(self.bucket_coord, C.c_int), _p(self.bucket_id, C.c_int), _p(self.vox_n, C.c_int), _p(self.vox_mean, C.c_float), _p(self.vox_cov, C.c_float))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_voxelmap_free(self._h)
            self._h = None

    def as_dict(self):
        """coord(tuple) -> (n, mean(3), cov(9)) keyed independent of voxel ids."""
        out = {}
        occ = np.flatnonzero(self.bucket_id >= 0)
        for b in occ:
            v = self.bucket_id[b]
            out[tuple(int(x) for x in self.bucket_coord[b])] = (int(self.vox_n[v]), self.vox_mean[v].copy(), self.vox_cov[v].copy())
        return out


def _pose_f32(T):
    return np.ascontiguousarray(np.asarray(T, dtype=np.float64).astype(np.float32).T).reshape(16)


def transform_points(T, pts):
    pts = _f32(pts)
    Tf = _pose_f32(T)
    out = np.empty_like(pts)
    lib().orc_transform_points(_p(Tf, C.c_float), _p(pts, C.c_float), len(pts), _p(out, C.c_float))
    return out


def find_correspondences(vmap, src, T, offs):
    src, offs = _f32(src), _i32(offs)
    cap = len(src) * len(offs)
    pairs = np.empty((max(cap, 1), 2), dtype=np.int32)
    Tf = _pose_f32(T)
    n = lib().orc_find_correspondences(vmap._h, _p(src, C.c_float), len(src), _p(Tf, C.c_float), _p(offs, C.c_int), len(offs), _p(pairs, C.c_int), cap)
    return pairs[:n].copy()


def evaluate(vmap, src, src_cov, offs, T_lin, T_eval, want_H=True, sum_float=False):
    """update_correspondences(T_lin) + compute_error(T_eval,H,b): returns (err, H(6,6), b(6), n_corr)."""
    src, src_cov, offs = _f32(src), _f32(src_cov), _i32(offs)
    H = np.zeros(36)
    b = np.zeros(6)
    nc = C.c_long(0)
    e = lib().orc_evaluate_f32(vmap._h, _p(src, C.c_float), _p(src_cov, C.c_float), len(src), _p(offs, C.c_int), len(offs), _p(_pose_in(T_lin), C.c_double),
                               _p(_pose_in(T_eval), C.c_double), _p(H, C.c_double) if want_H else None, _p(b, C.c_double) if want_H else None, int(sum_float), C.byref(nc))
    return e, H.reshape(6, 6).T.copy(), b, nc.value


class AlignResult:
    def __init__(self, r):
        self.T = _pose_out(r.T)
        self.H = np.array(r.H).reshape(6, 6).T.copy()
        self.iterations, self.converged = r.iterations, bool(r.converged)
        self.n_linearize, self.n_error = r.n_linearize, r.n_error


def align_f32(vmap, src, src_cov, offs, guess=None, params=None, sum_float=False):
    src, src_cov, offs = _f32(src), _f32(src_cov), _i32(offs)
    params = params or default_params()
    g = _pose_in(np.eye(4) if guess is None else guess)
    r = LsqResult()
    lib().orc_align_f32(vmap._h, _p(src, C.c_float), _p(src_cov, C.c_float), len(src), _p(offs, C.c_int), len(offs), C.byref(params), _p(g, C.c_double), int(sum_float), C.byref(r))
    return AlignResult(r)


P2D, D2D = 0, 1  # ndt_settings.hpp:6


def _ndt_source(source, res, mode, accum_double):
    if mode == P2D:
        return _f32(source), None, None
    sm = VoxelMap(source, None, res, accum_double=accum_double)
    return sm.vox_mean.copy(), sm.vox_cov.copy(), sm


def evaluate_ndt(vmap, src, src_cov, offs, T_lin, T_eval, want_H=True, sum_float=False):
    src, offs = _f32(src), _i32(offs)
    cov_p = _p(_f32(src_cov), C.c_float) if src_cov is not None else None
    H = np.zeros(36)
    b = np.zeros(6)
    nc = C.c_long(0)
    e = lib().orc_evaluate_ndt(vmap._h, _p(src, C.c_float), cov_p, len(src), _p(offs, C.c_int), len(offs), _p(_pose_in(T_lin), C.c_double), _p(_pose_in(T_eval), C.c_double),
                               _p(H, C.c_double) if want_H else None, _p(b, C.c_double) if want_H else None, int(sum_float), C.byref(nc))
    return e, H.reshape(6, 6).T.copy(), b, nc.value


def register_ndt(target, source, res=1.0, mode=D2D, method=DIRECT7, radius=-1.0, guess=None, params=None, accum_double=True):
    """Whole NDTCuda registration (ndt_cuda_impl.hpp:70-90) with the float oracle."""
    tm = VoxelMap(target, None, res, accum_double=accum_double)
    src, src_cov, _keep = _ndt_source(source, res, mode, accum_double)
    offs = _i32(offsets(method, radius))
    params = params or default_params()
    g = _pose_in(np.eye(4) if guess is None else guess)
    r = LsqResult()
    lib().orc_align_ndt(tm._h, _p(src, C.c_float), _p(_f32(src_cov), C.c_float) if src_cov is not None else None, len(src), _p(offs, C.c_int), len(offs), C.byref(params),
                        _p(g, C.c_double), 0, C.byref(r))
    return AlignResult(r)


def symmetrized(cov9):
    """(C + C^T)/2 in float: the packed 24-byte covariance the CUDA path stores (V L V^-1 is symmetric only up to rounding)."""
    m = _f32(cov9).reshape(-1, 3, 3)
    return (np.float32(0.5) * (m + m.transpose(0, 2, 1))).reshape(-1, 9)


def register_f32(target, source, k=20, reg=REG_PLANE, res=1.0, method=DIRECT1, radius=-1.0, guess=None, params=None, knn_method="kdtree", accum_double=False,
                 kernel_width=0.5, max_dist=3.0, symmetrize=False):
    """Whole FastVGICPCuda registration (setInputTarget + setInputSource + align) with the float oracle.  knn_method "rbf" =
    NearestNeighborMethod::GPU_RBF_KERNEL; symmetrize=True feeds the symmetric part of the regularised covariances (what the CUDA
    path stores)."""
    tc = estimate_covariances(target, k, reg, knn_method, kernel_width, max_dist)
    sc = estimate_covariances(source, k, reg, knn_method, kernel_width, max_dist)
    if symmetrize:
        tc, sc = symmetrized(tc), symmetrized(sc)
    vm = VoxelMap(target, tc, res, accum_double=accum_double)
    return align_f32(vm, source, sc, offsets(method, radius), guess, params)


def covariances_f64(pts, k=20, reg=REG_PLANE, threads=0):
    pts = _f32(pts)
    out = np.empty((len(pts), 9), dtype=np.float64)
    rc = lib().orc64_covariances(_p(pts, C.c_float), len(pts), k, int(reg), _p(out, C.c_double), threads)
    if rc:
        raise ValueError("bad k")
    return out


def align_f64(target, tgt_cov, source, src_cov, res=1.0, offs=None, guess=None, params=None, threads=0):
    target, source = _f32(target), _f32(source)
    tgt_cov = np.ascontiguousarray(tgt_cov, dtype=np.float64)
    src_cov = np.ascontiguousarray(src_cov, dtype=np.float64)
    offs = _i32(offsets(DIRECT1) if offs is None else offs)
    params = params or default_params()
    g = _pose_in(np.eye(4) if guess is None else guess)
    r = LsqResult()
    lib().orc64_align(_p(target, C.c_float), _p(tgt_cov, C.c_double), len(target), _p(source, C.c_float), _p(src_cov, C.c_double), len(source), float(res), _p(offs, C.c_int),
                      len(offs), C.byref(params), _p(g, C.c_double), threads, C.byref(r))
    return AlignResult(r)


def align_gicp_f64(target, tgt_cov, source, src_cov, max_corr_dist=-1.0, guess=None, params=None, threads=1):
    """FastGICP / FastGICPSingleThread (fast_gicp_impl.hpp:117-240) restated in double: BASELINE config 1 (CPU timing row)."""
    target, source = _f32(target), _f32(source)
    tgt_cov = np.ascontiguousarray(tgt_cov, dtype=np.float64)
    src_cov = np.ascontiguousarray(src_cov, dtype=np.float64)
    params = params or default_params()
    g = _pose_in(np.eye(4) if guess is None else guess)
    r = LsqResult()
    lib().orc64_align_gicp(_p(target, C.c_float), _p(tgt_cov, C.c_double), len(target), _p(source, C.c_float), _p(src_cov, C.c_double), len(source), float(max_corr_dist),
                           C.byref(params), _p(g, C.c_double), int(threads), C.byref(r))
    return AlignResult(r)


def remove_near_origin(pts):
    """src/align.cpp:128-133: drop points with squaredNorm() < 1e-3 (stable)."""
    p = _f32(np.asarray(pts)[:, :3])
    out = np.empty_like(p)
    m = lib().orc_remove_near_origin(_p(p, C.c_float), len(p), _p(out, C.c_float))
    return out[:m].copy()


def approximate_voxel_grid(pts, leaf, histsize=512):
    """pcl::ApproximateVoxelGrid<PointXYZ> restated (align.cpp:136-147, kitti.cpp:80-82, python/main.cpp:46-62)."""
    p = _f32(np.asarray(pts)[:, :3])
    out = np.empty_like(p)
    m = lib().orc_approximate_voxel_grid(_p(p, C.c_float), len(p), C.c_float(leaf), int(histsize), _p(out, C.c_float))
    return out[:m].copy()


def se3_exp(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    T = np.empty(16)
    lib().orc_se3_exp(_p(a, C.c_double), _p(T, C.c_double))
    return T.reshape(4, 4).T.copy()


def ldlt_solve6(A, rhs):
    A = np.ascontiguousarray(np.asarray(A, dtype=np.float64).T).reshape(36)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    x = np.empty(6)
    lib().orc_ldlt_solve6(_p(A, C.c_double), _p(rhs, C.c_double), _p(x, C.c_double))
    return x
