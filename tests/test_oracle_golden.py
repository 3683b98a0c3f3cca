"""CPU tests: pin the oracle (oracle/vgicp_oracle.c) against everything the reference's own tests/fixtures hold for
this path -- data/relative.txt under the four call orders of src/test/gicp_test.cpp:147-201 with its tolerances
(0.05 m / 1 deg / hasConverged) -- plus known-answer checks of the restated primitives.
"""
import numpy as np
import pytest

import oracle as O
from conftest import pose_error

T_TOL, R_TOL = 0.05, np.radians(1.0)  # gicp_test.cpp:148-149


def test_hash_known_answers():
    """vector3_hash.cuh:8-33 evaluated by hand in Python integers (incl. negative coords: int -> uint64 sign extension)."""
    M = 0xC6A4A7935BD1E995
    MASK = (1 << 64) - 1

    def combine(h, k):
        k &= MASK
        k = (k * M) & MASK
        k ^= k >> 47
        k = (k * M) & MASK
        h ^= k
        h = (h * M) & MASK
        return (h + 0xE6546B64) & MASK

    def ref(x, y, z):
        h = 0
        for v in (x, y, z):
            h = combine(h, v & MASK)  # two's complement == sign extension
        return h

    for c in [(0, 0, 0), (1, 2, 3), (-1, -2, -3), (2147483647, -2147483648, 5), (-70, 13, -2)]:
        assert O.vector3i_hash(*c) == ref(*c)
    assert O.vector3i_hash(1, -2, 3) == 0xC2EB94FAB88F7833


def test_voxel_coord_float_semantics():
    """calc_voxel_coord = floor(x/res - 0.5) in float (vector3_hash.cuh:35-38)."""
    pts = np.array([[0.49, 0.5, 0.51], [-0.49, -0.5, -0.51], [1.4999999, 1.5, 1.5000001], [70.3, -12.7, 0.0]], dtype=np.float32)
    for res in (1.0, 0.5, 0.3):
        want = np.floor(pts / np.float32(res) - np.float32(0.5)).astype(np.int32)
        assert np.array_equal(O.voxel_coords(pts, res), want)


def test_offsets_tables():
    """fast_vgicp_cuda.cu:42-95"""
    assert O.offsets(O.DIRECT1).tolist() == [[0, 0, 0]]
    assert O.offsets(O.DIRECT7).tolist() == [[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]
    d27 = O.offsets(O.DIRECT27)
    assert len(d27) == 27 and d27[0].tolist() == [-1, -1, -1] and d27[1].tolist() == [-1, -1, 0] and d27[26].tolist() == [1, 1, 1]
    r = O.offsets(O.DIRECT_RADIUS, 1.5)
    assert len(r) == 19  # |o| <= 1.501 within [-2,2]^3: centre + 6 faces + 12 edges


def test_knn_kdtree_equals_bruteforce(pair02):
    tgt, _ = pair02
    a = O.knn(tgt[:3000], 20, "kdtree")
    b = O.knn(tgt[:3000], 20, "bruteforce")
    assert np.array_equal(a, b)
    assert np.array_equal(a[:, 0], np.arange(3000))


def test_eigensolver_against_lapack(pair02):
    tgt, _ = pair02
    nbr = O.knn(tgt[:2000], 20)
    cov = O.covariances(tgt[:2000], nbr)
    for i in range(0, 2000, 37):
        ev, V = O.eig3_direct(cov[i])
        w = np.linalg.eigvalsh(cov[i].reshape(3, 3).astype(np.float64))
        assert np.abs(ev - w).max() < 2e-4 * max(1.0, np.abs(w).max())
        assert np.abs(V.T.astype(np.float64) @ V - np.eye(3)).max() < 1e-2


def test_plane_regularisation_identity(pair02):
    """C_reg = I - 0.999 n n^T for the direct eigen-solver restatement."""
    tgt, _ = pair02
    nbr = O.knn(tgt[:2000], 20)
    cov = O.covariances(tgt[:2000], nbr)
    reg = O.regularize(cov, O.REG_PLANE)
    for i in range(0, 2000, 41):
        _, V = O.eig3_direct(cov[i])
        n = V[:, 0].astype(np.float64)
        want = np.eye(3) - 0.999 * np.outer(n, n)
        assert np.abs(reg[i].reshape(3, 3).T - want).max() < 2e-3


def test_se3_exp_and_ldlt():
    rng = np.random.default_rng(0)
    for _ in range(20):
        a = rng.normal(size=6) * 0.3
        T = O.se3_exp(a)
        assert np.abs(T[:3, :3] @ T[:3, :3].T - np.eye(3)).max() < 1e-12
        # scipy cross-check of the rotation and of V*t
        th = np.linalg.norm(a[:3])
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        R = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * K @ K
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * K + (th - np.sin(th)) / th**3 * K @ K
        assert np.abs(T[:3, :3] - R).max() < 1e-12 and np.abs(T[:3, 3] - V @ a[3:]).max() < 1e-12
        A = rng.normal(size=(6, 6))
        A = A @ A.T + 1e-3 * np.eye(6)
        rhs = rng.normal(size=6)
        assert np.abs(O.ldlt_solve6(A, rhs) - np.linalg.solve(A, rhs)).max() < 1e-9
    assert np.array_equal(O.se3_exp(np.zeros(6)), np.eye(4))


@pytest.fixture(scope="module")
def covs(pair02):
    tgt, src = pair02
    return O.estimate_covariances(tgt), O.estimate_covariances(src)


@pytest.mark.parametrize("method", [O.DIRECT1, O.DIRECT7, O.DIRECT27])
def test_oracle_f32_reference_scenarios(pair02, relative_pose, covs, method):
    """gicp_test.cpp:157-200 scenarios on the float oracle (the swap scenarios only reorder the same stage calls, so
    forward and backward cover the numerics; the state machine itself is tested on the product wrapper)."""
    tgt, src = pair02
    tc, sc = covs
    fwd = O.align_f32(O.VoxelMap(tgt, tc), src, sc, O.offsets(method))
    e = pose_error(relative_pose, fwd.T)
    assert fwd.converged and e[0] < T_TOL and e[1] < R_TOL
    bwd = O.align_f32(O.VoxelMap(src, sc), tgt, tc, O.offsets(method))
    e = pose_error(relative_pose, np.linalg.inv(bwd.T))
    assert bwd.converged and e[0] < T_TOL and e[1] < R_TOL


def test_oracle_f64_twin_agrees(pair02, relative_pose, covs):
    """The double CPU twin (FastVGICP restated) lands on the same optimum as the float CUDA-path oracle."""
    tgt, src = pair02
    tc, sc = covs
    f32 = O.align_f32(O.VoxelMap(tgt, tc), src, sc, O.offsets(O.DIRECT1))
    r64 = O.align_f64(tgt, O.covariances_f64(tgt), src, O.covariances_f64(src))
    assert r64.converged
    e = pose_error(relative_pose, r64.T)
    assert e[0] < T_TOL and e[1] < R_TOL
    d = pose_error(f32.T, r64.T)
    assert d[0] < 1e-3 and d[1] < 1e-4


def test_voxelmap_invariants(pair02, covs):
    tgt, _ = pair02
    tc, _ = covs
    vm = O.VoxelMap(tgt, tc, 1.0)
    assert vm.num_buckets == 8192  # ~1.1k voxels fit the initial table (gaussian_voxelmap.cuh:20)
    assert vm.vox_n.sum() == len(tgt)  # nothing dropped
    coords = O.voxel_coords(tgt, 1.0)
    uniq = np.unique(coords, axis=0)
    assert vm.num_voxels == len(uniq)
    d = vm.as_dict()
    key = tuple(int(x) for x in coords[0])
    sel = (coords == coords[0]).all(axis=1)
    assert d[key][0] == sel.sum()
    assert np.abs(d[key][1] - tgt[sel].mean(axis=0)).max() < 1e-4


@pytest.mark.parametrize("mode,method", [(O.D2D, O.DIRECT7), (O.D2D, O.DIRECT1), (O.P2D, O.DIRECT1)])
def test_oracle_ndt_reference_gate(pair02, relative_pose, mode, method):
    """NDT_CUDA rows of gicp_test.cpp (defaults D2D + DIRECT7, ndt_cuda.cu:21-22): 0.05 m / 1 deg / hasConverged on the NDT oracle."""
    tgt, src = pair02
    fwd = O.register_ndt(tgt, src, mode=mode, method=method)
    e = pose_error(relative_pose, fwd.T)
    assert fwd.converged and e[0] < T_TOL and e[1] < R_TOL, e
    bwd = O.register_ndt(src, tgt, mode=mode, method=method)
    e = pose_error(relative_pose, np.linalg.inv(bwd.T))
    assert bwd.converged and e[0] < T_TOL and e[1] < R_TOL, e


def test_oracle_ndt_voxel_covariances_min_eig(pair02):
    tgt, _ = pair02
    vm = O.VoxelMap(tgt, None, 1.0, accum_double=True)
    w = np.linalg.eigvalsh(0.5 * (vm.vox_cov.reshape(-1, 3, 3) + vm.vox_cov.reshape(-1, 3, 3).transpose(0, 2, 1)).astype(np.float64))
    assert w.min() > 0.99e-3
    # a populated voxel: covariance equals the sample covariance of its points (population normalisation, :196)
    coords = O.voxel_coords(tgt, 1.0)
    v = int(np.argmax(vm.vox_n))
    b = int(np.flatnonzero(vm.bucket_id == v)[0])
    sel = (coords == vm.bucket_coord[b]).all(axis=1)
    P = tgt[sel].astype(np.float64)
    want = np.cov(P.T, bias=True)
    assert np.abs(vm.vox_cov[v].reshape(3, 3) - want).max() < 1e-3


def test_rbf_covariances_against_direct_evaluation(pair02, relative_pose):
    """covariance_estimation_rbf.cu:59-151 restated (orc_covariances_rbf): against the defining formula evaluated directly in
    double -- kernel-weighted mean/covariance over the points within max_dist, with the reference's quirk that the cloud is padded
    to a multiple of 512 with points at the origin (:126-129) which add to the weight sum of queries near the origin -- and end to
    end: a registration with RBF covariances meets the reference's gate against data/relative.txt."""
    rng = np.random.default_rng(2)
    pts = (rng.normal(size=(1500, 3)) * [3.0, 3.0, 0.2]).astype(np.float32)
    kw, md = 0.5, 3.0
    got = O.covariances_rbf(pts, kw, md).reshape(-1, 3, 3).transpose(0, 2, 1)  # column-major image -> [row, col]
    P = pts.astype(np.float64)
    npad = (-len(P)) % 512
    assert npad > 0
    ext = np.vstack([P, np.zeros((npad, 3))])
    near_origin = 0
    for i in range(0, len(P), 37):
        d2 = ((ext - P[i]) ** 2).sum(axis=1)
        w = np.where(d2 <= md * md, np.exp(-kw * d2), 0.0)
        near_origin += int(w[len(P):].sum() > 0)
        sw = w.sum()
        s1 = (w[:, None] * ext).sum(axis=0)
        s2 = (w[:, None, None] * ext[:, :, None] * ext[:, None, :]).sum(axis=0)
        want = (s2 - np.outer(s1 / sw, s1)) / sw
        assert np.abs(got[i] - want).max() < 2e-5 * max(1.0, np.abs(want).max()), i
    assert near_origin > 0  # the padding quirk is exercised
    tgt, src = pair02
    r = O.register_f32(tgt, src, method=O.DIRECT1, knn_method="rbf")
    dt, dr = pose_error(relative_pose, r.T)
    assert r.converged and dt < T_TOL and dr < R_TOL, (dt, dr)


def test_fastgicp_restatement_meets_the_reference_gate(pair02, relative_pose):
    """FastGICP (fast_gicp_impl.hpp:117-240; BASELINE config 1, the single-thread CPU row): the restatement converges to
    data/relative.txt within the reference test's tolerance (src/test/gicp_test.cpp:147-201 runs the same gate on GICP)."""
    tgt, src = pair02
    tc = O.covariances_f64(tgt, 20, O.REG_PLANE, 1)
    sc = O.covariances_f64(src, 20, O.REG_PLANE, 1)
    r = O.align_gicp_f64(tgt, tc, src, sc, threads=1)
    dt, dr = pose_error(relative_pose, r.T)
    assert r.converged and dt < T_TOL and dr < R_TOL, (dt, dr)
