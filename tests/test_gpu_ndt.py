"""NDTCuda (BASELINE config 5) on the same engine: GPU vs the oracle's restatement of ndt_cuda.cu /
ndt_compute_derivatives.cu, and the reference's alignment gate for NDT_CUDA (src/test/gicp_test.cpp:118-123,147-201)."""
import numpy as np
import pytest

import oracle as O
from conftest import pose_error, random_pose

pytestmark = pytest.mark.gpu


def sym(c9):
    m = c9.reshape(-1, 3, 3)
    return (0.5 * (m + m.transpose(0, 2, 1))).reshape(-1, 9)


@pytest.mark.parametrize("res", [1.0, 0.5])
def test_ndt_voxelmap_matches_oracle(pair02, res):
    """Points-only voxel Gaussians (gaussian_voxelmap.cu:122-148,178-198) + MIN_EIG (ndt_cuda.cu:129,140): same table, same counts,
    means/covariances to float rounding (sums in double on both sides)."""
    from fast_gicp_b200.core import Core

    tgt, src = pair02
    c = Core(0)
    c.set_problem(2)
    c.set_resolution(res)
    c.set_target_cloud(tgt)
    c.set_source_cloud(src)
    c.ndt_create_voxelmaps()
    vm = O.VoxelMap(tgt, None, res, accum_double=True)
    assert c.num_buckets() == vm.num_buckets and c.num_voxels() == vm.num_voxels
    coords, ids = c.get_voxel_buckets()
    assert np.array_equal(ids, vm.bucket_id) and np.array_equal(coords, vm.bucket_coord)
    assert np.array_equal(c.get_voxel_num_points(), vm.vox_n)
    # the points of a voxel are added in index order in double on both sides and the regulariser is the same restated eigen-solver:
    # bit for bit, including the rank-deficient few-point voxels where the closed-form solver amplifies any input difference
    assert np.array_equal(c.get_voxel_means(), vm.vox_mean)
    bad = np.flatnonzero((c.get_voxel_covs() != sym(vm.vox_cov).astype(np.float32)).any(axis=1))
    assert len(bad) <= 1, (len(bad), np.abs(c.get_voxel_covs() - sym(vm.vox_cov)).max())  # (one double-rounding tie of the trig step allowed)
    w = np.linalg.eigvalsh(c.get_voxel_covs().reshape(-1, 3, 3).astype(np.float64))
    assert w.min() > 0.99e-3  # eigenvalues clamped at 1e-3 (covariance_regularization.cu:84-101)
    c.close()


@pytest.mark.parametrize("problem,mode", [(1, O.P2D), (2, O.D2D)])
@pytest.mark.parametrize("method", [O.DIRECT1, O.DIRECT7])
def test_ndt_linear_system_matches_oracle(pair02, relative_pose, problem, mode, method):
    from fast_gicp_b200.core import Core

    tgt, src = pair02
    c = Core(0)
    c.set_problem(problem)
    c.set_neighbor_search_method(method)
    c.set_target_cloud(tgt)
    c.set_source_cloud(src)
    c.ndt_create_voxelmaps()
    tm = O.VoxelMap(tgt, None, 1.0, accum_double=True)
    if mode == O.D2D:
        sm = O.VoxelMap(src, None, 1.0, accum_double=True)
        s_pts, s_cov = sm.vox_mean, sym(sm.vox_cov).astype(np.float32)
    else:
        s_pts, s_cov = src, None
    offs = O.offsets(method)
    rng = np.random.default_rng(1)
    for T in (np.eye(4), relative_pose):
        c.update_correspondences(T)
        pairs = c.get_voxel_correspondences()
        want = O.find_correspondences(tm, s_pts, T, offs)
        assert np.array_equal(pairs, want)
        for Te in (T, T @ random_pose(rng, 0.004, 0.04)):
            err, H, b = c.compute_error(Te, True)
            e0, H0, b0, _ = O.evaluate_ndt(tm, s_pts, s_cov, offs, T, Te, True)
            # the few ill-conditioned voxel covariances (see the map test) enter both sides with tiny weight differences
            assert abs(err - e0) <= 2e-3 * abs(e0)
            assert np.abs(H - H0).max() <= 2e-3 * np.abs(H0).max()
            assert np.abs(b - b0).max() <= 2e-3 * max(np.abs(b0).max(), 1e-3 * np.abs(H0).max())
            e1, _, _ = c.compute_error(Te, False)
            assert abs(e1 - err) <= 1e-6 * abs(err)
    c.close()


@pytest.mark.parametrize("mode", ["P2D", "D2D"])
def test_ndt_alignment_reference_scenarios(pair02, relative_pose, mode):
    """gicp_test.cpp scenarios for NDT_CUDA (defaults: D2D, DIRECT7) + P2D/DIRECT1, gate 0.05 m / 1 deg / hasConverged."""
    from fast_gicp_b200 import NDTCuda, NDTDistanceMode

    target, source = pair02
    t_tol, r_tol = 0.05, np.radians(1.0)

    def make():
        reg = NDTCuda()
        if mode == "P2D":
            reg.setDistanceMode(NDTDistanceMode.P2D)
            reg.setNeighborSearchMethod("DIRECT1")
        return reg

    reg = make()
    reg.setInputTarget(target)
    reg.setInputSource(source)
    T = reg.align()
    e = pose_error(relative_pose, T)
    assert e[0] < t_tol and e[1] < r_tol and reg.hasConverged(), "FORWARD TEST"
    ref = O.register_ndt(target, source, mode=O.P2D if mode == "P2D" else O.D2D, method=O.DIRECT1 if mode == "P2D" else O.DIRECT7)
    d = pose_error(ref.T, T)
    assert d[0] < 2e-3 and d[1] < 2e-4, d  # same optimum as the oracle (ill-conditioned few-point voxels differ slightly)

    reg.setInputTarget(source)
    reg.setInputSource(target)
    T = reg.align()
    e = pose_error(relative_pose, np.linalg.inv(T.astype(np.float64)))
    assert e[0] < t_tol and e[1] < r_tol and reg.hasConverged(), "BACKWARD TEST"

    reg = make()
    reg.setInputSource(target)
    reg.swapSourceAndTarget()
    reg.setInputSource(source)
    T = reg.align()
    e = pose_error(relative_pose, T)
    assert e[0] < t_tol and e[1] < r_tol and reg.hasConverged(), "SWAP AND SET SOURCE TEST"

    reg = make()
    reg.setInputTarget(source)
    reg.swapSourceAndTarget()
    reg.setInputTarget(target)
    T = reg.align()
    e = pose_error(relative_pose, T)
    assert e[0] < t_tol and e[1] < r_tol and reg.hasConverged(), "SWAP AND SET TARGET TEST"


def test_vgicp_and_ndt_share_a_handle(pair02):
    """Switching the problem on one handle does not leak state between the VGICP and NDT maps."""
    from fast_gicp_b200.core import REG_PLANE, Core, pose_from_c

    tgt, src = pair02
    c = Core(0)
    c.set_target_cloud(tgt)
    c.find_target_neighbors(20)
    c.calculate_target_covariances(REG_PLANE)
    c.create_target_voxelmap()
    c.set_source_cloud(src)
    c.find_source_neighbors(20)
    c.calculate_source_covariances(REG_PLANE)
    a = pose_from_c(c.align().T)
    c.set_problem(2)
    c.set_neighbor_search_method("DIRECT7")
    n = pose_from_c(c.align().T)
    c.set_problem(0)
    c.set_neighbor_search_method("DIRECT1")
    a2 = pose_from_c(c.align().T)
    assert np.abs(a - a2).max() < 1e-9
    assert np.abs(a - n).max() < 0.05
    c.close()
