"""GPU parity of the hypothesis scoring between the hot-path calls of a SLAM-mode frame:
camera-pose vote and re-initialisation test, product (one suo_chi2_inlier_counts launch each) vs the numpy oracle.
Counts are integers: they must be IDENTICAL, provided no chi2 of the oracle sits within 1e-4 relative of the 5.991
gate (the reference inverts the covariance in float32, the kernel in FP64 — the only place they can differ)."""
import numpy as np
import pytest

from oracle import slam_oracle
from suo_slam_b200 import slam, synth

pytestmark = pytest.mark.gpu


def _decisive(chi2):
    return not np.any(np.abs(chi2 - 5.991) < 1e-4 * 5.991)


@pytest.mark.parametrize("seed,with_cov", [(3, True), (8, True), (5, False)])
def test_camera_vote_vs_oracle(seed, with_cov):
    sc = synth.make_slam_scene(seed, n_views=5, n_obj=7, bad_pnp=(1, 4), bad_estimate=(), with_cov=with_cov)
    cur = sc["detections"][sc["view_ids"][-1]]
    cur[13]["pose"] = None                                  # PnP failed for one object: it does not vote
    T_ref, c_ref, chi2 = slam_oracle.estimate_camera_pose(sc["obj_poses"], cur)
    assert _decisive(chi2)
    T, c = slam.estimate_camera_pose(sc["obj_poses"], cur, return_counts=True)
    assert np.array_equal(c, c_ref) and len(c) == 6
    np.testing.assert_array_equal(T, T_ref)                 # the same hypothesis, composed by the same numpy expression
    assert c.argmax() not in (1, 3)                         # (index 3 = object 14 after dropping 13): bad votes lose
    assert slam.estimate_camera_pose({}, cur) is None
    # raising the bar above the best count rejects every hypothesis (:1068)
    assert slam.estimate_camera_pose(sc["obj_poses"], cur, min_num_inliers=int(c.max()) + 1) is None


@pytest.mark.parametrize("seed,n_views", [(4, 6), (9, 20)])
def test_reinit_vs_oracle(seed, n_views):
    sc = synth.make_slam_scene(seed, n_views=n_views, n_obj=5, bad_pnp=(), bad_estimate=(2, 3))
    del sc["detections"][sc["view_ids"][1]][11]             # an object missed in one view
    args = (sc["obj_poses"], sc["cam_poses"], sc["detections"], sc["view_ids"], sc["view_ids"][-1])
    new_ref, num_ref, chi2 = slam_oracle.maybe_reinit_objects(*args)
    assert _decisive(chi2)
    new, num = slam.maybe_reinit_objects(*args, return_counts=True)
    assert num == num_ref
    assert sorted(new) == sorted(new_ref) == [12, 13]
    for o in new:
        np.testing.assert_array_equal(new[o], new_ref[o])
    assert slam.maybe_reinit_objects(sc["obj_poses"], sc["cam_poses"], sc["detections"], sc["view_ids"][:1], sc["view_ids"][0]) == {}
