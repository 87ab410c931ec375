"""CPU: the hypothesis-scoring oracle (oracle/slam_oracle.py) behaves as the reference describes:
the camera-pose vote picks a hypothesis from a well-localised object, the re-init test flags the badly mapped one."""
import numpy as np

from oracle import slam_oracle
from suo_slam_b200 import synth


def test_camera_vote_ignores_the_bad_pnp_pose():
    sc = synth.make_slam_scene(3, n_views=5, n_obj=6, bad_pnp=(1,), bad_estimate=())
    cur = sc["view_ids"][-1]
    T, counts, _ = slam_oracle.estimate_camera_pose(sc["obj_poses"], sc["detections"][cur])
    assert T is not None and counts.argmax() != 1 and counts[1] < 0.5 * counts.max()
    Tgt = sc["cam_poses"][cur]
    assert np.abs(T[:3, :3] - Tgt[:3, :3]).max() < 2e-2 and np.abs(T[:3, 3] - Tgt[:3, 3]).max() < 15.0
    none, _, _ = slam_oracle.estimate_camera_pose({}, sc["detections"][cur])
    assert none is None


def test_reinit_flags_the_badly_mapped_object():
    sc = synth.make_slam_scene(4, n_views=6, n_obj=5, bad_pnp=(), bad_estimate=(2,))
    new, num, _ = slam_oracle.maybe_reinit_objects(sc["obj_poses"], sc["cam_poses"], sc["detections"], sc["view_ids"], sc["view_ids"][-1])
    assert list(new) == [12] and num[12]["pnp"] > 3 * num[12]["estim"]
    assert all(num[o]["estim"] > num[o]["pnp"] * 0.5 for o in num if o != 12)
    assert slam_oracle.maybe_reinit_objects(sc["obj_poses"], sc["cam_poses"], sc["detections"], sc["view_ids"][:1], sc["view_ids"][0])[0] == {}
