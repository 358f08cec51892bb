"""Kinematic initialisation through files (kinematic_optimizer.py:30-224) on a synthetic video directory, and its hand-over to
prepare_input -- the reference's chain run_phys_mocap.py:95-153 up to the physics solve (which needs the GPU: see
tests/test_plumbing_gpu.py for the rest)."""
import os

import numpy as np


def test_kinematic_driver_to_phys_inputs(chd, tmp_path):
    F = 30
    vd = str(tmp_path / "walk")
    gt = chd.synth.write_mocap_clip(vd, F, seed=2)
    assert len(os.listdir(os.path.join(vd, "openpose_result"))) == F
    tc = chd.kinopt.load_totalcap_results(os.path.join(vd, "tracked_results.json"))
    assert tc["joint3d"].shape == (F, 25, 3) and tc["smpl_joint_angles"].shape == (F, 22, 3)
    np.testing.assert_allclose(tc["root_trans"], gt["root"], atol=1e-12)
    kin = os.path.join(vd, "kinematic_results")
    res = chd.kinopt.optimize_2d_3d(os.path.join(vd, "walk.mp4"), os.path.join(vd, "skeleton.bvh"), kin, 0, F)
    anim, new3d, proj, pn, pp, vel, info = res
    # floor: y-down camera frame, the clip's feet stand on y = floor_y
    assert pn[1] < -0.99
    feet = gt["joints"][:, [4, 5, 10, 11]].reshape(-1, 3)
    y_plane = pp[1] - (pn[0] * (feet[:, 0] - pp[0]) + pn[2] * (feet[:, 2] - pp[2])) / pn[1]     # plane height under the feet
    assert np.abs(y_plane - gt["floor_y"]).max() < 3.0
    fl = open(os.path.join(kin, "floor_out.txt")).read()
    assert len(fl.splitlines()) == 2 and not fl.endswith("\n")
    fc = np.load(os.path.join(kin, "foot_contacts.npy"))
    assert fc.shape == (F, 4) and fc.dtype.kind == "i" and set(np.unique(fc)) <= {0, 1}
    # the refined pose is closer to the truth than the noisy input was allowed to drift, and re-projects within a few px
    err = np.linalg.norm(new3d - gt["joints"][:, chd.kinopt.BACKWARD], axis=-1).mean()
    assert err < 4.0
    b = chd.prepare.load_bvh(os.path.join(kin, "final_test.bvh"))
    assert b.n_frames == F and len(b.names) == 28 and b.names == chd.prepare.COMBINED_NAMES
    R, T = chd.prepare.local_transforms(b)
    gp, _ = chd.prepare.forward_kinematics(b.parents, R, T)
    np.testing.assert_allclose(gp[:, chd.kinopt.BACKWARD], new3d, atol=1e-3)       # BVH carries six decimals
    # contact feet barely move in the result (cm per frame)
    toe = gp[:, 5]
    planted = (fc[1:, 1] == 1) & (fc[:-1, 1] == 1)
    assert np.linalg.norm(np.diff(toe, axis=0), axis=1)[planted].mean() < 0.5
    # hand-over: the four phys_optim input files from the kinematic result
    pin = str(tmp_path / "phys_in")
    p = chd.prepare.prepare_input(os.path.join(kin, "final_test.bvh"), os.path.join(kin, "floor_out.txt"), os.path.join(kin, "foot_contacts.npy"),
                                  pin, chd.prepare.combined_info(), 0, F, 1.0 / 30.0, False)
    q = chd.io_formats.read_phys_inputs(pin, F)
    assert q.n_frames == F and q.n_ee == 4 and q.floor_normal[2] > 0.99
    np.testing.assert_allclose(q.base_lin, p.base_lin, rtol=0, atol=1e-12)
    com_h = (q.base_lin - q.floor_point) @ q.floor_normal
    assert 0.7 < com_h.min() and com_h.max() < 1.2                                   # metres above the fitted floor
    for e in range(4):
        assert abs(sum(q.ee_durations[e]) - (F - 1) / 30.0) < 1e-9


def test_retargeted_character_to_phys_inputs(chd, tmp_path):
    """run_phys_mocap.py:117-153 for a Mixamo-style character: kinematic result -> re-targeting -> prepare_input with the
    character's tables (heel joints added on the fly)."""
    F = 24
    vd = str(tmp_path / "walk")
    chd.synth.write_mocap_clip(vd, F, seed=3)
    kin = os.path.join(vd, "kinematic_results")
    chd.kinopt.optimize_2d_3d(os.path.join(vd, "walk.mp4"), os.path.join(vd, "skeleton.bvh"), kin, 0, F)
    skel = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "towr", "retarget", "ybot_skel.bvh")
    info = chd.prepare.ybot_info()
    out = os.path.join(kin, "ybot_out.bvh")
    anim = chd.results.retarget(os.path.join(kin, "final_test.bvh"), skel, info, out, iterations=60)
    assert len(anim.names) == 67 and anim.rotations.shape[0] == F
    src = chd.prepare.load_bvh(os.path.join(kin, "final_test.bvh"))
    gs = chd.prepare.forward_kinematics(src.parents, *chd.prepare.local_transforms(src))[0]
    ga = anim.global_positions()
    # the character walks where the source walks (x / z of the hips are not scaled) and its feet follow the source's feet
    np.testing.assert_allclose(ga[:, 0, [0, 2]], gs[:, 0, [0, 2]], atol=2.0)       # cm: the IK may shift the root a little
    assert np.linalg.norm(ga[:, 65] - ga[:, 60], axis=1).max() < 80.0
    pin = str(tmp_path / "phys_in_ybot")
    p = chd.prepare.prepare_input(out, os.path.join(kin, "floor_out.txt"), os.path.join(kin, "foot_contacts.npy"), pin, info, 0, F, 1.0 / 30.0, False)
    q = chd.io_formats.read_phys_inputs(pin, F)
    assert q.n_ee == 4 and q.n_frames == F and 0.7 < q.max_leg_length < 1.2
    com_h = (q.base_lin - q.floor_point) @ q.floor_normal
    assert 0.6 < com_h.min() and com_h.max() < 1.4


def test_contact_label_mapping_round_trip(chd):
    rng = np.random.default_rng(0)
    fc = rng.integers(0, 2, (20, 4))
    vel = chd.kinopt.contacts_to_constraints(fc)
    assert vel.shape == (20, 28) and vel[:, :19].sum() == 0 and vel[:, 25:].sum() == 0
    np.testing.assert_array_equal(chd.kinopt.constraints_to_contacts(vel), fc)
