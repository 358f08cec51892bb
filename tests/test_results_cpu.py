"""prepare_input / load_results / apply_results against golden vectors produced by the reference's OWN functions
(tests/golden/make_towr_golden.py imports towr_utils.py and the BVH / Animation / IK library from the reference tree).
Tolerances: the reference normalises rotation axes with `axis / (|axis| + 1e-10)`, so its own output carries 1e-10-level
noise; 30 IK iterations amplify that to ~1e-8 cm."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "towr")
CASES = {"combined": (3, 41, False, 30.0), "ybot": (2, 38, False, 24.0), "ybot_noheel": (0, 36, True, 30.0)}


def qmat(q):
    w, x, y, z = [q[..., i] for i in range(4)]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0], R[..., 0, 1], R[..., 0, 2] = 1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)
    R[..., 1, 0], R[..., 1, 1], R[..., 1, 2] = 2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)
    R[..., 2, 0], R[..., 2, 1], R[..., 2, 2] = 2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)
    return R


@pytest.mark.parametrize("case", list(CASES))
def test_prepare_input_matches_reference_files(chd, case, tmp_path):
    s0, s1, comb, fps = CASES[case]
    d = os.path.join(G, case)
    info = chd.prepare.CHARACTERS[case.split("_")[0]]()
    chd.prepare.prepare_input(d + "/anim.bvh", d + "/floor.txt", d + "/foot_contacts.npy", str(tmp_path), info, s0, s1, 1.0 / fps, comb)
    for f in ("skel_info.txt", "motion_info.txt", "terrain_info.txt", "contact_info.txt"):
        got = np.array(open(os.path.join(str(tmp_path), f)).read().split(), dtype=np.float64)
        ref = np.array(open(os.path.join(d, "phys_in", f)).read().split(), dtype=np.float64)
        assert got.shape == ref.shape, f
        np.testing.assert_allclose(got, ref, rtol=0, atol=2e-9, err_msg=f)
        # same line structure (the reader of phys_optim.cpp:155-267 is whitespace driven, towr's tests are not: keep it anyway)
        assert len(open(os.path.join(str(tmp_path), f)).read().splitlines()) == len(open(os.path.join(d, "phys_in", f)).read().splitlines())


@pytest.mark.parametrize("case", list(CASES))
def test_load_results_matches_reference(chd, case):
    d = os.path.join(G, case)
    g = np.load(d + "/results.npz")
    r = chd.results.load_towr_results(d + "/sol_out.txt")
    assert r.num_feet == int(g["num_feet"]) and r.dt == float(g["dt"])
    for k in ("base_pos", "feet_pos", "feet_force"):
        np.testing.assert_array_equal(getattr(r, k), g[k])
    np.testing.assert_array_equal(r.feet_contact, g["feet_contact"])
    np.testing.assert_allclose(r.base_rot, g["base_rot"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(r.base_R, g["base_R"], rtol=0, atol=1e-9)


def test_load_results_without_flip_swaps_axis(chd):
    d = os.path.join(G, "combined")
    r = chd.results.load_towr_results(d + "/sol_out.txt", flip_coords=False)
    rf = chd.results.load_towr_results(d + "/sol_out.txt", flip_coords=True)
    np.testing.assert_array_equal(r.base_pos, -rf.base_pos)
    # rotation about the y/z-swapped axis by the same angle: equals the flipped result with the axis negated, i.e. transposed
    np.testing.assert_allclose(r.base_R, np.swapaxes(rf.base_R, -1, -2), atol=1e-12)


@pytest.mark.parametrize("case", list(CASES))
def test_apply_results_matches_reference_ik(chd, case):
    s0, s1, comb, fps = CASES[case]
    d = os.path.join(G, case)
    a = np.load(d + "/applied.npz")
    info = chd.prepare.CHARACTERS[case.split("_")[0]]()
    r = chd.results.load_towr_results(d + "/sol_out.txt")
    an0, names, og, com = chd.results.apply_results(r, d + "/anim.bvh", s0, s1, info, run_ik=False)
    np.testing.assert_allclose(an0.rotations, qmat(a["rot_q_noik"]), atol=2e-9)
    np.testing.assert_allclose(an0.positions, a["pos_noik"], atol=2e-9)
    np.testing.assert_allclose(com, a["com_og"], atol=2e-9)
    np.testing.assert_array_equal(an0.parents, a["parents"])
    np.testing.assert_allclose(an0.offsets, a["offsets"], atol=1e-12)
    hist = []
    an, _, _, _ = chd.results.apply_results(r, d + "/anim.bvh", s0, s1, info, run_ik=True)
    np.testing.assert_allclose(an.rotations, qmat(a["rot_q"]), atol=5e-8)
    np.testing.assert_allclose(an.positions, a["pos"], atol=5e-8)
    np.testing.assert_allclose(an.global_positions(), a["gpos"], atol=1e-6)
    # the IK actually moved the feet onto the optimised trajectories (cm)
    toe_err = np.linalg.norm(an.global_positions()[:, info.toes[0]] - r.feet_pos[:s1 - s0, 0] * 100.0, axis=1).mean()
    toe_err0 = np.linalg.norm(an0.global_positions()[:, info.toes[0]] - r.feet_pos[:s1 - s0, 0] * 100.0, axis=1).mean()
    assert toe_err < 0.25 * toe_err0


def test_save_bvh_round_trip(chd, tmp_path):
    d = os.path.join(G, "ybot")
    info = chd.prepare.ybot_info()
    r = chd.results.load_towr_results(d + "/sol_out.txt")
    an, names, _, _ = chd.results.apply_results(r, d + "/anim.bvh", 2, 38, info, run_ik=True, iterations=3)
    an = chd.results.remove_heel_from_anim(an)
    assert len(an.names) == 67
    out = str(tmp_path / "out.bvh")
    chd.results.save_bvh(out, an, an.names, frametime=1.0 / 24.0)
    b = chd.prepare.load_bvh(out)
    assert b.names == an.names and list(b.parents) == list(an.parents)
    R, T = chd.prepare.local_transforms(b)
    np.testing.assert_allclose(R, an.rotations, atol=5e-8)         # six decimals of a degree
    np.testing.assert_allclose(T[:, 0], an.positions[:, 0], atol=1e-6)
    assert b.channels[0][3:] == ["Zrotation", "Yrotation", "Xrotation"]


def test_oracle_euler_convention_is_the_reference_consumers(chd):
    """An anchor outside our own reading of TOWR: the reference's consumer of `base_ang` (towr_utils.load_results:115,
    `Quaternions.from_euler(order='xyz', world=True)`, executed by the golden generator) assigns the same rotation to the
    written angles as the oracle's EulerConverter restatement does (R = Rz Ry Rx; tests/test_oracle_cpu.py ties that to the
    oracle's `euler()`), so the orientation the solver optimises is the orientation the pipeline applies to the skeleton."""
    from scipy.spatial.transform import Rotation
    d = os.path.join(G, "combined")
    g = np.load(d + "/results.npz")
    ang = chd.io_formats.read_solution(d + "/sol_out.txt")["base_ang_deg"]            # degrees, as phys_optim writes them
    R = Rotation.from_euler("xyz", np.radians(ang)).as_matrix()                   # extrinsic xyz = Rz Ry Rx, the oracle's convention
    C = chd.prepare.C_BVH_TO_TOWR
    np.testing.assert_allclose(C @ R @ C.T, g["base_R"], atol=1e-9)


def test_retarget_matches_reference(chd):
    """combined_to_mixamo.retarget of the reference (golden: run on the `combined` clip towards the synthetic 67-joint
    skeleton, result read back from the BVH it saved: six decimals) vs chd.results.retarget."""
    g = np.load(os.path.join(G, "retarget", "retarget.npz"))
    a = chd.results.retarget(os.path.join(G, "combined", "anim.bvh"), os.path.join(G, "retarget", "ybot_skel.bvh"), chd.prepare.ybot_info())
    np.testing.assert_allclose(a.rotations, qmat(g["rot_q"]), atol=5e-7)
    np.testing.assert_allclose(a.positions, g["pos"], atol=2e-6)
    np.testing.assert_allclose(a.global_positions(), g["gpos"], atol=2e-5)
    np.testing.assert_allclose(a.positions[:, 1:], np.tile(a.offsets[None, 1:], (a.positions.shape[0], 1, 1)), atol=0)   # bones restored


@pytest.mark.parametrize("case", list(CASES))
def test_reader_parses_files_written_by_the_reference(chd, case):
    """Row A1 on the reference's own output: the four files its `prepare_input` wrote (Python `str(float)` tokens) go through
    `read_phys_inputs` -- the reader that mirrors phys_optim.cpp:155-267 -- and give the problem our own prepare step builds."""
    s0, s1, comb, fps = CASES[case]
    d = os.path.join(G, case)
    q = chd.io_formats.read_phys_inputs(os.path.join(d, "phys_in"), s1 - s0)
    b = chd.prepare.load_bvh(d + "/anim.bvh")
    n, pt = [[float(v) for v in l.split()] for l in open(d + "/floor.txt").read().splitlines()[:2]]
    p = chd.prepare.build_problem(b, n, pt, np.load(d + "/foot_contacts.npy"), chd.prepare.CHARACTERS[case.split("_")[0]](), s0, s1, 1.0 / fps, comb)
    assert q.n_frames == p.n_frames == s1 - s0 and q.n_ee == 4
    for k in ("hip_left", "hip_right", "inertia", "base_lin", "base_ang", "ee_pos", "floor_normal", "floor_point"):
        np.testing.assert_allclose(getattr(q, k), getattr(p, k), rtol=0, atol=2e-9, err_msg=k)
    for k in ("max_leg_length", "max_heel_length", "heel_dist", "body_mass", "dt"):
        assert abs(getattr(q, k) - getattr(p, k)) < 2e-9, k
    assert list(q.ee_start_contact) == list(p.ee_start_contact)
    for e in range(4):
        np.testing.assert_allclose(q.ee_durations[e], p.ee_durations[e], atol=1e-12)
