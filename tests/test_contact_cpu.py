"""CPU tests of the contact path against golden vectors produced by the reference's OWN code
(tests/golden/make_contact_golden.py): pins the oracle and the product's host-side preprocessing."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


@pytest.fixture(scope="module")
def golden():
    g = dict(np.load(os.path.join(HERE, "golden", "contact", "contact_golden.npz")))
    g["names"] = [str(n) for n in g["names"]]
    return g


def test_oracle_preprocessing_bit_exact(chd, golden):
    """the numpy checker of the product's preprocessing kernel against the arrays the reference's RealVideoDataset made"""
    from oracle import contact as oc
    raw = [golden["raw_" + n] for n in golden["names"]]
    raw = [r.copy() for r in raw]
    for r in raw:      # frame 5 of the longer clips had no detections in the JSON dir -> zeros (openpose_utils.py:60-62)
        if r.shape[0] > 45:
            r[5] = 0.0
    frames, seq_lens = oc.preprocess_videos(raw)
    assert list(seq_lens) == list(golden["seq_lens"])
    for i, n in enumerate(golden["names"]):
        np.testing.assert_array_equal(frames[i], golden["proc_" + n])


def test_keypoint_dir_loader(chd, golden, tmp_path):
    kp = golden["raw_vid_b"]
    d = tmp_path / "openpose_result"
    d.mkdir()
    for f in range(kp.shape[0]):
        people = [] if f == 3 else [{"pose_keypoints_2d": [float(x) for x in kp[f].reshape(-1)]}]
        (d / ("v_%012d_keypoints.json" % f)).write_text(json.dumps({"version": 1.3, "people": people}))
    got = chd.contact.load_keypoint_dir(str(d))
    exp = kp.copy()
    exp[3] = 0
    np.testing.assert_array_equal(got, exp)


def test_native_keypoint_reader_matches_json_load(chd, tmp_path):
    """chd_openpose_load (threaded C++ reader) == json.load + np.array(...).reshape(-1, 3) of openpose_utils.py:48-66, bit for
    bit, on OpenPose-style files: several people (first one counts), extra keys before / after, exponents, integers,
    negative values, pretty-printed and compact files, empty people lists; many files over several directories."""
    rng = np.random.default_rng(5)
    dirs, expect = [], []
    for v in range(3):
        d = tmp_path / ("vid%d" % v) / "openpose_result"
        d.mkdir(parents=True)
        frames = []
        for f in range(40 + 7 * v):
            kp = rng.normal(0, 300, (25, 3)) * rng.choice([1.0, 1e-7, 1e5], (25, 3))
            kp[rng.integers(0, 25)] = [0, 0, 0]
            kp[rng.integers(0, 25), 0] = 1234.0
            others = rng.normal(0, 1, (25, 3))
            if f % 11 == 3:
                doc = {"version": 1.3, "people": []}
            else:
                person = {"person_id": [-1], "pose_keypoints_2d": [float(x) for x in kp.reshape(-1)], "face_keypoints_2d": [],
                          "hand_left_keypoints_2d": [0.5, 1, 2]}
                doc = {"version": 1.3, "people": [person, {"person_id": [-1], "pose_keypoints_2d": [float(x) for x in others.reshape(-1)]}]}
            txt = json.dumps(doc, indent=2 if f % 2 else None)
            if f % 5 == 0:
                txt = txt.replace("1234.0", "1234")            # an integer literal
            path = d / ("v_%012d_keypoints.json" % f)
            path.write_text(txt)
            (d / ("v_%012d_rendered.png" % f)).write_text("not json")
            j = json.loads(txt)
            frames.append(np.zeros((25, 3)) if not j["people"] else np.array(j["people"][0]["pose_keypoints_2d"]).reshape(-1, 3))
        dirs.append(str(d))
        expect.append(np.stack(frames))
    got = chd.contact.load_keypoint_dirs(dirs, threads=4)
    for g, e in zip(got, expect):
        assert g.dtype == np.float64
        np.testing.assert_array_equal(g, e)
    np.testing.assert_array_equal(chd.contact.load_keypoint_dir(dirs[1]), expect[1])
    bad = tmp_path / "bad.json"
    bad.write_text(json.dumps({"people": [{"pose_keypoints_2d": [1.0, 2.0, 3.0]}]}))
    with pytest.raises(RuntimeError):
        chd.contact.load_keypoint_files([str(bad)])
    with pytest.raises(RuntimeError):
        chd.contact.load_keypoint_files([str(tmp_path / "missing.json")])


def test_oracle_windows_and_votes_match_reference(golden):
    from oracle import contact as oc
    frames = np.stack([golden["proc_" + n] for n in golden["names"]])
    win = oc.windows_from_frames(frames)
    np.testing.assert_array_equal(win, golden["windows"])
    for i, n in enumerate(golden["names"]):
        lab = oc.vote(golden["logits"][i], int(golden["seq_lens"][i]))
        np.testing.assert_array_equal(lab, golden["contacts_" + n])
        assert lab.dtype == np.int64


def test_oracle_forward_matches_reference_logits(golden):
    from make_contact_golden import contact_weights
    from oracle import contact as oc
    logits = oc.forward_torch(contact_weights(0), golden["windows"])
    np.testing.assert_allclose(logits, golden["logits"], rtol=0, atol=2e-5)


def test_pack_state_dict_shapes(chd):
    from make_contact_golden import contact_weights
    w, b, bn = chd.contact.pack_state_dict(contact_weights(0))
    assert w.size == 953984 and b.size == 1716 and bn.size == 4 * (1024 + 512 + 128 + 32)   # SURVEY D4 parameter counts
