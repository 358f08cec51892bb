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
