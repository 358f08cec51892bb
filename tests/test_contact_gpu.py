"""GPU parity tests of the contact classifier (CUDA kernels through the C ABI)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def test_golden_labels_bit_exact(chd):
    from make_contact_golden import contact_weights
    g = dict(np.load(os.path.join(HERE, "golden", "contact", "contact_golden.npz")))
    names = [str(n) for n in g["names"]]
    frames = np.stack([g["proc_" + n] for n in names])
    net = chd.contact.ContactNet(contact_weights(0))
    labels, logits, mabs = net.forward(frames, g["seq_lens"].astype(np.int32), want_logits=True)
    # fp32 logits: same math, different summation order than the reference's torch forward -> 2e-5 absolute
    np.testing.assert_allclose(logits, g["logits"], rtol=0, atol=2e-5)
    for i, n in enumerate(names):
        L = int(g["seq_lens"][i])
        np.testing.assert_array_equal(labels[i, :L], g["contacts_" + n])     # integer labels: bit exact
        assert (labels[i, L:] == 0).all()
    assert labels.dtype == np.int64 and mabs > 0
    assert net.launch_count() == 6      # gather, three tiled layers, tail, vote (one slab)


def test_large_batch_against_torch_reference(chd):
    from make_contact_golden import contact_weights, synth_keypoints
    from oracle import contact as oc
    sd = contact_weights(1)
    raw = [synth_keypoints(1000 + i, 60 + (i % 7)) for i in range(48)]
    net = chd.contact.ContactNet(sd)
    frames, seq_lens = net.preprocess(raw)
    f_ref, l_ref = oc.preprocess_videos(raw)
    np.testing.assert_array_equal(frames, f_ref)                       # device preprocessing: bit exact vs the numpy checker
    np.testing.assert_array_equal(seq_lens, l_ref)
    labels, logits, mabs = net.forward(frames, seq_lens, want_logits=True)
    det, _ = net.detect(raw)                                           # one-call path: same labels
    for i in range(len(raw)):
        np.testing.assert_array_equal(det[i], labels[i, :seq_lens[i]])
    ref_logits = oc.forward_torch(sd, oc.windows_from_frames(frames))
    np.testing.assert_allclose(logits, ref_logits, rtol=0, atol=5e-5)
    risky = 0
    for i in range(len(raw)):
        ref = oc.vote(ref_logits[i], int(seq_lens[i]))
        if not np.array_equal(labels[i, :seq_lens[i]], ref):
            # a flip is only admissible where a logit sits on the decision boundary
            assert np.abs(ref_logits[i]).min() < 1e-4
            risky += 1
    assert risky <= 1


def test_device_preprocessing_matches_reference_golden(chd):
    """chd_k_contact_prep (padding, scaling, low-confidence interpolation, normalisation on the device) against the arrays
    the reference's own RealVideoDataset produced (tests/golden/make_contact_golden.py): bit exact."""
    from make_contact_golden import contact_weights
    g = dict(np.load(os.path.join(HERE, "golden", "contact", "contact_golden.npz")))
    names = [str(n) for n in g["names"]]
    raw = [g["raw_" + n].copy() for n in names]
    for r in raw:      # frame 5 of the longer clips had no detections in the JSON dir -> zeros (openpose_utils.py:60-62)
        if r.shape[0] > 45:
            r[5] = 0.0
    net = chd.contact.ContactNet(contact_weights(0))
    frames, seq_lens = net.preprocess(raw)
    assert list(seq_lens) == list(g["seq_lens"])
    for i, n in enumerate(names):
        np.testing.assert_array_equal(frames[i], g["proc_" + n])
    det, _ = net.detect(raw)
    for i, n in enumerate(names):
        np.testing.assert_array_equal(det[i], g["contacts_" + n])


def test_trained_weights_drive_cuda_inference(chd):
    """Training (chd.train, torch on cuda:0) -> state_dict -> chd_contact_create: the CUDA inference path reproduces the
    eval-mode forward of the trained parameters (running statistics included) on the golden clips."""
    import torch
    T = chd.train
    g = dict(np.load(os.path.join(HERE, "golden", "contact", "contact_golden.npz")))
    names = [str(n) for n in g["names"]]
    frames = np.stack([g["proc_" + n] for n in names])
    tr = T.Trainer(seed=4, lr=1e-3, device="cuda")
    rng = np.random.default_rng(0)
    for _ in range(20):
        x = torch.as_tensor(rng.normal(0, 0.5, (64, 9, 13, 3)).astype(np.float32), device="cuda")
        y = torch.as_tensor((rng.uniform(size=(64, 5, 4)) < 0.5).astype(np.float32), device="cuda")
        tr.step(x, y)
    sd = tr.state_dict_numpy()
    assert int(sd["model.1.num_batches_tracked"]) == 20
    net = chd.contact.ContactNet(sd)
    labels, logits, mabs = net.forward(frames, g["seq_lens"].astype(np.int32), want_logits=True)
    win = torch.as_tensor(g["windows"].reshape(-1, 9, 13, 3), device="cuda")
    with torch.no_grad():
        ref = T.forward(tr.sd, win, False).cpu().numpy().reshape(logits.shape)
    np.testing.assert_allclose(logits, ref, rtol=0, atol=5e-5 * max(1.0, float(np.abs(ref).max())))
