"""Host side of the foot-contact classifier path (reference: scripts/run_detect_contacts.py ->
src/contact_learning/test.py --full-video --save-contacts --real-data).

Keypoint loading and the dataset preprocessing are restated in numpy with the reference's operation order (they
are fp64 on the host in the reference too); window construction, the MLP and the vote aggregation run in
hand-written CUDA behind `chd_contact_*` (include/chd.h).  No CPU fallback for the network.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Dict, List, Sequence

import numpy as np

from .phys import load_lib

TRAIN_DIM = (1280, 720)                      # real_video_dataset.py:17
TRAIN_NORMALIZATION = 200.4160302695367      # real_video_dataset.py:18
WINDOW, PRED = 9, 5
LIN_IDS, BN_IDS = [0, 3, 6, 10, 13], [1, 4, 7, 11]
DIMS = [351, 1024, 512, 128, 32, 20]


def load_keypoint_file(path: str, num_joints: int = 25) -> np.ndarray:
    """openpose_utils.py:48-66: first person's pose_keypoints_2d as (J,3); zeros if nobody was detected."""
    with open(path) as f:
        d = json.load(f)
    if len(d["people"]) == 0:
        return np.zeros((num_joints, 3))
    return np.array(d["people"][0]["pose_keypoints_2d"], dtype=np.float64).reshape(-1, 3)


def load_keypoint_dir(path: str) -> np.ndarray:
    """openpose_utils.py:68-76: all *.json of a directory in sorted order -> (F,25,3)."""
    files = sorted(os.path.join(path, f) for f in os.listdir(path) if f.split(".")[-1] == "json")
    return np.stack([load_keypoint_file(f) for f in files], axis=0)


def interpolate_low_confidence(seq: np.ndarray, thresh: float = 0.2) -> np.ndarray:
    """process_openpose_data (openpose_dataset.py:49-111) for one (F,J,3) sequence, in place on xy:
    leading / trailing low-confidence runs copy the nearest valid frame, interior runs are linearly interpolated
    with the reference's accumulating step."""
    xy, conf = seq[:, :, :2], seq[:, :, 2]
    F = seq.shape[0]
    for j in range(seq.shape[1]):
        t = 0
        while t < F:
            if conf[t, j] < thresh:
                nxt = t + 1
                while nxt < F and conf[nxt, j] < thresh:
                    nxt += 1
                init = t - 1
                if t == 0 and nxt == F:
                    pass
                elif t == 0:
                    xy[:nxt, j, :] = xy[nxt, j, :].reshape((1, 2))
                elif nxt == F:
                    xy[init:, j, :] = xy[init, j, :].reshape((1, 2))
                else:
                    step = 1.0 / (nxt - init)
                    cur = step
                    ct = t
                    while ct < nxt:
                        xy[ct, j, :] = (1.0 - cur) * xy[init, j, :] + cur * xy[nxt, j, :]
                        ct += 1
                        cur += step
                t = nxt
            else:
                t += 1
    return seq


def preprocess_videos(raw: Sequence[np.ndarray], dimensions=(1920, 1080)):
    """RealVideoDataset.__init__ (real_video_dataset.py:132-163): pad every video to the longest by repeating the
    last frame, scale xy by 1280/width, interpolate low-confidence joints, divide xy by the training normalisation.
    Returns (frames (V,Fmax,25,3) fp64, seq_lens (V,) int32)."""
    seq_lens = np.array([r.shape[0] for r in raw], dtype=np.int32)
    Fmax = int(seq_lens.max())
    out = np.zeros((len(raw), Fmax, 25, 3))
    scale = float(TRAIN_DIM[0]) / dimensions[0]
    for i, r in enumerate(raw):
        a = np.array(r, dtype=np.float64)
        if a.shape[0] < Fmax:
            a = np.concatenate([a, np.repeat(a[-1].reshape((1, 25, 3)), Fmax - a.shape[0], axis=0)], axis=0)
        a[:, :, :2] *= scale
        a = interpolate_low_confidence(a, 0.2)
        a[:, :, :2] /= TRAIN_NORMALIZATION
        out[i] = a
    return out, seq_lens


def pack_state_dict(sd: Dict[str, np.ndarray]):
    """Flattens a reference state_dict (numpy values) into the three arrays chd_contact_create takes."""
    w = np.concatenate([np.asarray(sd["model.%d.weight" % i], dtype=np.float32).reshape(-1) for i in LIN_IDS])
    b = np.concatenate([np.asarray(sd["model.%d.bias" % i], dtype=np.float32).reshape(-1) for i in LIN_IDS])
    bn = np.concatenate([np.concatenate([np.asarray(sd["model.%d.%s" % (i, k)], dtype=np.float32).reshape(-1)
                                         for k in ("weight", "bias", "running_mean", "running_var")]) for i in BN_IDS])
    return np.ascontiguousarray(w), np.ascontiguousarray(b), np.ascontiguousarray(bn)


def load_weights(path: str) -> Dict[str, np.ndarray]:
    """A reference `.pth` state_dict (torch.load) or an `.npz` with the same keys."""
    if path.endswith(".npz"):
        return dict(np.load(path))
    import torch
    sd = torch.load(path, map_location="cpu")
    return {k: v.numpy() for k, v in sd.items()}


class ContactNet:
    def __init__(self, state_dict: Dict[str, np.ndarray], device: int = -1, bn_eps: float = 1e-5):
        self.L = load_lib()
        L = self.L
        L.chd_contact_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.POINTER(C.c_void_p)]
        L.chd_contact_destroy.argtypes = [C.c_void_p]
        L.chd_contact_destroy.restype = None
        L.chd_contact_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.chd_contact_launch_count.argtypes = [C.c_void_p]
        L.chd_contact_launch_count.restype = C.c_int64
        w, b, bn = pack_state_dict(state_dict)
        assert w.size == sum(i * o for i, o in zip(DIMS[:-1], DIMS[1:])) and b.size == sum(DIMS[1:])
        h = C.c_void_p()
        rc = L.chd_contact_create(w.ctypes.data, b.ctypes.data, bn.ctypes.data, C.c_float(bn_eps), device, C.byref(h))
        if rc != 0:
            raise RuntimeError("chd_contact_create failed with code %d (no CUDA device? no CPU fallback exists)" % rc)
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.chd_contact_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, frames: np.ndarray, seq_lens: np.ndarray, want_logits: bool = False):
        """frames (V,Fmax,25,3) preprocessed fp64 -> labels (V,Fmax,4) int64 [, logits (V,Fmax-8,5,4)], min|logit|."""
        frames = np.ascontiguousarray(frames, dtype=np.float64)
        seq_lens = np.ascontiguousarray(seq_lens, dtype=np.int32)
        V, Fmax = frames.shape[:2]
        labels = np.zeros((V, Fmax, 4), dtype=np.int64)
        logits = np.zeros((V, Fmax - (WINDOW - 1), PRED, 4), dtype=np.float32) if want_logits else None
        mabs = np.zeros(1, dtype=np.float32)
        rc = self.L.chd_contact_forward(self.h, frames.ctypes.data, V, Fmax, seq_lens.ctypes.data, labels.ctypes.data,
                                        logits.ctypes.data if want_logits else None, mabs.ctypes.data)
        if rc != 0:
            raise RuntimeError("chd_contact_forward failed with code %d" % rc)
        return (labels, logits, float(mabs[0])) if want_logits else (labels, float(mabs[0]))

    def launch_count(self) -> int:
        return int(self.L.chd_contact_launch_count(self.h))


def detect_contacts(data_root: str, out_root: str, state_dict, dimensions=(1920, 1080)) -> List[str]:
    """`test.py --data D --out O --full-video --save-contacts --real-data`: for every video directory of D with an
    `openpose_result/` writes O/contact_results/<video>/foot_contacts.npy (int64, F x 4), test.py:143-152."""
    vids = sorted(d for d in os.listdir(data_root) if os.path.isdir(os.path.join(data_root, d)) and d[0] != ".")
    raw = [load_keypoint_dir(os.path.join(data_root, v, "openpose_result")) for v in vids]
    frames, seq_lens = preprocess_videos(raw, dimensions)
    net = ContactNet(state_dict)
    labels, _ = net.forward(frames, seq_lens)
    written = []
    for i, v in enumerate(vids):
        od = os.path.join(out_root, "contact_results", v)
        os.makedirs(od, exist_ok=True)
        np.save(os.path.join(od, "foot_contacts"), labels[i, :seq_lens[i]].astype(np.int64))
        written.append(os.path.join(od, "foot_contacts.npy"))
    return written
