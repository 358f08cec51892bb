"""Contact-classifier training: produces the `state_dict` that `chd.contact.ContactNet` (the CUDA
inference path) and the reference's `test.py --weights-path` consume, so weights can be made in-house when the pretrained
`.pth` is unavailable.

Reference: `src/contact_learning/train.py:45-185` (Adam lr 1e-4, betas (0.9, 0.999), eps 1e-8, weight decay 1e-4, batch 64,
mean BCE-with-logits over 5 target frames x 4 contacts, one random window per sequence per epoch),
`models/openpose_only.py:14-110` (network in training mode: batch-statistics BatchNorm with momentum 0.1, Dropout 0.3
after the third block, Xavier-uniform weights, biases 0.01), `data/openpose_dataset.py:277-363` (window construction with
the centre-frame root trick and N(0, 0.005) position noise).

Written on the flat parameter dictionary (reference key names) with functional torch ops -- plumbing; the GEMMs go to
cuBLAS, a library -- so that the trained parameters drop straight into `chd_contact_create`.  One training step is pinned
to the reference's own module (tests/golden/make_contact_train_golden.py, tests/test_train_cpu.py).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

LIN_IDS, BN_IDS = [0, 3, 6, 10, 13], [1, 4, 7, 11]
DIMS = [351, 1024, 512, 128, 32, 20]
LOWER_JOINTS = [8, 9, 10, 11, 12, 13, 14, 19, 20, 21, 22, 23, 24]     # openpose_dataset.py:38
ROOT_JOINT = 8
WINDOW, PRED = 9, 5


def init_state(seed: int = 0, device=None) -> Dict[str, "torch.Tensor"]:
    """Fresh parameters with the reference's initialisation (openpose_only.py:47-52) under its key names."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for li, i in enumerate(LIN_IDS):
        fan_in, fan_out = DIMS[li], DIMS[li + 1]
        a = (6.0 / (fan_in + fan_out)) ** 0.5                          # Xavier uniform, gain 1
        sd["model.%d.weight" % i] = (torch.rand(fan_out, fan_in, generator=g) * 2.0 - 1.0) * a
        sd["model.%d.bias" % i] = torch.full((fan_out,), 0.01)
    for bi, i in enumerate(BN_IDS):
        n = DIMS[bi + 1]
        sd["model.%d.weight" % i], sd["model.%d.bias" % i] = torch.ones(n), torch.zeros(n)
        sd["model.%d.running_mean" % i], sd["model.%d.running_var" % i] = torch.zeros(n), torch.ones(n)
        sd["model.%d.num_batches_tracked" % i] = torch.tensor(0, dtype=torch.long)
    return {k: v.to(device) if device is not None else v for k, v in sd.items()}


def trainable(sd):
    return [k for k in sd if k.endswith(".weight") or k.endswith(".bias")]


def forward(sd, x, training: bool, dropout_p: float = 0.3):
    """(B, 9, 13, 3) -> (B, 5, 4) logits.  Training mode updates the running statistics in `sd` in place (momentum 0.1)."""
    import torch
    import torch.nn.functional as Fn
    h = x.reshape(x.shape[0], -1)
    for li, i in enumerate(LIN_IDS):
        h = Fn.linear(h, sd["model.%d.weight" % i], sd["model.%d.bias" % i])
        if li < 4:
            b = BN_IDS[li]
            if training:
                sd["model.%d.num_batches_tracked" % b] += 1
            h = Fn.batch_norm(h, sd["model.%d.running_mean" % b], sd["model.%d.running_var" % b], sd["model.%d.weight" % b],
                              sd["model.%d.bias" % b], training, 0.1, 1e-5)
            h = torch.relu(h)
            if li == 2:
                h = Fn.dropout(h, dropout_p, training)
    return h.reshape(-1, PRED, 4)


def loss_fn(logits, labels):
    import torch.nn.functional as Fn
    return Fn.binary_cross_entropy_with_logits(logits.reshape(logits.shape[0], -1), labels.reshape(labels.shape[0], -1), reduction="none").mean()


def confusion(logits, labels, thresh: float = 0.5, tgt: Optional[int] = None):
    """openpose_only.py:80-110: counts (tp, fp, fn, tn) on the target (middle) frame."""
    import torch
    tgt = PRED // 2 if tgt is None else tgt
    pred = torch.sigmoid(logits[:, tgt]) > thresh
    lab = labels[:, tgt] > 0.5
    return np.array([int((pred & lab).sum()), int((pred & ~lab).sum()), int((~pred & lab).sum()), int((~pred & ~lab).sum())])


def metrics(c):
    """contact_learning/utils.py:73-96."""
    tp, fp, fn, tn = [float(v) for v in c]
    tot = tp + fp + fn + tn
    if tp + fp == 0:
        return (tp + tn) / tot, 0.0, 0.0, 0.0
    p, r = tp / (tp + fp), tp / (tp + fn) if tp + fn > 0 else 0.0
    return (tp + tn) / tot, p, r, (2 * p * r / (p + r) if p + r > 0 else 0.0)


class Trainer:
    def __init__(self, sd=None, seed: int = 0, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-4, device=None):
        import torch
        self.sd = sd if sd is not None else init_state(seed, device)
        self.params = [self.sd[k].requires_grad_(True) for k in trainable(self.sd)]
        self.opt = torch.optim.Adam(self.params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)

    def step(self, x, labels) -> Tuple[float, np.ndarray]:
        self.opt.zero_grad()
        out = forward(self.sd, x, True)
        loss = loss_fn(out, labels)
        loss.backward()
        self.opt.step()
        return float(loss.detach()), confusion(out.detach(), labels)

    def evaluate(self, x, labels):
        import torch
        with torch.no_grad():
            out = forward(self.sd, x, False)
            return float(loss_fn(out, labels)), confusion(out, labels)

    def state_dict_numpy(self) -> Dict[str, np.ndarray]:
        return {k: v.detach().cpu().numpy() for k, v in self.sd.items()}


def make_window(frames, contacts, tgt: int, noise_dev: float = 0.0, rng=None):
    """openpose_dataset.py:320-356 for one window centred on frame `tgt` of a preprocessed sequence (F, 25, 3):
    centre-frame MidHip subtracted from all xy and restored on the centre frame, 13 lower-body joints, optional noise on xy;
    labels = the middle 5 frames of the window."""
    h = WINDOW // 2
    w = np.array(frames[tgt - h:tgt + h + 1], dtype=np.float64)
    root = w[h, ROOT_JOINT, :2].copy()
    w[:, :, :2] -= root
    w[h, ROOT_JOINT, :2] = root
    w = w[:, LOWER_JOINTS]
    if noise_dev > 0:
        w[:, :, :2] += rng.normal(0.0, noise_dev, w[:, :, :2].shape)
    off = (WINDOW - PRED) // 2
    return w.astype(np.float32), np.asarray(contacts[tgt - h + off:tgt + h + 1 - off], dtype=np.float32)


def train(frames, contacts, epochs: int = 50, batch_size: int = 64, seed: int = 0, noise_dev: float = 0.005, device=None, val=None, log=None):
    """`frames`: list of preprocessed sequences (F_i, 25, 3) (what `ContactNet.preprocess` / the reference's dataset produce),
    `contacts`: list of (F_i, 4) labels [L heel, L toe, R heel, R toe].  One random window per sequence and epoch
    (train.py:99-118).  Returns the Trainer (its `state_dict_numpy()` feeds `chd.contact.ContactNet`)."""
    import torch
    rng = np.random.default_rng(seed)
    tr = Trainer(seed=seed, device=device)
    n = len(frames)
    h = WINDOW // 2
    for ep in range(epochs):
        order = rng.permutation(n)
        tot, cnt, conf = 0.0, 0, np.zeros(4, dtype=int)
        for s in range(0, n, batch_size):
            idx = order[s:s + batch_size]
            if len(idx) < 2:
                continue                   # batch statistics need at least two samples
            xs, ys = zip(*[make_window(frames[i], contacts[i], int(rng.integers(h, frames[i].shape[0] - h)), noise_dev, rng) for i in idx])
            x = torch.as_tensor(np.stack(xs), device=device)
            y = torch.as_tensor(np.stack(ys), device=device)
            l, c = tr.step(x, y)
            tot, cnt, conf = tot + l, cnt + 1, conf + c
        if log is not None and (ep % 5 == 0 or ep == epochs - 1):
            msg = "epoch %d  loss %.4f  acc %.3f" % (ep + 1, tot / max(cnt, 1), metrics(conf)[0])
            if val is not None:
                vl, vc = tr.evaluate(*val)
                msg += "  val loss %.4f  val f1 %.3f" % (vl, metrics(vc)[3])
            log(msg)
    return tr
