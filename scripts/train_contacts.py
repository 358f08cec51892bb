#!/usr/bin/env python
"""Contact-classifier training with the optimiser flags of the reference's `contact_learning/train.py:14-43`
(--batch-size --epochs --val-every --lr --beta1 --beta2 --eps --decay).  Data layout: the real-video layout the rest of this
repo uses, `<data>/<video>/openpose_result/*.json` + `<data>/<video>/foot_contacts.npy` (F x 4 labels L heel, L toe, R heel,
R toe) -- the reference's synthetic character / motion / view tree is not distributed with it.  The keypoints go through
the CUDA preprocessing of the inference path (`chd_contact_preprocess`), so training and inference see identical inputs.
Writes `<out>/op_only_weights_FINAL.pth` (a state_dict with the reference's key names: loads into the reference's
`OpenPoseModel` and into `scripts/detect_contacts.py --weights-path`) and the same as `.npz`."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--epochs", type=int, default=5000)
    ap.add_argument("--val-every", type=int, default=20)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--beta1", type=float, default=0.9)
    ap.add_argument("--beta2", type=float, default=0.999)
    ap.add_argument("--eps", type=float, default=1e-8)
    ap.add_argument("--decay", type=float, default=0.0001)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--val-frac", type=float, default=0.1)
    a = ap.parse_args(argv)
    import torch
    import chd
    T = chd.train
    vids = sorted(v for v in os.listdir(a.data) if os.path.isdir(os.path.join(a.data, v, "openpose_result")))
    raw = chd.contact.load_keypoint_dirs([os.path.join(a.data, v, "openpose_result") for v in vids])
    labels = [np.load(os.path.join(a.data, v, "foot_contacts.npy")).astype(np.float32) for v in vids]
    net = chd.contact.ContactNet(T.Trainer(seed=a.seed).state_dict_numpy())          # only its preprocessing kernel is used
    frames, lens = net.preprocess(raw)
    seqs = [frames[i, :lens[i]] for i in range(len(vids))]
    rng = np.random.default_rng(a.seed)
    order = rng.permutation(len(vids))
    n_val = max(1, int(a.val_frac * len(vids))) if len(vids) > 4 else 0
    val_i, tr_i = order[:n_val], order[n_val:]
    val = None
    if n_val:
        xs, ys = zip(*[T.make_window(seqs[i], labels[i], t) for i in val_i for t in range(4, seqs[i].shape[0] - 4)])
        val = (torch.as_tensor(np.stack(xs), device="cuda"), torch.as_tensor(np.stack(ys), device="cuda"))
    os.makedirs(a.out, exist_ok=True)
    tr = T.Trainer(seed=a.seed, lr=a.lr, betas=(a.beta1, a.beta2), eps=a.eps, weight_decay=a.decay, device="cuda")
    h = T.WINDOW // 2
    for ep in range(a.epochs):
        perm = rng.permutation(tr_i)
        tot, cnt = 0.0, 0
        for s in range(0, len(perm), a.batch_size):
            idx = perm[s:s + a.batch_size]
            if len(idx) < 2:
                continue
            xs, ys = zip(*[T.make_window(seqs[i], labels[i], int(rng.integers(h, seqs[i].shape[0] - h)), 0.005, rng) for i in idx])
            l, _ = tr.step(torch.as_tensor(np.stack(xs), device="cuda"), torch.as_tensor(np.stack(ys), device="cuda"))
            tot, cnt = tot + l, cnt + 1
        if ep % a.val_every == 0 or ep == a.epochs - 1:
            msg = "epoch %d  mean loss %.4f" % (ep + 1, tot / max(cnt, 1))
            if val is not None:
                vl, vc = tr.evaluate(*val)
                msg += "  val loss %.4f  acc %.3f  f1 %.3f" % (vl, T.metrics(vc)[0], T.metrics(vc)[3])
            print(msg)
    sd = tr.state_dict_numpy()
    torch.save({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, os.path.join(a.out, "op_only_weights_FINAL.pth"))
    np.savez(os.path.join(a.out, "op_only_weights_FINAL.npz"), **sd)
    print("FINISHED Training!")


if __name__ == "__main__":
    main()
