"""Multi-GPU plumbing of the phys-optim path: sequences are independent NLPs (the reference runs them as separate
processes, scripts/run_phys_mocap.py:80), so they shard embarrassingly -- one process per GPU, no collective inside
the solver, one gather of the fixed-size sampled solutions at the end.

Everything here is backend agnostic (`nccl` on the GPU box, `gloo` in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_by_work(work: Sequence[float], world: int) -> List[List[int]]:
    """Deals sequence indices to ranks: sorted by estimated work (number of variables ~ iterations x size),
    largest first, round robin with reversal every other pass (snake order) so that ragged batches balance.
    Deterministic; every rank computes the same assignment."""
    order = sorted(range(len(work)), key=lambda i: (-float(work[i]), i))
    shards: List[List[int]] = [[] for _ in range(world)]
    for k, idx in enumerate(order):
        r = k % world
        if (k // world) % 2 == 1:
            r = world - 1 - r
        shards[r].append(idx)
    for s in shards:
        s.sort()
    return shards


def pad_to(shards: List[List[int]]) -> int:
    """Per-rank slot count so that the gather has equal receive counts."""
    return max((len(s) for s in shards), default=0)


def gather_samples(local, world: int, group=None):
    """all_gather of a (slots, frames, stride) tensor -> (world*slots, frames, stride).  `local` is a torch tensor
    on the device the process group's backend expects (cuda for nccl, cpu for gloo)."""
    import torch
    import torch.distributed as dist
    if world == 1 or not dist.is_initialized():
        return local
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    return out


def unshard(gathered: np.ndarray, shards: List[List[int]], slots: int) -> np.ndarray:
    """Reorders gathered rows (rank major, `slots` rows per rank) back to the original sequence order."""
    n = sum(len(s) for s in shards)
    out = np.zeros((n,) + gathered.shape[1:], dtype=gathered.dtype)
    for r, s in enumerate(shards):
        for k, idx in enumerate(s):
            out[idx] = gathered[r * slots + k]
    return out


# ---------------------------------------------------------------------------------------------------------------------
# The product's multi-GPU entry point: shard -> solve -> sample into the send buffer -> ONE gather -> unshard.
# ---------------------------------------------------------------------------------------------------------------------
N_EXTRA = 16  # per-sequence trailer of the gathered block: frames, success[2], stage_status[6], stage_iters[6], pad


def work_estimate(problems) -> List[float]:
    """Predicted solver work of a sequence: interior-point iterations grow with the number of contact phases (every
    switch adds free swing / stance nodes and, in stage 3, a duration), the cost of one iteration with the frame count."""
    return [float(p.n_frames) * float(sum(len(d) for d in p.ee_durations)) * (p.n_ee / 2.0) for p in problems]


def frames_out(p) -> int:
    """SaveSolution frame count of a problem (phys_optim.cpp:84: t accumulates dt while t <= T + 1e-5)."""
    T = float(np.sum(np.asarray(p.ee_durations[0], dtype=np.float64)))
    return int((T + 1e-5) / p.dt) + 1


class ShardedSolver:
    """One process per GPU (torchrun).  Every rank holds the same problem list, solves its shard on its device and
    takes part in a single all_gather of the fixed-size result block (final SaveSolution snapshot + status trailer).

    `solve_fn(problems) -> dict(final=(n, fo, stride) array, frames, success, stage_status (6,n), stage_iters (6,n))`
    replaces the CUDA solve in the CPU (gloo) tests; the default drives `PhysBatch` on `device`."""

    def __init__(self, problems, weights=(0.4, 1.7, 0.3, 0.1, 0.1), device: int = 0, rank: int = 0, world: int = 1,
                 group=None, solve_fn=None, tensor_device=None):
        self.problems, self.rank, self.world, self.group = list(problems), rank, world, group
        self.shards = shard_by_work(work_estimate(self.problems), world)
        self.slots = pad_to(self.shards)
        self.mine = self.shards[rank]
        self.n_ee_max = max(p.n_ee for p in self.problems)
        self.stride = 6 + 7 * self.n_ee_max
        self.fo = max(frames_out(p) for p in self.problems)
        self.width = self.fo * self.stride + N_EXTRA
        self.solve_fn = solve_fn
        self.batch = None
        import torch
        if solve_fn is None:
            from . import phys
            self.batch = phys.PhysBatch([self.problems[i] for i in self.mine], weights=weights, device=device) if self.mine else None
            tensor_device = tensor_device or torch.device("cuda", device)
        self.tdev = tensor_device or torch.device("cpu")
        self.send = torch.zeros((self.slots, self.width), dtype=torch.float64, device=self.tdev)
        self.recv = torch.zeros((world * self.slots, self.width), dtype=torch.float64, device=self.tdev) if world > 1 else None
        self.last_ms = {}

    def close(self):
        if self.batch is not None:
            self.batch.close()
            self.batch = None

    def _solve_local(self, resident: bool):
        """fills self.send; returns the local status arrays"""
        import torch
        n = len(self.mine)
        trailer = np.zeros((self.slots, N_EXTRA))
        if n == 0:
            return trailer
        if self.solve_fn is not None:
            r = self.solve_fn([self.problems[i] for i in self.mine])
            blk = np.zeros((self.slots, self.fo, self.stride))
            f = np.asarray(r["final"])
            blk[:n, :f.shape[1], :f.shape[2]] = f
            self.send[:, :self.fo * self.stride] = torch.from_numpy(blk.reshape(self.slots, -1)).to(self.tdev)
            frames, success, sstat, siter = r["frames"], r["success"], r["stage_status"], r["stage_iters"]
        else:
            b = self.batch
            if resident:
                b.reset()
            B = b.B
            sstat, siter, success = np.zeros((6, B), np.int32), np.zeros((6, B), np.int32), np.zeros((B, 2), np.int32)
            b._chk(b.L.chd_phys_solve(b.h, None, None, success.ctypes.data, sstat.ctypes.data, siter.ctypes.data))
            # final iterate sampled on the device straight into the send buffer (no host round trip)
            fo_l, st_l = b.dims["frames_out_max"], 6 + 7 * b.n_ee_max
            if fo_l == self.fo and st_l == self.stride:
                view = self.send[:n, :self.fo * self.stride]
                if view.is_contiguous():
                    b._chk(b.L.chd_phys_sample_device(b.h, view.data_ptr(), torch.cuda.current_stream().cuda_stream))
                else:
                    tmp = torch.zeros((n, self.fo * self.stride), dtype=torch.float64, device=self.tdev)
                    b._chk(b.L.chd_phys_sample_device(b.h, tmp.data_ptr(), torch.cuda.current_stream().cuda_stream))
                    self.send[:n, :self.fo * self.stride] = tmp
            else:   # ragged shards: local frame / stride padding differs from the global one
                tmp = torch.zeros((n, fo_l, st_l), dtype=torch.float64, device=self.tdev)
                b._chk(b.L.chd_phys_sample_device(b.h, tmp.data_ptr(), torch.cuda.current_stream().cuda_stream))
                blk = torch.zeros((n, self.fo, self.stride), dtype=torch.float64, device=self.tdev)
                blk[:, :fo_l, :st_l] = tmp
                self.send[:n, :self.fo * self.stride] = blk.reshape(n, -1)
            frames = np.array([frames_out(self.problems[i]) for i in self.mine], np.int32)
        trailer[:n, 0] = frames
        trailer[:n, 1:3] = np.asarray(success).reshape(n, 2)
        trailer[:n, 3:9] = np.asarray(sstat).T
        trailer[:n, 9:15] = np.asarray(siter).T
        self.send[:, self.fo * self.stride:] = torch.from_numpy(trailer).to(self.tdev)
        return trailer

    def solve(self, resident: bool = False) -> dict:
        """Solves the shard (resident=True: device-side reset of an already uploaded batch) and gathers.  Every rank
        returns the full result in the original sequence order."""
        import time
        import torch
        import torch.distributed as dist
        t0 = time.perf_counter()
        self._solve_local(resident)
        if self.tdev.type == "cuda":
            torch.cuda.synchronize()
        t1 = time.perf_counter()
        if self.world > 1:
            dist.all_gather_into_tensor(self.recv.view(-1), self.send.view(-1), group=self.group)
            if self.tdev.type == "cuda":
                torch.cuda.synchronize()
            g = self.recv
        else:
            g = self.send
        t2 = time.perf_counter()
        self.last_ms = {"solve_ms": 1e3 * (t1 - t0), "gather_ms": 1e3 * (t2 - t1)}
        g = g.cpu().numpy()
        full = unshard(g, self.shards, self.slots)
        N = len(self.problems)
        tr = full[:, self.fo * self.stride:]
        return dict(samples=full[:, :self.fo * self.stride].reshape(N, self.fo, self.stride), frames=tr[:, 0].astype(np.int32),
                    success=tr[:, 1:3].astype(np.int32), stage_status=tr[:, 3:9].astype(np.int32).T,
                    stage_iters=tr[:, 9:15].astype(np.int32).T, d2h_bytes=int(g.nbytes))


def solve_sharded(problems, weights=(0.4, 1.7, 0.3, 0.1, 0.1), device: int = 0, rank: int = 0, world: int = 1, group=None,
                  solve_fn=None, tensor_device=None) -> dict:
    """shard -> solve -> one gather -> unshard for a list of `PhysProblem`s (see ShardedSolver)."""
    s = ShardedSolver(problems, weights, device, rank, world, group, solve_fn, tensor_device)
    try:
        return s.solve()
    finally:
        s.close()


# ---------------------------------------------------------------------------------------------------------------------
# Contact classifier: videos are independent too: shard by frame count, one gather of the int64 labels.
# ---------------------------------------------------------------------------------------------------------------------
def detect_contacts_sharded(raw, state_dict=None, device: int = 0, rank: int = 0, world: int = 1, group=None, detect_fn=None,
                            gather_device=None):
    """`raw`: list of (F_i, 25, 3) OpenPose keypoint arrays, identical on every rank.  Every rank runs
    `ContactNet.detect` (`chd_contact_detect`: preprocessing, windows, network, votes on its GPU) on its shard, the
    (slots, F_max, 4) int64 label blocks are gathered once and put back in input order.  Returns the list of (F_i, 4) labels.
    `detect_fn(list_of_raw) -> list of (F_i, 4)` replaces the CUDA path in the CPU tests."""
    import torch
    n = len(raw)
    shards = shard_by_work([float(r.shape[0]) for r in raw], world)
    slots = pad_to(shards)
    f_max = max(int(r.shape[0]) for r in raw)
    mine = shards[rank]
    # The reference pads every video of a batch to the batch's longest one by repeating the last frame
    # (real_video_dataset.py:165-191) and votes over the padded windows before trimming, so the last labels of a video
    # depend on the longest video it is batched with.  To return exactly what one unsharded call returns, a shard that does
    # not hold the globally longest video carries it along (its labels are dropped).
    longest = max(range(n), key=lambda i: (raw[i].shape[0], -i))
    batch = [raw[i] for i in mine] + ([raw[longest]] if mine and longest not in mine else [])
    if detect_fn is None:
        from .contact import ContactNet
        net = ContactNet(state_dict, device=device)
        labels = net.detect(batch)[0][:len(mine)] if mine else []
        net.close()
    else:
        labels = detect_fn(batch)[:len(mine)] if mine else []
    dev = gather_device if gather_device is not None else (torch.device("cuda", device) if detect_fn is None else torch.device("cpu"))
    block = torch.zeros((slots, f_max, 4), dtype=torch.int64, device=dev)
    for k, lab in enumerate(labels):
        block[k, :lab.shape[0]] = torch.as_tensor(np.asarray(lab, dtype=np.int64), device=dev)
    gathered = gather_samples(block, world, group).cpu().numpy()
    full = unshard(gathered, shards, slots)
    return [full[i, :raw[i].shape[0]] for i in range(n)]
