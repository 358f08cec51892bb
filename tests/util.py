"""Shared helpers of the parity tests: mapping between the product's master row order and the
oracle's ifopt row order (which follows phys_optim.cpp's per-stage AddConstraintSet order)."""
import numpy as np

ORACLE_PREFIX = {"acc": "splineacc", "terrain": "terrain-", "rom": "leg-length", "dyn": "dynamic", "force": "force-",
                 "heel": "ee-dist", "height": "height-", "tottime": "contactduration-"}


def oracle_type_blocks(o):
    """{type name: (start, stop)} of the oracle's rows for its current stage."""
    out, off = {}, 0
    for name, rows in o.constraint_sets():
        for k, pre in ORACLE_PREFIX.items():
            if name.startswith(pre):
                a, b = out.get(k, (off, off))
                out[k] = (a, off + rows)
        off += rows
    return out


def master_to_oracle_perm(master_slices, o):
    """index array idx such that g_master[idx_master] aligns with g_oracle[idx_oracle] for the types the
    oracle stage has; returns (idx_master, idx_oracle)."""
    blocks = oracle_type_blocks(o)
    im, io = [], []
    for name, a, b in master_slices:
        if name in blocks:
            oa, ob = blocks[name]
            assert ob - oa == b - a, (name, oa, ob, a, b)
            im.append(np.arange(a, b))
            io.append(np.arange(oa, ob))
    return np.concatenate(im), np.concatenate(io)


def to_tau(M, blocks):
    """Columns (last axis) with respect to phase durations d -> columns with respect to switch times tau (d = D tau):
    col(tau_k) = col(d_k) - col(d_{k+1}) inside every PhaseDurations block (offset, count)."""
    M = np.array(M, dtype=float, copy=True)
    for off, cnt in blocks:
        for k in range(cnt - 1):
            M[..., off + k] -= M[..., off + k + 1]
    return M


def duration_blocks(p, n):
    """(offset, count) of the PhaseDurations sets of problem p inside an x of length n (they are stacked last)."""
    cnts = [len(d) - 1 for d in p.ee_durations]
    off = n - sum(cnts)
    out = []
    for c in cnts:
        out.append((off, c))
        off += c
    return out


GPU_STAGE_IDS = {"1.1": 0, "1.2": 1, "2.1": 2, "2.2": 3, "3": 4, "4": 5}
