"""Shared helpers of the parity tests: mapping between the product's master row order and the
oracle's ifopt row order (which follows phys_optim.cpp's per-stage AddConstraintSet order)."""
import numpy as np

ORACLE_PREFIX = {"acc": "splineacc", "terrain": "terrain-", "rom": "leg-length", "dyn": "dynamic", "force": "force-",
                 "heel": "ee-dist", "height": "height-"}


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
