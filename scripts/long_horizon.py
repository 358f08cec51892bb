#!/usr/bin/env python
"""Long-horizon configuration of BASELINE.json (600-frame sequences, 4 end-effectors, dense contact switches):
solve a batch on cuda:0 and report per-stage status / iterations / residuals and the wall time."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=600)
    ap.add_argument("--n_ee", type=int, default=4)
    ap.add_argument("--sparse", action="store_true", help="walking gait instead of dense switches")
    args = ap.parse_args()
    import chd
    problems = [chd.synth.make_problem(s, n_frames=args.frames, n_ee=args.n_ee, dense=not args.sparse) for s in range(args.batch)]
    t0 = time.time()
    batch = chd.phys.PhysBatch(problems)
    batch.set_timing(True)
    t1 = time.time()
    out = batch.solve()
    t2 = time.time()
    d = batch.dims
    print(json.dumps({
        "config": "%d x %d frames, %d ee, %s" % (args.batch, args.frames, args.n_ee, "walk" if args.sparse else "dense switches"),
        "dims": {k: int(d[k]) for k in ("n_max", "m_max", "na_max", "nb_max", "w_max")},
        "create_s": t1 - t0, "solve_s": t2 - t1, "frames_per_s": args.batch * args.frames / (t2 - t1),
        "stage_status": out["stage_status"].tolist(), "stage_iters": out["stage_iters"].tolist(),
        "success": out["success"].tolist(), "launches": int(batch.launch_count()),
        "kernels": {k: list(v) for k, v in batch.kernel_times().items()} if hasattr(batch, "kernel_times") else None,
    }))


if __name__ == "__main__":
    main()
