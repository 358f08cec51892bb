#!/usr/bin/env python
"""Convergence check on sequences outside the benchmark batch: solve seeds [a, b) in one batch, print status counts
and iteration statistics per stage."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import chd
a, b, n_ee = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 2
ps = [chd.synth.make_problem(s, n_ee=n_ee) for s in range(a, b)]
batch = chd.phys.PhysBatch(ps)
t0 = time.time(); out = batch.solve(); dt = time.time() - t0
st, it = out["stage_status"][[0, 1, 2, 3, 5]], out["stage_iters"][[0, 1, 2, 3, 5]]
print(json.dumps({"seeds": [a, b], "n_ee": n_ee, "solve_s": dt, "frames_per_s": (b - a) * 120 / dt,
                  "not_converged": int((st != 0).sum()), "iters_median": np.median(it, axis=1).tolist(), "iters_max": it.max(axis=1).tolist(),
                  "total_iters_max": int(it.sum(axis=0).max())}))
