#!/usr/bin/env python
"""Per-phase cycle counters of chd_k_kkt (CHD_PROF=1): solve the schedule stage by stage for one configuration."""
import os, sys, time
os.environ["CHD_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chd
nf, ne, dense, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ps = [chd.synth.make_problem(s, n_frames=nf, n_ee=ne, dense=bool(dense)) for s in range(B)]
b = chd.phys.PhysBatch(ps)
print(b.dims)
b.set_timing(True)
for st, mi in [("1.1", 50), ("1.2", 300), ("2.1", 7000), ("2.2", 2500), ("3", 2000)]:
    t0 = time.time()
    r = b.solve_stage(st, mi)
    print(st, "status", r["status"].tolist(), "iters", r["iters"].tolist(), "%.3f s" % (time.time() - t0))
print(b.kernel_times())
