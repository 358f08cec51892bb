"""First-light script for the GPU box: solve a small batch, print per-stage statistics and kernel times."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import chd

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_ee = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ps = chd.synth.make_batch(B, 120, n_ee)
t0 = time.time()
b = chd.phys.PhysBatch(ps)
print("create %.3fs dims %s" % (time.time() - t0, b.dims))
print("sizes n,m,slots,Na,nb,w:", b.sizes[:4].tolist())
b.set_timing(True)
tot = 0
for st in ["1.1", "1.2", "2.1", "2.2", "3"]:
    t0 = time.time()
    r = b.solve_stage(st)
    dt = time.time() - t0
    tot += dt
    print("stage %s: %.3fs status %s iters %s" % (st, dt, np.bincount(r["status"] + 2, minlength=3).tolist(), r["iters"].tolist()[:16]))
    print("     f %s viol %s E0 %s lsfail %s" % (np.round(r["f"][:6], 5).tolist(), r["viol"][:6].tolist(), r["E0"][:6].tolist(), r["ls_fail"][:8].tolist()))
print("total solve time (timing mode, serialised) %.3fs" % tot)
print("kernel times (ms, launches):", b.kernel_times())
