import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, chd
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
b = chd.phys.PhysBatch(chd.synth.make_batch(B, 120, 2))
out = b.solve()
it = out["stage_iters"][[0, 1, 2, 3, 5]]
tot = it.sum(0)
print("total iterations per sequence: mean %.1f median %.1f max %d; sorted tail %s" % (tot.mean(), np.median(tot), tot.max(), np.sort(tot)[-8:].tolist()))
print("per stage max", it.max(1).tolist(), "mean", np.round(it.mean(1), 1).tolist())
worst = int(np.argmax(tot)); print("worst seq", worst, it[:, worst].tolist())
