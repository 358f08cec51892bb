#!/usr/bin/env python
"""Host-side breakdown of one end-to-end batch (bench.py's e2e leg): create (layout + allocations + H2D), solve, close."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chd
ps = chd.synth.make_batch(64)
for it in range(4):
    t0 = time.perf_counter(); b = chd.phys.PhysBatch(ps)
    t1 = time.perf_counter(); out = b.solve()
    t2 = time.perf_counter(); b.close()
    t3 = time.perf_counter()
    print("create %.1f ms  solve %.1f ms  close %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
