#!/usr/bin/env python
"""Drop-in for the reference's `./phys_optim` executable (towr_phys_optim/phys_optim.cpp), same flags
(gflags syntax `--flag value` or `--flag=value`, phys_optim.cpp:23-31), same four input files and the same four
output files, solved on the GPU.  `scripts/run_phys_mocap.py:159-174` can point its `--towr-phys-optim-path` here.

Extension: `--in_dir` / `--out_dir` / `--nframes` accept comma separated lists so that many clips are solved as one
batch (that is where the GPU pays off); `--n_ee 2` selects the toes-only parameterisation.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv=None):
    ap = argparse.ArgumentParser(allow_abbrev=False)
    ap.add_argument("--out_dir", default="sol_out")
    ap.add_argument("--in_dir", default="./")
    ap.add_argument("--nframes", default="100")
    ap.add_argument("--w_com_lin", type=float, default=0.4)
    ap.add_argument("--w_com_ang", type=float, default=1.7)
    ap.add_argument("--w_ee", type=float, default=0.3)
    ap.add_argument("--w_smooth", type=float, default=0.1)
    ap.add_argument("--w_dur", type=float, default=0.1)
    ap.add_argument("--n_ee", type=int, default=4)
    args = ap.parse_args(argv)
    import chd
    in_dirs = args.in_dir.split(",")
    out_dirs = args.out_dir.split(",")
    nframes = [int(x) for x in str(args.nframes).split(",")]
    if len(nframes) == 1:
        nframes = nframes * len(in_dirs)
    assert len(in_dirs) == len(out_dirs) == len(nframes)
    print("Out Dir: %s\nInput Directory: %s\nnum frames: %s" % (args.out_dir, args.in_dir, args.nframes))
    print("Optim weights (%g, %g, %g, %g)\nDuration cost weight %g" % (args.w_com_lin, args.w_com_ang, args.w_ee, args.w_smooth, args.w_dur))
    problems = [chd.io_formats.read_phys_inputs(d, f, n_ee=args.n_ee) for d, f in zip(in_dirs, nframes)]
    weights = (args.w_com_lin, args.w_com_ang, args.w_ee, args.w_smooth, args.w_dur)
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        # torchrun: one process per GPU, sequences sharded by predicted work, one gather of the final trajectories;
        # only the durations snapshot travels, so the sharded mode writes sol_out_durations.txt + success_log.txt
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        out = chd.parallel.solve_sharded(problems, weights=weights, device=local, rank=rank, world=world)
        dist.destroy_process_group()
        if rank == 0:
            ne_max = max(p.n_ee for p in problems)
            for i, (p, od) in enumerate(zip(problems, out_dirs)):
                nf, n_ee = int(out["frames"][i]), p.n_ee
                cols = list(range(6)) + [6 + 3 * e + d for e in range(n_ee) for d in range(3)] + \
                    [6 + 3 * ne_max + 3 * e + d for e in range(n_ee) for d in range(3)] + [6 + 6 * ne_max + e for e in range(n_ee)]
                chd.io_formats.write_solution(os.path.join(od, "sol_out_durations.txt"), p.dt, out["samples"][i, :nf][:, cols], n_ee)
                chd.io_formats.write_success_log(os.path.join(od, "success_log.txt"), out["success"][i, 0], out["success"][i, 1])
        return
    batch = chd.phys.PhysBatch(problems, weights=weights)
    out = batch.solve()
    for i, (p, od) in enumerate(zip(problems, out_dirs)):
        chd.phys.write_outputs(out, i, p, od, batch.n_ee_max)
        print("[%d] stages status %s iterations %s -> %s" % (i, out["stage_status"][:, i].tolist(), out["stage_iters"][:, i].tolist(), od))


if __name__ == "__main__":
    main()
