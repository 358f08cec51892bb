#!/usr/bin/env python
"""Benchmark of the phys-optim hot path (BASELINE.json metric: optimised frames/sec of the batched staged
physics optimisation).

    python bench.py --gpus N --steps K --warmup W            # product arm (hand-written sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the oracle port of the reference algorithm

One "step" = one full staged solve (stages 1.1, 1.2, 2.1, 2.2, 4 of phys_optim.cpp:554-749) of a batch of
synthetic 120-frame / 2-end-effector sequences (BASELINE.json configs[1]: batch 64 on one B200; the batch is
sharded 64 per GPU under torchrun -> weak scaling, sequences are independent NLPs).

value : whole-job frames/s with the problem tables already resident in HBM (device-side reset of the iterate).
e2e   : the same metric through the public host API with host buffers: layout build + H2D + solve + D2H of the
        three SaveSolution snapshots inside the timed region.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES, N_EE, PER_GPU = 120, 2, 64


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.reasons, self.stop_flag = gpu, [], set(), False
        self.max_mhz = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def result(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def _oracle_solve_one(seed):
    import chd
    from oracle.phys import OracleProblem
    p = chd.synth.make_problem(seed, FRAMES, N_EE)
    t0 = time.perf_counter()
    o = OracleProblem(p)
    r = o.solve()
    dt = time.perf_counter() - t0
    return dt, [s["status"] for s in r["stages"]], [s["iters"] for s in r["stages"]]


def cpu_arm(n_seq, cores, seed0):
    """Times the CPU oracle (a port of the reference algorithm, NOT TOWR/ifopt/IPOPT/MA57) on `n_seq`
    sequences spread over `cores` processes.  Returns frames/s, wall seconds."""
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_oracle_solve_one, [seed0 + i for i in range(n_seq)])
    wall = time.perf_counter() - t0
    ok = sum(all(s == 0 for s in r[1]) for r in res)
    return n_seq * FRAMES / wall, wall, ok, [sum(r[2]) for r in res]


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_seq = min(cores, PER_GPU * world)
    for _ in range(args.warmup):
        pass  # a CPU solve has no warm-up effects worth paying minutes for
    vals, walls = [], []
    for _ in range(args.steps):
        v, wall, ok, iters = cpu_arm(n_seq, cores, 0)
        vals.append(v)
        walls.append(wall)
    v = float(np.mean(vals))
    sample = "%d sequences x %d frames (seeds 0..%d of the GPU arm's batch) per step, one per core" % (n_seq, FRAMES, n_seq - 1)
    line = {"impl": "reference", "metric": "optimised frames/sec (batched phys-optim)", "value": v, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batch %d synthetic %d-frame sequences, %d foot end-effectors, staged phys-optim (1.1,1.2,2.1,2.2,4)"
                                   % (PER_GPU * world, FRAMES, N_EE)},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU oracle = our C++ restatement of the reference NLP + chd-ipm; the reference's own TOWR/ifopt/IPOPT/MA57 "
                    "stack cannot be built offline (DESIGN.md)"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="chd")
    ap.add_argument("--per-gpu", type=int, default=PER_GPU)
    ap.add_argument("--cpu-sample", type=int, default=8, help="sequences of the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank, world, local = _env_int("RANK", 0), _env_int("WORLD_SIZE", 1), _env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    import chd
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.per_gpu
    problems = chd.synth.make_batch(B, FRAMES, N_EE, seed0=rank * B)
    batch = chd.phys.PhysBatch(problems, device=local)
    stride = 6 + 7 * N_EE
    fo = batch.dims["frames_out_max"]
    gather_src = torch.empty((B, fo, stride), dtype=torch.float64, device="cuda")
    gather_dst = torch.empty((world * B, fo, stride), dtype=torch.float64, device="cuda") if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        """inputs resident: device-side reset, staged solve, final sampling + the one NCCL gather"""
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        batch.reset()
        sstat = np.zeros((6, B), np.int32)
        siter = np.zeros((6, B), np.int32)
        succ = np.zeros((B, 2), np.int32)
        batch._chk(batch.L.chd_phys_solve(batch.h, None, None, succ.ctypes.data, sstat.ctypes.data, siter.ctypes.data))
        batch.L.chd_phys_sample_device(batch.h, gather_src.data_ptr(), torch.cuda.current_stream().cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(gather_dst.view(-1), gather_src.view(-1))
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3, sstat, siter, succ

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = batch.launch_count()
    batch.set_timing(False)
    t_steps, last = [], None
    for _ in range(args.steps):
        dt, sstat, siter, succ = step_resident()
        t_steps.append(dt)
        last = (sstat, siter, succ)
    barrier()
    launches = batch.launch_count() - l0
    t_total = torch.tensor([sum(t_steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
    t_total = float(t_total.item())
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # per-kernel device time of one more (untimed) step, serialised with CUDA events around every launch
    batch.set_timing(True)
    batch.kernel_times(reset=True)
    step_resident()
    kt = batch.kernel_times(reset=True)
    batch.set_timing(False)

    # e2e: public API with host buffers (layout build + H2D + solve + D2H samples), timed on the host around
    # fully synchronous calls; max over ranks
    def step_e2e():
        t0 = time.perf_counter()
        b2 = chd.phys.PhysBatch(problems, device=local)
        out = b2.solve()
        h2d = b2.h2d_bytes()
        d2h = out["samples"].nbytes + out["frames"].nbytes + out["success"].nbytes
        b2.close()
        return time.perf_counter() - t0, h2d, d2h
    step_e2e()
    barrier()
    e2e_t, h2d, d2h = [], 0, 0
    for _ in range(args.steps):
        dt, h2d, d2h = step_e2e()
        e2e_t.append(dt)
    barrier()
    e2e_total = torch.tensor([sum(e2e_t)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_total, op=dist.ReduceOp.MAX)
    e2e_total = float(e2e_total.item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
        frames_total = world * B * FRAMES * args.steps
        value = frames_total / t_total
        # dominant kernel: chd_k_kkt.  Algorithmic bytes per launch: per sequence
        # 8*(nslots [J] + 2n [grad, dx] + 12m [row state in/out]); the band itself is scratch.
        sz = batch.sizes.astype(np.int64)
        # every sequence walks through the schedule at its own pace, so a launch factorises only the sequences that are
        # still iterating: units per launch = (sum of iterations over sequences and stages) / launches
        seq_iters = np.asarray(last[1])[[0, 1, 2, 3, 5]].sum(axis=0).astype(np.float64)      # per sequence
        act = seq_iters / max(kt["kkt"][1], 1)                                            # share of the launches a sequence is active in
        running_launch_bytes = float((8 * (sz[:, 2] + 2 * sz[:, 0] + 12 * sz[:, 1]) * act).sum())
        kkt_ms, kkt_n = kt["kkt"]
        eval_ms, eval_n = kt["eval"]
        ach = running_launch_bytes / (kkt_ms / max(kkt_n, 1) * 1e-3) / 1e9 if kkt_n else 0.0
        eval_bytes = float((8 * (2 * sz[:, 0] + 2 * sz[:, 1] + sz[:, 2] + (18 + 3 * N_EE) * FRAMES)).sum())
        eval_ach = eval_bytes / (eval_ms / max(eval_n, 1) * 1e-3) / 1e9 if eval_n else 0.0
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r1_kkt_ncu.json")))
            traffic = prof["dram_bytes_per_sequence_per_launch"] * float(act.sum())   # ncu --set full capture (profiles/r1_kkt_ncu_summary.md), per launch
        except Exception:
            pass
        kkt_flops = float((act * sz[:, 3] * (sz[:, 5].astype(np.float64) ** 2 + 2.0 * sz[:, 5] * (sz[:, 4] + 1) + (sz[:, 4] + 1.0) ** 2)).sum())
        kkt_gflops = kkt_flops * 2 / (kkt_ms / max(kkt_n, 1) * 1e-3) / 1e9
        try:
            dfma_peak, dmma_peak = chd.phys.measure_fp64_peak()
        except Exception:
            dfma_peak = dmma_peak = None
        line = {
            "metric": "optimised frames/sec (batched phys-optim)", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batch %d synthetic %d-frame sequences, %d foot end-effectors, staged phys-optim (1.1,1.2,2.1,2.2,4) on %d B200"
                                   % (B * world, FRAMES, N_EE, world),
                       "per_gpu_batch": B, "l2": "256 MiB flush buffer written before every timed step",
                       "stage3": "duration optimisation not implemented; schedule takes the reference's stage-4 path"},
            "clocks": sampler.result(),
            "e2e": {"value": world * B * FRAMES * args.steps / e2e_total, "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "chd_k_kkt", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                         "frac": ach / hbm_peak, "traffic": traffic, "algorithmic_bytes": running_launch_bytes, "peak_source": peak_src,
                         "note": "fp64-FMA / latency bound kernel (DESIGN.md); HBM fraction reported as the contract asks",
                         "ms_per_launch": kkt_ms / max(kkt_n, 1), "active_sequences_per_launch": float(act.sum()), "fp64_gflops": kkt_gflops},
            "roofline_fp64": {"bound": "fp64 tensor core (DMMA m8n8k4)", "kernel": "chd_k_kkt", "achieved": kkt_gflops, "unit": "GFLOP/s",
                              "peak": dmma_peak, "peak_dfma": dfma_peak, "frac": (kkt_gflops / dmma_peak) if dmma_peak else None,
                              "peak_source": "chd_measure_fp64_peak, measured in this run (all SMs)",
                              "note": "band-dense LDL^T flop count (2*Na*(w+nb+1)^2 per active sequence); one CTA per sequence, so at most "
                                      "active_sequences_per_launch of the 148 SMs work"},
            "kernels": {k: {"ms": v[0], "launches": v[1]} for k, v in kt.items()},
            "roofline_eval": {"bound": "hbm", "kernel": "chd_k_eval", "achieved": eval_ach, "peak": hbm_peak, "unit": "GB/s",
                              "frac": eval_ach / hbm_peak},
            "residual": {"stage_status_ok_frac": [float((last[0][s] == 0).mean()) for s in (0, 1, 2, 3, 5)],
                         "iters_mean": [float(last[1][s].mean()) for s in (0, 1, 2, 3, 5)],
                         "success_frac": [float(last[2][:, 0].mean()), float(last[2][:, 1].mean())]},
        }
        if not args.no_cpu:
            cores = os.cpu_count() or 1
            n_seq = min(args.cpu_sample, B)
            use = min(cores, n_seq)
            v, wall, ok, iters = cpu_arm(n_seq, use, 0)
            line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": use, "kind": "port",
                                    "sample": "seeds 0..%d of the same batch (%d x %d frames), full staged solve, %.1f s wall, %d/%d converged"
                                              % (n_seq - 1, n_seq, FRAMES, wall, ok, n_seq)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
