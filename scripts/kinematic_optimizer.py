#!/usr/bin/env python
"""Drop-in for the reference's `src/optimize/kinematic_optimizer.py` command line (flags :297-333): reads
`<dir>/openpose_result/*.json`, `<dir>/tracked_results.json`, `<dir>/foot_contacts.npy` next to --input_path and writes
`foot_contacts.npy`, `floor_out.txt`, `final_test.bvh` into --output_path.  Visualisation flags are accepted and ignored."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--input_path", default="../data/example_data/dance1/dance1.mp4")
    ap.add_argument("--output_path", default="../data/example_data/dance1/kinematic_results")
    ap.add_argument("--skel_path", default="skeleton_fitting/combined_body_25.bvh")
    ap.add_argument("--start", type=int, default=0)
    ap.add_argument("--end", type=int, default=100)
    ap.add_argument("--visualize", action="store_true")
    ap.add_argument("--viz-only", dest="viz_only", action="store_true")
    ap.add_argument("--gt-floor", dest="use_gt_floor", action="store_true")
    ap.add_argument("--character", default="ybot")
    ap.add_argument("--device", default=None, help="torch device of the batched solver (default: cuda if available)")
    a = ap.parse_args()
    import torch
    import chd
    dev = a.device or ("cuda" if torch.cuda.is_available() else None)
    chd.kinopt.optimize_2d_3d(a.input_path, a.skel_path, a.output_path, a.start, a.end, a.use_gt_floor, device=dev)


if __name__ == "__main__":
    main()
