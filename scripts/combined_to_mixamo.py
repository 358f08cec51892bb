#!/usr/bin/env python
"""Drop-in for `python skeleton_fitting/combined_to_mixamo.py --src_bvh --out_bvh --character` (run_phys_mocap.py:124-130).
The reference loads `<its directory>/<character>.bvh`; here the character's skeleton file is given with --skel_bvh."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None):
    ap = argparse.ArgumentParser("Retarget combined body25/smpl skeleton to a character skeleton")
    ap.add_argument("--src_bvh", required=True)
    ap.add_argument("--out_bvh", required=True)
    ap.add_argument("--character", default="ybot")
    ap.add_argument("--skel_bvh", required=True, help="the character's skeleton (the reference ships skeleton_fitting/<character>.bvh)")
    ap.add_argument("--device", default=None)
    a = ap.parse_args(argv)
    import torch
    import chd
    dev = a.device or ("cuda" if torch.cuda.is_available() else None)
    chd.results.retarget(a.src_bvh, a.skel_bvh, chd.prepare.CHARACTERS[a.character](), a.out_bvh, device=dev)
    print("Finished retargeting!")


if __name__ == "__main__":
    main()
