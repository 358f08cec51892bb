#!/usr/bin/env python
"""Drop-in for the two command-line modes of the reference's `src/utils/towr_utils.py` (:898-1039):

processing (default): --anim --floor --contacts --out --character --start --end --fps [--no-heel]
    -> skel_info.txt, motion_info.txt, terrain_info.txt, contact_info.txt  (prepare_input)
--viz: --data <sol_out_*.txt ...> --out-bvh <bvh ...> --anim --character --start --end [--no-ik]
    -> every result applied back onto the skeleton (apply_results, IK) and saved as BVH.  The matplotlib rendering flags
       of the reference (--out-vid, --name, --forces, --trace, --skel, --compare-og, --hide, --plots ...) are accepted and ignored."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--viz", action="store_true")
    ap.add_argument("--anim", required=True)
    ap.add_argument("--floor")
    ap.add_argument("--contacts", default=None)
    ap.add_argument("--out")
    ap.add_argument("--character", default=None)
    ap.add_argument("--start", type=int, default=None)
    ap.add_argument("--end", type=int, default=None)
    ap.add_argument("--fps", default=30.0)
    ap.add_argument("--no-heel", dest="heel", action="store_false")
    ap.add_argument("--data", nargs="+")
    ap.add_argument("--out-bvh", nargs="+", default=None)
    ap.add_argument("--no-ik", dest="ik", action="store_false")
    ap.add_argument("--device", default=None)
    a, _ = ap.parse_known_args()
    import torch
    import chd
    info = chd.prepare.CHARACTERS[a.character]()
    dev = a.device or ("cuda" if torch.cuda.is_available() else None)
    if not a.viz:
        chd.prepare.prepare_input(a.anim, a.floor, a.contacts, a.out, info, a.start, a.end, 1.0 / float(a.fps), not a.heel, device=dev)
        return
    for i, path in enumerate(a.data):
        res = chd.results.load_towr_results(path, flip_coords=True)
        anim, names, _, _ = chd.results.apply_results(res, a.anim, a.start, a.end, info, run_ik=a.ik, device=dev)
        if a.out_bvh and i < len(a.out_bvh):
            if info.heel_inds is None and res.feet_pos.shape[1] == 4:   # towr_utils.py:972-974
                anim = chd.results.remove_heel_from_anim(anim)
            chd.results.save_bvh(a.out_bvh[i], anim, anim.names)


if __name__ == "__main__":
    main()
