"""The step on either side of the phys-optim hot path: from a skeleton animation (BVH), a floor
fit and per-frame foot contacts to the four `phys_optim` input files, and the parser of its output files.

Reference: `src/utils/towr_utils.py:451-777` (prepare_input), `:51-122` (load_results).  The reference leans on an
un-vendored BVH / Animation / Quaternions library; this module brings its own BVH reader and a batched forward
kinematics (torch when a device is given, numpy otherwise), and keeps the reference's conventions:

* units cm -> m, coordinate change (x, y, z)_towr = (-x, -z, -y)_bvh (towr_utils.py:517-524, 556-559, 690-693),
  a proper rotation C, so orientations map as R' = C R C^T (the reference flips / swaps the rotation axis, :608-611);
* COM = mass-fraction weighted mean of the segment centres (segment centre = mean of its joints, :503-512),
  hip offsets relative to the COM with the root rotation and translation zeroed (:489-519),
  inertia = point masses at the segment centres about the COM (:526-539);
* root orientation as Euler angles of R' = Rz Ry Rx (what towr's EulerConverter and `load_results` assume), unwrapped
  frame to frame (:612-621);
* heels: a dummy joint below each ankle at the toes' height (`add_heel_to_anim`, towr_utils.py:401-423);
* contact schedule: toe start flag = max(heel, toe) while the toe durations use the toe column only unless
  `combined_contacts` (:719-737); `foot_contacts.npy` columns are L heel, L toe, R heel, R toe.

The skeleton-specific part is a `CharacterInfo` record (joint indices of the leg chains, segment -> joints map, mass
percentages); `simple_biped_info` matches the skeleton `write_test_bvh` generates for the tests.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .io_formats import PhysProblem, find_contact_durations, write_phys_inputs

C_BVH_TO_TOWR = np.array([[-1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, -1.0, 0.0]])   # (x, y, z) -> (-x, -z, -y)
MALE_MASS = 73.0   # character_info_utils.py:181


@dataclass
class CharacterInfo:
    """What `character_info_utils.py` stores per character, restricted to what prepare_input reads."""
    left_leg_chain: List[int]            # hip, knee, ankle, toe
    right_leg_chain: List[int]
    segment_to_joints: Dict[str, List[int]]
    segment_mass_percent: Dict[str, float]
    mass: float = MALE_MASS
    heel_inds: Optional[Tuple[int, int]] = None   # skeletons that already carry heel joints
    upper_body_joints: Optional[List[int]] = None  # incl. the root: what apply_results pins to the optimised COM
    to_combined: Optional[List[int]] = None        # character joint -> joint of the `combined` skeleton (-1: none), re-targeting
    ik_blacklist: Optional[List[int]] = None       # mapped joints that are not IK targets when re-targeting

    @property
    def hips(self):
        return [self.left_leg_chain[0], self.right_leg_chain[0]]

    @property
    def toes(self):
        return [self.left_leg_chain[-1], self.right_leg_chain[-1]]

    @property
    def ankles(self):
        return [self.left_leg_chain[-2], self.right_leg_chain[-2]]


# Zatsiorsky - de Leva segment mass percentages, male (character_info_utils.py:143-160)
MASS_PERCENT_MALE = {"head": 6.94, "upper_trunk": 15.96, "mid_trunk": 16.33, "lower_trunk": 11.17, "left_upper_arm": 2.71,
                     "left_forearm": 1.62, "left_hand": 0.61, "left_thigh": 14.16, "left_shank": 4.33, "left_foot": 1.37,
                     "right_upper_arm": 2.71, "right_forearm": 1.62, "right_hand": 0.61, "right_thigh": 14.16, "right_shank": 4.33,
                     "right_foot": 1.37}


def _segments(head, ut, mt, lt, lua, lfa, lh, lth, lsh, lf, rua, rfa, rh, rth, rsh, rf):
    keys = list(MASS_PERCENT_MALE.keys())
    return dict(zip(keys, [list(v) for v in (head, ut, mt, lt, lua, lfa, lh, lth, lsh, lf, rua, rfa, rh, rth, rsh, rf)]))


def combined_info() -> "CharacterInfo":
    """The reference's `combined` skeleton (OpenPose body + SMPL spine, 28 joints, own heel joints):
    character_info_utils.py:257-287."""
    return CharacterInfo(left_leg_chain=[1, 2, 3, 5], right_leg_chain=[7, 8, 9, 11],
                         segment_to_joints=_segments([17], [15, 16], [14, 15], [13, 14], [22, 23], [23, 24], [24], [1, 2], [2, 3], [3, 4, 5, 6],
                                                     [25, 26], [26, 27], [27], [7, 8], [8, 9], [9, 10, 11, 12]),
                         segment_mass_percent=dict(MASS_PERCENT_MALE), heel_inds=(4, 10), upper_body_joints=[0] + list(range(13, 28)))


def ybot_info() -> "CharacterInfo":
    """The reference's Mixamo `ybot` character (67 joints, no heel joints): character_info_utils.py:294-321."""
    return CharacterInfo(left_leg_chain=[62, 63, 64, 65], right_leg_chain=[57, 58, 59, 60],
                         segment_to_joints=_segments([5], [3], [2], [1], [10, 11], [11, 12], range(12, 33), [62, 63], [63, 64], [64, 65, 66],
                                                     [34, 35], [35, 36], range(36, 57), [57, 58], [58, 59], [59, 60, 61]),
                         segment_mass_percent=dict(MASS_PERCENT_MALE), upper_body_joints=list(range(0, 57)),
                         to_combined=[0, 13, 14, 15, 16, -1, -1, 18, 20, -1, 22, 23, 24] + [-1] * 21 + [25, 26, 27] + [-1] * 20 +
                                     [7, 8, 9, 11, -1, 1, 2, 3, 5, -1],             # character_info_utils.py:319-388
                         ik_blacklist=[10, 11, 12, 34, 35, 36])


CHARACTERS = {"combined": combined_info, "ybot": ybot_info}

# A 28-joint template with the `combined` joint numbering (own proportions, cm, y pointing down like the camera frame): used
# by the synthetic clips of the tests / plumbing configuration; the reference ships skeleton_fitting/combined_body_25.bvh.
COMBINED_NAMES = ["Hips", "LHip", "LKnee", "LAnkle", "LHeel", "LBigToe", "LSmallToe", "RHip", "RKnee", "RAnkle", "RHeel", "RBigToe",
                  "RSmallToe", "Spine", "Spine1", "Spine2", "Neck", "Nose", "LEye", "LEar", "REye", "REar", "LShoulder", "LElbow", "LWrist",
                  "RShoulder", "RElbow", "RWrist"]
COMBINED_PARENTS = [-1, 0, 1, 2, 3, 3, 3, 0, 7, 8, 9, 9, 9, 0, 13, 14, 15, 16, 17, 18, 17, 20, 16, 22, 23, 16, 25, 26]
COMBINED_OFFSETS = [[0, 0, 0], [9.5, 3, 0.5], [0.4, 41, 0.8], [0.2, 40, -1.2], [0.3, 7.5, -5.5], [1.5, 8, 13.5], [-3.5, 8.2, 11],
                    [-9.5, 3, 0.5], [-0.4, 41, 0.8], [-0.2, 40, -1.2], [-0.3, 7.5, -5.5], [-1.5, 8, 13.5], [3.5, 8.2, 11],
                    [0, -11, -1], [0, -13, 0.5], [0, -13, 0.5], [0, -15, 1], [0, -12, 9], [3, -3, -2], [5, 1, -8], [-3, -3, -2], [-5, 1, -8],
                    [17, 1, -1], [2, 27, 0], [1, 25, 2], [-17, 1, -1], [-2, 27, 0], [-1, 25, 2]]


def write_combined_template(path: str):
    """Rest pose of the template as a one-frame BVH (what `--skel_path` of the kinematic optimiser takes)."""
    write_bvh(path, COMBINED_NAMES, COMBINED_PARENTS, COMBINED_OFFSETS, np.zeros((1, 3 + 3 * len(COMBINED_NAMES))), 1.0 / 30.0, order="ZXY")


@dataclass
class Bvh:
    names: List[str]
    parents: np.ndarray          # (J,) int, -1 for the root
    offsets: np.ndarray          # (J, 3) cm
    channels: List[List[str]]    # per joint, e.g. ["Xposition", ..., "Zrotation", "Xrotation", "Yrotation"]
    frame_time: float
    motion: np.ndarray           # (F, total channels)
    chan_off: List[int] = field(default_factory=list)

    @property
    def n_frames(self):
        return self.motion.shape[0]


def load_bvh(path: str) -> Bvh:
    """Plain-text BVH reader (HIERARCHY + MOTION).  End sites are dropped (the reference's loader does the same)."""
    toks = open(path).read().split()
    names, parents, offsets, channels = [], [], [], []
    stack, i, end_site = [], 0, False
    while toks[i] != "MOTION":
        t = toks[i]
        if t in ("ROOT", "JOINT"):
            names.append(toks[i + 1])
            parents.append(stack[-1] if stack else -1)
            offsets.append([0.0, 0.0, 0.0])
            channels.append([])
            cur = len(names) - 1
            i += 2
            continue
        if t == "End":
            end_site = True
            i += 2
            continue
        if t == "{":
            if not end_site:
                stack.append(len(names) - 1)
            i += 1
            continue
        if t == "}":
            if end_site:
                end_site = False
            else:
                stack.pop()
            i += 1
            continue
        if t == "OFFSET":
            if not end_site:
                offsets[stack[-1]] = [float(toks[i + 1]), float(toks[i + 2]), float(toks[i + 3])]
            i += 4
            continue
        if t == "CHANNELS":
            n = int(toks[i + 1])
            channels[stack[-1]] = toks[i + 2:i + 2 + n]
            i += 2 + n
            continue
        i += 1
    assert toks[i + 1] == "Frames:"
    F = int(toks[i + 2])
    assert toks[i + 3] == "Frame" and toks[i + 4] == "Time:"
    ft = float(toks[i + 5])
    data = np.array(toks[i + 6:], dtype=np.float64)
    nch = sum(len(c) for c in channels)
    motion = data.reshape(F, nch)
    chan_off = list(np.cumsum([0] + [len(c) for c in channels])[:-1])
    return Bvh(names, np.array(parents), np.array(offsets), channels, ft, motion, chan_off)


def _axis_rot(axis: str, ang):
    c, s, o, z = np.cos(ang), np.sin(ang), np.ones_like(ang), np.zeros_like(ang)
    if axis == "X":
        m = [[o, z, z], [z, c, -s], [z, s, c]]
    elif axis == "Y":
        m = [[c, z, s], [z, o, z], [-s, z, c]]
    else:
        m = [[c, -s, z], [s, c, z], [z, z, o]]
    return np.stack([np.stack(r, axis=-1) for r in m], axis=-2)


def local_transforms(b: Bvh, zero_root: bool = False):
    """Per frame and joint: local rotation matrices (F,J,3,3) -- channel order = multiplication order, degrees -- and
    local translations (F,J,3) (the OFFSET unless the joint has position channels)."""
    F, J = b.n_frames, len(b.names)
    R = np.tile(np.eye(3), (F, J, 1, 1))
    T = np.tile(b.offsets[None], (F, 1, 1)).astype(np.float64)
    for j in range(J):
        o = b.chan_off[j]
        for k, ch in enumerate(b.channels[j]):
            v = b.motion[:, o + k]
            if ch.endswith("position"):
                T[:, j, "XYZ".index(ch[0])] = v
            else:
                R[:, j] = R[:, j] @ _axis_rot(ch[0], np.radians(v))
    if zero_root:
        R[:, 0] = np.eye(3)
        T[:, 0] = 0.0
    return R, T


def forward_kinematics(parents, R, T, device=None):
    """Global joint positions (F,J,3) and global rotations (F,J,3,3) from local transforms.  Joints are processed in
    file order (parents precede children in a BVH); every step is batched over the frames -- on `device` (torch) if given."""
    J = len(parents)
    if device is not None:
        import torch
        Rt = torch.as_tensor(R, dtype=torch.float64, device=device)
        Tt = torch.as_tensor(T, dtype=torch.float64, device=device)
        gR, gP = [None] * J, [None] * J
        for j in range(J):
            p = int(parents[j])
            if p < 0:
                gR[j], gP[j] = Rt[:, j], Tt[:, j]
            else:
                gR[j] = gR[p] @ Rt[:, j]
                gP[j] = gP[p] + (gR[p] @ Tt[:, j].unsqueeze(-1)).squeeze(-1)
        return torch.stack(gP, dim=1).cpu().numpy(), torch.stack(gR, dim=1).cpu().numpy()
    gR = np.zeros_like(R)
    gP = np.zeros_like(T)
    for j in range(J):
        p = int(parents[j])
        if p < 0:
            gR[:, j], gP[:, j] = R[:, j], T[:, j]
        else:
            gR[:, j] = gR[:, p] @ R[:, j]
            gP[:, j] = gP[:, p] + np.einsum("fab,fb->fa", gR[:, p], T[:, j])
    return gP, gR


def to_towr(v):
    """(.., 3) BVH frame [cm] -> towr frame [m]: flip all axes, swap y / z, scale (towr_utils.py:517-524)."""
    return -np.asarray(v, dtype=np.float64)[..., [0, 2, 1]] * 0.01


def euler_zyx_from_matrix(R):
    """Euler angles (x, y, z) with R = Rz(z) Ry(y) Rx(x) (towr EulerConverter), batched over the leading axes."""
    y = -np.arcsin(np.clip(R[..., 2, 0], -1.0, 1.0))
    x = np.arctan2(R[..., 2, 1], R[..., 2, 2])
    z = np.arctan2(R[..., 1, 0], R[..., 0, 0])
    return np.stack([x, y, z], axis=-1)


def unwrap_euler(e):
    """towr_utils.py:612-621: keeps every angle within pi of its predecessor by adding multiples of 2 pi in the direction
    of the predecessor's sign."""
    e = e.copy()
    for d in range(3):
        cur = e[0, d]
        for f in range(1, e.shape[0]):
            step = 1.0 if cur >= 0.0 else -1.0
            nxt = e[f, d]
            while abs(nxt - cur) > np.pi:
                nxt += step * 2 * np.pi
            e[f, d] = nxt
            cur = nxt
    return e


def segment_centres(pos, info: CharacterInfo):
    """(F, n_segments, 3) mean joint position per body segment and the (n_segments,) mass fractions."""
    keys = list(info.segment_to_joints.keys())
    cen = np.stack([pos[:, info.segment_to_joints[k], :].mean(axis=1) for k in keys], axis=1)
    frac = np.array([info.segment_mass_percent[k] * 0.01 for k in keys])
    return cen, frac


def build_problem(bvh: Bvh, floor_normal, floor_point_cm, foot_contacts, info: CharacterInfo, start_idx=0, end_idx=None,
                  dt=1.0 / 30.0, combined_contacts=False, device=None) -> PhysProblem:
    """Everything prepare_input computes, as a `PhysProblem` (solver end-effector order L toe, R toe, L heel, R heel)."""
    F_all = bvh.n_frames
    end_idx = F_all if end_idx is None else end_idx
    sl = slice(start_idx, end_idx)
    # --- skeleton quantities with the root rotation / translation zeroed (towr_utils.py:483-539) ---
    R0, T0 = local_transforms(bvh, zero_root=True)
    pos0, _ = forward_kinematics(bvh.parents, R0, T0, device)
    cen0, frac = segment_centres(pos0, info)
    com0 = np.einsum("s,fsd->fd", frac, cen0)                                   # cm, root frame
    hips = to_towr(pos0[:, info.hips, :] - com0[:, None, :])                    # (F, 2, 3) m
    cen_c = to_towr(cen0 - com0[:, None, :])                                    # segment centres about the COM, m
    m_seg = frac * info.mass
    d2 = np.einsum("fsd,fsd->fs", cen_c, cen_c)
    I = np.einsum("s,fs->f", m_seg, d2)[:, None, None] * np.eye(3) - np.einsum("s,fsa,fsb->fab", m_seg, cen_c, cen_c)
    inertia = np.stack([I[:, 0, 0], I[:, 1, 1], I[:, 2, 2], I[:, 0, 1], I[:, 0, 2], I[:, 1, 2]], axis=1)
    chain = info.left_leg_chain
    max_leg = np.linalg.norm(bvh.offsets[chain[1:]], axis=1).sum() * 0.01       # hip -> toe bone lengths (:494-496)
    # --- animated motion (towr_utils.py:541-583) ---
    R, T = local_transforms(bvh)
    parents, offsets = bvh.parents, bvh.offsets
    if info.heel_inds is None:   # add_heel_to_anim: a joint below each ankle at the toe's vertical offset
        heel_off = np.zeros((2, 3))
        heel_off[:, 1] = offsets[info.toes, 1]
        parents = np.concatenate([parents, info.ankles])
        offsets = np.concatenate([offsets, heel_off], axis=0)
        R = np.concatenate([R, np.tile(np.eye(3), (R.shape[0], 2, 1, 1))], axis=1)
        T = np.concatenate([T, np.tile(heel_off[None], (T.shape[0], 1, 1))], axis=1)
        heels = (len(parents) - 2, len(parents) - 1)
    else:
        heels = info.heel_inds
    pos, gR = forward_kinematics(parents, R, T, device)
    pos_t = to_towr(pos)
    cen, _ = segment_centres(pos_t[:, :len(bvh.names)], info)
    com = np.einsum("s,fsd->fd", frac, cen)
    l_toe, r_toe = pos_t[:, info.toes[0]], pos_t[:, info.toes[1]]
    l_heel, r_heel = pos_t[:, heels[0]], pos_t[:, heels[1]]
    heel_dist = float(np.mean(np.linalg.norm(l_toe - l_heel, axis=1)))
    max_heel = (np.linalg.norm(bvh.offsets[chain[1:-1]], axis=1).sum() + np.linalg.norm(offsets[heels[0]])) * 0.01
    # root orientation in the towr frame (:608-621)
    Rt = C_BVH_TO_TOWR @ gR[:, 0] @ C_BVH_TO_TOWR.T
    root_e = unwrap_euler(euler_zyx_from_matrix(Rt))
    # --- floor (:686-709) ---
    normal = -np.asarray(floor_normal, dtype=np.float64)[[0, 2, 1]]
    point = to_towr(np.asarray(floor_point_cm, dtype=np.float64))
    # --- contact schedule (:711-777) ---
    fc = np.asarray(foot_contacts)[sl]
    toe_heel = fc[:, [1, 0, 3, 2]]                         # L toe, L heel, R toe, R heel
    either_l, either_r = fc[:, [0, 1]].max(axis=1), fc[:, [2, 3]].max(axis=1)
    starts = [toe_heel[0, 0] if combined_contacts else either_l[0], toe_heel[0, 1],
              toe_heel[0, 2] if combined_contacts else either_r[0], toe_heel[0, 3]]
    durs = [find_contact_durations(either_l if combined_contacts else toe_heel[:, 0], dt), find_contact_durations(toe_heel[:, 1], dt),
            find_contact_durations(either_r if combined_contacts else toe_heel[:, 2], dt), find_contact_durations(toe_heel[:, 3], dt)]
    order = [0, 2, 1, 3]                                   # file order (L toe, L heel, R toe, R heel) -> solver order
    ee = [l_toe, l_heel, r_toe, r_heel]
    return PhysProblem(dt=float(dt), hip_left=hips[sl, 0], hip_right=hips[sl, 1], max_leg_length=float(max_leg),
                       max_heel_length=float(max_heel), heel_dist=heel_dist, body_mass=float(info.mass), inertia=inertia[sl],
                       base_lin=com[sl], base_ang=root_e[sl], ee_pos=np.stack([ee[k][sl] for k in order]),
                       floor_normal=normal, floor_point=point, ee_start_contact=[int(starts[k]) for k in order],
                       ee_durations=[list(durs[k]) for k in order])


def prepare_input(anim_bvh: str, floor_file: str, contacts_file: str, out_dir: str, info: CharacterInfo, start_idx=None,
                  end_idx=None, dt=1.0 / 30.0, combined_contacts=False, device=None) -> Optional[PhysProblem]:
    """towr_utils.prepare_input: writes skel_info.txt, motion_info.txt, terrain_info.txt, contact_info.txt into out_dir."""
    for f, what in ((anim_bvh, "animated bvh"), (floor_file, "floor"), (contacts_file, "contacts")):
        if not os.path.exists(f):
            print("Could not find %s file %s" % (what, f))
            return None
    bvh = load_bvh(anim_bvh)
    with open(floor_file) as f:
        normal = [float(x) for x in f.readline().split()]
        point = [float(x) for x in f.readline().split()]
    p = build_problem(bvh, normal, point, np.load(contacts_file), info, start_idx or 0, end_idx, dt, combined_contacts, device)
    write_phys_inputs(p, out_dir)
    return p


def load_results(res_dir: str, nframes: Optional[int] = None) -> dict:
    """towr_utils.load_results (:51-122): the three solution files and the success log of an output directory."""
    from .io_formats import read_solution
    out = {}
    for key, name in (("no_dynamics", "sol_out_no_dynamics.txt"), ("dynamics", "sol_out_dynamics.txt"), ("durations", "sol_out_durations.txt")):
        path = os.path.join(res_dir, name)
        if os.path.exists(path):
            out[key] = read_solution(path)
    log = os.path.join(res_dir, "success_log.txt")
    if os.path.exists(log):
        vals = open(log).read().split()
        out["success"] = {vals[0]: int(vals[1]), vals[2]: int(vals[3])}
    return out


def write_bvh(path: str, names, parents, offsets, rows, frame_time: float, order: str = "ZXY", end_site=(0.0, 0.0, 0.0)):
    """BVH writer: root with 6 channels (XYZ position + rotations in `order`), 3 rotation channels per joint, joints in
    depth-first order (which must be the index order, as every BVH reader assumes); `rows` = (F, 3 + 3 J) channel values
    in degrees.  Six decimals like the reference's `BVH.save` (BVH.py:174-248), which also writes root-only positions."""
    J = len(names)
    children = {j: [c for c in range(J) if parents[c] == j] for j in range(J)}
    chan = " ".join(a + "rotation" for a in order)
    lines, seen = ["HIERARCHY"], []

    def emit(j, depth):
        seen.append(j)
        ind = "\t" * depth
        lines.append("%s%s %s" % (ind, "ROOT" if depth == 0 else "JOINT", names[j]))
        lines.append(ind + "{")
        lines.append("%s\tOFFSET %f %f %f" % ((ind,) + tuple(float(v) for v in offsets[j])))
        lines.append("%s\tCHANNELS %s" % (ind, ("6 Xposition Yposition Zposition " if depth == 0 else "3 ") + chan))
        if not children[j]:
            lines.extend([ind + "\tEnd Site", ind + "\t{", "%s\t\tOFFSET %f %f %f" % ((ind,) + tuple(end_site)), ind + "\t}"])
        for c in children[j]:
            emit(c, depth + 1)
        lines.append(ind + "}")

    emit(0, 0)
    if seen != list(range(J)):
        raise ValueError("joint indices must follow the depth-first order of the hierarchy")
    rows = np.asarray(rows, dtype=np.float64)
    lines += ["MOTION", "Frames: %d" % rows.shape[0], "Frame Time: %f" % frame_time]
    lines += [" ".join("%f" % v for v in r) for r in rows]
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


# ---------------------------------------------------------------------------------------------------------------------
# A small biped for the tests / the plumbing configuration (no motion-capture data ships with the reference).
# ---------------------------------------------------------------------------------------------------------------------
# joints in depth-first (= file) order
SIMPLE_JOINTS = ["Hips", "Spine", "Chest", "Head", "LeftArm", "LeftHand", "RightArm", "RightHand", "LeftUpLeg", "LeftLeg",
                 "LeftFoot", "LeftToe", "RightUpLeg", "RightLeg", "RightFoot", "RightToe"]
SIMPLE_PARENTS = [-1, 0, 1, 2, 2, 4, 2, 6, 0, 8, 9, 10, 0, 12, 13, 14]
# offsets in cm; like the reference's data the BVH frame has y pointing DOWN (the coordinate change (x, y, z) -> (-x, -z, -y)
# of towr_utils.py:517-524 then yields z up)
SIMPLE_OFFSETS = [[0, 0, 0], [0, -12, 0], [0, -25, 0], [0, -25, 0], [18, 0, 0], [0, 55, 0], [-18, 0, 0], [0, 55, 0],
                  [9, 4, 0], [0, 42, 0], [0, 42, 0], [0, 8, 14], [-9, 4, 0], [0, 42, 0], [0, 42, 0], [0, 8, 14]]


def simple_biped_info() -> CharacterInfo:
    return CharacterInfo(left_leg_chain=[8, 9, 10, 11], right_leg_chain=[12, 13, 14, 15],
                         segment_to_joints={"head": [3], "trunk": [0, 1, 2], "l_thigh": [8, 9], "l_shank": [9, 10], "l_foot": [10, 11],
                                            "r_thigh": [12, 13], "r_shank": [13, 14], "r_foot": [14, 15], "l_arm": [4, 5], "r_arm": [6, 7]},
                         segment_mass_percent={"head": 7.0, "trunk": 43.0, "l_thigh": 14.0, "l_shank": 4.5, "l_foot": 1.5,
                                               "r_thigh": 14.0, "r_shank": 4.5, "r_foot": 1.5, "l_arm": 5.0, "r_arm": 5.0})


def write_test_bvh(path: str, n_frames: int, seed: int = 0, fps: float = 30.0):
    """A walking-like clip of the simple biped: root translation + ZXY rotations per joint (degrees)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_frames) / fps
    lines = ["HIERARCHY"]
    children = {j: [c for c, p in enumerate(SIMPLE_PARENTS) if p == j] for j in range(len(SIMPLE_JOINTS))}

    def emit(j, depth):
        ind = "  " * depth
        lines.append("%s%s %s" % (ind, "ROOT" if j == 0 else "JOINT", SIMPLE_JOINTS[j]))
        lines.append(ind + "{")
        lines.append("%s  OFFSET %g %g %g" % ((ind,) + tuple(SIMPLE_OFFSETS[j])))
        if j == 0:
            lines.append(ind + "  CHANNELS 6 Xposition Yposition Zposition Zrotation Xrotation Yrotation")
        else:
            lines.append(ind + "  CHANNELS 3 Zrotation Xrotation Yrotation")
        if not children[j]:
            lines.append(ind + "  End Site")
            lines.append(ind + "  {")
            lines.append(ind + "    OFFSET 0 0 4")
            lines.append(ind + "  }")
        for c in children[j]:
            emit(c, depth + 1)
        lines.append(ind + "}")

    emit(0, 0)
    J = len(SIMPLE_JOINTS)
    rows = np.zeros((n_frames, 6 + 3 * (J - 1)))
    rows[:, 0] = 3.0 * np.sin(2 * np.pi * 0.9 * t)
    rows[:, 1] = -96.0 - 1.5 * np.sin(2 * np.pi * 1.8 * t)
    rows[:, 2] = 110.0 * t
    rows[:, 3:6] = np.stack([4 * np.sin(2 * np.pi * 0.9 * t), 3 * np.sin(2 * np.pi * 1.8 * t + 0.3), 12 * np.sin(2 * np.pi * 0.45 * t)], axis=1)
    swing = 28 * np.sin(2 * np.pi * 0.9 * t)
    col = lambda j: 6 + 3 * (j - 1)
    rows[:, col(8) + 1] = swing
    rows[:, col(12) + 1] = -swing
    rows[:, col(9) + 1] = 20 + 18 * np.cos(2 * np.pi * 0.9 * t)
    rows[:, col(13) + 1] = 20 - 18 * np.cos(2 * np.pi * 0.9 * t)
    rows[:, 6:] += rng.normal(0, 0.4, rows[:, 6:].shape)
    lines += ["MOTION", "Frames: %d" % n_frames, "Frame Time: %.7f" % (1.0 / fps)]
    lines += [" ".join("%.6f" % v for v in r) for r in rows]
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
