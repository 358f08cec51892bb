"""A 69-joint (67 + two heel joints) skeleton with the Mixamo-style layout used by scripts/bench_kinematic.py."""


def ybot_like():
    names = ["Hips", "Spine", "Spine1", "Spine2", "Neck", "Head", "HeadTop_End", "LeftEye", "RightEye"]
    parents = [-1, 0, 1, 2, 3, 4, 5, 5, 5]
    off = [[0, 0, 0], [0, -10, -1], [0, -12, 0], [0, -13.5, 0], [0, -15, 0.5], [0, -10, 3], [0, -18, 0], [3, -8, 9], [-3, -8, 9]]
    for side, sx in (("Left", 1.0), ("Right", -1.0)):
        base = len(names)
        names += [side + n for n in ("Shoulder", "Arm", "ForeArm", "Hand")]
        parents += [3, base, base + 1, base + 2]
        off += [[sx * 6, -12, 0], [sx * 13, 0, 0], [sx, 27, 0], [sx * 0.5, 27, 1]]
        hand = base + 3
        for fi in range(5):
            for k in range(4):
                names.append("%sF%d_%d" % (side, fi, k))
                parents.append(hand if k == 0 else len(names) - 2)
                off.append([sx * (fi - 2) * 2.0, 7.0 - k, 0.5] if k == 0 else [0, 3.0, 0])
    for side, sx in (("Right", -1.0), ("Left", 1.0)):
        base = len(names)
        names += [side + n for n in ("UpLeg", "Leg", "Foot", "ToeBase", "Toe_End")]
        parents += [0, base, base + 1, base + 2, base + 3]
        off += [[sx * 9.2, 5.5, 0.2], [sx * 0.3, 40.5, 0.4], [sx * 0.1, 42, -0.8], [0, 10.5, 12.5], [0, 0, 7]]
    names += ["LeftHeel", "RightHeel"]
    parents += [64, 59]
    off += [[0, 10.5, 0], [0, 10.5, 0]]
    return names, parents, off
