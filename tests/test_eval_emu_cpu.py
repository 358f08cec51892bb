"""CPU check of the product's evaluation source (csrc/chd_eval.cuh compiled for the host with a one-thread shim,
tests/emu) against the oracle, including stage 3: run-time polynomial location, switch-time columns, TotalTime /
duration-bound rows and DurationCost.  The oracle differentiates with respect to the phase durations d
(the reference's variables); the product works with switch times tau (d = D tau), so oracle columns are mapped
with  col(tau_k) = col(d_k) - col(d_{k+1})."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SET_PREFIX = {0: "splineacc", 1: "terrain-", 2: "leg-length", 3: "dynamic", 4: "force-", 5: "ee-dist", 6: "height-",
              7: "contactduration-"}


def emu_lib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    L = C.CDLL(os.path.join(HERE, "emu", "libchd_emu.so"))
    L.chd_emu_eval.argtypes = [C.c_void_p] * 2 + [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 11
    return L


def emu_eval(chd, p, stage, x=None, dyn=0):
    L = emu_lib()
    b = chd.phys.PhysBatch.__new__(chd.phys.PhysBatch)
    arr, keep = chd.phys.make_problem_array([p])
    dims = np.zeros(4, np.int32)
    assert L.chd_emu_eval(C.addressof(arr), None, stage, None, 0, *([None] * 5), dims.ctypes.data, *([None] * 5)) == 0
    n, m, ns, nd = [int(v) for v in dims]
    out = dict(n=n, m=m, nslots=ns, n_dur=nd, cost=np.zeros(1), grad=np.zeros(n), g=np.zeros(m), Jv=np.zeros(ns),
               ent_col=np.zeros(ns, np.int32), ent_ptr=np.zeros(m + 1, np.int32), row_set=np.zeros(m, np.int32),
               row_lo=np.zeros(m), row_hi=np.zeros(m), dur_xoff=np.zeros(p.n_ee, np.int32))
    xx = None if x is None else np.ascontiguousarray(x, np.float64)
    ptr = lambda a: a.ctypes.data
    rc = L.chd_emu_eval(C.addressof(arr), None, stage, None if xx is None else ptr(xx), dyn, ptr(out["cost"]), ptr(out["grad"]),
                        ptr(out["g"]), ptr(out["Jv"]), ptr(out["ent_col"]), dims.ctypes.data, ptr(out["ent_ptr"]),
                        ptr(out["row_set"]), ptr(out["row_lo"]), ptr(out["row_hi"]), ptr(out["dur_xoff"]))
    assert rc == 0
    return out


from tests.util import to_tau  # noqa: E402


def compare(chd, p, stage_name, rng, perturb_dur):
    from oracle.phys import OracleProblem, STAGES
    o = OracleProblem(p)
    o.set_stage(stage_name)
    st = STAGES[stage_name]
    e0 = emu_eval(chd, p, st)
    n, nd = e0["n"], e0["n_dur"]
    xo = o.get_x()
    x = np.zeros(n)
    nn = n - nd
    x[:nn] = xo[:nn] + rng.normal(0, 0.02, nn)
    d0 = np.concatenate([np.asarray(d)[:-1] for d in p.ee_durations])
    x[nn:] = d0 + (rng.uniform(-0.04, 0.04, nd) if perturb_dur else 0.0)
    if stage_name == "3":
        o.set_x(x)
    else:
        o.set_x(x[:nn])
    e = emu_eval(chd, p, st, x, dyn=1 if (perturb_dur or stage_name == "3") else 0)
    blocks = []
    if stage_name == "3":
        off = nn
        for d in p.ee_durations:
            blocks.append((off, len(d) - 1))
            off += len(d) - 1
    # cost / gradient
    assert abs(e["cost"][0] - o.cost()) <= 1e-10 * max(1.0, abs(o.cost()))
    go = o.grad()
    ge = e["grad"][:len(go)]
    np.testing.assert_allclose(ge, to_tau(go, blocks), rtol=0, atol=1e-9 * max(1.0, np.abs(go).max()))
    # rows by set
    J = to_tau(np.asarray(o.jac().todense()), blocks)
    co = o.cons()
    names = o.constraint_sets()
    offs = np.cumsum([0] + [r for _, r in names])
    checked = 0
    for t, pre in SET_PREFIX.items():
        rows_m = np.nonzero(e["row_set"] == t)[0]
        rows_o = np.concatenate([np.arange(offs[i], offs[i + 1]) for i, (nm, _) in enumerate(names) if nm.startswith(pre)] or
                                [np.zeros(0, int)]).astype(int)
        if len(rows_o) == 0:
            continue
        assert len(rows_m) == len(rows_o), (t, len(rows_m), len(rows_o))
        np.testing.assert_allclose(e["g"][rows_m], co[rows_o], rtol=0, atol=1e-9 * max(1.0, np.abs(co[rows_o]).max()))
        for rm, ro in zip(rows_m, rows_o):
            dense = np.zeros(J.shape[1])
            for q in range(e["ent_ptr"][rm], e["ent_ptr"][rm + 1]):
                c = e["ent_col"][q]
                if 0 <= c < J.shape[1]:
                    dense[c] += e["Jv"][q]
                elif c >= J.shape[1]:
                    assert e["Jv"][q] == 0.0
            np.testing.assert_allclose(dense, J[ro], rtol=0, atol=1e-9 * max(1.0, np.abs(J[ro]).max()), err_msg="set %d row %d" % (t, rm))
            checked += 1
    assert checked == len(co)
    if stage_name == "3":   # duration lower bounds as rows
        rows = np.nonzero(e["row_set"] == 8)[0]
        np.testing.assert_allclose(e["g"][rows], x[nn:], rtol=0, atol=0)
    return e


@pytest.mark.parametrize("n_ee,seed", [(2, 0), (2, 5), (4, 1)])
@pytest.mark.parametrize("stage", ["2.2", "3"])
def test_eval_source_matches_oracle(chd, n_ee, seed, stage):
    p = chd.synth.make_problem(seed, n_ee=n_ee)
    rng = np.random.default_rng(100 + seed)
    compare(chd, p, stage, rng, perturb_dur=False)
    if stage == "3":
        compare(chd, p, stage, rng, perturb_dur=True)     # polynomial boundaries moved: run-time columns


def _fd_check(chd, p, seed):
    """Central differences of the product's own g(x) / cost(x) with respect to every switch time tau_k (moving
    tau_k by h changes d_k by +h and d_{k+1} by -h) against its analytic switch-time columns and gradient: an anchor
    that does not involve the oracle."""
    rng = np.random.default_rng(seed)
    e0 = emu_eval(chd, p, 4)
    n, nd = e0["n"], e0["n_dur"]
    nn = n - nd
    b = chd.phys.PhysBatch([p], host_only=True)
    x = b.get_x()[0, :n].copy()
    x[:nn] += rng.normal(0, 0.02, nn)
    x[nn:] += rng.uniform(-0.01, 0.01, nd)
    e = emu_eval(chd, p, 4, x, dyn=1)
    h = 1e-6
    cnts = [len(d) - 1 for d in p.ee_durations]
    off = nn
    worst = 0.0
    for c in cnts:
        for k in range(c):
            xp, xm = x.copy(), x.copy()
            xp[off + k] += h
            xm[off + k] -= h
            if k + 1 < c:
                xp[off + k + 1] -= h
                xm[off + k + 1] += h
            ep, em = emu_eval(chd, p, 4, xp, dyn=1), emu_eval(chd, p, 4, xm, dyn=1)
            col_fd = (ep["g"] - em["g"]) / (2 * h)
            col = np.zeros(e["m"])
            for r in range(e["m"]):
                for q in range(e["ent_ptr"][r], e["ent_ptr"][r + 1]):
                    if e["ent_col"][q] == off + k:
                        col[r] += e["Jv"][q]
            scale = max(1.0, np.abs(col_fd).max())
            worst = max(worst, np.abs(col - col_fd).max() / scale)
            g_fd = (ep["cost"][0] - em["cost"][0]) / (2 * h)
            assert abs(g_fd - e["grad"][off + k]) <= 1e-5 * max(1.0, abs(g_fd)), (k, g_fd, e["grad"][off + k])
        off += c
    assert worst <= 1e-5, worst


def test_switch_time_columns_finite_differences(chd):
    _fd_check(chd, chd.synth.make_problem(3, n_frames=60, n_ee=2), 0)


def test_switch_time_columns_finite_differences_random_points(chd):
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=4, deadline=None, derandomize=True)
    @given(st.integers(min_value=0, max_value=10_000))
    def run(seed):
        _fd_check(chd, chd.synth.make_problem(seed % 7, n_frames=40, n_ee=2 if seed % 2 else 4), seed)
    run()
