#!/usr/bin/env python
"""Generate golden vectors by running the REAL reference implementation.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (where /root/reference
is mounted); its outputs (small .npz files next to this script) are committed
and are what travels to the GPU box.  Nothing here is imported by the product.

The reference (Sarrasor/RoboticManipulators, pure Python: NumPy + SymPy) does
not run unmodified under numpy>=1.24 / without matplotlib / with its stale
SymPy-1.6.2 pickles, so we work on a scratch COPY under a temp dir (never in
this repo) with three non-arithmetic shims:
  1. numpy.float = float                       (np.float was removed)
  2. stub matplotlib / matplotlib.pyplot       (plotting side-car only)
  3. delete robots/data/*.pkl, data/*.pkl      (unpickle to zero matrices)

Usage:  python tests/golden/make_golden.py [--only fanuc,jac,...] [--jobs 8]
        python tests/golden/make_golden.py --only config1 [--config1-n 10000]   (config-1 set, ~10 min)
        python tests/golden/make_golden.py --only general --general-n 4000      (all-custom robot, ~3 min)
        python tests/golden/make_golden.py --only armslarge,trajlarge,chainrandom   (the *_large / random sets, < 1 min)
"""
import argparse
import os
import shutil
import sys
import tempfile
import types

import numpy as np

REF_SRC = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

CANONICAL_8 = [(m, k, w) for m in (1, -1) for k in (1, -1) for w in (0, 1)]
ALL_12 = [(m, k, w) for m in (1, -1) for k in (1, -1) for w in (0, 1, -1)]


def _install_shims():
    np.float = float  # shim 1
    mpl = types.ModuleType("matplotlib")
    mpl.rc = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")

    class _Ax:
        def __getattr__(self, name):
            return lambda *a, **k: None

    class _Fig:
        def add_subplot(self, *a, **k):
            return _Ax()

        def __getattr__(self, name):
            return lambda *a, **k: None

    plt.figure = lambda *a, **k: _Fig()
    plt.subplots = lambda *a, **k: (_Fig(), _Ax())
    for name in ("show", "ion", "ioff", "pause", "rc", "close", "plot"):
        setattr(plt, name, lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt  # shim 2
    m3d = types.ModuleType("mpl_toolkits")
    m3d2 = types.ModuleType("mpl_toolkits.mplot3d")
    m3d2.Axes3D = object
    sys.modules["mpl_toolkits"] = m3d
    sys.modules["mpl_toolkits.mplot3d"] = m3d2


def prepare_reference(workdir=None):
    """Copy the reference to a scratch dir, drop stale pickles, chdir there."""
    if workdir is None:
        workdir = tempfile.mkdtemp(prefix="rm_ref_")
    dst = os.path.join(workdir, "ref")
    if not os.path.isdir(dst):
        shutil.copytree(REF_SRC, dst,
                        ignore=shutil.ignore_patterns("images", ".git"))
        for root, _, files in os.walk(dst):
            for f in files:
                if f.endswith(".pkl"):
                    os.remove(os.path.join(root, f))  # shim 3
    _install_shims()
    os.chdir(dst)  # the robots open robots/data/*.pkl relative to the CWD
    if dst not in sys.path:
        sys.path.insert(0, dst)
    return dst


# --------------------------------------------------------------------------
# workers (module level so multiprocessing can pickle them)
# --------------------------------------------------------------------------
_ROBOTS = {}


def _fanuc(tag):
    from robots.fanuc import Fanuc165F
    if tag not in _ROBOTS:
        if tag == "default":
            _ROBOTS[tag] = Fanuc165F(save=False)
        elif tag == "general":
            # everything non-stock at once: lengths, offsets, directions, rotating base and tool
            from utils.robo_math import Transformation as tf
            Tb = tf("TxTyTzRzRyRx", [-350, 420, 150, -0.4, 0.15, 0.25])
            Tt = tf("TxTyTzRyRz", [60, -40, 210, 0.35, -0.2])
            _ROBOTS[tag] = Fanuc165F(T_base=np.array(Tb.transformation, dtype=float),
                                     T_tool=np.array(Tt.transformation, dtype=float), save=False,
                                     lengths=CUSTOM_LENGTHS, offsets=np.array(CUSTOM_OFFSETS),
                                     directions=np.array(CUSTOM_DIRECTIONS))
        elif tag == "custom":
            # non-stock link lengths, joint offsets and directions (fanuc.py:72-96)
            _ROBOTS[tag] = Fanuc165F(save=False, lengths=CUSTOM_LENGTHS,
                                     offsets=np.array(CUSTOM_OFFSETS),
                                     directions=np.array(CUSTOM_DIRECTIONS))
        else:
            T_base = np.array([[1, 0, 0, 500], [0, 1, 0, 300],
                               [0, 0, 1, -100], [0, 0, 0, 1]], dtype=float)
            T_tool = np.array([[1, 0, 0, 100], [0, 1, 0, 0],
                               [0, 0, 1, -300], [0, 0, 0, 1]], dtype=float)
            if tag == "rotated":
                # general (rotating) base and tool to exercise the full path
                from utils.robo_math import Transformation as tf
                Tb = tf("TxTyTzRzRyRx", [500, 300, -100, 0.3, -0.2, 0.1])
                Tt = tf("TxTzRxRz", [100, -300, 0.25, -0.4])
                T_base = np.array(Tb.transformation, dtype=float)
                T_tool = np.array(Tt.transformation, dtype=float)
            _ROBOTS[tag] = Fanuc165F(T_base=T_base, T_tool=T_tool, save=False)
    return _ROBOTS[tag]


def _init_worker(dst):
    _install_shims()
    os.chdir(dst)
    if dst not in sys.path:
        sys.path.insert(0, dst)
    import io
    sys.stdout = io.StringIO()  # the reference prints [INFO] lines on IK failure


def _w_fanuc_fk_ik(args):
    tag, q = args
    robot = _fanuc(tag)
    T = robot.forward_kinematics(list(q), plot=False)
    out = np.zeros((12, 6))
    for b, (m, k, w) in enumerate(ALL_12):
        out[b] = robot.inverse_kinematics(T, m=m, k=k, w=w)
    sys.stdout.seek(0), sys.stdout.truncate(0)
    return T, out


def _w_fanuc_ik_only(args):
    tag, T = args
    robot = _fanuc(tag)
    out = np.zeros((12, 6))
    for b, (m, k, w) in enumerate(ALL_12):
        out[b] = robot.inverse_kinematics(T, m=m, k=k, w=w)
    sys.stdout.seek(0), sys.stdout.truncate(0)
    return out


_JC = {}


def _w_fanuc_jac(args):
    tag, qm = args
    import sympy as sp
    from utils.jacobians import JacobianCalculator
    if tag not in _JC:
        indices = [0, 1, 0, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0]
        seq = "TzRzTzTxRyTxRyTzTxRxRyRxTx"
        robot = _fanuc(tag)
        _JC[tag] = JacobianCalculator(seq, indices, f_of_t=False,
                                      T_base=robot.T_base, T_tool=robot.T_tool)
    ls = (346.0, 324.0, 312.0, 1075.0, 225.0, 1280.0, 215.0)
    vals = [ls[0], qm[0], ls[1], ls[2], qm[1], ls[3], qm[2], ls[4], ls[5],
            qm[3], qm[4], qm[5], ls[6]]
    J = np.array(_JC[tag].calculate_numerically(vals), dtype=float)
    return J


def _rng_in_limits(rng, lim_deg, n):
    lim = np.deg2rad(np.array(lim_deg))
    return rng.uniform(lim[:, 0], lim[:, 1], size=(n, lim.shape[0]))


def _w_config1(args):
    """One config-1 pose through the real reference: FK, all 12 (m, k, w) IK branches of that
    pose, and calculate_numerically's 6x6 Jacobian at the model angles q' = q*dir + off."""
    q = args
    T, ik = _w_fanuc_fk_ik(("default", q))
    off = np.array([0.0, -np.pi / 2, np.pi / 2, 0.0, 0.0, 0.0])
    dirs = np.array([1.0, 1.0, -1.0, -1.0, -1.0, -1.0])
    J = _w_fanuc_jac(("default", q * dirs + off))
    return T, ik, J


def gen_config1(pool, out_dir, n):
    """BASELINE.json configs[0]: FK -> IK round trip + Jacobian on n (1e4) uniform-in-limits joint
    vectors, seed 0, through the REAL reference.  Stored: q, T[:, :3, :] (the 4th row is 0 0 0 1),
    the 12-branch IK answers (zeros = the reference's failure sentinel) and J."""
    rng = np.random.default_rng(0)
    q = _rng_in_limits(rng, FANUC_LIM_DEG, n)
    T = np.zeros((n, 3, 4))
    ik = np.zeros((n, 12, 6))
    J = np.zeros((n, 6, 6))
    done = 0
    for i, (Ti, iki, Ji) in enumerate(pool.imap(_w_config1, list(q), chunksize=8)):
        assert np.array_equal(Ti[3], [0, 0, 0, 1])
        T[i], ik[i], J[i] = Ti[:3], iki, Ji
        done += 1
        if done % 500 == 0:
            print("config1", done, "/", n, flush=True, file=sys.__stdout__)
    np.savez_compressed(os.path.join(out_dir, "fanuc_config1.npz"), q=q, T34=T, ik12=ik, J=J,
                        branches12=np.array(ALL_12), seed=np.array(0))
    print("config1 done", n, flush=True, file=sys.__stdout__)


FANUC_LIM_DEG = ((-185.0, 185.0), (-60.0, 76.0), (-156.0, 156.0),
                 (-360.0, 360.0), (-125.0, 125.0), (-360.0, 360.0))


CUSTOM_LENGTHS = (300.0, 350.0, 280.0, 900.0, 200.0, 1100.0, 180.0)
CUSTOM_OFFSETS = (0.1, -1.2, 1.4, 0.0, 0.2, -0.3)
CUSTOM_DIRECTIONS = (1, -1, 1, -1, 1, 1)


def gen_fanuc_custom(pool, out_dir, n_rand):
    """FK + 12-branch IK of a Fanuc165F built with custom lengths / offsets / directions."""
    rng = np.random.default_rng(5)
    q = np.vstack([np.zeros((1, 6)), rng.uniform(-1.0, 1.0, size=(n_rand, 6))])
    r = pool.map(_w_fanuc_fk_ik, [("custom", qi) for qi in q], chunksize=2)
    robot = _fanuc("custom")
    res = {"q": q, "T": np.array([x[0] for x in r]), "ik12": np.array([x[1] for x in r]),
           "branches12": np.array(ALL_12), "lengths": np.array(CUSTOM_LENGTHS),
           "offsets": np.array(CUSTOM_OFFSETS), "directions": np.array(CUSTOM_DIRECTIONS, dtype=float),
           "d": np.array(robot.d), "dq": np.array(robot.dq)}
    print("fanuc custom", len(q), flush=True, file=sys.__stdout__)
    np.savez_compressed(os.path.join(out_dir, "fanuc_custom.npz"), **res)


def _w_general(q):
    return _w_fanuc_fk_ik(("general", q))


def gen_fanuc_general(pool, out_dir, n):
    """FK + 12-branch IK through the REAL reference for a robot with NOTHING at its default: custom
    link lengths, joint offsets and directions, rotating base and tool transforms; n joint vectors
    uniform in the joint limits (seed 11)."""
    rng = np.random.default_rng(11)
    q = _rng_in_limits(rng, FANUC_LIM_DEG, n)
    T = np.zeros((n, 3, 4))
    ik = np.zeros((n, 12, 6))
    for i, (Ti, iki) in enumerate(pool.imap(_w_general, list(q), chunksize=8)):
        T[i], ik[i] = Ti[:3], iki
        if (i + 1) % 500 == 0:
            print("general", i + 1, "/", n, flush=True, file=sys.__stdout__)
    robot = _fanuc("general")
    np.savez_compressed(os.path.join(out_dir, "fanuc_general.npz"), q=q, T34=T, ik12=ik,
                        branches12=np.array(ALL_12), lengths=np.array(CUSTOM_LENGTHS),
                        offsets=np.array(CUSTOM_OFFSETS), directions=np.array(CUSTOM_DIRECTIONS, dtype=float),
                        T_base=np.array(robot.T_base, dtype=float), T_tool=np.array(robot.T_tool, dtype=float))
    print("general done", n, flush=True, file=sys.__stdout__)


def gen_fanuc(pool, out_dir, n_rand):
    rng = np.random.default_rng(0)
    structured = np.array([
        [0, 0, 0, 0, 0, 0],
        [np.pi / 2, 0, 0, 0, 0, 0],
        [-0.485, 0.5, -0.23, 0.23, 0.21, 0.213],       # fanuc_kinematics.py
        [0.3, -0.7, 0.9, -1.1, 0.8, 2.0],
        [1.570796, -0.28, 0.1, 0, 0.2, 0],             # tests/test_robots.py:36
        [0.1, 0.2, -0.3, 0.0, 0.0, 0.0],               # wrist-singular (r11 = 1)
        [0.5, 0.3, 0.2, 1.0, 0.0, -1.0],               # q4 = 0: singular, q3+q5 = 0
        [0.0, 0.0, 0.0, np.pi / 2, np.pi / 2, 0.0],    # r13 ~ 0 branch
    ], dtype=float)
    q = np.vstack([structured, _rng_in_limits(rng, FANUC_LIM_DEG, n_rand)])
    res = {}
    for tag in ("default", "translated", "rotated"):
        qs = q if tag == "default" else q[: 8 + n_rand // 4]
        r = pool.map(_w_fanuc_fk_ik, [(tag, qi) for qi in qs], chunksize=2)
        res[f"{tag}_q"] = qs
        res[f"{tag}_T"] = np.array([x[0] for x in r])
        res[f"{tag}_ik12"] = np.array([x[1] for x in r])
        print("fanuc", tag, len(qs), flush=True, file=sys.__stdout__)
    # IK-only targets from the reference's own tests (unreachable / limits)
    T_extra = np.array([
        [[1, 0, 0, 100000], [0, 1, 0, 2882], [0, 0, 1, 895], [0, 0, 0, 1]],
        [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 100], [0, 0, 0, 1]],
        [[1, 0, 0, 2000], [0, 1, 0, 0], [0, 0, 1, 900], [0, 0, 0, 1]],
        [[0, 0, 1, 1500], [0, 1, 0, 200], [-1, 0, 0, 1200], [0, 0, 0, 1]],
    ], dtype=float)
    r = pool.map(_w_fanuc_ik_only, [("default", T) for T in T_extra])
    res["extra_T"] = T_extra
    res["extra_ik12"] = np.array(r)
    res["branches12"] = np.array(ALL_12)
    # robot constants as the reference computed them
    robot = _fanuc("default")
    res["d"] = np.array(robot.d)
    res["dq"] = np.array(robot.dq)
    res["qs_lim_rad"] = np.array(robot.qs_lim_rad)
    for tag in ("translated", "rotated"):
        rb = _fanuc(tag)
        res[f"{tag}_T_base"] = np.array(rb.T_base, dtype=float)
        res[f"{tag}_T_tool"] = np.array(rb.T_tool, dtype=float)
    np.savez_compressed(os.path.join(out_dir, "fanuc_fk_ik.npz"), **res)


def gen_fanuc_jac(pool, out_dir, n_rand):
    import sympy as sp
    rng = np.random.default_rng(1)
    structured = np.array([
        [0.3, -0.2, 0.1, 0.23, 1.57, -0.2],                       # fanuc_jacobians.py:39
        [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],                           # :41 rotation singularity
        [np.pi, -np.pi / 2, 0.0, 0, np.pi, 0.0],                  # :43 full extension
        [0, -0.812734284248, -1.526091868267, np.pi, 2.37356282787, np.pi],  # :45 ballerina
    ])
    qm = np.vstack([structured, rng.uniform(-np.pi, np.pi, size=(n_rand, 6))])
    res = {"qm": qm}
    for tag in ("default", "rotated"):
        qs = qm if tag == "default" else qm[: 4 + n_rand // 4]
        J = pool.map(_w_fanuc_jac, [(tag, x) for x in qs], chunksize=2)
        res[f"{tag}_J"] = np.array(J)
        res[f"{tag}_n"] = np.array(len(qs))
        print("jac", tag, len(qs), flush=True, file=sys.__stdout__)
    # the reference's own singularity measure: symbolic J.evalf(subs).det()
    from utils.jacobians import JacobianCalculator
    variables = ['l_0', 'q_0', 'l_1', 'l_2', 'q_1', 'l_3', 'q_2', 'l_4', 'l_5',
                 'q_3', 'q_4', 'q_5', 'l_6']
    indices = [0, 1, 0, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0]
    jc = JacobianCalculator("TzRzTzTxRyTxRyTzTxRxRyRxTx", indices,
                            variables=variables, f_of_t=False, simplify=False)
    J_num = jc.calculate_numeric()
    ls = (346.0, 324.0, 312.0, 1075.0, 225.0, 1280.0, 215.0)
    dets = []
    for x in qm[:6]:
        vd = {sp.symbols(f"l_{i}"): ls[i] for i in range(7)}
        vd.update({sp.symbols(f"q_{i}"): x[i] for i in range(6)})
        dets.append(float(J_num.evalf(subs=vd).det()))
    res["det_symbolic_first6"] = np.array(dets)
    np.savez_compressed(os.path.join(out_dir, "fanuc_jacobian.npz"), **res)


def gen_generic_chain(out_dir):
    """Transformation + JacobianCalculator.calculate_numerically on assorted
    sequences (incl. prismatic joints, inverse tokens, base/tool)."""
    import sympy as sp
    from utils.robo_math import Transformation as tf
    from utils.jacobians import JacobianCalculator
    rng = np.random.default_rng(2)
    T_base = np.array([[1, 0, 0, 3], [0, 1, 0, 6], [0, 0, 1, -9], [0, 0, 0, 1.0]])
    T_tool = np.array([[0, 1, 0, 2], [1, 0, 0, -4], [0, 0, 1, 6], [0, 0, 0, 1.0]])
    cases = [
        ("RzTx", [1, 0], False),
        ("RzTxRzTx", [1, 0, 1, 0], False),
        ("RzTxRzTx", [1, 0, 1, 0], True),
        ("RzTxRzTxTyRz", [1, 1, 1, 1, 0, 1], False),
        ("RzTzRyTxRyTx", [1, 0, 1, 0, 1, 0], True),
        ("TzRzTzTxRyTxRyTzTxRxRyRxTx", [0, 1, 0, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0], False),
        ("RxiTyRziTxiRyTzi", [0, 1, 0, 0, 1, 0], True),
        ("TxTyTzRxRyRz", [1, 1, 1, 1, 1, 1], True),
    ]
    res = {"n_cases": np.array(len(cases))}
    for c, (seq, ind, with_bt) in enumerate(cases):
        ntok = len(ind)
        vals = rng.uniform(-2, 2, size=(6, ntok))
        Ts, Js = [], []
        for v in vals:
            T = np.array(tf(seq, list(v)).transformation, dtype=float)
            kw = dict(T_base=sp.Matrix(T_base), T_tool=sp.Matrix(T_tool)) if with_bt else {}
            jc = JacobianCalculator(seq, ind, **kw)
            J = np.array(jc.calculate_numerically(list(v)), dtype=float)
            Ts.append(T)
            Js.append(J)
        res[f"c{c}_seq"] = np.array(seq)
        res[f"c{c}_ind"] = np.array(ind)
        res[f"c{c}_bt"] = np.array(with_bt)
        res[f"c{c}_vals"] = vals
        res[f"c{c}_T"] = np.array(Ts)
        res[f"c{c}_J"] = np.array(Js)
    res["T_base"] = T_base
    res["T_tool"] = T_tool
    # tests/test_jacobians.py:59-75
    jc = JacobianCalculator("RzTxRzTx", [1, 0, 1, 0], simplify=False)
    res["test_numeric_values_J"] = np.array(
        jc.calculate_numerically((0.0, 1.0, 0.0, 2.0)), dtype=float)
    # derivative tokens and elementary matrices
    for tok in ("Tx", "Ty", "Tz", "Rx", "Ry", "Rz", "Txi", "Tyi", "Tzi",
                "Rxi", "Ryi", "Rzi", "Txd", "Tyd", "Tzd", "Rxd", "Ryd", "Rzd"):
        res[f"elem_{tok}"] = np.array(tf(tok, [0.7]).transformation, dtype=float)
    np.savez_compressed(os.path.join(out_dir, "generic_chain.npz"), **res)
    print("generic chain done", flush=True, file=sys.__stdout__)


def gen_generic_chain_random(out_dir, ncases=28):
    """Random token sequences (3 - 12 tokens, translations and rotations about every axis, inverse
    tokens as passive elements, random joint masks, with and without base / tool) through the REAL
    reference's Transformation and JacobianCalculator.calculate_numerically: every structure is a
    different specialised kernel on the GPU side."""
    import sympy as sp
    from utils.robo_math import Transformation as tf
    from utils.jacobians import JacobianCalculator
    rng = np.random.default_rng(21)
    T_base = np.array(tf("TxTyTzRzRx", [0.4, -0.3, 0.2, 0.5, -0.25]).transformation, dtype=float)
    T_tool = np.array(tf("TzRyTx", [0.15, 0.35, -0.1]).transformation, dtype=float)
    plain = ["Tx", "Ty", "Tz", "Rx", "Ry", "Rz"]
    res = {"n_cases": np.array(ncases), "T_base": T_base, "T_tool": T_tool}
    for c in range(ncases):
        ntok = int(rng.integers(3, 13))
        toks, ind = [], []
        for _ in range(ntok):
            t = plain[int(rng.integers(0, 6))]
            if rng.uniform() < 0.15:
                toks.append(t + "i"), ind.append(0)            # inverse tokens are never joints
            else:
                toks.append(t), ind.append(int(rng.uniform() < 0.55))
        if not any(ind):
            ind[[i for i, t in enumerate(toks) if not t.endswith("i")][0]] = 1
        seq, with_bt = "".join(toks), bool(c % 2)
        vals = rng.uniform(-2, 2, size=(8, ntok))
        kw = dict(T_base=sp.Matrix(T_base), T_tool=sp.Matrix(T_tool)) if with_bt else {}
        jc = JacobianCalculator(seq, ind, **kw)
        res[f"c{c}_seq"], res[f"c{c}_ind"], res[f"c{c}_bt"] = np.array(seq), np.array(ind), np.array(with_bt)
        res[f"c{c}_vals"] = vals
        res[f"c{c}_T"] = np.array([np.array(tf(seq, list(v)).transformation, dtype=float) for v in vals])
        res[f"c{c}_J"] = np.array([np.array(jc.calculate_numerically(list(v)), dtype=float) for v in vals])
    np.savez_compressed(os.path.join(out_dir, "generic_chain_random.npz"), **res)
    print("generic chain random done", flush=True, file=sys.__stdout__)


def _w_rrr(args):
    q = args
    from robots.rrr_robot import RRRRobot
    if "rrr" not in _ROBOTS:
        _ROBOTS["rrr"] = RRRRobot()
    r = _ROBOTS["rrr"]
    T = r.forward_kinematics(list(q), plot=False)
    out = np.zeros((4, 3))
    for b, (m, k) in enumerate([(1, 1), (1, -1), (-1, 1), (-1, -1)]):
        out[b] = r.inverse_kinematics(T, m=m, k=k)
    return T, out


def _w_rr(args):
    q, lengths = args
    from robots.rr_robot import RRRobot
    key = ("rr", lengths)
    if key not in _ROBOTS:
        _ROBOTS[key] = RRRobot(lengths=lengths, save=False)
    r = _ROBOTS[key]
    T = r.forward_kinematics(list(q), plot=False)
    out = np.zeros((2, 2))
    for b, m in enumerate((1, -1)):
        out[b] = r.inverse_kinematics(T, m=m)
    return T, out


def gen_rrr_rr(pool, out_dir, n_rand):
    rng = np.random.default_rng(3)
    q3 = np.vstack([[np.pi / 2, 0, 0.2], [0, 0, 0], [0.0, -np.pi / 2, 0.0],
                    rng.uniform(-np.pi, np.pi, size=(n_rand, 3))])
    r = pool.map(_w_rrr, list(q3), chunksize=4)
    res = {"rrr_q": q3, "rrr_T": np.array([x[0] for x in r]),
           "rrr_ik4": np.array([x[1] for x in r]),
           "rrr_branches": np.array([(1, 1), (1, -1), (-1, 1), (-1, -1)])}
    q2 = np.vstack([[0.3, 0.4], [0, 0], rng.uniform(-np.pi, np.pi, size=(n_rand, 2))])
    for name, lengths in (("rr", None), ("rr_l", (0.5, 1.25))):
        r = pool.map(_w_rr, [(x, lengths) for x in q2], chunksize=4)
        res[f"{name}_q"] = q2
        res[f"{name}_T"] = np.array([x[0] for x in r])
        res[f"{name}_ik2"] = np.array([x[1] for x in r])
    np.savez_compressed(os.path.join(out_dir, "rrr_rr.npz"), **res)
    print("rrr/rr done", flush=True, file=sys.__stdout__)


ARM_T_BASE_ARGS = ("TxTyTzRzRyRx", [0.5, 0.3, -0.1, 0.3, -0.2, 0.1])
ARM_T_TOOL_ARGS = ("TxTzRxRz", [0.1, -0.3, 0.25, -0.4])


def _arm_bt():
    from utils.robo_math import Transformation as tf
    return (np.array(tf(*ARM_T_BASE_ARGS).transformation, dtype=float),
            np.array(tf(*ARM_T_TOOL_ARGS).transformation, dtype=float))


def _w_rrr_bt(q):
    from robots.rrr_robot import RRRRobot
    if "rrr_bt" not in _ROBOTS:
        Tb, Tt = _arm_bt()
        _ROBOTS["rrr_bt"] = RRRRobot(T_base=Tb, T_tool=Tt)
    r = _ROBOTS["rrr_bt"]
    T = r.forward_kinematics(list(q), plot=False)
    out = np.zeros((4, 3))
    for b, (m, k) in enumerate([(1, 1), (1, -1), (-1, 1), (-1, -1)]):
        out[b] = r.inverse_kinematics(T, m=m, k=k)
    return T, out


def _w_rr_bt(q):
    from robots.rr_robot import RRRobot
    if "rr_bt" not in _ROBOTS:
        Tb, Tt = _arm_bt()
        _ROBOTS["rr_bt"] = RRRobot(T_base=Tb, T_tool=Tt, lengths=(0.5, 1.25), save=False)
    r = _ROBOTS["rr_bt"]
    T = r.forward_kinematics(list(q), plot=False)
    out = np.zeros((2, 2))
    for b, m in enumerate((1, -1)):
        out[b] = r.inverse_kinematics(T, m=m)
    return T, out


def gen_arms_base_tool(pool, out_dir, n_rand, name="arms_base_tool.npz"):
    """RRR and RR (lengths 0.5, 1.25) with rotating T_base / T_tool (robot.py:39-57)."""
    rng = np.random.default_rng(6)
    Tb, Tt = _arm_bt()
    q3 = np.vstack([[np.pi / 2, 0, 0.2], rng.uniform(-np.pi, np.pi, size=(n_rand, 3))])
    q2 = np.vstack([[0.3, 0.4], rng.uniform(-np.pi, np.pi, size=(n_rand, 2))])
    r3 = pool.map(_w_rrr_bt, list(q3), chunksize=4)
    r2 = pool.map(_w_rr_bt, list(q2), chunksize=4)
    res = {"T_base": Tb, "T_tool": Tt, "rr_lengths": np.array((0.5, 1.25)),
           "rrr_q": q3, "rrr_T": np.array([x[0] for x in r3]), "rrr_ik4": np.array([x[1] for x in r3]),
           "rrr_branches": np.array([(1, 1), (1, -1), (-1, 1), (-1, -1)]),
           "rr_q": q2, "rr_T": np.array([x[0] for x in r2]), "rr_ik2": np.array([x[1] for x in r2])}
    np.savez_compressed(os.path.join(out_dir, name), **res)
    print("arms base/tool done", flush=True, file=sys.__stdout__)


def gen_trajectories(out_dir):
    from utils.trapezoidal_generator import TrapezoidalGenerator
    from utils.trajectory_generator import TrajectoryGenerator
    from robots.rrr_robot import RRRRobot
    rng = np.random.default_rng(4)
    res = {}
    # coefficients (tests/test_trapezoidal_generator.py + random)
    cases = [(2, 2, 0, 0.0, 5.0), (2, 2, 5, 0.0, 5.0), (1, 20, 10, 0.0, -2.3),
             (1, 20, 0, 0.0, 0.01), (1, 10, 10, 1.0, 1.0 + 1e-3)]
    for _ in range(40):
        cases.append((float(rng.uniform(0.5, 3)), float(rng.uniform(1, 30)),
                      int(rng.choice([0, 5, 10, 100])),
                      float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3))))
    coef = []
    for v, a, f, q0, qf in cases:
        tg = TrapezoidalGenerator(v, a, control_freq=f)
        coef.append(tg.generate_coefficients(q0, qf))
    res["coef_cases"] = np.array(cases)
    res["coef_out"] = np.array(coef, dtype=float)

    # p2p joint trajectories, variable-length outputs
    p2p = []
    specs = [((0.0, 0.0, 0.0), (-0.9, -2.3, 1.2), n, cf)
             for n in (37, 50, 100, 1000, 1001) for cf in (10, 0)]
    for _ in range(12):
        specs.append((tuple(rng.uniform(-np.pi, np.pi, 3)),
                      tuple(rng.uniform(-np.pi, np.pi, 3)),
                      int(rng.choice([50, 200, 1000])), int(rng.choice([0, 10]))))
    for i, (q0, qf, n, cf) in enumerate(specs):
        tgen = TrajectoryGenerator(1, 20, 1, 10, control_freq=cf)
        tgen._tg.set_frequency(cf)
        tgen._tg.set_limits(1, 20)
        res[f"p2p{i}_spec"] = np.array(list(q0) + list(qf) + [n, cf], dtype=float)
        try:
            qs, dqs, ddqs, ts = tgen._generate_equalized_profiles(list(q0), list(qf), n)
        except (IndexError, ValueError):
            # too few samples for the acceleration phase: the reference raises
            # (IndexError at trapezoidal_generator.py:142, before its own ValueError)
            res[f"p2p{i}_error"] = np.array(1)
            continue
        res[f"p2p{i}_error"] = np.array(0)
        res[f"p2p{i}_q"] = np.array(qs).T
        res[f"p2p{i}_dq"] = np.array(dqs).T
        res[f"p2p{i}_ddq"] = np.array(ddqs).T
        res[f"p2p{i}_t"] = np.array(ts)
    res["p2p_count"] = np.array(len(specs))

    # Cartesian linear trajectory -> RRR IK (m=k=1) (rrr_robot_planning.py:44-48)
    robot = RRRRobot()
    lin_specs = [((1.0, 0.0, 2.0), (np.sqrt(2) / 2, np.sqrt(2) / 2, 1.2), 50, 10),
                 ((0.8, -0.5, 1.5), (-0.4, 1.1, 0.6), 120, 0),
                 ((1.2, 0.3, 0.4), (0.2, -1.0, 1.9), 200, 10)]
    import io
    for i, (p0, pf, n, cf) in enumerate(lin_specs):
        tgen = TrajectoryGenerator(1, 20, 1, 10, control_freq=cf)
        xs, dxs, ts = tgen.generate_lin_trajectory(list(p0), list(pf), n=n)
        qs = []
        for pt in xs:
            T_IK = np.array([[1, 0, 0, pt[0]], [0, 1, 0, pt[1]],
                             [0, 0, 1, pt[2]], [0, 0, 0, 1]])
            old = sys.stdout
            sys.stdout = io.StringIO()
            qs.append(robot.inverse_kinematics(T_IK))
            sys.stdout = old
        qs = np.array(qs)
        # dq = J_v(q)^-1 dx through the reference's numeric Jacobian
        from utils.jacobians import JacobianCalculator
        jc = JacobianCalculator("RzTzRyTxRyTx", [1, 0, 1, 0, 1, 0])
        dqs = []
        for q, dx in zip(qs, dxs):
            J = np.array(jc.calculate_numerically(
                [q[0], 1.0, q[1], 1.0, q[2], 1.0]), dtype=float)[:3, :3]
            try:
                dqs.append(np.linalg.inv(J).dot(dx))
            except np.linalg.LinAlgError:
                dqs.append(np.full(3, np.nan))
        res[f"lin{i}_spec"] = np.array(list(p0) + list(pf) + [n, cf], dtype=float)
        res[f"lin{i}_x"] = xs
        res[f"lin{i}_dx"] = dxs
        res[f"lin{i}_t"] = ts
        res[f"lin{i}_q"] = qs
        res[f"lin{i}_dq"] = np.array(dqs)
    res["lin_count"] = np.array(len(lin_specs))
    np.savez_compressed(os.path.join(out_dir, "trajectories.npz"), **res)
    print("trajectories done", flush=True, file=sys.__stdout__)


def gen_trajectories_large(out_dir, count=480):
    """`count` random point-to-point moves through the REAL reference's equalised trapezoid profiles
    (TrajectoryGenerator._generate_equalized_profiles, trajectory_generator.py:71-92), grouped by
    (n, control_freq) so that a batched call reproduces a whole group: sample COUNTS, times and
    q / dq / ddq are compared bit for bit.  Padded to n samples; `length` 0 = the reference raised."""
    from utils.trajectory_generator import TrajectoryGenerator
    rng = np.random.default_rng(14)
    res = {}
    groups = [(n, cf) for n in (40, 64, 100) for cf in (0, 10)]
    per = count // len(groups)
    for gi, (n, cf) in enumerate(groups):
        q0 = rng.uniform(-np.pi, np.pi, size=(per, 3))
        qf = rng.uniform(-np.pi, np.pi, size=(per, 3))
        qf[0] = q0[0] + rng.uniform(-2e-3, 2e-3, size=3)        # short moves: few samples or an error
        qf[1] = q0[1] + np.array([1e-2, -3.0, 0.5])
        q = np.zeros((per, n, 3))
        dq, ddq, t = np.zeros_like(q), np.zeros_like(q), np.zeros((per, n))
        length = np.zeros(per, dtype=np.int32)
        for i in range(per):
            tgen = TrajectoryGenerator(1, 20, 1, 10, control_freq=cf)
            tgen._tg.set_frequency(cf)
            tgen._tg.set_limits(1, 20)
            try:
                qs, dqs, ddqs, ts = tgen._generate_equalized_profiles(list(q0[i]), list(qf[i]), n)
            except (IndexError, ValueError):
                continue
            L = len(ts)
            length[i] = L
            q[i, :L], dq[i, :L], ddq[i, :L], t[i, :L] = np.array(qs).T, np.array(dqs).T, np.array(ddqs).T, ts
        res[f"g{gi}_spec"] = np.array([n, cf])
        res[f"g{gi}_q0"], res[f"g{gi}_qf"] = q0, qf
        res[f"g{gi}_q"], res[f"g{gi}_dq"], res[f"g{gi}_ddq"], res[f"g{gi}_t"] = q, dq, ddq, t
        res[f"g{gi}_len"] = length
    res["groups"] = np.array(len(groups))
    np.savez_compressed(os.path.join(out_dir, "trajectories_large.npz"), **res)
    print("trajectories large done", flush=True, file=sys.__stdout__)


def gen_calibration(out_dir, rows):
    """fanuc_calibration.py:11-91 on the real dataset (or its first `rows`)."""
    import io
    import scipy.io
    import sympy as sp
    from utils.model_calibration import ModelCalibrator
    import utils.model_calibration as mcmod
    mcmod.tqdm = lambda it, **k: it
    mat = scipy.io.loadmat('data/calibration_dataset.mat')
    qs = np.array(mat['q'], dtype=float)
    pts = np.concatenate([np.array(mat[k], dtype=float)[:, :, np.newaxis]
                          for k in ('mA', 'mB', 'mC')], axis=1) * 1000
    offsets_nominal = np.array([0.0, -np.pi / 2, np.pi / 2, 0.0, 0.0, 0.0])
    lengths_nominal = [('d_1', 670), ('d_2', 312), ('d_3', 1075),
                       ('d_5', 225), ('d_4', 1280), ("d_6", 215)]
    sequence = "TzRzTxRyTxRyTxTzRxRyRxTx"
    variables = ['d_1', 'q_1', 'd_2', 'q_2', 'd_3', 'q_3',
                 'd_4', 'd_5', 'q_4', 'q_5', 'q_6', 'd_6']
    joint_indices = [0, 1, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0]
    directions = np.array([1, 1, -1, -1, -1, -1])
    old = sys.stdout
    sys.stdout = io.StringIO()
    mc = ModelCalibrator(sequence, joint_indices=joint_indices,
                         variables=variables, link_lengths=lengths_nominal,
                         offsets=offsets_nominal, directions=directions)
    r_seq, r_vars = mc.get_reduced_model()
    T_base_0 = sp.Matrix([[1, 0, 0, -563.145620015451], [0, 1, 0, -559.28130331986],
                          [0, 0, 1, 554.98685687043], [0, 0, 0, 1]])
    T_tools_0 = [
        sp.Matrix([[1, 0, 0, 416.687065414667], [0, 1, 0, -17.0660344931335],
                   [0, 0, 1, 59.6890106614009], [0, 0, 0, 1]]),
        sp.Matrix([[1, 0, 0, 420.667529952858], [0, 1, 0, 57.1117905464326],
                   [0, 0, 1, -15.7733691549466], [0, 0, 0, 1]]),
        sp.Matrix([[1, 0, 0, 420.483792669375], [0, 1, 0, -43.7121193657286],
                   [0, 0, 1, -42.4708914215651], [0, 0, 0, 1]])]
    res = {"r_seq": np.array(r_seq), "r_vars": np.array(r_vars),
           "r_par": np.array(mc._r_par), "r_j_ind": np.array(mc._r_j_ind),
           "r_nom_par": np.array([str(x) for x in mc._r_nom_par]),
           "qs": qs, "pts": pts[:, :, 0], "rows": np.array(rows)}
    qs_i, pts_i = qs[:rows], pts[:rows]
    ev0 = mc.evaluate(qs_i, pts_i, T_base_0, T_tools_0)
    res["eval_initial"] = np.array([float(ev0[k]) for k in sorted(ev0)])
    res["eval_keys"] = np.array(sorted(ev0))
    # first-row probes
    res["pose_row0"] = np.array(mc.get_pose(qs[0], T_base_0, T_tools_0[0]), dtype=float)
    res["jac_row0"] = np.array(mc.get_jacobian(qs[0], T_base_0, T_tools_0[0]), dtype=float)

    # capture the normal equations exactly as the reference accumulated them
    captured = {}
    real_inv = np.linalg.inv

    def spy_inv(a):
        captured.setdefault("mats", []).append(np.array(a, dtype=float))
        return real_inv(a)

    np.linalg.inv = spy_inv
    T_base, T_tools = mc.estimate_tool_base(qs_i, pts_i)
    res["A_sum"] = captured["mats"][-1]
    res["base_est"] = np.array(mc._base_est, dtype=float)
    res["T_base"] = np.array(T_base, dtype=float)
    res["T_tools"] = np.array([np.array(t, dtype=float) for t in T_tools])
    ev1 = mc.evaluate(qs_i, pts_i, T_base, T_tools)
    res["eval_after_base"] = np.array([float(ev1[k]) for k in sorted(ev1)])
    params = mc.estimate_parameters(qs_i, pts_i, T_base, T_tools)
    np.linalg.inv = real_inv
    res["J_sum"] = captured["mats"][-1]
    res["params"] = np.array(params, dtype=float)
    res["Jdp_sum"] = res["J_sum"].dot(res["params"])  # g = H * (H^-1 g); stored for reference only
    ev2 = mc.evaluate(qs_i, pts_i, T_base, T_tools)
    res["eval_after_params"] = np.array([float(ev2[k]) for k in sorted(ev2)])
    sys.stdout = old
    name = "calibration.npz" if rows >= 240 else f"calibration_{rows}.npz"
    np.savez_compressed(os.path.join(out_dir, name), **res)
    print("calibration done", rows, flush=True, file=sys.__stdout__)

    # other models through the reducer
    sys.stdout = io.StringIO()
    red = {}
    for nm, seq, ind in (("rr", "RzTxRzTx", [1, 0, 1, 0]),
                         ("rrr", "RzTzRyTxRyTx", [1, 0, 1, 0, 1, 0]),
                         ("fanuc13", "TzRzTzTxRyTxRyTzTxRxRyRxTx",
                          [0, 1, 0, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0])):
        m2 = ModelCalibrator(seq, ind, None, None)
        s, v = m2.get_reduced_model()
        red[f"{nm}_seq"] = np.array(s)
        red[f"{nm}_vars"] = np.array(v)
        red[f"{nm}_par"] = np.array(m2._r_par)
        red[f"{nm}_j_ind"] = np.array(m2._r_j_ind)
        red[f"{nm}_nom"] = np.array([str(x) for x in m2._r_nom_par])
    sys.stdout = old
    np.savez_compressed(os.path.join(out_dir, "reduced_models.npz"), **red)


def gen_poly(out_dir):
    """utils/polynomial_generator.py + TrajectoryGenerator.generate_joint_poly_trajectory."""
    from utils.polynomial_generator import PolynomialGenerator
    from utils.trajectory_generator import TrajectoryGenerator
    rng = np.random.default_rng(6)
    pg = PolynomialGenerator()
    res = {}
    cases = [([0, 0, 0], [1, 0, 0], 1.0, 0.0)]        # tests/test_polynomial_generator.py
    for _ in range(30):
        cases.append((list(rng.uniform(-2, 2, 3)), list(rng.uniform(-2, 2, 3)),
                      float(rng.uniform(0.5, 4.0)), float(rng.choice([0.0, 0.0, 0.3, -0.5]))))
    coefs, q, dq, ddq = [], [], [], []
    n = 64
    for q0, qf, tf, t0 in cases:
        c = pg.generate_coefficients(q0, qf, t0 + tf, t0)
        coefs.append(c)
        q.append(pg.polynomial_from_coefs(c, t0, t0 + tf, n))
        dq.append(pg.dpolynomial_from_coefs(1, c, t0, t0 + tf, n))
        ddq.append(pg.dpolynomial_from_coefs(2, c, t0, t0 + tf, n))
    res["q0"] = np.array([c[0] for c in cases])
    res["qf"] = np.array([c[1] for c in cases])
    res["tf"] = np.array([c[2] + c[3] for c in cases])
    res["t0"] = np.array([c[3] for c in cases])
    res["coefs"], res["q"], res["dq"], res["ddq"] = map(np.array, (coefs, q, dq, ddq))
    tg = TrajectoryGenerator(1, 20, 1, 10)
    q_0 = [[0.5, 0, 0], [-0.6, 0, 0], [0, 0, 0]]       # rrr_robot_planning.py:22-25
    q_f = [[1.57, 0, 0], [0.5, 0, 0], [-2.0, 0, 0]]
    res["planning_qs"] = tg.generate_joint_poly_trajectory(q_0, q_f, 2, t_0=0, n=50, plot=False)
    np.savez_compressed(os.path.join(out_dir, "polynomial.npz"), **res)
    print("poly done", flush=True, file=sys.__stdout__)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="fanuc,custom,jac,chain,rrr,armsbt,traj,calib,poly")
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    ap.add_argument("--n-rand", type=int, default=56)
    ap.add_argument("--calib-rows", type=int, default=240)
    ap.add_argument("--out", default=HERE)
    ap.add_argument("--config1-n", type=int, default=10000,
                    help="poses of the config-1 set (only with --only config1; ~10 min on 8 cores)")
    ap.add_argument("--general-n", type=int, default=2000,
                    help="poses of the all-custom robot set (only with --only general; ~3 min on 8 cores)")
    args = ap.parse_args()
    only = set(args.only.split(","))
    dst = prepare_reference()
    import multiprocessing as mp
    pool = mp.Pool(args.jobs, initializer=_init_worker, initargs=(dst,))
    if "poly" in only:
        gen_poly(args.out)
    if "chain" in only:
        gen_generic_chain(args.out)
    if "chainrandom" in only:
        gen_generic_chain_random(args.out)
    if "traj" in only:
        gen_trajectories(args.out)
    if "trajlarge" in only:
        gen_trajectories_large(args.out)
    if "rrr" in only:
        gen_rrr_rr(pool, args.out, args.n_rand)
    if "armsbt" in only:
        gen_arms_base_tool(pool, args.out, args.n_rand)
    if "fanuc" in only:
        gen_fanuc(pool, args.out, args.n_rand)
    if "custom" in only:
        gen_fanuc_custom(pool, args.out, args.n_rand // 2)
    if "jac" in only:
        gen_fanuc_jac(pool, args.out, args.n_rand // 2)
    if "config1" in only:
        gen_config1(pool, args.out, args.config1_n)
    if "general" in only:
        gen_fanuc_general(pool, args.out, args.general_n)
    if "armslarge" in only:   # the same recipe as armsbt at 3000 joint vectors per arm (~2 min)
        gen_arms_base_tool(pool, args.out, 3000, name="arms_base_tool_large.npz")
    pool.close()
    if "calib" in only:
        gen_calibration(args.out, args.calib_rows)


if __name__ == "__main__":
    main()
