#!/usr/bin/env python
"""TEST / BASELINE INFRASTRUCTURE -- never imported by the product package.

Installs the UNMODIFIED reference (Sarrasor/RoboticManipulators, pure Python) into the git-ignored
directory baseline/_ref/ so that it travels to the GPU box and can be timed there on the host
cores (bench.py --impl reference).  The reference is a script collection without setup.py /
pyproject.toml, so "installing" it is a copy of its .py and data files; nothing is edited.  The
three things that keep it from running under numpy >= 1.24 / without matplotlib / with SymPy 1.14
 are handled OUTSIDE its sources, at run time, by `activate()`:
  1. numpy.float = float                  (np.float was removed from NumPy)
  2. stub matplotlib / mpl_toolkits       (plotting side-car, imported unconditionally)
  3. the stale SymPy-1.6.2 pickles (robots/data/*.pkl, data/*.pkl) are NOT copied: the classes
     regenerate them on first use (fanuc.py:120-149)

  python oracle/make_ref.py            # build baseline/_ref from /root/reference (if present)
"""
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference"
REF_DST = os.path.join(ROOT, "baseline", "_ref")


def build(force=False):
    """Copy the reference into baseline/_ref (no-op when the source tree is absent, e.g. on the
    GPU box, or when the copy exists).  Returns the install path or None."""
    if os.path.isdir(REF_DST) and not force:
        return REF_DST
    if not os.path.isdir(REF_SRC):
        return None
    if os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    shutil.copytree(REF_SRC, REF_DST,
                    ignore=shutil.ignore_patterns("images", ".git", "*.pkl", "__pycache__", "docs"))
    os.makedirs(os.path.join(REF_DST, "robots", "data"), exist_ok=True)
    os.makedirs(os.path.join(REF_DST, "data"), exist_ok=True)
    return REF_DST


def available():
    return os.path.isfile(os.path.join(REF_DST, "robots", "fanuc.py"))


def activate():
    """Make `import robots.fanuc` etc. resolve to baseline/_ref in THIS process (shims 1 and 2, the
    reference's relative data paths need the CWD)."""
    import numpy as np
    np.float = float
    mpl = types.ModuleType("matplotlib")
    mpl.rc = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")

    class _Any:
        def __getattr__(self, name):
            return lambda *a, **k: _Any()

    plt.figure = lambda *a, **k: _Any()
    plt.subplots = lambda *a, **k: (_Any(), _Any())
    for name in ("show", "ion", "ioff", "pause", "rc", "close", "plot"):
        setattr(plt, name, lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    tk = types.ModuleType("mpl_toolkits")
    tk3 = types.ModuleType("mpl_toolkits.mplot3d")
    tk3.Axes3D = object
    sys.modules["mpl_toolkits"] = tk
    sys.modules["mpl_toolkits.mplot3d"] = tk3
    os.chdir(REF_DST)
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)


# ------------------------------------------------------------------------------------------
# timing of the real reference on the host cores (bench.py --impl reference)
# ------------------------------------------------------------------------------------------
_STATE = {}


def _worker_init():
    import io
    activate()
    sys.stdout = io.StringIO()            # the reference prints [INFO] lines on IK failures
    from robots.fanuc import Fanuc165F
    from utils.jacobians import JacobianCalculator
    _STATE["robot"] = Fanuc165F(save=False)
    _STATE["jc"] = JacobianCalculator("TzRzTzTxRyTxRyTzTxRxRyRxTx",
                                      [0, 1, 0, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0], f_of_t=False)


def _worker_poses(qs):
    """The bench step on the reference's own API, pose by pose: forward_kinematics (fanuc.py:151),
    calculate_numerically Jacobian (jacobians.py:166) + determinant, and inverse_kinematics over
    the canonical 8 (m, k, w) branches (fanuc.py:196)."""
    import time
    import numpy as np
    robot, jc = _STATE["robot"], _STATE["jc"]
    ls = (346.0, 324.0, 312.0, 1075.0, 225.0, 1280.0, 215.0)
    off = np.array([0.0, -np.pi / 2, np.pi / 2, 0.0, 0.0, 0.0])
    dirs = np.array([1.0, 1.0, -1.0, -1.0, -1.0, -1.0])
    t0 = time.perf_counter()
    acc = 0.0
    for q in qs:
        T = robot.forward_kinematics(list(q), plot=False)
        qm = q * dirs + off
        vals = [ls[0], qm[0], ls[1], ls[2], qm[1], ls[3], qm[2], ls[4], ls[5], qm[3], qm[4], qm[5], ls[6]]
        J = np.array(jc.calculate_numerically(vals), dtype=float)
        acc += float(np.linalg.det(J)) * 0.0 + float(T[0, 3])
        for m in (1, -1):
            for k in (1, -1):
                for w in (0, 1):
                    acc += float(robot.inverse_kinematics(T, m=m, k=k, w=w)[0])
        sys.stdout.seek(0), sys.stdout.truncate(0)
    return time.perf_counter() - t0, acc


def time_reference(poses_per_core=3, procs=None, seed=0):
    """poses/s of the real reference on `procs` host processes (default: all cores); the robot
    construction (symbolic pre-expansion, ~3 s) is outside the timed region."""
    import multiprocessing as mp
    import time
    import numpy as np
    procs = procs or os.cpu_count() or 1
    rng = np.random.default_rng(seed)
    lim = np.deg2rad(np.array(((-185.0, 185.0), (-60.0, 76.0), (-156.0, 156.0), (-360.0, 360.0),
                               (-125.0, 125.0), (-360.0, 360.0))))
    q = rng.uniform(lim[:, 0], lim[:, 1], size=(procs * poses_per_core, 6))
    cwd = os.getcwd()
    try:
        with mp.get_context("fork").Pool(procs, initializer=_worker_init) as pool:
            pool.map(_worker_poses, [q[:1]] * procs)          # warm: first-call caches
            t0 = time.perf_counter()
            res = pool.map(_worker_poses, [q[i::procs] for i in range(procs)])
            wall = time.perf_counter() - t0
    finally:
        os.chdir(cwd)
    return {"value": len(q) / wall, "unit": "poses/s", "cores": procs, "kind": "reference",
            "sample": (f"{len(q)} uniform-in-limits poses ({poses_per_core} per process): the UNMODIFIED "
                       f"reference from baseline/_ref -- Fanuc165F.forward_kinematics + "
                       f"JacobianCalculator.calculate_numerically + det + 8 x inverse_kinematics per pose, "
                       f"{procs} processes, {wall:.1f} s wall (per-process busy {max(r[0] for r in res):.1f} s)")}


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    if "--time" in sys.argv:
        print(time_reference())
