"""calibration_dataset.mat I/O (fanuc_calibration.py:13-20 of the reference): MAT-v5 file with
`q` (m x 6 joint angles, rad) and `mA`, `mB`, `mC` (m x 3 measured points, METRES).  Loading
returns the arrays the calibrator consumes -- qs [m,6] and pts [m,9] in MILLIMETRES (x1000,
:20) -- staged through pinned host memory onto the GPU."""
import numpy as np
import scipy.io
import torch

POINT_KEYS = ("mA", "mB", "mC")


def load_calibration_mat(path, device="cuda", point_keys=POINT_KEYS):
    """-> (qs [m,k], pts [m,3n]) float64 tensors on `device` (NumPy arrays if device is None)."""
    mat = scipy.io.loadmat(path)
    qs = np.ascontiguousarray(mat["q"], dtype=np.float64)
    pts = np.ascontiguousarray(
        np.concatenate([np.asarray(mat[k], dtype=np.float64) for k in point_keys], axis=1) * 1000)
    if device is None:
        return qs, pts
    out = []
    for a in (qs, pts):
        h = torch.from_numpy(a)
        if torch.cuda.is_available():
            h = h.pin_memory()
        out.append(h.to(device, non_blocking=True))
    return tuple(out)


def save_calibration_mat(path, qs, pts, point_keys=POINT_KEYS):
    """Write (qs [m,k], pts [m,3n] in millimetres) in the reference's file layout (metres)."""
    qs = qs.detach().cpu().numpy() if isinstance(qs, torch.Tensor) else np.asarray(qs)
    pts = pts.detach().cpu().numpy() if isinstance(pts, torch.Tensor) else np.asarray(pts)
    pts = pts.reshape(len(qs), -1)
    data = {"q": qs}
    for i, k in enumerate(point_keys):
        data[k] = pts[:, 3 * i:3 * i + 3] / 1000.0
    scipy.io.savemat(path, data)
