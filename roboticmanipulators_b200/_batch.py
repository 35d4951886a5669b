"""Layout helpers for the batched entry points.

Batched arrays come in two storage layouts (include/rk_b200.h):
  'aos'  pose-major, the reference's shapes on a leading batch axis ([N,6], [N,4,4], ...)
  'soa'  component-major planes ([6,N], [16,N], ...) -- the fast layout on B200: every
         warp access is a contiguous 256-byte segment.  SOA results are returned as
         strided VIEWS with the logical reference shape ([N,4,4] etc.), so indexing code
         written against the reference keeps working; `.contiguous()` gives AOS.
"""
import torch

from . import _lib


def detect_layout(x, ncomp):
    """'soa' if x ([N, ncomp] logical) is a transposed view of component planes [ncomp, ld]."""
    if x.dim() == 2 and x.shape[1] == ncomp and x.stride(0) == 1 and x.shape[0] > 1 \
            and x.stride(1) >= x.shape[0]:
        return "soa"
    return "aos"


def prep_input(x, ncomp, layout, dtype):
    """Return (tensor, ld) with storage in the requested layout (copying if needed).
    x logical shape [N, ncomp]."""
    _lib.require_cuda(x, "input")
    n = x.shape[0]
    if layout != "soa" and x.dtype is dtype and x.dim() == 2 and x.shape[1] == ncomp and x.is_contiguous():
        return x, n                     # the common case costs one check (call-latency path)
    x = x.to(dtype).reshape(n, ncomp)
    if layout == "soa":
        # Zero-copy only for DENSE planes (pitch == n).  The C ABI has ONE leading dimension for the
        # inputs and the outputs of a call and every caller allocates its outputs with pitch n, so a
        # column slice of wider planes (Q[:, lo:hi].t(), pitch > n) is compacted first: 8 bytes per
        # component and pose, against the >= 100 bytes per pose every kernel writes.
        if n > 1 and x.stride(0) == 1 and x.stride(1) == n:
            return x, n
        store = x.t().contiguous()          # [ncomp, N]
        return store.t(), n
    return x.contiguous(), n


def alloc_view(n, shape, layout, dtype, device):
    """One output with logical shape [n, *shape] in ONE torch call (call-latency path: measured on the
    GPU box, an allocation costs ~2.3 us of host time and every Python-level slice / view / permute
    another 1-2 us -- carving several outputs out of one slab was SLOWER than allocating each).
    SOA storage is component-major planes [prod(shape), n], returned directly as the strided view
    with the reference's logical shape; its data_ptr() is the start of the storage."""
    if layout != "soa" or not shape:
        return torch.empty((n, *shape), dtype=dtype, device=device)
    strides, acc = [], n
    for s_ in reversed(shape):
        strides.append(acc)
        acc *= s_
    return torch.empty_strided((n, *shape), (1, *reversed(strides)), dtype=dtype, device=device)


def flags_of(layout):
    if layout not in ("aos", "soa"):
        raise ValueError("layout must be 'aos', 'soa' or 'auto'")
    return _lib.RK_SOA if layout == "soa" else _lib.RK_AOS
