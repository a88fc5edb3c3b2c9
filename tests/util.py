"""Parity metrics shared by the CPU and GPU tests."""
import numpy as np


def rel_err(got, ref):
    """norm-wise relative error over finite entries; non-finite / DBL_MAX entries must match exactly."""
    got, ref = np.asarray(got, float), np.asarray(ref, float)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    fin = np.isfinite(ref) & (np.abs(ref) < 1e300)
    assert np.array_equal(got[~fin], ref[~fin]), "infinite / DBL_MAX entries differ"
    if not fin.any():
        return 0.0
    return float(np.abs(got[fin] - ref[fin]).max() / max(np.abs(ref[fin]).max(), 1e-300))


def zeros_preserved(got, ref):
    """structural zeros of the oracle must be exact zeros"""
    got, ref = np.asarray(got), np.asarray(ref)
    return bool(np.all(got[ref == 0.0] == 0.0))


def x_err(got, ref):
    got, ref = np.asarray(got, float), np.asarray(ref, float)
    return float(np.abs(got - ref).max() / max(1.0, np.abs(ref).max()))


def active_set(iact, nact=None):
    iact = np.asarray(iact)
    if nact is not None:
        iact = iact[:nact]
    return set(int(i) for i in iact if i > 0)
