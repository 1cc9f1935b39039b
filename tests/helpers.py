"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Stated tolerances: max-abs-diff / max-abs-ref per output column,
# against the reference's fp32 result, with its fp64 result as tie-breaker.
TOL_ORDER = {0: 1e-5, 1: 1e-5, 2: 1e-5, 3: 2e-5, 4: 5e-5}
TOL_RESIDUAL = 1e-4
TOL_GRAD = 1e-4
TOL_LOSS = 1e-5


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def split_state(z, prefix):
    """npz keys 'sol.Layers.0.weight' -> {'Layers.0.weight': tensor}."""
    out = {}
    for k in z.files:
        if k.startswith(prefix) and not k.endswith(".f64"):
            out[k[len(prefix):]] = torch.from_numpy(np.asarray(z[k]))
    return out


def rel_err(a, b):
    """max|a-b| / max|b| (b is the reference)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    if den == 0:
        return float(np.max(np.abs(a - b)))
    return float(np.max(np.abs(a - b)) / den)


def assert_close_to_reference(ours, ref32, ref64, tol, what):
    """Pass if within `tol` of the fp32 reference, or if closer to the fp64 reference than
    the fp32 reference itself is (then the disagreement is the reference's own rounding)."""
    e32 = rel_err(ours, ref32)
    if e32 <= tol:
        return
    if ref64 is not None:
        e64 = rel_err(ours, ref64)
        noise = rel_err(ref32, ref64)
        assert e64 <= max(tol, 1.5 * noise), \
            "%s: rel err vs fp32 ref %.3e, vs fp64 ref %.3e (ref noise %.3e), tol %.1e" % (what, e32, e64, noise, tol)
        return
    raise AssertionError("%s: rel err %.3e > tol %.1e" % (what, e32, tol))
