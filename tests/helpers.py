"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import os
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

RTOL = 1e-5  # north_star: "agreement within 1e-5 relative (fp32)"


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def sub(d, prefix):
    return OrderedDict((k[len(prefix):], v) for k, v in d.items() if k.startswith(prefix))


def rel_err(a, ref):
    """Norm-wise relative error ||a-ref||_2 / ||ref||_2 (SURVEY.md section 7 'hard parts')."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.linalg.norm(ref.ravel())
    return np.linalg.norm((a - ref).ravel()) / (den if den > 0 else 1.0)


def max_err(a, ref):
    """Max abs error relative to the largest reference magnitude."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.abs(ref).max()
    return np.abs(a - ref).max() / (den if den > 0 else 1.0)


def assert_close(a, ref, tol=RTOL, what=""):
    assert np.asarray(a).shape == np.asarray(ref).shape, (what, np.asarray(a).shape, np.asarray(ref).shape)
    r, m = rel_err(a, ref), max_err(a, ref)
    assert r <= tol and m <= tol, "%s: rel %.3e max %.3e > %.1e" % (what, r, m, tol)
