"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def assert_same_bits(a, b, what=""):
    """Bit-exact comparison (NaNs compare equal when their payload bits match)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} != {b.shape}"
    assert a.dtype == b.dtype, f"{what}: dtype {a.dtype} != {b.dtype}"
    if a.size == 0:
        return
    av = a.view(np.uint8)
    bv = b.view(np.uint8)
    if not np.array_equal(av, bv):
        bad = np.argwhere(a.view(f"u{a.dtype.itemsize}") != b.view(f"u{b.dtype.itemsize}"))
        raise AssertionError(f"{what}: {len(bad)} elements differ, first at {bad[0].tolist()}: "
                             f"{a[tuple(bad[0])]} vs {b[tuple(bad[0])]}")
