"""Helpers shared by the parity tests: load the committed golden fixtures."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def msda_cases():
    return sorted(os.path.basename(p)[len("msda_"):-len(".npz")]
                  for p in glob.glob(os.path.join(GOLDEN, "msda_*.npz")))


def load_msda(name):
    with np.load(os.path.join(GOLDEN, f"msda_{name}.npz")) as z:
        return {k: z[k] for k in z.files}


def load_golden(filename):
    """any fixture under tests/golden/ as an NpzFile (lazy per-key loading)"""
    return np.load(os.path.join(GOLDEN, filename))
