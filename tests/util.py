"""Helpers shared by the tests: golden fixtures, oracle systems, engine systems."""
import json
import os

import numpy as np

from mcsolver_b200.lattice import build_tables
from tests.specs import spec_of

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name))


def tables_for(meta):
    return build_tables(spec_of(meta["spec"], tuple(meta["L"])), meta["T"], meta["model"])


def oracle_system(t, h_over_T=0.0, **flags):
    from oracle import oracle as orc
    return orc.System.from_tables(t, h_over_T=h_over_T, **flags)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) if a.size else 0.0


# result-tuple slots compared on one configuration / one trajectory; 7 (autoCorr) needs two sweeps.
ON_CORE_SLOTS = [0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 20, 21, 22, 23, 24, 25, 26]
ON_RG_SLOTS = [11, 12, 13, 14, 15, 16, 17, 18, 19]    # block-spin statistics (table path)
