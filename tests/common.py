import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def golden_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "random_*.npz"))) + [os.path.join(GOLDEN, "config1.npz")]


def load(path):
    return np.load(path, allow_pickle=False)


def assert_matches_expected(g, got, label=""):
    """got: dict with score,best_dfs,best_j,num_best,has_unique (+ optional best_set*, node_scores)."""
    for k in ("score", "best_dfs", "best_j", "num_best", "has_unique"):
        a, b = np.asarray(got[k]).astype(np.int64), np.asarray(g["exp_" + k]).astype(np.int64)
        assert np.array_equal(a, b), f"{label}: {k} differs at samples {np.flatnonzero(a != b)[:8]}: {a[a != b][:8]} vs {b[a != b][:8]}"
    if "best_set" in got:
        assert np.array_equal(np.asarray(got["best_set_ptr"]).astype(np.int64), g["exp_best_set_ptr"].astype(np.int64)), f"{label}: best_set_ptr"
        assert np.array_equal(np.asarray(got["best_set"]).astype(np.int64), g["exp_best_set"].astype(np.int64)), f"{label}: best_set"
        assert np.array_equal(np.asarray(got["best_set_unique"]).astype(np.int64), g["exp_best_set_unique"].astype(np.int64)), f"{label}: best_set_unique"
    if "node_scores" in got:
        assert np.array_equal(got["node_scores"], g["exp_node_scores"]), f"{label}: node_scores"


def placements_to_dict(res):
    p = res["placements"]
    out = {"score": p["score"], "best_dfs": p["best_node"], "best_j": p["best_j"], "num_best": p["num_best"],
           "has_unique": p["has_unique"]}
    for k in ("best_set", "best_set_ptr", "best_set_unique", "node_scores"):
        if k in res:
            out[k] = res[k]
    return out
