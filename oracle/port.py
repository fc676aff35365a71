"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/liboracle_port.so (oracle/usher_port.c, the plain-C
restatement of mapper2_body + the per-sample search).  Same call shape as oracle/ref.py's RefTree.search.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle_port.so")
MUT_DTYPE = np.dtype(
    [("position", "<i4"), ("ref_nuc", "u1"), ("par_nuc", "u1"), ("mut_nuc", "u1"), ("is_missing", "u1")]
)
_lib = None


def build():
    src = os.path.join(HERE, "usher_port.c")
    if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        L.port_tree_create.restype = vp
        L.port_tree_create.argtypes = [u32, vp, vp, vp]
        L.port_tree_free.argtypes = [vp]
        L.port_tree_bfs.argtypes = [vp, vp, vp]
        L.port_search.restype = C.c_int
        L.port_search.argtypes = [vp, u32, vp, vp, C.c_int] + [vp] * 9 + [u64]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class PortTree:
    def __init__(self, parent, row_ptr, muts):
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        self.muts = np.ascontiguousarray(muts, dtype=MUT_DTYPE)
        self.n = len(self.parent)
        self.h = lib().port_tree_create(self.n, _p(self.parent), _p(self.row_ptr), _p(self.muts))

    def close(self):
        if self.h:
            lib().port_tree_free(self.h)
            self.h = None

    def bfs(self):
        bfs = np.zeros(self.n, np.uint32)
        nl = np.zeros(self.n, np.uint32)
        lib().port_tree_bfs(self.h, _p(bfs), _p(nl))
        return bfs, nl

    def search(self, s_ptr, sm, per_node=False, want_set=True, set_cap=None):
        s_ptr = np.ascontiguousarray(s_ptr, dtype=np.uint64)
        sm = np.ascontiguousarray(sm, dtype=MUT_DTYPE)
        B = len(s_ptr) - 1
        out = {
            "score": np.zeros(B, np.int32),
            "best_dfs": np.zeros(B, np.uint32),
            "best_j": np.zeros(B, np.uint32),
            "num_best": np.zeros(B, np.uint32),
            "has_unique": np.zeros(B, np.uint8),
        }
        node_scores = np.zeros((B, self.n), np.int32) if per_node else None
        cap = int(set_cap if set_cap is not None else max(1024, 64 * B)) if want_set else 0
        bset = np.zeros(cap, np.uint32) if want_set else None
        bset_u = np.zeros(cap, np.uint8) if want_set else None
        bptr = np.zeros(B + 1, np.uint64) if want_set else None
        rc = lib().port_search(
            self.h, B, _p(s_ptr), _p(sm), 1 if per_node else 0, _p(out["score"]), _p(out["best_dfs"]),
            _p(out["best_j"]), _p(out["num_best"]), _p(out["has_unique"]), _p(node_scores), _p(bset),
            _p(bset_u), _p(bptr), cap,
        )
        if want_set:
            if rc != 0:
                return self.search(s_ptr, sm, per_node, True, int(bptr[-1]) + 16)
            out["best_set_ptr"] = bptr
            out["best_set"] = bset[: int(bptr[-1])]
            out["best_set_unique"] = bset_u[: int(bptr[-1])]
        if per_node:
            out["node_scores"] = node_scores
        return out
