"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/_ref/libusher_ref.so (the reference's own
mapper2_body + usher_common compiled from /root/reference/src, see oracle/Makefile and oracle/ref_driver.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arm may import this.
Nothing under usher_b200/ does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libusher_ref.so")

# numpy view of `struct ref_mut` (oracle/ref_driver.cpp) == `ub200_mutation` (include/usher_b200.h)
MUT_DTYPE = np.dtype(
    [("position", "<i4"), ("ref_nuc", "u1"), ("par_nuc", "u1"), ("mut_nuc", "u1"), ("is_missing", "u1")]
)

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(
                f"{LIB_PATH} missing: run `make -C oracle ref` in a container that mounts /root/reference"
            )
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
        L.usher_ref_tree_from_flat.restype = vp
        L.usher_ref_tree_from_flat.argtypes = [u32, vp, vp, vp]
        L.usher_ref_tree_free.argtypes = [vp]
        L.usher_ref_tree_from_newick_vcf.restype = vp
        L.usher_ref_tree_from_newick_vcf.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        L.usher_ref_tree_parsimony.restype = u64
        L.usher_ref_tree_parsimony.argtypes = [vp]
        L.usher_ref_tree_export.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
        L.usher_ref_read_samples.restype = u32
        L.usher_ref_read_samples.argtypes = [vp, C.c_char_p]
        L.usher_ref_samples_export.argtypes = [vp, vp, vp, vp, vp, vp]
        L.usher_ref_usher_common.restype = C.c_int
        L.usher_ref_usher_common.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.usher_ref_search.restype = C.c_int
        L.usher_ref_search.argtypes = [vp, u32, vp, vp, C.c_int, C.c_int] + [vp] * 9 + [u64, vp]
        L.usher_ref_condensed_export.restype = u64
        L.usher_ref_condensed_export.argtypes = [vp, vp, u64]
        L.usher_ref_search_strided.restype = C.c_double
        L.usher_ref_search_strided.argtypes = [vp, u64, vp, u32, u32, C.c_int, vp]
        L.usher_ref_search_strided2.restype = C.c_double
        L.usher_ref_search_strided2.argtypes = [vp, u64, vp, u32, u32, C.c_int, C.c_int32, vp]
        L.usher_ref_score_nodes.restype = C.c_int
        L.usher_ref_score_nodes.argtypes = [vp, u64, vp, u64, vp, C.c_int, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RefTree:
    """A reference MAT::Tree living inside libusher_ref.so."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def from_flat(cls, parent, row_ptr, muts):
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        muts = np.ascontiguousarray(muts, dtype=MUT_DTYPE)
        return cls(lib().usher_ref_tree_from_flat(len(parent), _p(parent), _p(row_ptr), _p(muts)))

    @classmethod
    def from_newick_vcf(cls, newick, vcf, condense_and_round_trip=True, threads=1):
        return cls(
            lib().usher_ref_tree_from_newick_vcf(
                newick.encode(), vcf.encode(), int(condense_and_round_trip), threads
            )
        )

    def close(self):
        if self.h:
            lib().usher_ref_tree_free(self.h)
            self.h = None

    def parsimony(self):
        return int(lib().usher_ref_tree_parsimony(self.h))

    def export(self):
        n, m, nl = C.c_uint32(), C.c_uint64(), C.c_uint64()
        lib().usher_ref_tree_export(self.h, C.byref(n), C.byref(m), C.byref(nl), None, None, None, None)
        parent = np.zeros(n.value, np.int32)
        row_ptr = np.zeros(n.value + 1, np.uint64)
        muts = np.zeros(m.value, MUT_DTYPE)
        names = np.zeros(nl.value, np.uint8)
        lib().usher_ref_tree_export(
            self.h, C.byref(n), C.byref(m), C.byref(nl), _p(parent), _p(row_ptr), _p(muts), _p(names)
        )
        return parent, row_ptr, muts, names.tobytes().decode().split("\n")[:-1]

    def condensed(self):
        n = lib().usher_ref_condensed_export(self.h, None, 0)
        buf = C.create_string_buffer(int(n) + 1)
        lib().usher_ref_condensed_export(self.h, buf, n)
        out = {}
        for line in buf.raw[: int(n)].decode().splitlines():
            name, members = line.split("\t")
            out[name] = members.split(",")
        return out

    def read_samples(self, vcf):
        ns = lib().usher_ref_read_samples(self.h, vcf.encode())
        ne, nl = C.c_uint64(), C.c_uint64()
        lib().usher_ref_samples_export(self.h, C.byref(ne), C.byref(nl), None, None, None)
        s_ptr = np.zeros(ns + 1, np.uint64)
        sm = np.zeros(ne.value, MUT_DTYPE)
        names = np.zeros(nl.value, np.uint8)
        lib().usher_ref_samples_export(self.h, C.byref(ne), C.byref(nl), _p(s_ptr), _p(sm), _p(names))
        return s_ptr, sm, names.tobytes().decode().split("\n")[:-1]

    def usher_common(self, outdir, threads=1, print_parsimony_scores=False, no_add=False):
        return lib().usher_ref_usher_common(
            self.h, outdir.encode(), threads, int(print_parsimony_scores), int(no_add)
        )

    def usher_common2(self, outdir, threads=1, no_add=False, sort1=False, sort2=False, sort3=False, reverse=False,
                      max_uncertainty=1000000, max_parsimony=1000000, uncondensed=False):
        f = lib().usher_ref_usher_common2
        f.argtypes = [C.c_void_p, C.c_char_p] + [C.c_int] * 6 + [C.c_uint32, C.c_uint32, C.c_int]
        return f(self.h, outdir.encode(), threads, int(no_add), int(sort1), int(sort2), int(sort3), int(reverse),
                 max_uncertainty, max_parsimony, int(uncondensed))

    def search_strided(self, calls, stride, offset=0, threads=1, seed_best=-1):
        """Seconds the reference's two-pass search of one sample takes over every `stride`-th BFS node.  With
        seed_best >= 0 the running best starts from the (known) true best score: see usher_ref_search_strided2."""
        calls = np.ascontiguousarray(calls, dtype=MUT_DTYPE)
        best = C.c_int32()
        sec = lib().usher_ref_search_strided2(self.h, len(calls), _p(calls), stride, offset, threads, int(seed_best),
                                              C.byref(best))
        return float(sec), int(best.value)

    def score_nodes(self, calls, dfs_nodes, threads=1):
        """The reference's -p score (mapper2_body(inp, true, false); +1 on invalid nodes) of one sample at the
        listed DFS nodes of a from_flat tree, and per node 0 = not a valid placement, 1 = valid, 3 = valid with
        has_unique (sibling placement)."""
        calls = np.ascontiguousarray(calls, dtype=MUT_DTYPE)
        nodes = np.ascontiguousarray(dfs_nodes, dtype=np.uint32)
        out = np.zeros(len(nodes), np.int32)
        valid = np.zeros(len(nodes), np.uint8)
        rc = lib().usher_ref_score_nodes(self.h, len(calls), _p(calls), len(nodes), _p(nodes), threads, _p(out), _p(valid))
        if rc != 0:
            raise ValueError("usher_ref_score_nodes: node index out of range")
        return out, valid

    def search(self, s_ptr, sm, n_nodes, threads=1, per_node=False, want_set=True, set_cap=None):
        """Reference two-pass search (or -p single pass when per_node) of each sample on the frozen tree."""
        s_ptr = np.ascontiguousarray(s_ptr, dtype=np.uint64)
        sm = np.ascontiguousarray(sm, dtype=MUT_DTYPE)
        B = len(s_ptr) - 1
        out = {
            "score": np.zeros(B, np.int32),
            "best_dfs": np.zeros(B, np.uint32),
            "best_j": np.zeros(B, np.uint32),
            "num_best": np.zeros(B, np.uint32),
            "has_unique": np.zeros(B, np.uint8),
            "seconds": np.zeros(B, np.float64),
        }
        node_scores = np.zeros((B, n_nodes), np.int32) if per_node else None
        cap = int(set_cap if set_cap is not None else max(1024, 64 * B)) if want_set else 0
        bset = np.zeros(cap, np.uint32) if want_set else None
        bset_u = np.zeros(cap, np.uint8) if want_set else None
        bptr = np.zeros(B + 1, np.uint64) if want_set else None
        rc = lib().usher_ref_search(
            self.h, B, _p(s_ptr), _p(sm), threads, 1 if per_node else 0,
            _p(out["score"]), _p(out["best_dfs"]), _p(out["best_j"]), _p(out["num_best"]),
            _p(out["has_unique"]), _p(node_scores), _p(bset), _p(bset_u), _p(bptr), cap, _p(out["seconds"]),
        )
        if want_set:
            if rc != 0:  # grow and retry
                return self.search(s_ptr, sm, n_nodes, threads, per_node, True, int(bptr[-1]) + 16)
            out["best_set_ptr"] = bptr
            out["best_set"] = bset[: int(bptr[-1])]
            out["best_set_unique"] = bset_u[: int(bptr[-1])]
        if per_node:
            out["node_scores"] = node_scores
        return out
