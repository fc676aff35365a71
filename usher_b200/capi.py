"""ctypes binding of include/usher_b200.h (libusher_b200.so) and include/usher_b200_synth.h.

This is plumbing for tests/ and bench.py; the product is the C ABI itself.  Nothing here computes scores:
every placement goes through the CUDA kernels, and the calls raise if the library or a GPU is missing.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

MUT_DTYPE = np.dtype(
    [("position", "<i4"), ("ref_nuc", "u1"), ("par_nuc", "u1"), ("mut_nuc", "u1"), ("is_missing", "u1")]
)
PLACEMENT_DTYPE = np.dtype(
    [("score", "<i4"), ("best_node", "<u4"), ("best_j", "<u4"), ("num_best", "<u4"), ("has_unique", "<u4"),
     ("best_num_leaves", "<u4"), ("reserved", "<u4", (2,))]
)
HDR_DTYPE = np.dtype([("g", "<i4"), ("tiekey", "<u4"), ("level_flags", "<u4"), ("nmut_c0", "<u4")])

WANT_NODE_SCORES = 1
WANT_BEST_SET = 2
E_CAPACITY = -6
E_NO_DEVICE = -7


class FlatMat(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("n_mutations", C.c_uint64), ("parent", C.c_void_p),
                ("row_ptr", C.c_void_p), ("mutations", C.c_void_p), ("tie_index", C.c_void_p)]


class MatInfo(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("max_level", C.c_uint32), ("n_mutations", C.c_uint64),
                ("genome_len", C.c_uint32), ("n_tiles", C.c_uint32), ("device_bytes", C.c_uint64),
                ("algorithmic_bytes", C.c_uint64), ("device", C.c_int32), ("reserved", C.c_uint32)]


class Timing(C.Structure):
    _fields_ = [("prep_ms", C.c_float), ("score_ms", C.c_float), ("reduce_ms", C.c_float),
                ("score_launches", C.c_uint32), ("total_launches", C.c_uint32), ("score_bytes", C.c_uint64)]


class DerivedView(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("genome_len", C.c_uint32), ("max_level", C.c_uint32),
                ("n_tiles", C.c_uint32), ("n_mutations", C.c_uint64)] + [
        (k, C.c_void_p) for k in ("level", "tie_index", "num_leaves", "tiekey", "key_to_node", "row32", "mutw",
                                  "hdr", "ref_of", "tile_start", "anc_ptr", "anc")] + [
        ("n_tiles3", C.c_uint32), ("n_seed_segs", C.c_uint32), ("narrow3", C.c_uint32), ("reserved3", C.c_uint32),
        ("stream_words", C.c_uint64)] + [
        (k, C.c_void_p) for k in ("stream", "hdr3", "tile3_start", "tile3_w0", "tile3_lvl", "tile3_sseg", "seed_end", "blk_words",
                                  "blk_rec")]


class UB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"usher_b200 error {code}: {msg}")
        self.code = code


_lib = None
_synth = None

EXPORTS = [
    "ub200_last_error", "ub200_abi_version", "ub200_device_count", "ub200_mat_create", "ub200_mat_destroy",
    "ub200_mat_info_get", "ub200_mat_node_arrays", "ub200_mat_set_pass_samples", "ub200_place_batch",
    "ub200_samples_upload", "ub200_samples_free", "ub200_place_resident", "ub200_results_download",
    "ub200_results_device_ptr", "ub200_node_scores_download", "ub200_best_set_download", "ub200_mat_set_stream",
    "ub200_mat_synchronize", "ub200_last_timing", "ub200_results_copy_device", "ub200_mat_set_scan_sharing",
    "ub200_fs_tree_create", "ub200_fs_tree_destroy", "ub200_fs_sites",
    "ub200_multi_create", "ub200_multi_destroy", "ub200_multi_size", "ub200_multi_mat", "ub200_multi_place_batch",
]


def lib():
    """Load libusher_b200.so (building it in-tree if sources are newer).  Raises if it cannot be loaded."""
    global _lib
    if _lib is None:
        _build.build()
        L = C.CDLL(_build.LIB)
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        L.ub200_last_error.restype = C.c_char_p
        L.ub200_mat_create.argtypes = [C.POINTER(FlatMat), C.c_int, C.POINTER(vp)]
        L.ub200_mat_destroy.argtypes = [vp]
        L.ub200_mat_destroy.restype = None
        L.ub200_mat_info_get.argtypes = [vp, C.POINTER(MatInfo)]
        L.ub200_mat_node_arrays.argtypes = [vp, vp, vp, vp]
        L.ub200_mat_set_pass_samples.argtypes = [vp, u32]
        L.ub200_mat_set_scan_sharing.argtypes = [vp, u32]
        L.ub200_place_batch.argtypes = [vp, u32, vp, vp, u32, vp, vp, vp, vp, u64]
        L.ub200_samples_upload.argtypes = [vp, u32, vp, vp, C.POINTER(vp)]
        L.ub200_samples_free.argtypes = [vp]
        L.ub200_samples_free.restype = None
        L.ub200_place_resident.argtypes = [vp, vp, u32, C.c_int]
        L.ub200_results_download.argtypes = [vp, vp, vp]
        L.ub200_results_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
        L.ub200_results_copy_device.argtypes = [vp, vp, vp]
        L.ub200_node_scores_download.argtypes = [vp, vp, vp]
        L.ub200_best_set_download.argtypes = [vp, vp, vp, vp, u64]
        L.ub200_mat_set_stream.argtypes = [vp, vp]
        L.ub200_mat_synchronize.argtypes = [vp]
        L.ub200_last_timing.argtypes = [vp, C.POINTER(Timing)]
        L.ub200_multi_create.argtypes = [C.POINTER(FlatMat), C.c_int, vp, C.POINTER(vp)]
        L.ub200_multi_destroy.argtypes = [vp]
        L.ub200_multi_destroy.restype = None
        L.ub200_multi_size.argtypes = [vp]
        L.ub200_multi_mat.argtypes = [vp, C.c_int]
        L.ub200_multi_mat.restype = vp
        L.ub200_multi_place_batch.argtypes = [vp, u32, vp, vp, u32, vp, vp, vp, vp, u64]
        if hasattr(L, "ub200_fs_tree_create"):   # (developer A/B runs load older builds of the library)
            L.ub200_fs_tree_create.argtypes = [u32, vp, C.c_int, C.POINTER(vp)]
            L.ub200_fs_tree_destroy.argtypes = [vp]
            L.ub200_fs_tree_destroy.restype = None
            L.ub200_fs_sites.argtypes = [vp, u32, vp, vp, vp, vp, u64, vp, vp, vp, C.POINTER(u64)]
        L.ub200_debug_derive.argtypes = [C.POINTER(FlatMat), u32, u32, C.POINTER(vp), C.POINTER(DerivedView),
                                         C.c_char_p, C.c_size_t]
        L.ub200_debug_derive_free.argtypes = [vp]
        L.ub200_debug_derive_free.restype = None
        _lib = L
    return _lib


def synth_lib():
    global _synth
    if _synth is None:
        _build.build()
        L = C.CDLL(_build.SYNTH)
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        L.ub200_synth_mat_create.argtypes = [u32, C.c_double, u32, C.c_int, u64, C.POINTER(vp)]
        L.ub200_synth_free.argtypes = [vp]
        L.ub200_synth_free.restype = None
        L.ub200_synth_flat.argtypes = [vp, C.POINTER(FlatMat)]
        L.ub200_synth_reference.argtypes = [vp]
        L.ub200_synth_reference.restype = vp
        L.ub200_synth_samples.argtypes = [vp, u32, C.c_int, u64, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        _synth = L
    return _synth


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check(rc):
    if rc != 0:
        raise UB200Error(rc, lib().ub200_last_error().decode())


def make_flat(parent, row_ptr, muts, tie_index=None):
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
    muts = np.ascontiguousarray(muts, dtype=MUT_DTYPE)
    tie = None if tie_index is None else np.ascontiguousarray(tie_index, dtype=np.uint32)
    f = FlatMat(len(parent), len(muts), _p(parent).value, _p(row_ptr).value,
                _p(muts).value if len(muts) else None, None if tie is None else _p(tie).value)
    f._keep = (parent, row_ptr, muts, tie)
    return f


def _view(ptr, dtype, count):
    if count == 0:
        return np.zeros(0, dtype)
    buf = (C.c_char * (np.dtype(dtype).itemsize * count)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count)


def debug_derive(parent, row_ptr, muts, tie_index=None, target_tiles=64, min_tile_cost=0):
    """Host-only: run the derivation and return copies of the derived arrays (no GPU needed)."""
    f = make_flat(parent, row_ptr, muts, tie_index)
    h = C.c_void_p()
    v = DerivedView()
    err = C.create_string_buffer(512)
    rc = lib().ub200_debug_derive(C.byref(f), target_tiles, min_tile_cost, C.byref(h), C.byref(v), err, 512)
    if rc != 0:
        raise UB200Error(rc, err.value.decode())
    n, T = v.n_nodes, v.n_tiles
    out = {
        "n": n, "L": v.genome_len, "max_level": v.max_level, "m": v.n_mutations,
        "level": _view(v.level, np.uint32, n).copy(), "tie_index": _view(v.tie_index, np.uint32, n).copy(),
        "num_leaves": _view(v.num_leaves, np.uint32, n).copy(), "tiekey": _view(v.tiekey, np.uint32, n).copy(),
        "key_to_node": _view(v.key_to_node, np.uint32, n).copy(), "row32": _view(v.row32, np.uint32, n + 1).copy(),
        "mutw": _view(v.mutw, np.uint32, int(v.n_mutations)).copy(), "hdr": _view(v.hdr, HDR_DTYPE, n).copy(),
        "ref_of": _view(v.ref_of, np.uint8, v.genome_len).copy(),
        "tile_start": _view(v.tile_start, np.uint32, T + 1).copy(),
        "anc_ptr": _view(v.anc_ptr, np.uint32, T + 1).copy(),
    }
    out["anc"] = _view(v.anc, np.uint32, int(out["anc_ptr"][-1])).copy()
    T3 = v.n_tiles3
    if T3:   # k_score3 layout
        out.update({
            "narrow3": int(v.narrow3),
            "stream": _view(v.stream, np.uint32, int(v.stream_words)).copy(),
            "hdr3": _view(v.hdr3, HDR_DTYPE, n).copy(),
            "tile3_start": _view(v.tile3_start, np.uint32, T3 + 1).copy(),
            "tile3_w0": _view(v.tile3_w0, np.uint32, T3 + 1).copy(),
            "tile3_lvl": _view(v.tile3_lvl, np.uint32, T3).copy(),
            "tile3_sseg": _view(v.tile3_sseg, np.uint32, T3 + 1).copy(),
            "seed_end": _view(v.seed_end, np.uint32, v.n_seed_segs).copy(),
            "blk_words": _view(v.blk_words, np.uint32, (n + 31) // 32).copy(),
            "blk_rec": _view(v.blk_rec, np.uint32, 4 * ((n + 31) // 32)).copy().reshape(-1, 4),
        })
    lib().ub200_debug_derive_free(h)
    return out


class Mat:
    """A tree resident on one GPU (ub200_mat)."""

    def __init__(self, parent, row_ptr, muts, tie_index=None, device=0):
        self._flat = make_flat(parent, row_ptr, muts, tie_index)
        self.h = C.c_void_p()
        _check(lib().ub200_mat_create(C.byref(self._flat), device, C.byref(self.h)))
        self.info = MatInfo()
        _check(lib().ub200_mat_info_get(self.h, C.byref(self.info)))
        self.n = self.info.n_nodes

    @classmethod
    def from_flat_struct(cls, flat, device=0):
        self = cls.__new__(cls)
        self._flat = flat
        self.h = C.c_void_p()
        _check(lib().ub200_mat_create(C.byref(flat), device, C.byref(self.h)))
        self.info = MatInfo()
        _check(lib().ub200_mat_info_get(self.h, C.byref(self.info)))
        self.n = self.info.n_nodes
        return self

    def close(self):
        if self.h:
            lib().ub200_mat_destroy(self.h)
            self.h = None

    def node_arrays(self):
        bfs, nl, lv = (np.zeros(self.n, np.uint32) for _ in range(3))
        _check(lib().ub200_mat_node_arrays(self.h, _p(bfs), _p(nl), _p(lv)))
        return bfs, nl, lv

    def set_pass_samples(self, n):
        _check(lib().ub200_mat_set_pass_samples(self.h, n))

    def set_scan_sharing(self, groups_per_scan):
        _check(lib().ub200_mat_set_scan_sharing(self.h, groups_per_scan))

    def set_stream(self, cuda_stream_ptr):
        _check(lib().ub200_mat_set_stream(self.h, cuda_stream_ptr))

    def synchronize(self):
        _check(lib().ub200_mat_synchronize(self.h))

    def timing(self):
        t = Timing()
        _check(lib().ub200_last_timing(self.h, C.byref(t)))
        return t

    def place_batch(self, s_ptr, calls, node_scores=False, best_set=False):
        """ub200_place_batch: host buffers in, host buffers out."""
        s_ptr = np.ascontiguousarray(s_ptr, dtype=np.uint64)
        calls = np.ascontiguousarray(calls, dtype=MUT_DTYPE)
        B = len(s_ptr) - 1
        out = np.zeros(B, PLACEMENT_DTYPE)
        ns = np.zeros((B, self.n), np.int32) if node_scores else None
        flags = (WANT_NODE_SCORES if node_scores else 0) | (WANT_BEST_SET if best_set else 0)
        bptr = np.zeros(B + 1, np.uint64) if best_set else None
        cap = max(1024, 4 * B)
        while True:
            bset = np.zeros(cap, np.uint32) if best_set else None
            rc = lib().ub200_place_batch(self.h, B, _p(s_ptr), _p(calls), flags, _p(out), _p(ns), _p(bset),
                                         _p(bptr), cap if best_set else 0)
            if rc == E_CAPACITY:
                cap = int(bptr[B]) + 16
                continue
            _check(rc)
            break
        res = {"placements": out}
        if node_scores:
            res["node_scores"] = ns
        if best_set:
            res["best_set_ptr"] = bptr
            raw = bset[: int(bptr[B])]
            res["best_set"] = raw & np.uint32(0x7FFFFFFF)
            res["best_set_unique"] = (raw >> np.uint32(31)).astype(np.uint8)
        return res

    def upload(self, s_ptr, calls):
        return Samples(self, s_ptr, calls)


class MultiMat:
    """One replica of a tree per GPU of this box (ub200_multi); place_batch shards the samples over the replicas."""

    def __init__(self, flat, n_devices=0, devices=None):
        self._flat = flat
        self.h = C.c_void_p()
        ids = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        _check(lib().ub200_multi_create(C.byref(flat), len(ids) if ids is not None else n_devices, _p(ids), C.byref(self.h)))
        self.size = lib().ub200_multi_size(self.h)
        self.n = flat.n_nodes

    def set_pass_samples(self, n, sharing=0):
        for i in range(self.size):
            m = lib().ub200_multi_mat(self.h, i)
            _check(lib().ub200_mat_set_pass_samples(m, n))
            _check(lib().ub200_mat_set_scan_sharing(m, sharing))

    def place_batch(self, s_ptr, calls, best_set=False):
        s_ptr = np.ascontiguousarray(s_ptr, dtype=np.uint64)
        calls = np.ascontiguousarray(calls, dtype=MUT_DTYPE)
        B = len(s_ptr) - 1
        out = np.zeros(B, PLACEMENT_DTYPE)
        bptr = np.zeros(B + 1, np.uint64) if best_set else None
        cap = max(1024, 4 * B)
        while True:
            bset = np.zeros(cap, np.uint32) if best_set else None
            rc = lib().ub200_multi_place_batch(self.h, B, _p(s_ptr), _p(calls), WANT_BEST_SET if best_set else 0, _p(out),
                                               None, _p(bset), _p(bptr), cap if best_set else 0)
            if rc == E_CAPACITY:
                cap = int(bptr[B]) + 16
                continue
            _check(rc)
            break
        res = {"placements": out}
        if best_set:
            res["best_set_ptr"] = bptr
            raw = bset[: int(bptr[B])]
            res["best_set"] = raw & np.uint32(0x7FFFFFFF)
            res["best_set_unique"] = (raw >> np.uint32(31)).astype(np.uint8)
        return res

    def close(self):
        if self.h:
            lib().ub200_multi_destroy(self.h)
            self.h = None


def fitch_sankoff(parent_bfs, ref_code, var_ptr, var_node, var_nuc, device=0):
    """ub200_fs_*: per-site parsimony assignment on a tree given in BFS order.  Returns (site, node, par_state, state)
    arrays sorted by (site, node)."""
    parent_bfs = np.ascontiguousarray(parent_bfs, dtype=np.int32)
    ref_code = np.ascontiguousarray(ref_code, dtype=np.uint8)
    var_ptr = np.ascontiguousarray(var_ptr, dtype=np.uint64)
    var_node = np.ascontiguousarray(var_node, dtype=np.uint32)
    var_nuc = np.ascontiguousarray(var_nuc, dtype=np.uint8)
    h = C.c_void_p()
    _check(lib().ub200_fs_tree_create(len(parent_bfs), _p(parent_bfs), device, C.byref(h)))
    try:
        cap = max(1024, 4 * len(var_node))
        while True:
            o_site, o_node, o_st = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32), np.zeros(cap, np.uint8)
            cnt = C.c_uint64()
            rc = lib().ub200_fs_sites(h, len(ref_code), _p(ref_code), _p(var_ptr), _p(var_node), _p(var_nuc), cap,
                                      _p(o_site), _p(o_node), _p(o_st), C.byref(cnt))
            if rc == E_CAPACITY:
                cap = int(cnt.value) + 16
                continue
            _check(rc)
            break
    finally:
        lib().ub200_fs_tree_destroy(h)
    k = int(cnt.value)
    order = np.lexsort((o_node[:k], o_site[:k]))
    return o_site[:k][order], o_node[:k][order], (o_st[:k][order] >> 4), (o_st[:k][order] & 15)


class Samples:
    """A sample batch resident on the GPU (ub200_samples)."""

    def __init__(self, mat, s_ptr, calls):
        self.mat = mat
        s_ptr = np.ascontiguousarray(s_ptr, dtype=np.uint64)
        calls = np.ascontiguousarray(calls, dtype=MUT_DTYPE)
        self.n = len(s_ptr) - 1
        self.h = C.c_void_p()
        _check(lib().ub200_samples_upload(mat.h, self.n, _p(s_ptr), _p(calls), C.byref(self.h)))

    def place(self, flags=0, sync=True):
        _check(lib().ub200_place_resident(self.mat.h, self.h, flags, int(sync)))

    def download(self):
        out = np.zeros(self.n, PLACEMENT_DTYPE)
        _check(lib().ub200_results_download(self.mat.h, self.h, _p(out)))
        return out

    def copy_results_to(self, dev_ptr):
        _check(lib().ub200_results_copy_device(self.mat.h, self.h, dev_ptr))

    def device_ptr(self):
        p, b = C.c_void_p(), C.c_size_t()
        _check(lib().ub200_results_device_ptr(self.h, C.byref(p), C.byref(b)))
        return p.value, b.value

    def close(self):
        if self.h:
            lib().ub200_samples_free(self.h)
            self.h = None


class Synth:
    """Seeded synthetic MAT of SURVEY.md §8(d) (libub200_synth.so)."""
    UNIFORM, SC2 = 0, 1
    SNV40, LEAF, AMBIG = 0, 1, 2

    def __init__(self, n_nodes, mu, genome_len, shape, seed):
        self.h = C.c_void_p()
        rc = synth_lib().ub200_synth_mat_create(n_nodes, mu, genome_len, shape, seed, C.byref(self.h))
        if rc:
            raise UB200Error(rc, "ub200_synth_mat_create failed")
        self.flat = FlatMat()
        synth_lib().ub200_synth_flat(self.h, C.byref(self.flat))
        self.n = self.flat.n_nodes
        self.m = self.flat.n_mutations

    def arrays(self):
        """numpy views (no copy) of parent / row_ptr / mutations."""
        return (_view(self.flat.parent, np.int32, self.n), _view(self.flat.row_ptr, np.uint64, self.n + 1),
                _view(self.flat.mutations, MUT_DTYPE, int(self.m)))

    def samples(self, n, family, seed):
        sp, sc, so = C.c_void_p(), C.c_void_p(), C.c_void_p()
        rc = synth_lib().ub200_synth_samples(self.h, n, family, seed, C.byref(sp), C.byref(sc), C.byref(so))
        if rc:
            raise UB200Error(rc, "ub200_synth_samples failed")
        s_ptr = _view(sp.value, np.uint64, n + 1).copy()
        calls = _view(sc.value, MUT_DTYPE, int(s_ptr[-1])).copy()
        origin = _view(so.value, np.uint32, n).copy()
        return s_ptr, calls, origin

    def close(self):
        if self.h:
            synth_lib().ub200_synth_free(self.h)
            self.h = None
