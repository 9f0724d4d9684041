"""ctypes loader for the CPU checkers.  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs -- never by
vdjer_b200/.

  build("port")       -> oracle/liboracle.so        (oracle/vdj_oracle.c, our restatement)
  build("reference")  -> oracle/_ref/libvdjref.so   (the reference's own functions, compiled
                         from /root/reference by oracle/Makefile; prebuilt copy travels to the GPU box)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, "liboracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libvdjref.so")
REF_LIB_G = os.path.join(HERE, "_ref", "libvdjref_g.so")


class _Res(C.Structure):
    _fields_ = [
        ("n_records", C.c_uint64), ("n_windows", C.c_uint64), ("n_gated", C.c_uint64),
        ("n_pre_total", C.c_uint64), ("n_pre", C.c_uint64),
        ("pre_first_pos", C.POINTER(C.c_uint64)), ("pre_freq", C.POINTER(C.c_uint16)),
        ("pre_qual_sums", C.POINTER(C.c_uint8)),
        ("n_nodes", C.c_uint64), ("n_hits", C.c_uint64),
        ("node_first_pos", C.POINTER(C.c_uint64)), ("node_freq", C.POINTER(C.c_uint16)),
        ("out_deg", C.POINTER(C.c_uint8)), ("out_succ", C.POINTER(C.c_uint32)),
        ("in_deg", C.POINTER(C.c_uint8)), ("in_pred", C.POINTER(C.c_uint32)),
        ("t_pass1", C.c_double), ("t_prune", C.c_double), ("t_pass2", C.c_double),
    ]


def have_reference() -> bool:
    return os.path.exists(REF_LIB)


def _arr(ptr, n, dtype, cols=1):
    if n == 0:
        return np.zeros((0, cols) if cols > 1 else (0,), dtype=dtype)
    a = np.ctypeslib.as_array(ptr, shape=(n * cols,)).copy()
    return a.reshape(n, cols) if cols > 1 else a


def _cbuf(buf):
    a = buf if isinstance(buf, np.ndarray) else np.frombuffer(bytes(buf), dtype=np.uint8)
    a = np.ascontiguousarray(a.view(np.uint8))
    if a.size == 0 or a[-1] != 0:
        a = np.concatenate([a, np.zeros(1, np.uint8)])
    return a


_libs = {}


def _load(kind: str, variant: str = "O2"):
    key = (kind, variant)
    if key in _libs:
        return _libs[key]
    if kind == "port":
        lib = C.CDLL(PORT_LIB)
        lib.vdj_oracle_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_Res)]
        lib.vdj_oracle_free.argtypes = [C.POINTER(_Res)]
    elif kind == "reference":
        lib = C.CDLL(REF_LIB if variant == "O2" else REF_LIB_G)
        lib.vdjref_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(_Res)]
        lib.vdjref_free.argtypes = [C.POINTER(_Res)]
    else:
        raise ValueError(kind)
    _libs[key] = lib
    return lib


def build(primary, secondary, read_length: int, k: int = 35, mf: int = 3, mq: int = 90,
          kind: str = "port", variant: str = "O2", scratch_dir: str | None = None) -> dict:
    """Run the CPU graph build; returns numpy arrays in the layout of vdj_oracle.h."""
    p, s = _cbuf(primary), _cbuf(secondary)
    lib = _load(kind, variant)
    r = _Res()
    if kind == "port":
        rc = lib.vdj_oracle_build(p.ctypes.data, s.ctypes.data, read_length, k, mf, mq, C.byref(r))
    else:
        scratch = scratch_dir or os.path.join(HERE, "_ref")
        rc = lib.vdjref_build(p.ctypes.data, s.ctypes.data, read_length, k, mf, mq, scratch.encode(), C.byref(r))
    if rc != 0:
        raise RuntimeError(f"{kind} oracle failed with {rc}")
    n_pre, n = int(r.n_pre), int(r.n_nodes)
    out = dict(
        kind=kind, n_records=int(r.n_records), n_windows=int(r.n_windows), n_gated=int(r.n_gated),
        n_pre_total=int(r.n_pre_total), n_pre=n_pre, n_nodes=n, n_hits=int(r.n_hits),
        pre_first_pos=_arr(r.pre_first_pos, n_pre, np.uint64), pre_freq=_arr(r.pre_freq, n_pre, np.uint16),
        pre_qual_sums=_arr(r.pre_qual_sums, n_pre, np.uint8, k),
        first_pos=_arr(r.node_first_pos, n, np.uint64), frequency=_arr(r.node_freq, n, np.uint16),
        out_deg=_arr(r.out_deg, n, np.uint8), out_succ=_arr(r.out_succ, n, np.uint32, 4),
        in_deg=_arr(r.in_deg, n, np.uint8), in_pred=_arr(r.in_pred, n, np.uint32, 4),
        t_pass1=r.t_pass1, t_prune=r.t_prune, t_pass2=r.t_pass2,
    )
    (lib.vdj_oracle_free if kind == "port" else lib.vdjref_free)(C.byref(r))
    return out


GRAPH_KEYS = ["n_pre_total", "n_pre", "n_nodes", "pre_first_pos", "pre_freq", "pre_qual_sums",
              "first_pos", "frequency", "out_deg", "out_succ", "in_deg", "in_pred"]


def diff(a: dict, b: dict, keys=GRAPH_KEYS) -> list[str]:
    """Names of the fields in which two oracle outputs differ."""
    bad = []
    for k in keys:
        x, y = a[k], b[k]
        same = np.array_equal(x, y) if isinstance(x, np.ndarray) else x == y
        if not same:
            bad.append(k)
    return bad


# --------------------------------------------------------------------------------------------
# glue check: the reference's downstream code (identify_root_nodes, condense_graph, dump_graph)
# on a graph rebuilt by glue/vdjgraph_glue.inc from vdjgraph_result arrays
# --------------------------------------------------------------------------------------------
GLUE_LIB = os.path.join(HERE, "_ref", "libvdjglue.so")


def have_glue() -> bool:
    return os.path.exists(GLUE_LIB)


def _result_from(graph):
    """(struct, keep-alive list): a vdjgraph_result over the arrays of a Graph / oracle dict; `hm_slots`
    (the layout of the reference's `nodes` map) is passed on when the graph has it."""
    from vdjer_b200.graph import _Result   # struct layout of include/vdjgraph.h
    def get(n):
        return graph.get(n) if isinstance(graph, dict) else getattr(graph, n, None)
    r = _Result()
    r.n_nodes = len(get("first_pos"))
    keep = []
    for name, dt in [("first_pos", np.uint64), ("frequency", np.uint16), ("out_deg", np.uint8),
                     ("in_deg", np.uint8), ("out_succ", np.uint32), ("in_pred", np.uint32)]:
        a = np.ascontiguousarray(get(name), dtype=dt)
        keep.append(a)
        setattr(r, name, a.ctypes.data_as(type(getattr(r, name))))
    hm = get("hm_slots")
    if hm is not None:
        a = np.ascontiguousarray(hm, dtype=np.uint32)
        keep.append(a)
        r.hm_slots = a.ctypes.data_as(type(r.hm_slots))
        r.hm_buckets = a.size
    keep.append(r)
    return r, keep


def glue_dot(primary, secondary, read_length: int, k: int, mf: int, mq: int, dot_path: str,
             graph=None, scratch_dir: str | None = None):
    """Write the reference's vdjer.dot for these records.  graph=None: built by the reference's own
    functions; otherwise an object/dict with first_pos, frequency, out_deg, in_deg, out_succ,
    in_pred (the vdjgraph_result arrays), rebuilt through the glue.  Returns (n_nodes, n_roots)."""
    p, s = _cbuf(primary), _cbuf(secondary)
    lib = C.CDLL(GLUE_LIB)
    lib.vdjglue_dot.restype = C.c_long
    lib.vdjglue_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p,
                                C.c_void_p, C.c_char_p, C.POINTER(C.c_long)]
    res_ptr, keep = None, []
    if graph is not None:
        r, keep = _result_from(graph)
        res_ptr = C.addressof(r)
    n_roots = C.c_long(0)
    scratch = scratch_dir or os.path.join(HERE, "_ref")
    n = lib.vdjglue_dot(p.ctypes.data, s.ctypes.data, read_length, k, mf, mq, scratch.encode(), res_ptr,
                        dot_path.encode(), C.byref(n_roots))
    if n < 0:
        raise RuntimeError(f"vdjglue_dot failed with {n}")
    return int(n), int(n_roots.value)


def glue_rebuild_ms(primary, secondary, read_length: int, k: int, graph, scratch_dir: str | None = None) -> float:
    """Wall time (ms) of glue/vdjgraph_glue.inc's vdjgraph_rebuild_nodes on these result arrays:
    what the reference-side glue costs after vdjgraph_build has returned."""
    p, s = _cbuf(primary), _cbuf(secondary)
    lib = C.CDLL(GLUE_LIB)
    lib.vdjglue_rebuild_ms.restype = C.c_double
    lib.vdjglue_rebuild_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_void_p]
    r, keep = _result_from(graph)
    scratch = scratch_dir or os.path.join(HERE, "_ref")
    ms = lib.vdjglue_rebuild_ms(p.ctypes.data, s.ctypes.data, read_length, k, scratch.encode(), C.addressof(r))
    if ms < 0:
        raise RuntimeError(f"vdjglue_rebuild_ms failed with {ms}")
    return float(ms)


def reference_iteration_order(primary, secondary, read_length: int, k: int, mf: int, mq: int, n_nodes: int,
                              scratch_dir: str | None = None):
    """(ids, bucket_count): node creation ranks in the iteration order of the reference's own `nodes`
    dense_hash_map after build_graph2, and its bucket count."""
    p, s = _cbuf(primary), _cbuf(secondary)
    lib = C.CDLL(GLUE_LIB)
    lib.vdjglue_iteration_order.restype = C.c_long
    lib.vdjglue_iteration_order.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p,
                                            C.c_void_p, C.c_long, C.POINTER(C.c_uint64)]
    out = np.zeros(max(n_nodes, 1), np.uint32)
    nb = C.c_uint64(0)
    scratch = scratch_dir or os.path.join(HERE, "_ref")
    n = lib.vdjglue_iteration_order(p.ctypes.data, s.ctypes.data, read_length, k, mf, mq, scratch.encode(),
                                    out.ctypes.data, n_nodes, C.byref(nb))
    if n != n_nodes:
        raise RuntimeError(f"vdjglue_iteration_order returned {n}, expected {n_nodes}")
    return out[:n_nodes], int(nb.value)
