"""ctypes binding of the C ABI in include/ktn.h.

The same binding drives the CUDA library (the product) and, from tests only, the CPU
oracle: both export identical symbols.  `load_cuda_library()` is the only loader the
product uses; it fails loudly when libktn.so is missing -- there is no CPU fallback.
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SYNTH_LIB_PATH = os.path.join(_HERE, "libktn_synth.so")
CUDA_LIB_PATH = os.environ.get("KTN_LIB") or os.path.join(_HERE, "libktn.so")      # KTN_LIB: an A/B variant build of the CUDA library (scripts/build_variants.sh)

# wire-format constants (include/ktn.h)
(OP_CONST, OP_VAR, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_NEG, OP_EXP, OP_LOG, OP_SQRT, OP_ABS, OP_SIN, OP_COS,
 OP_IFELSE, OP_LE, OP_LT, OP_GE, OP_GT, OP_EQ) = range(20)
ROW_NL, ROW_DENSE = 1, 2
KTN_OK, KTN_NUMERIC_NONFINITE = 0, 1
FLAG_LEAN_VIEW = 1          # ktn_options.flags: cut views carry only what the LP needs
FLAG_TIME_KERNELS = 2       # compaction and cut kernel timed separately (one more event per round)
FLAG_EAGER_VIEW = 4         # multi-device / pipelined handles: every shard's cuts are downloaded as soon as that shard has finished
FLAG_DIRECT_VIEW = 8        # single-device handles: the kernels store the cut batch straight into mapped pinned host memory
SYNTH_QCQP, SYNTH_LSE, SYNTH_SOC = 0, 1, 2


class KtnError(RuntimeError):
    pass


class ktn_options(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("f_tol", C.c_double),
                ("cut_coef_rng", C.c_double), ("topk", C.c_int64), ("flags", C.c_int32), ("reserved", C.c_int32),
                ("ngpus", C.c_int32), ("devices", C.c_int32 * 16)]


class ktn_timings(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("exchange_ms", C.c_double),
                ("d2h_ms", C.c_double), ("launches", C.c_int64), ("rounds", C.c_int64), ("eval_ms", C.c_double), ("compact_ms", C.c_double),
                ("eval_ms_sum", C.c_double), ("compact_ms_sum", C.c_double), ("rounds_timed", C.c_int64), ("cut_ms", C.c_double), ("cut_ms_sum", C.c_double), ("exchange_ms_sum", C.c_double), ("exchanges_timed", C.c_int64)]


class ktn_cut_view(C.Structure):
    _fields_ = [("n_cuts", C.c_int64), ("nnz", C.c_int64), ("row_id", C.c_void_p), ("row_ptr", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p),
                ("lo", C.c_void_p), ("hi", C.c_void_p), ("g", C.c_void_p), ("viol", C.c_void_p), ("bconst", C.c_void_p)]


_P = C.c_void_p
_memoryview_from_memory = C.pythonapi.PyMemoryView_FromMemory      # (address, size, PyBUF_WRITE) -> memoryview, no copy
_memoryview_from_memory.restype = C.py_object
_memoryview_from_memory.argtypes = (C.c_void_p, C.c_ssize_t, C.c_int)
_SIGS = {
    "ktn_create": (C.c_int, [C.POINTER(ktn_options), C.POINTER(_P)]),
    "ktn_destroy": (None, [_P]),
    "ktn_last_error": (C.c_char_p, [_P]),
    "ktn_backend": (C.c_char_p, []),
    "ktn_set_params": (C.c_int, [_P, C.c_double, C.c_double, C.c_int64]),
    "ktn_load_begin": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "ktn_add_rows": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "ktn_load_end": (C.c_int, [_P]),
    "ktn_set_bounds": (C.c_int, [_P, _P, _P]),
    "ktn_jac_structure": (C.c_int, [_P, _P, _P]),
    "ktn_num_rows": (C.c_int64, [_P]),
    "ktn_jac_nnz": (C.c_int64, [_P]),
    "ktn_separate": (C.c_int, [_P, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ktn_gencut_rows": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ktn_separate_ladder": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ktn_fetch_cuts": (C.c_int, [_P] + [_P] * 9),
    "ktn_fetch_cuts_view": (C.c_int, [_P, C.POINTER(ktn_cut_view)]),
    "ktn_get_g": (C.c_int, [_P, _P]),
    "ktn_eval_g": (C.c_int, [_P, _P, _P]),
    "ktn_timings_get": (C.c_int, [_P, C.POINTER(ktn_timings)]),
    "ktn_set_stream": (C.c_int, [_P, _P]),
    "ktn_separate_device_async": (C.c_int, [_P, _P]),
    "ktn_sync_counts": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ktn_algorithmic_bytes": (C.c_int64, [_P]),
    "ktn_comm_unique_id": (C.c_int, [_P]),
    "ktn_comm_init": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "ktn_set_row_offset": (C.c_int, [_P, C.c_int64]),
    "ktn_allgather_cuts_async": (C.c_int, [_P]),
    "ktn_sync_gathered": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ktn_exchange_transport": (C.c_int, [_P]),
    "ktn_fetch_gathered": (C.c_int, [_P] + [_P] * 9),
    "ktn_gathered_error_row": (C.c_int, [_P, C.POINTER(C.c_int64)]),
}
# exported by the CUDA library only (test / bench support, include/ktn.h bottom)
_SYNTH_SIGS = {
    "ktn_synth_rows": (C.c_int, [C.c_int32, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_int64)] + [_P] * 7),
    "ktn_synth_point": (C.c_int, [C.c_int32, C.c_uint64, C.c_int64, _P]),
}
ABI_SYMBOLS = sorted(list(_SIGS) + list(_SYNTH_SIGS))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


class _SynthMixin:
    """ktn_synth_rows / ktn_synth_point of a loaded library (self.dll)."""

    def _bind_synth(self):
        for name, (res, args) in _SYNTH_SIGS.items():
            fn = getattr(self.dll, name)
            fn.restype, fn.argtypes = res, args

    # ---- synthetic instances ----
    def synth_rows(self, kind, seed, num_var, row_begin, nrows):
        """Rows [row_begin, row_begin+nrows) of a synthetic instance as a WireRows batch."""
        nn = C.c_int64(0)
        rc = self.dll.ktn_synth_rows(kind, seed, num_var, row_begin, nrows, C.byref(nn), None, None, None, None, None, None, None)
        if rc != 0:
            raise KtnError(f"ktn_synth_rows failed: {rc}")
        n = nn.value
        w = WireRows(np.empty(nrows + 1, np.int64), np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float64),
                     np.empty(nrows, np.float64), np.empty(nrows, np.float64), np.empty(nrows, np.uint8))
        rc = self.dll.ktn_synth_rows(kind, seed, num_var, row_begin, nrows, C.byref(nn), _ptr(w.expr_ptr), _ptr(w.op), _ptr(w.arg),
                                     _ptr(w.val), _ptr(w.lb), _ptr(w.ub), _ptr(w.flags))
        if rc != 0:
            raise KtnError(f"ktn_synth_rows failed: {rc}")
        return w

    def synth_point(self, kind, seed, num_var):
        x = np.empty(num_var, np.float64)
        rc = self.dll.ktn_synth_point(kind, seed, num_var, _ptr(x))
        if rc != 0:
            raise KtnError(f"ktn_synth_point failed: {rc}")
        return x


class SynthLibrary(_SynthMixin):
    """libktn_synth.so: the deterministic instance generators alone (no CUDA, no separator).  Tests, the CPU baseline and
    `bench.py --impl reference` take their inputs from here, so that they never map the product library."""

    def __init__(self, path=None):
        path = path or SYNTH_LIB_PATH
        if not os.path.exists(path):
            raise KtnError(f"shared library not found: {path} (run `python -c 'import __graft_entry__ as g; g.build()'`)")
        self.path = path
        self.dll = C.CDLL(path)
        self._bind_synth()


class KtnLibrary(_SynthMixin):
    """A loaded shared library exporting include/ktn.h."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise KtnError(f"shared library not found: {path} (run `python -c 'import __graft_entry__ as g; g.build()'`)")
        self.path = path
        self.dll = C.CDLL(path)      # RTLD_LOCAL: the CUDA library, the oracle and the emulator export the same symbols
        for name, (res, args) in _SIGS.items():
            fn = getattr(self.dll, name)
            fn.restype, fn.argtypes = res, args
        self.has_synth = hasattr(self.dll, "ktn_synth_rows")
        if self.has_synth:
            self._bind_synth()
        self.backend = self.dll.ktn_backend().decode()

    def create(self, f_tol=1e-6, cut_coef_rng=1e9, topk=0, device=-1, flags=0, ngpus=0, devices=None):
        """ngpus > 1: ONE handle that shards the rows over `ngpus` devices (ktn_options.ngpus / devices, include/ktn.h)."""
        return Handle(self, f_tol, cut_coef_rng, topk, device, flags, ngpus, devices)


@dataclass
class WireRows:
    """A batch of rows in the expression wire format of include/ktn.h."""
    expr_ptr: np.ndarray  # int64 [nrows+1]
    op: np.ndarray        # int32 [n_nodes]
    arg: np.ndarray       # int32 [n_nodes]
    val: np.ndarray       # float64 [n_nodes]
    lb: np.ndarray        # float64 [nrows]
    ub: np.ndarray        # float64 [nrows]
    flags: np.ndarray     # uint8 [nrows]

    @property
    def nrows(self):
        return len(self.lb)


@dataclass
class CutBatch:
    """Cuts of one round in CSR form (ascending row order); see ktn_fetch_cuts."""
    status: int
    err_row: int
    row_id: np.ndarray
    row_ptr: np.ndarray
    col: np.ndarray
    val: np.ndarray
    lo: np.ndarray
    hi: np.ndarray
    g: np.ndarray
    viol: np.ndarray
    bconst: np.ndarray

    @property
    def n_cuts(self):
        return len(self.row_id)

    def row(self, c):
        s, e = self.row_ptr[c], self.row_ptr[c + 1]
        return self.col[s:e], self.val[s:e]


class Handle:
    def __init__(self, lib, f_tol, cut_coef_rng, topk, device, flags=0, ngpus=0, devices=None):
        self.lib, self.dll = lib, lib.dll
        dev = (C.c_int32 * 16)(*([-1] * 16))
        for s, d in enumerate(devices or []):
            dev[s] = d
        o = ktn_options(C.sizeof(ktn_options), device, f_tol, cut_coef_rng, topk, flags, 0, ngpus, dev)
        p = _P()
        rc = self.dll.ktn_create(C.byref(o), C.byref(p))
        if rc != 0 or not p:
            raise KtnError(f"ktn_create failed with status {rc} ({lib.backend})")
        self.h = p
        self.num_var = self.num_constr = 0

    def close(self):
        if getattr(self, "h", None):
            self.dll.ktn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what, numeric_ok=False):
        if rc < 0 or (rc > 0 and not numeric_ok):
            msg = self.dll.ktn_last_error(self.h)
            raise KtnError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
        return rc

    def set_params(self, f_tol, cut_coef_rng, topk=0):
        self._ck(self.dll.ktn_set_params(self.h, f_tol, cut_coef_rng, topk), "ktn_set_params")

    # ---- loading ----
    def load_begin(self, num_var, num_constr):
        self.num_var, self.num_constr = int(num_var), int(num_constr)
        self._ck(self.dll.ktn_load_begin(self.h, num_var, num_constr), "ktn_load_begin")

    def add_rows(self, first_row, w):
        for a, t in ((w.expr_ptr, np.int64), (w.op, np.int32), (w.arg, np.int32), (w.val, np.float64),
                     (w.lb, np.float64), (w.ub, np.float64), (w.flags, np.uint8)):
            assert a.dtype == t and a.flags.c_contiguous
        n, ne = w.nrows, int(w.expr_ptr[-1])       # the C side reads exactly these many entries: short arrays would be read past their end
        if not (len(w.lb) == len(w.ub) == len(w.flags) == n == len(w.expr_ptr) - 1 and len(w.op) == len(w.arg) == len(w.val) == ne):
            raise ValueError(f"wire rows: {n} rows / {ne} nodes, but lb {len(w.lb)}, ub {len(w.ub)}, flags {len(w.flags)}, op {len(w.op)}, arg {len(w.arg)}, val {len(w.val)}")
        self._ck(self.dll.ktn_add_rows(self.h, first_row, w.nrows, _ptr(w.expr_ptr), _ptr(w.op), _ptr(w.arg), _ptr(w.val),
                                       _ptr(w.lb), _ptr(w.ub), _ptr(w.flags)), "ktn_add_rows")

    def load_end(self):
        self._ck(self.dll.ktn_load_end(self.h), "ktn_load_end")

    def load(self, num_var, w):
        self.load_begin(num_var, w.nrows)
        self.add_rows(0, w)
        self.load_end()

    def set_bounds(self, lb, ub):
        lb = np.ascontiguousarray(lb, np.float64); ub = np.ascontiguousarray(ub, np.float64)
        assert len(lb) == self.num_constr == len(ub)
        self._ck(self.dll.ktn_set_bounds(self.h, _ptr(lb), _ptr(ub)), "ktn_set_bounds")

    def jac_structure(self):
        m = self.dll.ktn_num_rows(self.h)
        rp = np.empty(m + 1, np.int64)
        self._ck(self.dll.ktn_jac_structure(self.h, _ptr(rp), None), "ktn_jac_structure")
        cols = np.empty(int(rp[-1]), np.int32)
        self._ck(self.dll.ktn_jac_structure(self.h, _ptr(rp), _ptr(cols)), "ktn_jac_structure")
        return rp, cols

    # ---- rounds ----
    def _fetch(self, status, nc, nz, err, gathered=False):
        b = CutBatch(status, err, np.empty(nc, np.int64), np.empty(nc + 1, np.int64), np.empty(nz, np.int32), np.empty(nz, np.float64),
                     np.empty(nc, np.float64), np.empty(nc, np.float64), np.empty(nc, np.float64), np.empty(nc, np.float64), np.empty(nc, np.float64))
        fn = self.dll.ktn_fetch_gathered if gathered else self.dll.ktn_fetch_cuts
        self._ck(fn(self.h, _ptr(b.row_id), _ptr(b.row_ptr), _ptr(b.col), _ptr(b.val), _ptr(b.lo), _ptr(b.hi), _ptr(b.g), _ptr(b.viol), _ptr(b.bconst)),
                 "ktn_fetch_cuts", numeric_ok=True)
        return b

    def _fetch_view(self, status, err):
        """Zero-copy CutBatch: numpy views of the library's pinned buffer (ktn_fetch_cuts_view).  Valid until the second
        later view fetch on this handle; copy what must live longer."""
        v = ktn_cut_view()
        self._ck(self.dll.ktn_fetch_cuts_view(self.h, C.byref(v)), "ktn_fetch_cuts_view")
        nc, nz = v.n_cuts, v.nnz
        # ONE memoryview over the span of the sections, the arrays are offsets into it: 6 us for the lean view where nine
        # np.ctypeslib.as_array calls took 40 (this call sits inside the end-to-end round)
        secs = ((v.row_id, nc, 8, np.int64), (v.row_ptr, nc + 1, 8, np.int64), (v.col, nz, 4, np.int32), (v.val, nz, 8, np.float64),
                (v.lo, nc, 8, np.float64), (v.hi, nc, 8, np.float64), (v.g, nc, 8, np.float64), (v.viol, nc, 8, np.float64), (v.bconst, nc, 8, np.float64))
        live = [(p, n * sz) for p, n, sz, _ in secs if p and n]
        out = []
        if live:
            base = min(p for p, _ in live)
            mv = _memoryview_from_memory(base, max(p + b for p, b in live) - base, 0x200)
        for p, n, sz, dt in secs:
            out.append(np.frombuffer(mv, dtype=dt, count=n, offset=p - base) if p and n else np.empty(0, dt))
        return CutBatch(status, err, *out)

    def separate(self, xstar, fetch=True, view=False):
        """One round at x*.  view=True returns zero-copy views of the library's pinned buffer (the hot path of optimize!);
        the default copies into fresh arrays."""
        x = np.ascontiguousarray(xstar, np.float64)
        assert len(x) == self.num_var
        nc, nz, er = C.c_int64(), C.c_int64(), C.c_int64()
        st = self._ck(self.dll.ktn_separate(self.h, _ptr(x), C.byref(nc), C.byref(nz), C.byref(er)), "ktn_separate", numeric_ok=True)
        if not fetch:
            return st, nc.value, nz.value, er.value
        if view:
            return self._fetch_view(st, er.value)
        return self._fetch(st, nc.value, nz.value, er.value)

    def separate_ladder(self, ray, n_first=2, n_last=1023, view=False):
        """boundroutine's search (src/model.jl:175-197): the cuts at the first point 2^n * ray that violates a row.  Returns (n_hit, CutBatch)."""
        r = np.ascontiguousarray(ray, np.float64)
        assert len(r) == self.num_var
        hit, nc, nz, er = C.c_int32(-1), C.c_int64(), C.c_int64(), C.c_int64()
        st = self._ck(self.dll.ktn_separate_ladder(self.h, _ptr(r), n_first, n_last, C.byref(hit), C.byref(nc), C.byref(nz), C.byref(er)),
                      "ktn_separate_ladder", numeric_ok=True)
        return hit.value, (self._fetch_view(st, er.value) if view else self._fetch(st, nc.value, nz.value, er.value))

    def gencut_rows(self, x, rows, round_coefs=False):
        x = np.ascontiguousarray(x, np.float64); rows = np.ascontiguousarray(rows, np.int64)
        assert len(x) == self.num_var
        nc, nz, er = C.c_int64(), C.c_int64(), C.c_int64()
        st = self._ck(self.dll.ktn_gencut_rows(self.h, _ptr(x), _ptr(rows), len(rows), int(bool(round_coefs)),
                                               C.byref(nc), C.byref(nz), C.byref(er)), "ktn_gencut_rows", numeric_ok=True)
        return self._fetch(st, nc.value, nz.value, er.value)

    def get_g(self):
        g = np.empty(self.num_constr, np.float64)
        self._ck(self.dll.ktn_get_g(self.h, _ptr(g)), "ktn_get_g")
        return g

    def eval_g(self, x):
        x = np.ascontiguousarray(x, np.float64); g = np.empty(self.num_constr, np.float64)
        self._ck(self.dll.ktn_eval_g(self.h, _ptr(x), _ptr(g)), "ktn_eval_g")
        return g

    def timings(self):
        t = ktn_timings()
        self._ck(self.dll.ktn_timings_get(self.h, C.byref(t)), "ktn_timings_get")
        return {k: getattr(t, k) for k, _ in ktn_timings._fields_}

    def algorithmic_bytes(self):
        return int(self.dll.ktn_algorithmic_bytes(self.h))

    # ---- device-resident / sharded ----
    def set_stream(self, stream_ptr):
        self._ck(self.dll.ktn_set_stream(self.h, _P(stream_ptr)), "ktn_set_stream")

    def separate_device_async(self, d_x_ptr):
        self._ck(self.dll.ktn_separate_device_async(self.h, _P(d_x_ptr)), "ktn_separate_device_async")

    def sync_counts(self):
        nc, nz, er = C.c_int64(), C.c_int64(), C.c_int64()
        st = self._ck(self.dll.ktn_sync_counts(self.h, C.byref(nc), C.byref(nz), C.byref(er)), "ktn_sync_counts", numeric_ok=True)
        return st, nc.value, nz.value, er.value

    def fetch_last(self):
        st, nc, nz, er = self.sync_counts()
        return self._fetch(st, nc, nz, er)

    def comm_init(self, nranks, rank, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._ck(self.dll.ktn_comm_init(self.h, nranks, rank, buf), "ktn_comm_init")

    def set_row_offset(self, first_global_row):
        self._ck(self.dll.ktn_set_row_offset(self.h, first_global_row), "ktn_set_row_offset")

    def allgather_cuts_async(self):
        self._ck(self.dll.ktn_allgather_cuts_async(self.h), "ktn_allgather_cuts_async")

    def sync_gathered(self):
        nc, nz = C.c_int64(), C.c_int64()
        self.gathered_status = self._ck(self.dll.ktn_sync_gathered(self.h, C.byref(nc), C.byref(nz)), "ktn_sync_gathered", numeric_ok=True)
        return nc.value, nz.value

    def gathered_error_row(self):
        er = C.c_int64(-1)
        self._ck(self.dll.ktn_gathered_error_row(self.h, C.byref(er)), "ktn_gathered_error_row")
        return er.value

    def exchange_transport(self):
        return {0: "none", 1: "nccl", 2: "peer-push"}[int(self.dll.ktn_exchange_transport(self.h))]

    def fetch_gathered(self):
        nc, nz = self.sync_gathered()
        return self._fetch(self.gathered_status, nc, nz, self.gathered_error_row(), gathered=True)


def comm_unique_id(lib):
    buf = C.create_string_buffer(128)
    rc = lib.dll.ktn_comm_unique_id(buf)
    if rc != 0:
        raise KtnError(f"ktn_comm_unique_id failed: {rc}")
    return buf.raw


_cuda_lib = None


def load_cuda_library():
    """The product's only backend.  Raises if libktn.so has not been built; never falls back to a CPU path."""
    global _cuda_lib
    if _cuda_lib is None:
        lib = KtnLibrary(CUDA_LIB_PATH)
        if lib.backend != "cuda":
            raise KtnError(f"{CUDA_LIB_PATH} is not the CUDA backend (got {lib.backend!r})")
        _cuda_lib = lib
    return _cuda_lib
