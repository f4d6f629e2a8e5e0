"""ctypes binding of include/bydb_synth.h: the host-side part writer and the synthetic generator."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import capi

F_LATENCY, F_WALK3, F_INT1000, F_UNIFORM = 1, 2, 3, 4
I_DELTA, I_FLUCT, I_RANDOM100, I_COUNTER = 10, 11, 12, 13


class _WColumn(C.Structure):
    _fields_ = [("name", C.c_char_p), ("value_type", C.c_int32), ("dec_digits", C.c_int32), ("i64", C.c_void_p),
                ("f64", C.c_void_p), ("dec_k", C.c_void_p), ("str_idx", C.c_void_p), ("str_values", C.POINTER(C.c_char_p)),
                ("n_str_values", C.c_uint32), ("reserved", C.c_uint32)]


class _WriteInput(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("series_ids", C.c_void_p), ("timestamps", C.c_void_p), ("versions", C.c_void_p),
                ("n_fields", C.c_uint32), ("fields", C.POINTER(_WColumn)), ("tag_family", C.c_char_p), ("n_tags", C.c_uint32),
                ("tags", C.POINTER(_WColumn)), ("threads", C.c_uint32), ("reserved", C.c_uint32)]


class _SynthField(C.Structure):
    _fields_ = [("name", C.c_char_p), ("kind", C.c_int32), ("reserved", C.c_int32)]


class _SynthSpec(C.Structure):
    _fields_ = [("n_series", C.c_uint64), ("n_points", C.c_uint64), ("sid0", C.c_uint64), ("sid_step", C.c_uint64),
                ("t0", C.c_int64), ("t_step", C.c_int64), ("n_fields", C.c_uint32), ("fields", C.POINTER(_SynthField)),
                ("region_values", C.c_uint32), ("region_run", C.c_uint32), ("code_tag", C.c_uint32), ("threads", C.c_uint32),
                ("seed", C.c_uint64)]


_bound = False


def _lib():
    global _bound
    L = capi.load_library()
    if not _bound:
        L.bydb_part_write.argtypes = [C.POINTER(_WriteInput), C.POINTER(C.c_void_p)]
        L.bydb_synth_part.argtypes = [C.POINTER(_SynthSpec), C.POINTER(C.c_void_p)]
        L.bydb_part_image_n_files.argtypes = [C.c_void_p]
        L.bydb_part_image_n_files.restype = C.c_uint32
        L.bydb_part_image_file_name.argtypes = [C.c_void_p, C.c_uint32]
        L.bydb_part_image_file_name.restype = C.c_char_p
        L.bydb_part_image_file_data.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]
        L.bydb_part_image_file_data.restype = C.c_void_p
        L.bydb_part_image_counts.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.bydb_part_image_free.argtypes = [C.c_void_p]
        _bound = True
    return L


class PartImage:
    """A finished part in host memory; file views are zero-copy numpy arrays valid while the image lives."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None):
            _lib().bydb_part_image_free(self._h)
            self._h = None

    def files(self) -> Dict[str, np.ndarray]:
        out = {}
        L = _lib()
        for i in range(L.bydb_part_image_n_files(self._h)):
            ln = C.c_uint64(0)
            p = L.bydb_part_image_file_data(self._h, i, C.byref(ln))
            name = L.bydb_part_image_file_name(self._h, i).decode()
            if ln.value:
                out[name] = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (ln.value,))
            else:
                out[name] = np.zeros(0, dtype=np.uint8)
        return out

    def counts(self) -> Tuple[int, int]:
        a, b = C.c_uint64(), C.c_uint64()
        _lib().bydb_part_image_counts(self._h, C.byref(a), C.byref(b))
        return a.value, b.value


def _wcolumn(name, vt, values, keep, dec_digits=-1, str_values=None) -> _WColumn:
    c = _WColumn()
    nb = name.encode()
    keep.append(nb)
    c.name, c.value_type, c.dec_digits = nb, vt, -1
    if vt == capi.VT_INT64:
        a = np.ascontiguousarray(values, dtype=np.int64)
        keep.append(a)
        c.i64 = a.ctypes.data
    elif vt == capi.VT_FLOAT64 and dec_digits >= 0:
        a = np.ascontiguousarray(values, dtype=np.int64)
        keep.append(a)
        c.dec_k, c.dec_digits = a.ctypes.data, dec_digits
    elif vt == capi.VT_FLOAT64:
        a = np.ascontiguousarray(values, dtype=np.float64)
        keep.append(a)
        c.f64 = a.ctypes.data
    else:
        a = np.ascontiguousarray(values, dtype=np.uint32)
        sv = (C.c_char_p * len(str_values))(*[s if isinstance(s, bytes) else s.encode() for s in str_values])
        keep.extend([a, sv])
        c.str_idx, c.str_values, c.n_str_values = a.ctypes.data, sv, len(str_values)
    return c


def write_part(series_ids, timestamps, versions, fields: Sequence[tuple], tag_family: Optional[str] = None,
               tags: Sequence[tuple] = (), threads: int = 0) -> PartImage:
    """fields / tags: (name, value_type, values[, dec_digits | str_values]).  Rows must be sorted by (sid, ts)."""
    keep: list = []
    sid = np.ascontiguousarray(series_ids, dtype=np.uint64)
    ts = np.ascontiguousarray(timestamps, dtype=np.int64)
    ver = np.ascontiguousarray(versions, dtype=np.int64)
    inp = _WriteInput()
    inp.n_rows, inp.series_ids, inp.timestamps, inp.versions = sid.size, sid.ctypes.data, ts.ctypes.data, ver.ctypes.data
    fa = (_WColumn * max(len(fields), 1))()
    for i, f in enumerate(fields):
        fa[i] = _wcolumn(f[0], f[1], f[2], keep, dec_digits=f[3] if len(f) > 3 and f[1] == capi.VT_FLOAT64 else -1)
    ta = (_WColumn * max(len(tags), 1))()
    for i, t in enumerate(tags):
        ta[i] = _wcolumn(t[0], t[1], t[2], keep, str_values=t[3] if len(t) > 3 else None)
    inp.n_fields, inp.fields = len(fields), fa
    fam = tag_family.encode() if tag_family else None
    inp.tag_family, inp.n_tags, inp.tags, inp.threads = fam, len(tags), ta, threads
    out = C.c_void_p()
    rc = _lib().bydb_part_write(C.byref(inp), C.byref(out))
    if rc != 0:
        raise capi.BydbError(rc, "bydb_part_write failed (rows must be sorted by (sid, ts), unique, ts != 0)")
    return PartImage(out)


def synth_part(n_series: int, n_points: int, fields: Sequence[Tuple[str, int]], sid0: int = 1, sid_step: int = 1,
               t0: int = 1_700_000_000_000_000_000, t_step: int = 60_000_000_000, region_values: int = 0, region_run: int = 0,
               code_tag: bool = False, seed: int = 0xB200, threads: int = 0, zone_tag: bool = False) -> PartImage:
    keep = []
    fa = (_SynthField * max(len(fields), 1))()
    for i, (name, kind) in enumerate(fields):
        nb = name.encode()
        keep.append(nb)
        fa[i].name, fa[i].kind = nb, kind
    sp = _SynthSpec(n_series, n_points, sid0, sid_step, t0, t_step, len(fields), fa, region_values, region_run, int(bool(code_tag)) | (2 if zone_tag else 0), threads, seed)
    out = C.c_void_p()
    rc = _lib().bydb_synth_part(C.byref(sp), C.byref(out))
    if rc != 0:
        raise capi.BydbError(rc, "bydb_synth_part failed")
    return PartImage(out)
