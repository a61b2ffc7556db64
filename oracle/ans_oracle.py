"""ctypes front-end of oracle/ans_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Mirrors the pybind11 API of the reference coder (cbench/csrc/ans/rans64.hpp:127-149,
tans.hpp:147-157) so parity tests read like the reference's tests/ans_test.py.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Parity status: pinned against oracle/_ref (the unmodified reference) and tests/golden/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libans_oracle.so")
_lib = None

_ERRORS = {
    -1: "Error (generic)",
    -2: "Destination buffer is too small",
    -3: "Src size incorrect",
    -4: "tableLog requires too much memory : unsupported",
    -5: "Unsupported max Symbol Value : too large",
    -6: "symbol or index out of range",
    -7: "output capacity too small",
}


def build():
    """Compile the C oracle (gcc, seconds)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "ans_oracle.c")
        if not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_rans64_put_rcp.restype = C.c_uint64
        _lib.orc_rans64_put_rcp.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        _lib.orc_bls_bound.restype = C.c_int64
        _lib.orc_bls_bound.argtypes = [C.c_int64, C.c_int64]
        _lib.orc_bls_num_chunks.restype = C.c_int64
        _lib.orc_bls_num_chunks.argtypes = [C.c_int64, C.c_int64]
    return _lib


def _check(rc):
    if rc != 0:
        raise ValueError(_ERRORS.get(rc, f"oracle error {rc}"))


def _i32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.int32)


def _p(a, t=C.c_int32):
    return a.ctypes.data_as(C.POINTER(t))


def pmf_to_quantized_cdf(pmf, precision):
    pmf = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    cdf = np.zeros(len(pmf) + 1, dtype=np.int32)
    _check(lib().orc_pmf_to_quantized_cdf(_p(pmf, C.c_float), len(pmf), int(precision), _p(cdf)))
    return cdf.tolist()


class _R64Tables(C.Structure):
    _fields_ = [("T", C.c_int), ("stride", C.c_int), ("precision", C.c_int), ("bypass", C.c_int),
                ("bypass_precision", C.c_int), ("cdfs", C.POINTER(C.c_int32)),
                ("sizes", C.POINTER(C.c_int32)), ("offsets", C.POINTER(C.c_int32))]


class _Rans64Base:
    def __init__(self, freq_precision=16, bypass_coding=True, bypass_precision=4):
        self.freq_precision = int(freq_precision)
        self.bypass_coding = bool(bypass_coding)
        self.bypass_precision = int(bypass_precision)
        self._init = False

    def init_params(self, freqs, num_symbols, offsets):
        freqs, num_symbols, offsets = _i32(freqs), _i32(num_symbols), _i32(offsets)
        if freqs.ndim != 2 or freqs.shape[0] != num_symbols.size:
            raise ValueError("freqs should be 2-dimensional with shape (num_symbols.size(), >num_symbols.max())")
        T, M = freqs.shape
        stride = int(num_symbols.max()) + 2
        self._cdfs = np.zeros((T, stride), dtype=np.int32)
        self._sizes = np.zeros(T, dtype=np.int32)
        _check(lib().orc_rans64_init_params(_p(freqs), T, M, _p(num_symbols), self.freq_precision,
                                            _p(self._cdfs), stride, _p(self._sizes)))
        self._offsets = offsets.reshape(-1).copy()
        self._make_tb()

    def init_cdf_params(self, cdfs, cdfs_sizes, offsets):
        cdfs, cdfs_sizes, offsets = _i32(cdfs), _i32(cdfs_sizes), _i32(offsets)
        if cdfs.ndim != 2 or cdfs.shape[0] != cdfs_sizes.size:
            raise ValueError("cdfs should be 2-dimensional with shape (cdfs_sizes.size(), >cdfs_sizes.max())")
        self._cdfs = cdfs.copy()
        self._sizes = cdfs_sizes.reshape(-1).copy()
        self._offsets = offsets.reshape(-1).copy()
        self._make_tb()

    def _make_tb(self):
        self._tb = _R64Tables(self._cdfs.shape[0], self._cdfs.shape[1], self.freq_precision,
                              int(self.bypass_coding), self.bypass_precision, _p(self._cdfs), _p(self._sizes),
                              _p(self._offsets))
        self._init = True

    def get_cdfs(self):
        if not self._init:
            return np.zeros((0,), dtype=np.int32)
        m = int(self._sizes.max())
        return self._cdfs[:, :m].copy()

    def _need_init(self):
        if not self._init:
            raise ValueError("ANS not initialized!")


class Rans64Encoder(_Rans64Base):
    def encode_with_indexes(self, symbols, indexes, ar_indexes=None, ar_offsets=None, cache=0):
        self._need_init()
        assert not cache, "cache mode is not part of the oracle"
        sym, idx = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1)
        n = sym.size
        cap = n * 12 + 16
        scratch = np.empty(cap, dtype=np.uint32)
        first = C.c_int64(0)
        _check(lib().orc_rans64_encode(C.byref(self._tb), _p(sym), _p(idx), C.c_int64(n), _p(scratch, C.c_uint32),
                                       C.c_int64(cap), C.byref(first)))
        return scratch[first.value:].tobytes()

    # -- this repo's multi-lane segment format (CPU specification) --
    def encode_lanes(self, symbols, indexes, chunk_syms):
        self._need_init()
        sym, idx = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1)
        n = sym.size
        cap = lib().orc_bls_bound(n, chunk_syms)
        out = np.empty(cap, dtype=np.uint8)
        out_len = C.c_int64(0)
        _check(lib().orc_bls_encode(C.byref(self._tb), _p(sym), _p(idx), C.c_int64(n), C.c_int64(chunk_syms),
                                    _p(out, C.c_uint8), C.c_int64(cap), C.byref(out_len)))
        return out[:out_len.value].tobytes()

    def encode_lanes_slices(self, symbols, indexes, slice_n, n_chunks):
        """Multi-slice segment (the y path: one slice per coding group, lane states carried across groups):
        chunk_syms[g] = ceil(slice_n[g] / n_chunks) rounded up to 128."""
        self._need_init()
        sym, idx = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1)
        sn = np.ascontiguousarray(slice_n, dtype=np.int64)
        cs = np.maximum(128, ((-(-sn // n_chunks) + 127) // 128) * 128).astype(np.int64)
        n_chunks = int(max(1 if sn.size == 0 else 0, (-(-sn // cs)).max() if sn.size else 0))  # drop chunks that own nothing
        cap = 64 + 4 * sn.size + n_chunks * 132 + int(sn.sum()) * 24 + 64
        out = np.empty(cap, dtype=np.uint8)
        out_len = C.c_int64(0)
        _check(lib().orc_bls_encode_slices(C.byref(self._tb), _p(sym), _p(idx), C.c_int(sn.size), _p(sn, C.c_int64),
                                           _p(cs, C.c_int64), C.c_int64(n_chunks), _p(out, C.c_uint8), C.c_int64(cap),
                                           C.byref(out_len)))
        return out[:out_len.value].tobytes()


class _DState(C.Structure):
    _fields_ = [("x", C.c_uint64), ("pos", C.c_int64)]


class Rans64Decoder(_Rans64Base):
    def set_stream(self, stream):
        self._stream = np.frombuffer(bytes(stream), dtype=np.uint32).copy()
        self._st = _DState()
        lib().orc_rans64_set_stream(C.byref(self._st), _p(self._stream, C.c_uint32))

    def decode_stream(self, indexes, ar_indexes=None, ar_offsets=None):
        self._need_init()
        idx = _i32(indexes)
        out = np.empty(idx.shape, dtype=np.int32)
        _check(lib().orc_rans64_decode_stream(C.byref(self._tb), C.byref(self._st), _p(self._stream, C.c_uint32),
                                              _p(idx), C.c_int64(idx.size), _p(out)))
        return out

    def decode_with_indexes(self, encoded, indexes, ar_indexes=None, ar_offsets=None):
        self._need_init()
        self.set_stream(encoded)
        return self.decode_stream(indexes)

    def decode_lanes(self, encoded, indexes, chunk_syms):
        self._need_init()
        idx = _i32(indexes)
        enc = np.frombuffer(bytes(encoded), dtype=np.uint8)
        out = np.empty(idx.shape, dtype=np.int32)
        consumed = C.c_int64(0)
        _check(lib().orc_bls_decode(C.byref(self._tb), _p(enc, C.c_uint8), C.c_int64(enc.size), _p(idx),
                                    C.c_int64(idx.size), C.c_int64(chunk_syms), _p(out), C.byref(consumed)))
        return out, consumed.value

    def decode_lanes_slices(self, encoded, indexes, slice_n):
        self._need_init()
        idx = _i32(indexes).reshape(-1)
        sn = np.ascontiguousarray(slice_n, dtype=np.int64)
        enc = np.frombuffer(bytes(encoded), dtype=np.uint8)
        out = np.empty(idx.shape, dtype=np.int32)
        consumed = C.c_int64(0)
        _check(lib().orc_bls_decode_slices(C.byref(self._tb), _p(enc, C.c_uint8), C.c_int64(enc.size), _p(idx),
                                           C.c_int(sn.size), _p(sn, C.c_int64), _p(out), C.byref(consumed)))
        return out, consumed.value


# ---------------------------------------------------------------------------------------------- tANS
class _DEntry(C.Structure):
    _fields_ = [("newState", C.c_uint32), ("symbol", C.c_uint16), ("nbBits", C.c_uint16)]


class _TansTables(C.Structure):
    _fields_ = [("T", C.c_int), ("tableLog", C.c_int), ("bypass", C.c_int), ("bypass_precision", C.c_int),
                ("max_nsym", C.c_int),
                ("nsym", C.POINTER(C.c_int32)), ("offsets", C.POINTER(C.c_int32)),
                ("ct_state", C.POINTER(C.c_uint16)), ("ct_nb", C.POINTER(C.c_uint32)), ("ct_fs", C.POINTER(C.c_int32)),
                ("dt", C.POINTER(_DEntry)), ("dt_fast", C.POINTER(C.c_int32)),
                ("bct_state", C.POINTER(C.c_uint16)), ("bct_nb", C.POINTER(C.c_uint32)), ("bct_fs", C.POINTER(C.c_int32)),
                ("bdt", C.POINTER(_DEntry)), ("bdt_fast", C.c_int)]


DENTRY_DTYPE = np.dtype([("newState", np.uint32), ("symbol", np.uint16), ("nbBits", np.uint16)])


def tans_normalize(count, table_log):
    count = np.ascontiguousarray(np.asarray(count), dtype=np.uint32)
    norm = np.zeros(count.size, dtype=np.int16)
    _check(lib().orc_tans_normalize(_p(norm, C.c_int16), C.c_uint32(table_log), _p(count, C.c_uint32), count.size))
    return norm


def tans_build_ctable(norm, table_log):
    norm = np.ascontiguousarray(norm, dtype=np.int16)
    st = np.zeros(1 << table_log, dtype=np.uint16)
    nb = np.zeros(norm.size, dtype=np.uint32)
    fs = np.zeros(norm.size, dtype=np.int32)
    _check(lib().orc_tans_build_ctable(_p(norm, C.c_int16), norm.size, C.c_uint32(table_log), _p(st, C.c_uint16),
                                       _p(nb, C.c_uint32), _p(fs)))
    return st, nb, fs


def tans_build_dtable(norm, table_log):
    norm = np.ascontiguousarray(norm, dtype=np.int16)
    # the reference checks tableLog before touching the table (tans.cpp:274)
    dt = np.zeros(1 << min(table_log, 16), dtype=DENTRY_DTYPE)
    fast = C.c_int(0)
    _check(lib().orc_tans_build_dtable(_p(norm, C.c_int16), norm.size, C.c_uint32(table_log),
                                       dt.ctypes.data_as(C.POINTER(_DEntry)), C.byref(fast)))
    return dt, fast.value


class _TansBase:
    def __init__(self, table_log=11, max_symbol_value=255, bypass_coding=False, bypass_precision=4):
        self.table_log = int(table_log)
        self.max_symbol_value = int(max_symbol_value)
        self.bypass_coding = bool(bypass_coding)
        self.bypass_precision = int(bypass_precision)
        self._init = False

    def init_params(self, freqs, num_symbols, offsets):
        freqs, num_symbols, offsets = _i32(freqs), _i32(num_symbols), _i32(offsets)
        if freqs.ndim != 2 or freqs.shape[0] != num_symbols.size:
            raise ValueError("freqs should be 2-dimensional with shape (num_symbols.size(), >num_symbols.max())")
        T = freqs.shape[0]
        tl, tsz = self.table_log, 1 << self.table_log
        self._nsym = num_symbols.reshape(-1).copy()
        self._offsets = offsets.reshape(-1).copy()
        self._max_nsym = int(self._nsym.max())
        self._ct_state = np.zeros((T, tsz), dtype=np.uint16)
        self._ct_nb = np.zeros((T, self._max_nsym), dtype=np.uint32)
        self._ct_fs = np.zeros((T, self._max_nsym), dtype=np.int32)
        self._dt = np.zeros((T, tsz), dtype=DENTRY_DTYPE)
        self._dt_fast = np.zeros(T, dtype=np.int32)
        for t in range(T):
            n = int(self._nsym[t])
            norm = tans_normalize(freqs[t, :n].astype(np.uint32), tl)
            if self._is_encoder:
                st, nb, fs = tans_build_ctable(norm, tl)
                self._ct_state[t], self._ct_nb[t, :n], self._ct_fs[t, :n] = st, nb, fs
            else:
                dt, fast = tans_build_dtable(norm, tl)
                self._dt[t], self._dt_fast[t] = dt, fast
        nb_ = 1 << self.bypass_precision
        self._bct = (np.zeros(tsz, np.uint16), np.zeros(nb_, np.uint32), np.zeros(nb_, np.int32))
        self._bdt, self._bdt_fast = np.zeros(tsz, dtype=DENTRY_DTYPE), 0
        if self.bypass_coding:
            norm = tans_normalize(np.ones(nb_, dtype=np.uint32), tl)
            if self._is_encoder:
                self._bct = tans_build_ctable(norm, tl)
            else:
                self._bdt, self._bdt_fast = tans_build_dtable(norm, tl)
        self._tb = _TansTables(T, tl, int(self.bypass_coding), self.bypass_precision, self._max_nsym,
                               _p(self._nsym), _p(self._offsets),
                               _p(self._ct_state, C.c_uint16), _p(self._ct_nb, C.c_uint32), _p(self._ct_fs),
                               self._dt.ctypes.data_as(C.POINTER(_DEntry)), _p(self._dt_fast),
                               _p(self._bct[0], C.c_uint16), _p(self._bct[1], C.c_uint32), _p(self._bct[2]),
                               self._bdt.ctypes.data_as(C.POINTER(_DEntry)), int(self._bdt_fast))
        self._init = True


class TansEncoder(_TansBase):
    _is_encoder = True

    def encode_with_indexes(self, symbols, indexes, ar_indexes=None, ar_offsets=None, cache=0):
        if not self._init:
            raise ValueError("ANS not initialized!")
        sym, idx = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1)
        n = sym.size
        cap = max(n * self.table_log // 8, 1) + 16
        out = np.zeros(cap, dtype=np.uint8)
        out_len = C.c_int64(0)
        _check(lib().orc_tans_encode(C.byref(self._tb), _p(sym), _p(idx), C.c_int64(n), _p(out, C.c_uint8),
                                     C.c_int64(cap), C.byref(out_len)))
        return out[:out_len.value].tobytes()


class TansDecoder(_TansBase):
    _is_encoder = False

    def decode_with_indexes(self, encoded, indexes, ar_indexes=None, ar_offsets=None):
        if not self._init:
            raise ValueError("ANS not initialized!")
        idx = _i32(indexes)
        enc = np.frombuffer(bytes(encoded), dtype=np.uint8)
        out = np.empty(idx.shape, dtype=np.int32)
        _check(lib().orc_tans_decode(C.byref(self._tb), _p(enc, C.c_uint8), C.c_int64(enc.size), _p(idx),
                                     C.c_int64(idx.size), _p(out)))
        return out
