"""Drop-in for the reference's native coder module ``cbench.ans`` (cbench/csrc/ans/lib.cpp:10-33), backed by
the sm_100a CUDA library through the C ABI (include/basic_b200.h).

Same classes, method names, argument meaning and error behaviour as the pybind11 module
(rans64.hpp:127-149, tans.hpp:147-157):

    Rans64Encoder / Rans64Decoder(freq_precision=16, bypass_coding=True, bypass_precision=4)
    TansEncoder / TansDecoder(table_log=11, max_symbol_value=255, bypass_coding=False, bypass_precision=4)
    .init_params(freqs, num_symbols, offsets)   .init_cdf_params(cdfs, cdfs_sizes, offsets)   .get_cdfs()
    .encode_with_indexes(symbols, indexes, ar_indexes=None, ar_offsets=None, cache=0) -> bytes    .flush() -> bytes
    .decode_with_indexes(encoded, indexes, ...) -> int32 array   .set_stream(bytes)   .decode_stream(indexes, ...)
    pmf_to_quantized_cdf(pmf, precision) -> list[int]

Two additions: arrays may be numpy arrays OR torch CUDA tensors (then nothing crosses PCIe), and every
class takes ``lanes=`` (1 = the reference bitstream byte for byte -- the default; 0 = multi-lane container
sized for <= 0.5 % overhead; N > 1 = ceil(N / 32) chunks of 32 interleaved lanes) and ``device=``.
The in-coder autoregressive table lookup (``init_ar_params`` + ``ar_indexes`` / ``ar_offsets``, ans_interface.hpp:58-105,
the table branch) is provided for the rANS coder in the reference-compatible single-stream mode (lanes = 1:
the lookup makes every symbol depend on earlier ones).  Custom AR ops (``init_custom_ar_ops``) are not.
"""
import ctypes as C

import numpy as np

from . import _native as N

try:  # torch is only needed when CUDA tensors are passed
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_tensor(a):
    return torch is not None and isinstance(a, torch.Tensor)


class _Arg:
    """int32, contiguous view of a numpy array / torch tensor + its raw pointer (keeps the storage alive)."""

    def __init__(self, a, device_index):
        if _is_tensor(a):
            if a.is_cuda and a.device.index != device_index:
                raise ValueError("tensor lives on another CUDA device than the coder")
            self.obj = a.detach().to(torch.int32).contiguous()
            self.ptr, self.size, self.shape = self.obj.data_ptr(), self.obj.numel(), tuple(self.obj.shape)
            self.cuda = self.obj.is_cuda
        else:
            self.obj = np.ascontiguousarray(np.asarray(a), dtype=np.int32)  # forcecast like py::array_t<int32_t>
            self.ptr, self.size, self.shape = self.obj.ctypes.data, self.obj.size, self.obj.shape
            self.cuda = False


def _stream(*args):
    for a in args:
        if isinstance(a, _Arg) and a.cuda:
            return torch.cuda.current_stream(a.obj.device).cuda_stream
    return 0


def _no_ar(ar_indexes, ar_offsets):
    if ar_indexes is not None or ar_offsets is not None:
        raise NotImplementedError("ar_indexes / ar_offsets need init_ar_params (rANS, lanes=1)")


def pmf_to_quantized_cdf(pmf, precision, device=0):
    """rans64.cpp:69-126, computed on the GPU."""
    pmf = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    out = np.zeros(pmf.size + 1, dtype=np.int32)
    N.check(N.lib().basic_pmf_to_quantized_cdf(pmf.ctypes.data, pmf.size, int(precision), int(device), out.ctypes.data))
    return out.tolist()


class _Coder:
    _kind = N.KIND_RANS64
    _role = N.ROLE_BOTH

    def __init__(self, precision, max_symbol_value, bypass_coding, bypass_precision, lanes, device):
        N.require_gpu()
        self.lanes = int(lanes)
        self.device = int(device if device is not None else 0)
        self._h = C.c_void_p()
        N.check(N.lib().basic_coder_create(self._kind | (self._role << 4), int(precision), int(max_symbol_value),
                                           int(bool(bypass_coding)), int(bypass_precision), self.device, C.byref(self._h)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                N.lib().basic_coder_destroy(h)
            except Exception:  # interpreter shutdown
                pass

    # ---- tables
    def init_params(self, freqs, num_symbols, offsets):
        freqs = np.ascontiguousarray(np.asarray(freqs), dtype=np.int32)
        num_symbols = np.ascontiguousarray(np.asarray(num_symbols), dtype=np.int32).reshape(-1)
        offsets = np.ascontiguousarray(np.asarray(offsets), dtype=np.int32).reshape(-1)
        if freqs.ndim != 2 or freqs.shape[0] != num_symbols.size:
            raise ValueError("freqs should be 2-dimensional with shape (num_symbols.size(), >num_symbols.max())")
        if offsets.size != num_symbols.size:
            raise ValueError("offsets should have one entry per table")
        N.check(N.lib().basic_coder_init_params(self._h, freqs.ctypes.data, freqs.shape[0], freqs.shape[1],
                                                num_symbols.ctypes.data, offsets.ctypes.data))

    # ---- in-coder autoregressive table lookup (ans_interface.cpp:75-135)
    _ar_order = 0

    def init_ar_params(self, ar_tables, ar_offsets):
        tab = np.ascontiguousarray(np.asarray(ar_tables), dtype=np.int32)
        off = np.asarray(ar_offsets)
        norder = tab.ndim - 2
        if off.ndim != 3 or off.shape[1] != norder or off.shape[0] != tab.shape[0]:
            raise ValueError("ar_offset should be 3-dimensional with shape (ar_tables_size, ar_order, <=data_dims)")
        if norder <= 0:
            raise ValueError("ar_tables should be at least 3-dimensional with shape (ar_tables_size, index_dim, *ar_order_dims)")
        if norder > 2:
            raise ValueError("Too many dimensions!")
        if self._kind != N.KIND_RANS64 or self.lanes != 1:
            raise NotImplementedError("the AR lookup is provided for the rANS coder at lanes=1 (every symbol depends on earlier ones)")
        N.check(N.lib().basic_coder_init_ar_params(self._h, tab.ctypes.data, tab.shape[0], tab.shape[1], tab.shape[2],
                                                   tab.shape[3] if norder == 2 else 0))
        self._ar_order = norder

    def _ar_args(self, ar_indexes, ar_offsets, n):
        if ar_offsets is None:
            raise ValueError("ar_offsets is required for ar coding!")
        off = _Arg(ar_offsets, self.device)
        if len(off.shape) != 2 or off.shape[0] != self._ar_order or off.shape[1] != n:
            raise ValueError("ar_offsets should have shape (ar_order, number of symbols)")
        ai = _Arg(ar_indexes, self.device) if ar_indexes is not None else None
        if ai is not None and ai.size != n:
            raise ValueError("ar_indexes should have one entry per symbol")
        return ai, off

    def init_custom_ar_ops(self, *a, **k):
        raise NotImplementedError("custom AR ops (ar_funcs.hpp) are not provided; the table lookup (init_ar_params) is")

    create_ar_ptrs = init_custom_ar_ops

    # ---- C-ABI handle for the fused y path
    @property
    def handle(self):
        return self._h


class _Rans64(_Coder):
    _kind = N.KIND_RANS64

    def __init__(self, freq_precision=16, bypass_coding=True, bypass_precision=4, lanes=1, device=0):
        super().__init__(freq_precision, 0, bypass_coding, bypass_precision, lanes, device)

    def init_cdf_params(self, cdfs, cdfs_sizes, offsets):
        cdfs = np.ascontiguousarray(np.asarray(cdfs), dtype=np.int32)
        cdfs_sizes = np.ascontiguousarray(np.asarray(cdfs_sizes), dtype=np.int32).reshape(-1)
        offsets = np.ascontiguousarray(np.asarray(offsets), dtype=np.int32).reshape(-1)
        if cdfs.ndim != 2 or cdfs.shape[0] != cdfs_sizes.size:
            raise ValueError("cdfs should be 2-dimensional with shape (cdfs_sizes.size(), >cdfs_sizes.max())")
        N.check(N.lib().basic_coder_init_cdf_params(self._h, cdfs.ctypes.data, cdfs.shape[0], cdfs.shape[1],
                                                    cdfs_sizes.ctypes.data, offsets.ctypes.data))

    def get_cdfs(self):
        T, M = C.c_int(0), C.c_int(0)
        N.check(N.lib().basic_coder_cdfs_shape(self._h, C.byref(T), C.byref(M)))
        if T.value == 0:
            return np.zeros((0,), dtype=np.int32)
        out = np.zeros((T.value, M.value), dtype=np.int32)
        N.check(N.lib().basic_coder_get_cdfs(self._h, out.ctypes.data))
        return out


class _EncoderMixin:
    def encode_with_indexes(self, symbols, indexes, ar_indexes=None, ar_offsets=None, cache=0):
        sym, idx = _Arg(symbols, self.device), _Arg(indexes, self.device)
        if sym.size != idx.size:
            raise ValueError("symbols and indexes must have the same number of elements")
        n = sym.size
        out_len = C.c_int64(0)
        if self._ar_order:   # rans64.cpp:218-228, :259-263
            if cache:
                raise NotImplementedError("cache=True with AR tables")
            ai, off = self._ar_args(ar_indexes, ar_offsets, n)
            N.check(N.lib().basic_coder_encode_ar(self._h, sym.ptr, idx.ptr, n, ai.ptr if ai else None, off.ptr, self._ar_order,
                                                  None, 0, C.byref(out_len), _stream(sym, idx)))
            return N.last_output(self._h)
        _no_ar(ar_indexes, ar_offsets)
        N.check(N.lib().basic_coder_encode(self._h, sym.ptr, idx.ptr, n, self.lanes, int(bool(cache)), None, 0,
                                           C.byref(out_len), _stream(sym, idx)))
        self._cached = getattr(self, "_cached", 0) + (n if cache else 0)
        return N.last_output(self._h) if not cache else b""

    def flush(self):
        out_len = C.c_int64(0)
        N.check(N.lib().basic_coder_flush(self._h, self.lanes, None, 0, C.byref(out_len), 0))
        self._cached = 0
        return N.last_output(self._h)


class _DecoderMixin:
    def _out_like(self, idx, indexes):
        if idx.cuda:
            out = torch.empty(idx.shape, dtype=torch.int32, device=idx.obj.device)
            return out, out.data_ptr()
        out = np.empty(idx.shape, dtype=np.int32)
        return out, out.ctypes.data

    def decode_with_indexes(self, encoded, indexes, ar_indexes=None, ar_offsets=None):
        idx = _Arg(indexes, self.device)
        enc = np.frombuffer(bytes(encoded), dtype=np.uint8)
        out, optr = self._out_like(idx, indexes)
        if self._ar_order:   # rans64.cpp:406-416, :439-443
            ai, off = self._ar_args(ar_indexes, ar_offsets, idx.size)
            N.check(N.lib().basic_coder_decode_ar(self._h, enc.ctypes.data if enc.size else 0, enc.size, idx.ptr, idx.size,
                                                  ai.ptr if ai else None, off.ptr, self._ar_order, optr, _stream(idx)))
            return out
        _no_ar(ar_indexes, ar_offsets)
        N.check(N.lib().basic_coder_decode(self._h, enc.ctypes.data if enc.size else 0, enc.size, idx.ptr, idx.size,
                                           self.lanes, optr, _stream(idx)))
        return out

    def set_stream(self, stream):
        enc = np.frombuffer(bytes(stream), dtype=np.uint8)
        N.check(N.lib().basic_coder_set_stream(self._h, enc.ctypes.data if enc.size else 0, enc.size, self.lanes, 0))

    def decode_stream(self, indexes, ar_indexes=None, ar_offsets=None):
        _no_ar(ar_indexes, ar_offsets)
        idx = _Arg(indexes, self.device)
        out, optr = self._out_like(idx, indexes)
        N.check(N.lib().basic_coder_decode_stream(self._h, idx.ptr, idx.size, optr, _stream(idx)))
        return out


class Rans64Encoder(_EncoderMixin, _Rans64):
    _role = N.ROLE_BOTH

    def encode_batch(self, symbols, indexes):
        """[B, n] symbols / indexes -> B reference (lanes=1) streams, coded by one CTA each in a single launch -- what
        B calls of encode_with_indexes return (the z node codes one stream per image)."""
        sym, idx = _Arg(symbols, self.device), _Arg(indexes, self.device)
        if sym.shape != idx.shape or len(sym.shape) != 2:
            raise ValueError("symbols and indexes must be [n_streams, n] arrays of the same shape")
        B, n = sym.shape
        lens = (C.c_int64 * max(B, 1))()
        N.check(N.lib().basic_coder_encode_batch(self._h, sym.ptr, idx.ptr, n, B, None, 0, lens, _stream(sym, idx)))
        blob = N.last_output(self._h)
        out, at = [], 0
        for b in range(B):
            out.append(blob[at:at + lens[b]])
            at += lens[b]
        return out

    def peek_cache(self):
        raise NotImplementedError("peek_cache exposes the reference's internal symbol list; not provided")


class Rans64Decoder(_DecoderMixin, _Rans64):
    _role = N.ROLE_BOTH

    def decode_batch(self, streams, indexes):
        """B reference (lanes=1) streams + [B, n] indexes -> [B, n] symbols in a single launch (one CTA per stream)."""
        idx = _Arg(indexes, self.device)
        if len(idx.shape) != 2 or idx.shape[0] != len(streams):
            raise ValueError("indexes must be [n_streams, n]")
        B, n = idx.shape
        lens = (C.c_int64 * max(B, 1))(*[len(s) for s in streams])
        enc = np.frombuffer(b"".join(bytes(s) for s in streams), dtype=np.uint8)
        out, optr = self._out_like(idx, indexes)
        N.check(N.lib().basic_coder_decode_batch(self._h, enc.ctypes.data if enc.size else 0, lens, B, idx.ptr, n, optr,
                                                 _stream(idx)))
        return out


class _Tans(_Coder):
    _kind = N.KIND_TANS

    def __init__(self, table_log=11, max_symbol_value=255, bypass_coding=False, bypass_precision=4, lanes=1, device=0):
        if int(lanes) != 1:
            raise ValueError("tANS is available in the reference-compatible single-stream mode only (lanes=1)")
        super().__init__(table_log, max_symbol_value, bypass_coding, bypass_precision, 1, device)


class TansEncoder(_EncoderMixin, _Tans):
    _role = N.ROLE_ENCODER


class TansDecoder(_DecoderMixin, _Tans):
    _role = N.ROLE_DECODER
