"""ctypes binding of the C ABI (include/basic_b200.h).  Loading fails loudly: there is no CPU fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libbasic_b200.so")
_lib = None

OK, ERR_VALUE, ERR_CUDA, ERR_CAPACITY, ERR_STREAM = 0, -1, -2, -3, -4
KIND_RANS64, KIND_TANS = 0, 1
ROLE_BOTH, ROLE_ENCODER, ROLE_DECODER = 0, 1, 2
LANES_REFERENCE, LANES_AUTO = 1, 0
CTX_FP32, CTX_TF32X3, CTX_FP16X3 = 0, 1, 2

# every extern "C" symbol declared in include/basic_b200.h
SYMBOLS = [
    "basic_last_error", "basic_device_count", "basic_coder_create", "basic_coder_destroy", "basic_coder_init_params",
    "basic_coder_init_cdf_params", "basic_coder_cdfs_shape", "basic_coder_get_cdfs", "basic_pmf_to_quantized_cdf",
    "basic_coder_encode_bound", "basic_coder_encode", "basic_coder_flush", "basic_coder_last_output", "basic_coder_output_size", "basic_coder_take_output", "basic_coder_decode", "basic_coder_set_stream",
    "basic_coder_decode_stream", "basic_coder_init_ar_params", "basic_coder_encode_ar", "basic_coder_decode_ar", "basic_coder_encode_batch", "basic_coder_decode_batch", "basic_coder_set_scale_table", "basic_gauss_quantize_index", "basic_gauss_dequantize",
    "basic_ctx_create", "basic_ctx_destroy", "basic_ctx_set_weights", "basic_ctx_set_weights_internal", "basic_ctx_set_map", "basic_ctx_num_stages",
    "basic_ctx_set_precision", "basic_ctx_stage_positions", "basic_ctx_stage_params", "basic_ypath_encode_bound", "basic_ypath_encode",
    "basic_ypath_decode", "basic_profile_enable", "basic_profile_read", "basic_debug_mma_bench", "basic_launch_count",
]


class NativeLibraryError(RuntimeError):
    pass


def lib():
    """Loads (building first if the sources are newer and nvcc is available) the CUDA library."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if _build.stale():
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            if not os.path.exists(LIB_PATH):
                raise NativeLibraryError(
                    f"cbench_basic_b200: CUDA library missing and could not be built ({e}). "
                    "There is no CPU fallback; run `python cbench_basic_b200/build.py`.") from e
            # a library that does not match the sources is only used on request: results would not be those of HEAD
            if os.environ.get("BASIC_ALLOW_STALE_LIB") != "1":
                raise NativeLibraryError(
                    f"cbench_basic_b200: {LIB_PATH} was built from other sources and could not be rebuilt ({e}). "
                    "Run `python cbench_basic_b200/build.py`, or set BASIC_ALLOW_STALE_LIB=1 to load it anyway.") from e
            import warnings
            warnings.warn(f"cbench_basic_b200: loading a stale library ({e})")
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as e:
        raise NativeLibraryError(f"cbench_basic_b200: cannot load {LIB_PATH}: {e}") from e
    vp, i32p, f32p, u8p, i64 = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64
    L.basic_last_error.restype = C.c_char_p
    L.basic_coder_create.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_uint, C.c_int, C.POINTER(vp)]
    L.basic_coder_destroy.argtypes = [vp]
    L.basic_coder_destroy.restype = None
    L.basic_coder_init_params.argtypes = [vp, i32p, C.c_int, C.c_int, i32p, i32p]
    L.basic_coder_init_cdf_params.argtypes = [vp, i32p, C.c_int, C.c_int, i32p, i32p]
    L.basic_coder_cdfs_shape.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.basic_coder_get_cdfs.argtypes = [vp, i32p]
    L.basic_pmf_to_quantized_cdf.argtypes = [f32p, C.c_int, C.c_int, C.c_int, i32p]
    L.basic_coder_encode_bound.argtypes = [vp, i64, C.c_int]
    L.basic_coder_encode_bound.restype = i64
    L.basic_coder_encode.argtypes = [vp, i32p, i32p, i64, C.c_int, C.c_int, u8p, i64, C.POINTER(i64), vp]
    L.basic_coder_flush.argtypes = [vp, C.c_int, u8p, i64, C.POINTER(i64), vp]
    L.basic_coder_last_output.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
    L.basic_coder_output_size.argtypes = [vp]
    L.basic_coder_output_size.restype = i64
    L.basic_coder_take_output.argtypes = [vp, vp, i64]
    L.basic_coder_decode.argtypes = [vp, u8p, i64, i32p, i64, C.c_int, i32p, vp]
    L.basic_coder_set_stream.argtypes = [vp, u8p, i64, C.c_int, vp]
    L.basic_coder_decode_stream.argtypes = [vp, i32p, i64, i32p, vp]
    L.basic_coder_init_ar_params.argtypes = [vp, i32p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.basic_coder_encode_ar.argtypes = [vp, i32p, i32p, i64, i32p, i32p, C.c_int, u8p, i64, C.POINTER(i64), vp]
    L.basic_coder_decode_ar.argtypes = [vp, u8p, i64, i32p, i64, i32p, i32p, C.c_int, i32p, vp]
    L.basic_coder_encode_batch.argtypes = [vp, i32p, i32p, i64, C.c_int, u8p, i64, C.POINTER(i64), vp]
    L.basic_coder_decode_batch.argtypes = [vp, u8p, C.POINTER(i64), C.c_int, i32p, i64, i32p, vp]
    L.basic_coder_set_scale_table.argtypes = [vp, f32p, C.c_int]
    L.basic_gauss_quantize_index.argtypes = [vp, f32p, f32p, i32p, i64, C.c_int, C.c_int, C.c_int, i32p, i32p, f32p, vp]
    L.basic_gauss_dequantize.argtypes = [vp, i32p, f32p, i32p, i64, C.c_int, C.c_int, C.c_int, f32p, vp]
    L.basic_ctx_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.basic_ctx_destroy.argtypes = [vp]
    L.basic_ctx_destroy.restype = None
    L.basic_ctx_set_weights.argtypes = [vp] + [f32p] * 8
    L.basic_ctx_set_weights_internal.argtypes = [vp] + [f32p] * 12 + [C.c_int]
    L.basic_ctx_set_map.argtypes = [vp, i32p, C.c_int, C.c_int]
    L.basic_ctx_num_stages.argtypes = [vp]
    L.basic_ctx_set_precision.argtypes = [vp, C.c_int, C.c_int]
    L.basic_ctx_stage_positions.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(i64)]
    L.basic_ctx_stage_params.argtypes = [vp, C.c_int, f32p, f32p, C.c_int, f32p, vp]
    L.basic_ypath_encode_bound.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.basic_ypath_encode_bound.restype = i64
    L.basic_ypath_encode.argtypes = [vp, vp, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, i64,
                                     C.POINTER(i64), f32p, vp]
    L.basic_ypath_decode.argtypes = [vp, vp, u8p, i64, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, vp]
    L.basic_profile_enable.argtypes = [C.c_int]
    L.basic_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(i64)]
    L.basic_debug_mma_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    L.basic_launch_count.argtypes = [C.c_int]
    L.basic_launch_count.restype = i64
    _lib = L
    return L


class CudaError(RuntimeError):
    pass


class StreamError(ValueError):
    pass


def check(rc):
    """Maps C-ABI status codes to exceptions: ValueError exactly where the reference raises py::value_error."""
    if rc == OK:
        return
    msg = lib().basic_last_error().decode("utf-8", "replace")
    if rc == ERR_VALUE:
        raise ValueError(msg)
    if rc == ERR_STREAM:
        raise StreamError(msg)
    if rc == ERR_CAPACITY:
        raise BufferError(msg)
    raise CudaError(msg or f"basic_b200 error {rc}")


def require_gpu():
    if lib().basic_device_count() < 1:
        raise CudaError("cbench_basic_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


_new_bytes = C.pythonapi.PyBytes_FromStringAndSize   # (NULL, n): an uninitialised bytes object to be filled in place
_new_bytes.restype = C.py_object
_new_bytes.argtypes = [C.c_void_p, C.c_ssize_t]
_bytes_ptr = C.pythonapi.PyBytes_AsString
_bytes_ptr.restype = C.c_void_p
_bytes_ptr.argtypes = [C.py_object]


def last_output(handle, prefix=b""):
    """The stream of the last encoding call made with out == NULL, as a bytes object, behind `prefix` (a framing header).
    The object is allocated at its final size and the library copies into it directly (chunk by chunk while the rest is
    still in flight): no second pass over a multi-megabyte stream to prepend a few header bytes."""
    n = int(lib().basic_coder_output_size(handle))
    if n == 0:
        return bytes(prefix)
    k = len(prefix)
    out = _new_bytes(None, k + n)
    base = _bytes_ptr(out)
    if k:
        C.memmove(base, prefix, k)
    check(lib().basic_coder_take_output(handle, base + k, n))
    return out


def last_output_view(handle):
    """Zero-copy form of last_output(): a read-only memoryview over the coder's page-locked staging buffer (the device-to-host
    copy is awaited, nothing else moves).  Valid until the next encoding call on the same coder; a decoder given this view
    uploads straight out of it (no staging copy either: the library recognises page-locked memory)."""
    ptr, n = C.c_void_p(), C.c_int64()
    check(lib().basic_coder_last_output(handle, C.byref(ptr), C.byref(n)))
    if not n.value:
        return memoryview(b"")
    return memoryview((C.c_uint8 * n.value).from_address(ptr.value)).toreadonly()


PHASES = ("context_model", "quantise", "coder_encode", "coder_decode", "host_set_stream", "host_staging")


def profile(on):
    check(lib().basic_profile_enable(1 if on else 0))


def profile_read():
    """{phase: (total ms, spans)} accumulated since the last read (CUDA events on the launching stream)."""
    ms, n = (C.c_double * 8)(), (C.c_int64 * 8)()
    check(lib().basic_profile_read(ms, n))
    return {name: (ms[i], n[i]) for i, name in enumerate(PHASES)}


def launch_count(reset=False):
    return int(lib().basic_launch_count(1 if reset else 0))
