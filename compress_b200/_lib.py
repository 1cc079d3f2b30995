"""ctypes binding of libb200comp.so (the C ABI in include/b2c.h).

There is deliberately NO fallback: if the CUDA library has not been built, or no CUDA device is
present, importing / using the package raises.  The CPU oracle under oracle/ is test
infrastructure and is never imported from here.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2C_LIB") or os.path.join(_HERE, "_lib", "libb200comp.so")   # B2C_LIB: tuning builds


class B2CError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  compress_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    lib.b2c_device_count.restype = c.c_int
    lib.b2c_ctx_create.restype = c.c_void_p
    lib.b2c_ctx_create.argtypes = [c.c_int, c.c_size_t]
    lib.b2c_ctx_destroy.argtypes = [c.c_void_p]
    lib.b2c_strerror.restype = c.c_char_p
    lib.b2c_strerror.argtypes = [c.c_int]
    lib.b2c_last_cuda_error.restype = c.c_char_p
    lib.b2c_last_cuda_error.argtypes = [c.c_void_p]
    lib.b2c_sm_count.restype = c.c_int
    lib.b2c_sm_count.argtypes = [c.c_void_p]
    lib.b2c_launch_count.restype = c.c_uint64
    lib.b2c_launch_count.argtypes = [c.c_void_p]
    lib.b2c_zstd_bound.restype = c.c_size_t
    lib.b2c_zstd_bound.argtypes = [c.c_size_t, c.c_int]
    lib.b2c_zstd_encode_device.restype = c.c_int
    lib.b2c_zstd_encode_device.argtypes = [
        c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_size_t, c.c_void_p, c.c_uint32, c.c_void_p, c.c_size_t,
        c.c_void_p, c.c_uint32, c.c_void_p]
    lib.b2c_zstd_encode_device_debug.restype = c.c_int
    lib.b2c_zstd_encode_device_debug.argtypes = [
        c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_size_t, c.c_void_p, c.c_uint32, c.c_void_p, c.c_size_t, c.c_void_p,
        c.c_uint32, c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint32, c.c_void_p]
    lib.b2c_zstd_encode_chunks.restype = c.c_int
    lib.b2c_zstd_encode_chunks.argtypes = [
        c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_size_t]
    lib.b2c_s2_decode_staged_count.restype = c.c_int
    lib.b2c_s2_decode_staged_count.argtypes = [c.c_void_p, c.c_uint32, c.c_void_p]
    lib.b2c_s2_stream_bound.restype = c.c_size_t
    lib.b2c_s2_stream_bound.argtypes = [c.c_size_t, c.c_size_t]
    lib.b2c_s2_encode_stream_device.restype = c.c_int
    lib.b2c_s2_encode_stream_device.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_uint64, c.c_uint32, c.c_void_p,
                                                c.c_uint64, c.c_void_p, c.c_void_p, c.c_void_p]
    lib.b2c_s2_encode_stream.restype = c.c_int
    lib.b2c_s2_encode_stream.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_size_t, c.c_uint32, c.c_void_p, c.c_size_t,
                                         c.c_void_p]
    lib.b2c_s2_decode_stream.restype = c.c_int
    lib.b2c_s2_decode_stream.argtypes = [c.c_void_p, c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.c_void_p]
    lib.b2c_zstd_frame_bound.restype = c.c_size_t
    lib.b2c_zstd_frame_bound.argtypes = [c.c_size_t, c.c_int]
    lib.b2c_zstd_encode_frames_device.restype = c.c_int
    lib.b2c_zstd_encode_frames_device.argtypes = [
        c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint32, c.c_void_p, c.c_uint64, c.c_void_p,
        c.c_void_p, c.c_void_p]
    lib.b2c_zstd_encode_frames.restype = c.c_int
    lib.b2c_zstd_encode_frames.argtypes = [
        c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_size_t]
    lib.b2c_zstd_encode_packed.restype = c.c_int
    lib.b2c_zstd_encode_packed.argtypes = [
        c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_size_t, c.c_uint32, c.c_void_p, c.c_size_t, c.c_void_p,
        c.c_void_p, c.c_void_p]
    lib.b2c_zstd_encode_device_timed.restype = c.c_int
    lib.b2c_zstd_encode_device_timed.argtypes = [
        c.c_void_p, c.c_int, c.c_void_p, c.c_size_t, c.c_uint32, c.c_void_p, c.c_size_t, c.c_void_p, c.c_uint32,
        c.c_void_p, c.c_void_p]
    lib.b2c_profile_enable.restype = c.c_int
    lib.b2c_profile_enable.argtypes = [c.c_void_p, c.c_int]
    lib.b2c_profile_read.restype = c.c_int
    lib.b2c_profile_read.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p]
    lib.b2c_decode_profile_enable.restype = c.c_int
    lib.b2c_decode_profile_enable.argtypes = [c.c_void_p, c.c_int]
    lib.b2c_decode_profile_read.restype = c.c_int
    lib.b2c_decode_profile_read.argtypes = [c.c_void_p, c.c_void_p]
    lib.b2c_decode_staged_count.restype = c.c_int
    lib.b2c_decode_staged_count.argtypes = [c.c_void_p, c.c_uint32, c.c_void_p]
    lib.b2c_zstd_decode_device.restype = c.c_int
    lib.b2c_zstd_decode_device.argtypes = [
        c.c_void_p, c.c_void_p, c.c_size_t, c.c_void_p, c.c_void_p, c.c_void_p, c.c_size_t, c.c_void_p, c.c_uint32,
        c.c_void_p, c.c_uint32, c.c_void_p]
    lib.b2c_zstd_decode_chunks.restype = c.c_int
    lib.b2c_zstd_decode_chunks.argtypes = [
        c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_size_t]
    lib.b2c_s2_bound.restype = c.c_size_t
    lib.b2c_s2_bound.argtypes = [c.c_size_t]
    lib.b2c_s2_encode_device.restype = c.c_int
    lib.b2c_s2_encode_device.argtypes = [
        c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_size_t, c.c_void_p, c.c_uint32, c.c_void_p, c.c_size_t,
        c.c_void_p, c.c_uint32, c.c_void_p]
    lib.b2c_s2_decode_device.restype = c.c_int
    lib.b2c_s2_decode_device.argtypes = [
        c.c_void_p, c.c_void_p, c.c_size_t, c.c_void_p, c.c_void_p, c.c_void_p, c.c_size_t, c.c_void_p, c.c_uint32,
        c.c_void_p, c.c_uint32, c.c_void_p]
    lib.b2c_s2_encode_chunks.restype = c.c_int
    lib.b2c_s2_encode_chunks.argtypes = [
        c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_size_t]
    lib.b2c_s2_decode_chunks.restype = c.c_int
    lib.b2c_s2_decode_chunks.argtypes = [
        c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_size_t]
    lib.b2c_huf_compress_device.restype = c.c_int
    lib.b2c_huf_compress_device.argtypes = [
        c.c_void_p, c.c_int, c.c_void_p, c.c_size_t, c.c_void_p, c.c_uint32, c.c_void_p, c.c_size_t, c.c_void_p,
        c.c_uint32, c.c_void_p]
    lib.b2c_huf_decompress_device.restype = c.c_int
    lib.b2c_huf_decompress_device.argtypes = [
        c.c_void_p, c.c_int, c.c_void_p, c.c_size_t, c.c_void_p, c.c_void_p, c.c_size_t, c.c_void_p, c.c_void_p,
        c.c_uint32, c.c_void_p]
    for nm in ("b2c_huf_compress_chunks", "b2c_huf_decompress_chunks"):
        getattr(lib, nm).restype = c.c_int
        getattr(lib, nm).argtypes = [c.c_void_p, c.c_int, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_size_t]
    lib.b2c_huf_read_table.restype = c.c_int
    lib.b2c_huf_read_table.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_size_t]
    lib.b2c_queue_create.restype = c.c_void_p
    lib.b2c_queue_create.argtypes = [c.c_int, c.c_size_t, c.c_uint]
    lib.b2c_queue_destroy.argtypes = [c.c_void_p]
    for nm in ("b2c_queue_zstd_encode", "b2c_queue_s2_encode"):
        getattr(lib, nm).restype = c.c_int64
        getattr(lib, nm).argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t]
    for nm in ("b2c_queue_zstd_decode", "b2c_queue_s2_decode"):
        getattr(lib, nm).restype = c.c_int64
        getattr(lib, nm).argtypes = [c.c_void_p, c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t]
    lib.b2c_queue_stats.restype = c.c_int
    lib.b2c_queue_stats.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p]
    return lib


lib = _load()

# every symbol include/b2c.h declares; tests assert the built library exports all of them
EXPORTED_SYMBOLS = [
    "b2c_device_count", "b2c_ctx_create", "b2c_ctx_destroy", "b2c_strerror", "b2c_last_cuda_error",
    "b2c_sm_count", "b2c_launch_count", "b2c_zstd_bound", "b2c_zstd_encode_device",
    "b2c_zstd_encode_chunks", "b2c_zstd_encode_device_debug", "b2c_zstd_encode_packed", "b2c_zstd_encode_device_timed",
    "b2c_zstd_decode_device", "b2c_zstd_decode_chunks", "b2c_profile_enable", "b2c_profile_read", "b2c_decode_profile_enable", "b2c_decode_profile_read", "b2c_decode_staged_count", "b2c_s2_decode_staged_count",
    "b2c_huf_compress_device", "b2c_huf_decompress_device",
    "b2c_s2_bound", "b2c_s2_encode_device", "b2c_s2_decode_device", "b2c_s2_encode_chunks", "b2c_s2_decode_chunks",
    "b2c_huf_compress_chunks", "b2c_huf_decompress_chunks", "b2c_huf_read_table",
    "b2c_queue_create", "b2c_queue_destroy", "b2c_queue_zstd_encode", "b2c_queue_zstd_decode", "b2c_queue_s2_encode",
    "b2c_queue_s2_decode", "b2c_queue_stats",
    "b2c_zstd_frame_bound", "b2c_zstd_encode_frames_device", "b2c_zstd_encode_frames",
    "b2c_s2_stream_bound", "b2c_s2_encode_stream_device", "b2c_s2_encode_stream", "b2c_s2_decode_stream",
]


def check(rc, ctx=None):
    if rc != 0:
        msg = lib.b2c_strerror(rc).decode()
        if ctx:
            msg += ": " + lib.b2c_last_cuda_error(ctx).decode()
        raise B2CError(f"libb200comp error {rc}: {msg}")
