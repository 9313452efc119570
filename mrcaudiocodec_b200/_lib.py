"""ctypes binding of libmrc.so (include/mrc.h).  There is no fallback: if the shared library is missing or no
CUDA device is usable, importing callers get an exception."""
import ctypes as C
import os

_here = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_here, "libmrc.so")

MRC_OK, MRC_E_INVALID, MRC_E_CUDA, MRC_E_NOSPACE, MRC_E_FORMAT, MRC_E_STATE = 0, -1, -2, -3, -4, -5
PRECISION_FP64, PRECISION_FP32 = 0, 1
FLAG_SPREAD_SEQUENTIAL = 1
FLAG_NO_CHAIN_TABLES = 2
FLAG_BLOCK_SWITCHING = 4
NO_TABLE = 15

c_i16p = C.POINTER(C.c_int16)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)
c_u16p = C.POINTER(C.c_uint16)
c_f64p = C.POINTER(C.c_double)


class MrcConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("sample_rate", C.c_int32), ("n_mdct_lines", C.c_int32),
                ("n_scale_bits", C.c_int32), ("n_mant_size_bits", C.c_int32), ("joint", C.c_int32),
                ("precision", C.c_int32), ("flags", C.c_int32), ("target_bits_per_sample", C.c_double),
                ("reserved1", C.c_int64 * 4)]


class MrcTables(C.Structure):
    _fields_ = [("n_bands", C.c_int32), ("n_huff_tables", C.c_int32), ("band_nlines", c_i32p),
                ("kbd_window", c_f64p), ("hann_window", c_f64p), ("bark", c_f64p), ("quiet_intensity", c_f64p),
                ("huff_escape", c_i32p), ("huff_len", c_u8p), ("huff_code", c_u16p)]


class MrcBlockTables(C.Structure):
    _fields_ = [("a", C.c_int32), ("b", C.c_int32), ("n_bands", C.c_int32), ("band_nlines", c_i32p),
                ("window", c_f64p), ("hann_window", c_f64p), ("bark", c_f64p), ("quiet_intensity", c_f64p)]


class MrcError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "libmrc error %d: %s" % (code, msg))
        self.code = code


EXPORTS = ["mrc_version", "mrc_last_error", "mrc_create", "mrc_destroy", "mrc_set_tables", "mrc_host_alloc",
           "mrc_host_free", "mrc_encode_batch", "mrc_encode_batch_device", "mrc_decode_batch",
           "mrc_decode_batch_device", "mrc_encode_block", "mrc_decode_block", "mrc_stage_analysis",
           "mrc_stage_alloc_quant", "mrc_mantissa_histogram", "mrc_last_timing", "mrc_measure_peaks",
           "mrc_set_switch_tables", "mrc_encode_block_ab", "mrc_decode_block_ab", "mrc_detect_transients",
           "mrc_encode_shard", "mrc_encode_shard_device"]

# int32_t (*mrc_reservoir_exchange)(void* user, int32_t have_result, int32_t* reservoir)
RESERVOIR_EXCHANGE = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32))

_lib = None


def load():
    """Load libmrc.so and declare every prototype of include/mrc.h.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libmrc.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "-- there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.mrc_version.restype = C.c_int32
    lib.mrc_last_error.restype = C.c_char_p
    lib.mrc_last_error.argtypes = [vp]
    lib.mrc_create.argtypes = [C.POINTER(MrcConfig), C.POINTER(vp)]
    lib.mrc_destroy.argtypes = [vp]
    lib.mrc_set_tables.argtypes = [vp, C.POINTER(MrcTables)]
    lib.mrc_host_alloc.argtypes = [C.POINTER(vp), C.c_int64]
    lib.mrc_host_free.argtypes = [vp]
    lib.mrc_encode_batch.argtypes = [vp, vp, vp, C.c_int32, vp, C.c_int64, vp]
    lib.mrc_encode_batch_device.argtypes = [vp, vp, vp, C.c_int32, vp, C.c_int64, vp]
    lib.mrc_decode_batch.argtypes = [vp, vp, vp, C.c_int32, vp, C.c_int64, vp]
    lib.mrc_decode_batch_device.argtypes = [vp, vp, vp, vp, C.c_int32, vp, C.c_int64, vp]
    lib.mrc_encode_block.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.mrc_decode_block.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp, vp]
    lib.mrc_stage_analysis.argtypes = [vp, vp, vp, C.c_int32, vp, vp, vp, vp, vp]
    lib.mrc_stage_alloc_quant.argtypes = [vp, vp, vp, C.c_int32, vp, vp, vp, vp, vp, vp]
    lib.mrc_set_switch_tables.argtypes = [vp, C.POINTER(MrcBlockTables), vp, C.c_int32, C.c_double, C.c_double]
    lib.mrc_encode_block_ab.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.mrc_decode_block_ab.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp]
    lib.mrc_detect_transients.argtypes = [vp, vp, vp, C.c_int32, vp, vp, C.c_int32, vp]
    lib.mrc_encode_shard.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                     vp, C.c_int64, vp, RESERVOIR_EXCHANGE, vp]
    lib.mrc_encode_shard_device.argtypes = lib.mrc_encode_shard.argtypes
    lib.mrc_last_timing.argtypes = [vp, vp, vp]
    lib.mrc_measure_peaks.argtypes = [vp, vp]
    for name in EXPORTS:
        f = getattr(lib, name)
        if name not in ("mrc_last_error",):
            f.restype = C.c_int32
    _lib = lib
    return lib
