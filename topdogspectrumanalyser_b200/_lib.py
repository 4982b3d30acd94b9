"""ctypes binding of libtdsa.so (C ABI in include/tdsa.h).

There is no CPU fallback: if the shared library is missing this module raises,
and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TDSA_LIB", os.path.join(_HERE, "libtdsa.so"))   # TDSA_LIB: A/B builds of the same ABI

# constants mirrored from include/tdsa.h
WINDOW_HANN, WINDOW_HAMMING, WINDOW_RECT, WINDOW_BLACKMAN, WINDOW_CUSTOM = 0, 1, 2, 3, 4
NORM_NONE, NORM_RMS_F32 = 0, 1
MODE_POWER, MODE_PSD, MODE_MAG20 = 0, 1, 2
PREC_F64, PREC_F32 = 0, 1
AVG_OFF, AVG_EXP, AVG_LIN = 0, 1, 2

WINDOW_IDS = {"hanning": WINDOW_HANN, "hann": WINDOW_HANN, "hamming": WINDOW_HAMMING,
              "rectangle": WINDOW_RECT, "rect": WINDOW_RECT, "blackman": WINDOW_BLACKMAN}
MODE_IDS = {"power": MODE_POWER, "psd": MODE_PSD, "mag20": MODE_MAG20}
PREC_IDS = {"f64": PREC_F64, "float64": PREC_F64, "f32": PREC_F32, "float32": PREC_F32}
AVG_IDS = {"off": AVG_OFF, "exp": AVG_EXP, "lin": AVG_LIN}

_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double
_pi32 = C.POINTER(C.c_int32)

# name -> (restype, argtypes); every symbol include/tdsa.h declares
SIGNATURES = {
    "tdsa_version": (_i32, []),
    "tdsa_last_error": (C.c_char_p, []),
    "tdsa_launch_count": (_i64, []),
    "tdsa_create": (_i32, [_i32, _i32, _i32, _i32, _f64, _f64, _i32, C.POINTER(_vp)]),
    "tdsa_destroy": (_i32, [_vp]),
    "tdsa_set_stream": (_i32, [_vp, _vp]),
    "tdsa_set_window": (_i32, [_vp, _i32, _i32]),
    "tdsa_set_window_table_host": (_i32, [_vp, _vp]),
    "tdsa_get_window_table_host": (_i32, [_vp, _vp]),
    "tdsa_set_mode": (_i32, [_vp, _i32, _f64, _f64]),
    "tdsa_set_precision": (_i32, [_vp, _i32]),
    "tdsa_psd_db_batch": (_i32, [_vp, _vp, _i64, _i64, _vp]),
    "tdsa_power_linear_batch": (_i32, [_vp, _vp, _i64, _i64, _vp]),
    "tdsa_psd_db_batch_dc": (_i32, [_vp, _vp, _i64, _i64, _f64, _vp, _vp, _vp]),
    "tdsa_psd_db_avg_hold": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _vp, _pi32, _vp, _vp, _pi32, _i32, _vp]),
    "tdsa_psd_db_avg_hold_dc": (_i32, [_vp, _vp, _i64, _i64, _f64, _vp, _vp, _i32, _i32, _vp, _pi32, _vp, _vp, _pi32,
                                       _i32, _vp]),
    "tdsa_psd_db_avg_hold_dev": (_i32, [_vp, _vp, _i64, _i64, _i32, _f64, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32,
                                        _vp]),
    "tdsa_trace_update_dev": (_i32, [_vp, _i64, _i64, _f64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "tdsa_group_avg_db": (_i32, [_vp, _vp, _i64, _i64, _vp]),
    "tdsa_group_avg_db_peers": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _i32]),
    "tdsa_welch": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "tdsa_trace_update": (_i32, [_vp, _i64, _i64, _f64, _i32, _i32, _vp, _pi32, _vp, _vp, _pi32, _vp, _vp, _vp]),
    "tdsa_trace_update_tare": (_i32, [_vp, _i64, _i64, _f64, _i32, _i32, _vp, _pi32, _vp, _vp, _pi32, _vp, _vp, _vp,
                                      _pi32, _pi32, _i32, _vp, _vp]),
    "tdsa_colormap_rgba": (_i32, [_vp, _i64, C.c_float, C.c_float, _vp, _vp, _vp]),
    "tdsa_density_update": (_i32, [_vp, _i64, _f64, _vp, _vp]),
    "tdsa_band_power": (_i32, [_vp, _vp, _i64, _f64, _f64, _vp, _vp]),
    "tdsa_top_peaks": (_i32, [_vp, _i64, _i32, _i32, C.c_float, _vp, _vp, _vp, _vp]),
    "tdsa_parse_sweep_csv_host": (_i32, [C.c_char_p, _i64, _i64, _i64, _vp, _vp, _vp, _vp, C.POINTER(C.c_int64),
                                         C.POINTER(C.c_int64)]),
    "tdsa_parse_sweep_binary_host": (_i32, [C.c_char_p, _i64, _i64, _i64, _vp, _vp, _vp, _vp, C.POINTER(C.c_int64),
                                            C.POINTER(C.c_int64)]),
    "tdsa_stitch": (_i32, [_vp, _vp, _f64, _i64, _i64, _f64, _f64, _i64, _vp, _vp, _vp]),
    "tdsa_stitch_range": (_i32, [_vp, _vp, _f64, _i64, _i64, _f64, _f64, _i64, _i64, _i64, _vp, _vp, _vp]),
    "tdsa_ring_push": (_i32, [_vp, _i64, _vp, _i64, _i64, C.POINTER(C.c_int64), _vp]),
    "tdsa_ring_push_dev": (_i32, [_vp, _i64, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _vp, _vp]),
    "tdsa_ring_image_rgba": (_i32, [_vp, _i64, _i64, _vp, C.c_float, C.c_float, _vp, _vp, _vp]),
    "tdsa_find_peaks_snap": (_i32, [_vp, _i64, C.c_float, C.c_float, _i32, _vp, _vp]),
    "tdsa_h2d_async": (_i32, [_vp, _vp, C.c_size_t, _vp, _vp]),
    "tdsa_psd_db_batch_host": (_i32, [_vp, _vp, _i64, _i64, _vp, _i64]),
    "tdsa_plan_info": (_i32, [_vp, _pi32, _pi32, _pi32, _pi32, _pi32]),
}

_lib = None


class TdsaError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libtdsa.so (built by ``__graft_entry__.build()`` / ``csrc/build.sh``). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TdsaError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or topdogspectrumanalyser_b200/csrc/build.sh). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().tdsa_last_error()
        raise TdsaError(f"libtdsa error {rc}: {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().tdsa_launch_count())
