"""ctypes bindings of the C ABI in include/psxgpu.h (libpsxgpu.so).  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpsxgpu.so")


class Opts(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("modulation_crop", C.c_uint32), ("show_vram", C.c_uint32),
        ("display_format", C.c_uint32), ("clut_slots", C.c_uint32), ("encode_threads", C.c_uint32),
        ("profile", C.c_uint32), ("reserved", C.c_uint32 * 9),
    ]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "wire_commands", "packed_prims", "epochs", "waves", "cuts_raw", "cuts_war", "cuts_clut",
        "serial_self_sampling", "serial_copy", "serial_upload", "clut_ops", "clut_cache_hits", "rejected",
        "kernel_launches", "h2d_bytes", "d2h_bytes", "flushes")] + [("reserved", C.c_uint64 * 7)]


class Timing(C.Structure):
    _fields_ = [
        ("last_flush_device_ms", C.c_float), ("last_kernels_ms", C.c_float), ("last_raster_ms", C.c_float),
        ("last_raster_max_ms", C.c_float), ("last_raster_launches", C.c_uint32), ("last_waves", C.c_uint32),
        ("last_hash_ms", C.c_float), ("last_encode_ms", C.c_float), ("last_stage_ms", C.c_float),
        ("last_scan_ms", C.c_float), ("last_pinned_wait_ms", C.c_float), ("last_launch_ms", C.c_float),
        ("reserved", C.c_float * 4),
    ]


class FrontendStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "gp0_words", "gp1_words", "commands", "unknown_commands", "culled_primitives", "frames", "idle")] + [
        ("reserved", C.c_uint64 * 5)]


class StateFile(C.Structure):
    _fields_ = [("version", C.c_uint32), ("data_compression", C.c_uint32), ("data_compressed_size", C.c_uint32),
                ("data_uncompressed_size", C.c_uint32), ("offset_to_data", C.c_uint32), ("title", C.c_char * 128),
                ("serial", C.c_char * 32)]


FRAME_CALLBACK = C.CFUNCTYPE(None, C.c_void_p, C.c_long)

_lib = None


def load() -> C.CDLL:
    """Loads libpsxgpu.so.  Raises if it has not been built: the CUDA extension IS the product."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m duckstation_b200.build` "
            "(there is no CPU fallback for the rasteriser)")
    lib = C.CDLL(LIB_PATH)
    P = C.c_void_p
    u32, u64p = C.c_uint32, C.POINTER(C.c_uint64)
    lib.psxgpu_abi_version.restype = C.c_int
    lib.psxgpu_create.argtypes = [C.c_int, u32, C.POINTER(Opts), C.POINTER(P)]
    lib.psxgpu_destroy.argtypes = [P]
    lib.psxgpu_destroy.restype = None
    lib.psxgpu_last_error.argtypes = [P]
    lib.psxgpu_last_error.restype = C.c_char_p
    lib.psxgpu_submit_wire.argtypes = [P, u32, C.c_char_p, C.c_size_t]
    lib.psxgpu_submit_batch.argtypes = [P, u32, u32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.psxgpu_flush.argtypes = [P]
    lib.psxgpu_sync.argtypes = [P]
    lib.psxgpu_run_staged.argtypes = [P]
    lib.psxgpu_set_batch_chunk.argtypes = [P, u32]
    lib.psxgpu_read_vram.argtypes = [P, u32, u32, u32, u32, u32, C.c_void_p, u32]
    lib.psxgpu_shadow_vram.argtypes = [P, u32]
    lib.psxgpu_shadow_vram.restype = C.POINTER(C.c_uint16)
    lib.psxgpu_clear_vram.argtypes = [P, u32]
    lib.psxgpu_load_state.argtypes = [P, u32, C.c_void_p, C.c_void_p]
    lib.psxgpu_save_state.argtypes = [P, u32, C.c_void_p, C.c_void_p]
    lib.psxgpu_scanout.argtypes = [P, u32, C.c_void_p, C.c_void_p, u32, C.POINTER(u32), C.POINTER(u32)]
    lib.psxgpu_display.argtypes = [P, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
    lib.psxgpu_presenter_create.argtypes = [P, u32, u32, C.POINTER(C.c_void_p)]
    lib.psxgpu_presenter_destroy.argtypes = [C.c_void_p]
    lib.psxgpu_presenter_destroy.restype = None
    lib.psxgpu_present.argtypes = [C.c_void_p, u32, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(u32), C.POINTER(u32)]
    lib.psxgpu_display.restype = C.c_void_p
    lib.psxgpu_hash.argtypes = [P, u32, u32, u64p]
    lib.psxgpu_hash_device.argtypes = [P, u32, u32, C.c_void_p]
    lib.psxgpu_get_stats.argtypes = [P, C.POINTER(Stats)]
    lib.psxgpu_reset_stats.argtypes = [P]
    lib.psxgpu_get_timing.argtypes = [P, C.POINTER(Timing)]
    lib.psxgpu_get_raster_launch_ms.argtypes = [P, C.POINTER(C.c_float), u32]
    lib.psxgpu_frontend_create.argtypes = [u32, u32, C.POINTER(P)]
    lib.psxgpu_frontend_destroy.argtypes = [P]
    lib.psxgpu_frontend_destroy.restype = None
    lib.psxgpu_frontend_gp0.argtypes = [P, C.c_void_p, C.c_size_t]
    lib.psxgpu_frontend_gp1.argtypes = [P, u32]
    lib.psxgpu_frontend_vsync.argtypes = [P]
    lib.psxgpu_frontend_finish_read.argtypes = [P]
    lib.psxgpu_frontend_records.argtypes = [P, C.POINTER(C.c_size_t)]
    lib.psxgpu_frontend_records.restype = C.c_void_p
    lib.psxgpu_frontend_clear.argtypes = [P]
    lib.psxgpu_frontend_get_stats.argtypes = [P, C.POINTER(FrontendStats)]
    lib.psxgpu_frontend_feed_dump.argtypes = [P, C.c_char_p, C.c_size_t]
    lib.psxgpu_frontend_feed_dump.restype = C.c_long
    lib.psxgpu_replay_dump.argtypes = [P, u32, C.c_char_p, C.c_size_t, u32, u32, FRAME_CALLBACK, C.c_void_p]
    lib.psxgpu_replay_dump.restype = C.c_long
    lib.psxgpu_state_file_info.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(StateFile)]
    lib.psxgpu_state_file_info.restype = C.c_int
    lib.psxgpu_state_find_gpu_section.argtypes = [C.c_char_p, C.c_size_t, u32]
    lib.psxgpu_state_find_gpu_section.restype = C.c_long
    lib.psxgpu_frontend_load_state.argtypes = [P, C.c_void_p, u32, C.c_char_p, C.c_size_t, u32, C.POINTER(C.c_size_t)]
    lib.psxgpu_frontend_load_state.restype = C.c_int
    lib.psxgpu_frontend_save_state.argtypes = [P, C.c_void_p, u32, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.psxgpu_frontend_save_state.restype = C.c_int
    for name in ("psxgpu_frontend_create", "psxgpu_frontend_gp0", "psxgpu_frontend_gp1", "psxgpu_frontend_vsync",
                 "psxgpu_frontend_finish_read", "psxgpu_frontend_clear", "psxgpu_frontend_get_stats"):
        getattr(lib, name).restype = C.c_int
    for name in ("psxgpu_create", "psxgpu_submit_wire", "psxgpu_submit_batch", "psxgpu_flush", "psxgpu_sync",
                 "psxgpu_run_staged", "psxgpu_set_batch_chunk", "psxgpu_read_vram", "psxgpu_clear_vram", "psxgpu_load_state",
                 "psxgpu_save_state", "psxgpu_scanout", "psxgpu_hash", "psxgpu_hash_device", "psxgpu_get_stats",
                 "psxgpu_reset_stats", "psxgpu_get_timing", "psxgpu_get_raster_launch_ms", "psxgpu_presenter_create",
                 "psxgpu_present"):
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


EXPORTS = (
    "psxgpu_abi_version", "psxgpu_create", "psxgpu_destroy", "psxgpu_last_error", "psxgpu_submit_wire",
    "psxgpu_submit_batch", "psxgpu_flush", "psxgpu_sync", "psxgpu_run_staged", "psxgpu_set_batch_chunk", "psxgpu_read_vram",
    "psxgpu_shadow_vram", "psxgpu_clear_vram", "psxgpu_load_state", "psxgpu_save_state", "psxgpu_scanout",
    "psxgpu_display", "psxgpu_hash", "psxgpu_hash_device", "psxgpu_get_stats", "psxgpu_reset_stats",
    "psxgpu_get_timing", "psxgpu_get_raster_launch_ms",
    "psxgpu_frontend_create", "psxgpu_frontend_destroy", "psxgpu_frontend_gp0", "psxgpu_frontend_gp1",
    "psxgpu_frontend_vsync", "psxgpu_frontend_finish_read", "psxgpu_frontend_records", "psxgpu_frontend_clear",
    "psxgpu_frontend_get_stats", "psxgpu_frontend_feed_dump", "psxgpu_replay_dump",
    "psxgpu_presenter_create", "psxgpu_presenter_destroy", "psxgpu_present",
    "psxgpu_state_file_info", "psxgpu_state_find_gpu_section", "psxgpu_frontend_load_state", "psxgpu_frontend_save_state",
)
