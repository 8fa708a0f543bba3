"""ctypes binding of include/pcfe.h (detmatch_b200/lib/libpcfe.so).

This is the whole Python<->native boundary: plain device pointers, sizes, a device index and a
cudaStream_t.  torch is used only to own device memory and to name the current stream.
The library must exist -- there is deliberately no fallback path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PCFE_LIB selects another build of the same library (kernel tuning experiments only)
LIB_PATH = os.environ.get("PCFE_LIB") or os.path.join(_HERE, "lib", "libpcfe.so")

_lib = None
ERR_SHAPE = -2  # PCFE_ERR_SHAPE

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_int64 = ctypes.c_int64
c_size_t = ctypes.c_size_t
_f3 = ctypes.c_float * 3
_f6 = ctypes.c_float * 6


class Frame(ctypes.Structure):
    """pcfe_frame_t"""
    _fields_ = [("points", c_void_p), ("n", c_int64), ("voxels", c_void_p), ("coors", c_void_p),
                ("num_points", c_void_p)]


# every symbol include/pcfe.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "pcfe_version": (c_int, []),
    "pcfe_error_string": (ctypes.c_char_p, [c_int]),
    "pcfe_launch_count": (ctypes.c_uint64, []),
    "pcfe_profile_enable": (c_int, [c_int]),
    "pcfe_profile_report": (c_int, [ctypes.c_char_p, c_size_t]),
    "pcfe_debug_set": (c_int, [ctypes.c_char_p, c_int]),
    "pcfe_grid_size": (c_int, [_f3, _f6, ctypes.POINTER(ctypes.c_int32)]),
    "pcfe_dynamic_voxelize_f32": (c_int, [c_void_p, c_int64, c_int, _f3, _f6, c_void_p, c_int, c_void_p]),
    "pcfe_dynamic_voxelize_batch_f32": (c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(c_int64), c_int, c_int,
                                                _f3, _f6, ctypes.POINTER(c_void_p), c_int, c_void_p]),
    "pcfe_hard_voxelize_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, _f3, _f6, c_int, c_int]),
    "pcfe_hard_voxelize_f32": (c_int, [c_void_p, c_int64, c_int, _f3, _f6, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "pcfe_hard_voxelize_batch_f32": (c_int, [ctypes.POINTER(Frame), c_int, c_int, _f3, _f6, c_int, c_int,
                                             c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "pcfe_hard_voxelize_batch_filtered_f32": (c_int, [ctypes.POINTER(Frame), c_int, c_int, _f3, _f6, _f6, c_int, c_int,
                                                      c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "pcfe_voxel_mean_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_void_p]),
    "pcfe_hard_voxelize_mean_batch_f32": (c_int, [ctypes.POINTER(Frame), c_int, c_int, _f3, _f6, ctypes.POINTER(ctypes.c_float),
                                                  c_int, c_int, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "pcfe_hard_voxelize_packed_batch_f32": (c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(c_int64), c_int, c_int, _f3, _f6,
                                                    ctypes.POINTER(ctypes.c_float), c_int, c_int, c_int, c_void_p, c_void_p,
                                                    c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "pcfe_dynamic_scatter_workspace_bytes": (c_size_t, [ctypes.POINTER(ctypes.c_int32), c_int, c_int64]),
    "pcfe_dynamic_scatter_map_i32": (c_int, [c_void_p, c_int64, c_int, ctypes.POINTER(ctypes.c_int32), c_void_p, c_void_p,
                                             c_void_p, c_size_t, c_int, c_void_p]),
    "pcfe_dynamic_scatter_reduce_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int64,
                                                c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pcfe_dynamic_scatter_backward_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                                                  c_int, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "pcfe_points_in_boxes_workspace_bytes": (c_size_t, [c_int, c_int]),
    "pcfe_points_in_boxes_part_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p,
                                              c_size_t, c_int, c_void_p]),
    "pcfe_points_in_boxes_all_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p,
                                             c_size_t, c_int, c_void_p]),
    "pcfe_points_in_boxes_boxmajor_f32": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p,
                                                  c_size_t, c_int, c_void_p]),
    "pcfe_pcdet_points_in_boxes_gpu_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p,
                                                   c_size_t, c_int, c_void_p]),
    "pcfe_pcdet_points_in_boxes_cpu_f32": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p,
                                                   c_size_t, c_int, c_void_p]),
    "pcfe_roiaware_pool3d_forward_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_int, c_int,
                                                 c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pcfe_roiaware_pool3d_backward_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                                  c_int64, c_void_p, c_int, c_void_p]),
    "pcfe_debug_sincosf": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    "pcfe_debug_axis_sweep": (c_int, [ctypes.c_float, ctypes.c_float, ctypes.c_float, c_void_p, c_int, c_void_p]),
}


def lib():
    """Loads libpcfe.so.  Raises if it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m detmatch_b200.build` "
                "(or __graft_entry__.build()).  detmatch_b200 has no CPU / PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().pcfe_error_string(rc).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")


def f3(v):
    assert len(v) == 3, "voxel_size must have 3 entries"
    return _f3(*[float(x) for x in v])  # narrowed to float32 like pybind's std::vector<float>


def f6(v):
    assert len(v) == 6, "coors_range must have 6 entries"
    return _f6(*[float(x) for x in v])


def grid_size(voxel_size, coors_range):
    g = (ctypes.c_int32 * 3)()
    check(lib().pcfe_grid_size(f3(voxel_size), f6(coors_range), g), "pcfe_grid_size")
    return [int(g[0]), int(g[1]), int(g[2])]


def profile(enable):
    check(lib().pcfe_profile_enable(1 if enable else 0), "pcfe_profile_enable")


def profile_report():
    """{kernel name: (total_ms, launches)} since the last report."""
    buf = ctypes.create_string_buffer(8192)
    lib().pcfe_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, ms, cnt = line.split()
        out[name] = (float(ms), int(cnt))
    return out


def debug_set(name, value):
    check(lib().pcfe_debug_set(name.encode(), int(value)), f"pcfe_debug_set({name})")
