"""ctypes binding of libcc3d_b200.so (C-ABI declared in include/cc3d_b200.h).

There is no CPU fallback: if the CUDA library is missing or fails, the error is raised to the caller.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CC3D_B200_LIB") or os.path.join(_HERE, "libcc3d_b200.so")   # env: experimental builds only

HOST, DEVICE = 0, 1
U8, U16, U32, U64, F32, F64 = range(6)


class ResolveInfo(ctypes.Structure):
  _fields_ = [
    ("N", ctypes.c_uint64),
    ("epl", ctypes.c_uint64),
    ("first_foreground_row", ctypes.c_int64),
    ("last_foreground_row", ctypes.c_int64),
  ]


class CC3DB200Error(RuntimeError):
  def __init__(self, code, message):
    super().__init__(f"cc3d_b200 error {code}: {message}")
    self.code = code


_lib = None


def lib():
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise ImportError(
      f"{LIB_PATH} not found: build it with `python connected-components-3d_b200/build.py` "
      "(cc3d_b200 has no CPU fallback)"
    )
  L = ctypes.CDLL(LIB_PATH)
  vp, i64, u64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int
  p = ctypes.POINTER
  L.cc3d_b200_last_error.restype = ctypes.c_char_p
  L.cc3d_b200_version.restype = ctypes.c_char_p
  L.cc3d_b200_prepass.restype = ci
  L.cc3d_b200_prepass.argtypes = [vp, ci, i64, i64, i64, ci, p(u64), p(i64), p(i64), vp, vp, vp]
  L.cc3d_b200_label_resolve.restype = ci
  L.cc3d_b200_label_resolve.argtypes = [vp, ci, i64, i64, i64, ci, vp, ci, ci, ci, vp, p(ResolveInfo), p(vp)]
  L.cc3d_b200_label_write.restype = ci
  L.cc3d_b200_label_write.argtypes = [vp, vp, ci, ci, vp]
  L.cc3d_b200_label_write_rows.restype = ci
  L.cc3d_b200_label_write_rows.argtypes = [vp, i64, i64, vp, ci, vp]
  L.cc3d_b200_label_write_remap.restype = ci
  L.cc3d_b200_label_write_remap.argtypes = [vp, vp, ci, u64, vp, ci, ci, vp]
  L.cc3d_b200_slab_begin.restype = ci
  L.cc3d_b200_slab_begin.argtypes = [vp, ci, i64, i64, i64, ci, vp, ci, vp, p(vp), vp, vp, vp]
  L.cc3d_b200_face_pairs_async.restype = ci
  L.cc3d_b200_face_pairs_async.argtypes = [vp, vp, vp, vp, ci, i64, i64, ci, vp, ci, vp, u64, vp, vp]
  L.cc3d_b200_slab_finish.restype = ci
  L.cc3d_b200_slab_finish.argtypes = [vp, vp, ci, vp, ci, vp]
  L.cc3d_b200_face_pairs.restype = ci
  L.cc3d_b200_face_pairs.argtypes = [vp, vp, vp, vp, ci, i64, i64, ci, vp, ci, vp, u64, p(u64), vp]
  L.cc3d_b200_solve_pairs.restype = ci
  L.cc3d_b200_solve_pairs.argtypes = [vp, i64, vp, vp, i64, vp]
  L.cc3d_b200_merge_slabs.restype = ci
  L.cc3d_b200_merge_slabs.argtypes = [ci, vp, vp, vp, ci, vp, p(i64)]
  L.cc3d_b200_merge_workspace_bytes.restype = ctypes.c_size_t
  L.cc3d_b200_merge_workspace_bytes.argtypes = [u64]
  L.cc3d_b200_merge_slabs_device.restype = ci
  L.cc3d_b200_merge_slabs_device.argtypes = [vp, ci, i64, ci, u64, vp, u64, p(vp), p(vp), vp]
  L.cc3d_b200_merge_slabs_device_small.restype = ci
  L.cc3d_b200_merge_slabs_device_small.argtypes = [vp, ci, i64, ci, u64, vp, u64, p(vp), p(vp), vp]
  L.cc3d_b200_session_release.restype = None
  L.cc3d_b200_session_release.argtypes = [vp]
  L.cc3d_b200_label.restype = ci
  L.cc3d_b200_label.argtypes = [vp, ci, i64, i64, i64, ci, vp, ci, ci, vp, ci, ci, p(u64), vp]
  L.cc3d_b200_label_with_info.restype = ci
  L.cc3d_b200_label_with_info.argtypes = [vp, ci, i64, i64, i64, ci, vp, ci, ci, vp, ci, ci, p(ResolveInfo), vp]
  L.cc3d_b200_statistics.restype = ci
  L.cc3d_b200_statistics.argtypes = [vp, ci, i64, i64, i64, u64, vp, vp, vp, ci, vp]
  L.cc3d_b200_statistics_auto.restype = ci
  L.cc3d_b200_statistics_auto.argtypes = [vp, ci, i64, i64, i64, u64, p(u64), vp, vp, vp, ci, vp]
  L.cc3d_b200_voxel_connectivity_graph.restype = ci
  L.cc3d_b200_voxel_connectivity_graph.argtypes = [vp, ci, i64, i64, i64, ci, vp, ci, vp]
  L.cc3d_b200_color_connectivity_graph.restype = ci
  L.cc3d_b200_color_connectivity_graph.argtypes = [vp, ci, i64, i64, i64, ci, vp, p(u64), ci, vp]
  L.cc3d_b200_contacts.restype = ci
  L.cc3d_b200_contacts.argtypes = [vp, ci, i64, i64, i64, ci, vp, vp, u64, p(u64), ci, vp]
  L.cc3d_b200_crackle_v0_decode.restype = ci
  L.cc3d_b200_crackle_v0_decode.argtypes = [vp, vp, i64, i64, i64, vp, u64, vp, p(u64), ci, vp]
  L.cc3d_b200_remap_labels.restype = ci
  L.cc3d_b200_remap_labels.argtypes = [vp, ci, i64, vp, u64, vp, ci, ci, vp]
  L.cc3d_b200_runs.restype = ci
  L.cc3d_b200_runs.argtypes = [vp, ci, i64, vp, vp, vp, u64, p(u64), ci, vp]
  L.cc3d_b200_draw.restype = ci
  L.cc3d_b200_draw.argtypes = [vp, ci, i64, u64, vp, vp, u64, ci, vp]
  L.cc3d_b200_dust.restype = ci
  L.cc3d_b200_dust.argtypes = [vp, vp, ci, i64, i64, i64, ci, ci, i64, i64, ci, ci, p(u64), p(u64), vp]
  L.cc3d_b200_mask_by_label.restype = ci
  L.cc3d_b200_mask_by_label.argtypes = [vp, ci, vp, ci, i64, vp, u64, ci, vp]
  L.cc3d_b200_workspace_bytes.restype = ctypes.c_size_t
  L.cc3d_b200_release_workspace.restype = None
  L.cc3d_b200_launch_count.restype = ctypes.c_ulonglong
  L.cc3d_b200_debug_set_queue_capacity.restype = None
  L.cc3d_b200_debug_set_queue_capacity.argtypes = [u64]
  L.cc3d_b200_debug_set_big_tiles.restype = None
  L.cc3d_b200_debug_set_big_tiles.argtypes = [ci]
  L.cc3d_b200_set_timing.restype = None
  L.cc3d_b200_set_timing.argtypes = [ci]
  L.cc3d_b200_last_timings.restype = ci
  L.cc3d_b200_last_timings.argtypes = [p(ctypes.c_char_p), p(ctypes.c_float), ci]
  _lib = L
  return L


def check(rc):
  if rc != 0:
    raise CC3DB200Error(rc, lib().cc3d_b200_last_error().decode())


def last_timings():
  L = lib()
  names = (ctypes.c_char_p * 64)()
  ms = (ctypes.c_float * 64)()
  n = L.cc3d_b200_last_timings(names, ms, 64)
  return [(names[i].decode(), float(ms[i])) for i in range(n)]
