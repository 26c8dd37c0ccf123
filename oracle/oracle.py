"""oracle.py — TEST INFRASTRUCTURE ONLY. Never imported by the product package.

Python front end of the CPU oracle: a ctypes binding of oracle/cc3d_oracle.c (plain-C restatement
of the reference's two-pass union-find labelling) plus numpy restatements of the reference's
boundary logic, `statistics` and `dust`:

  connected_components  <- fastcc3d.pyx:245-626 (validation, layout normalisation, out-dtype rule A.2)
  estimate_provisional_labels <- fastcc3d.pyx:169-242 / cc3d.hpp:287-315
  statistics            <- fastcc3d.pyx:682-938
  dust                  <- cc3d/__init__.py:71-155

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
Parity pin: tests/test_oracle.py checks it against the reference build (oracle/_ref) when that is
present and against tests/golden/*.npz (vectors generated from the reference) always.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libcc3d_oracle.so")
_lib = None

_KIND = {1: {"u": 0}, 2: {"u": 1}, 4: {"u": 2, "f": 4}, 8: {"u": 3, "f": 5}}


def build(force: bool = False) -> str:
  """Compile the C restatement with gcc into oracle/_build/ (git-ignored, travels with gpurun)."""
  src = os.path.join(_HERE, "cc3d_oracle.c")
  if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _LIB_PATH, src, "-lm"])
  return _LIB_PATH


def _load():
  global _lib
  if _lib is None:
    build()
    lib = ctypes.CDLL(_LIB_PATH)
    i64, vp, u64p = ctypes.c_int64, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)
    lib.cc3d_oracle_label.restype = ctypes.c_int
    lib.cc3d_oracle_label.argtypes = [vp, ctypes.c_int, i64, i64, i64, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, vp, u64p]
    lib.cc3d_oracle_epl.restype = ctypes.c_uint64
    lib.cc3d_oracle_epl.argtypes = [vp, ctypes.c_int, i64, i64, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.cc3d_oracle_statistics.restype = None
    lib.cc3d_oracle_statistics.argtypes = [vp, ctypes.c_int, i64, i64, i64, ctypes.c_uint64, vp, vp, vp]
    _lib = lib
  return _lib


def reference_module():
  """The UNMODIFIED reference built by oracle/build_ref.sh, or None if oracle/_ref is absent."""
  ref_dir = os.path.join(_HERE, "_ref")
  if not os.path.isdir(ref_dir):
    return None
  if ref_dir not in sys.path:
    sys.path.insert(0, ref_dir)
  try:
    import fastcc3d  # type: ignore
    return fastcc3d
  except ImportError:
    return None


def _as_unsigned_or_float(data: np.ndarray) -> np.ndarray:
  dt = data.dtype
  if dt == bool:
    return data.view(np.uint8)
  if np.issubdtype(dt, np.signedinteger):
    return data.view(f"u{dt.itemsize}")
  return data


def _kind(dtype) -> int:
  dtype = np.dtype(dtype)
  cls = "f" if np.issubdtype(dtype, np.floating) else "u"
  try:
    return _KIND[dtype.itemsize][cls]
  except KeyError:
    raise TypeError(f"Type {dtype} is not currently supported.")


def estimate_provisional_labels(data: np.ndarray):
  lib = _load()
  sx = data.shape[0] if data.flags.f_contiguous else data.shape[-1]
  lin = np.ascontiguousarray(_as_unsigned_or_float(data).reshape(-1, order="F" if data.flags.f_contiguous else "C"))
  first, last = ctypes.c_int64(0), ctypes.c_int64(0)
  epl = lib.cc3d_oracle_epl(lin.ctypes.data, _kind(lin.dtype), sx, lin.size, ctypes.byref(first), ctypes.byref(last))
  return int(epl), int(first.value), int(last.value)


def _even_ceil(n: int) -> int:
  return n << 1 if n & 1 else n  # (sic) fastcc3d.pyx:163-166


def connected_components(data, connectivity=26, return_N=False, delta=0, out_dtype=None,
                         periodic_boundary=False, binary_image=False):
  lib = _load()
  dims = data.ndim
  if dims not in (1, 2, 3):
    raise ValueError("Only 1D, 2D, and 3D arrays supported.")
  if dims == 2 and connectivity not in (4, 8, 6, 18, 26):
    raise ValueError("bad connectivity")
  if dims != 2 and connectivity not in (6, 18, 26):
    raise ValueError("bad connectivity")
  if periodic_boundary and connectivity not in (4, 8, 6):
    raise ValueError("periodic")
  if periodic_boundary and delta != 0:
    raise ValueError("periodic continuous")
  if data.size == 0:
    out = np.zeros((0,), dtype=out_dtype if out_dtype is not None else data.dtype)
    return (out, 0) if return_N else out

  order = "F" if data.flags.f_contiguous else "C"
  while data.ndim < 3:
    data = data[np.newaxis, ...] if order == "C" else data[..., np.newaxis]
  if not data.flags.c_contiguous and not data.flags.f_contiguous:
    data = np.copy(data, order=order)
  if data.dtype == np.float16:
    if delta == 0:
      data = data.view(np.uint16)
    else:
      raise TypeError("float16 is not supported for continuous images (delta != 0).")
  shape = list(data.shape)
  if order == "C":
    shape.reverse()
  sx, sy, sz = shape
  voxels = sx * sy * sz
  dtype = data.dtype
  binary_image = bool(binary_image) or dtype == bool
  if binary_image:
    epl, first_row, last_row = voxels, 0, sy
  else:
    epl, first_row, last_row = estimate_provisional_labels(data)
  max_labels = min(epl, voxels)
  if np.issubdtype(dtype, np.floating):
    delta = float(delta)
    binary_image = binary_image or (delta == np.finfo(dtype).max)
  else:
    delta = int(delta)
    binary_image = binary_image or (delta == np.iinfo(np.uint8 if dtype == bool else dtype).max)
  uf = _even_ceil(data.shape[0]) * _even_ceil(data.shape[1]) * _even_ceil(data.shape[2])
  if binary_image:
    if connectivity in (4, 6):
      max_labels = min(max_labels, uf // 2 + 1)
    else:  # (sic) 8 and 18 also land here, fastcc3d.pyx:412
      max_labels = min(max_labels, uf // 8 + 1)
  if out_dtype is not None:
    out_dtype = np.dtype(out_dtype)
    if out_dtype not in (np.uint16, np.uint32, np.uint64):
      raise ValueError("out_dtype must be one of uint16, uint32, uint64")
    if np.iinfo(out_dtype).max < max_labels:
      raise ValueError("out_dtype too small")
  elif max_labels < np.iinfo(np.uint16).max:
    out_dtype = np.dtype(np.uint16)
  elif max_labels < np.iinfo(np.uint32).max:
    out_dtype = np.dtype(np.uint32)
  else:
    out_dtype = np.dtype(np.uint64)

  udata = _as_unsigned_or_float(data)
  lin = np.ascontiguousarray(udata.reshape(-1, order=order))
  # single-foreground-row fast path ignores periodic_boundary (fastcc3d.pyx:469-470, 644-679)
  special_row = (delta == 0 and first_row == last_row and first_row >= 0)
  periodic = bool(periodic_boundary) and not special_row
  out32 = np.zeros(voxels, dtype=np.uint32)
  N = ctypes.c_uint64(0)
  d = np.array([delta]).astype(lin.dtype)
  rc = lib.cc3d_oracle_label(lin.ctypes.data, _kind(lin.dtype), sx, sy, sz, int(connectivity), d.ctypes.data,
                             int(binary_image and not special_row), int(periodic), out32.ctypes.data, ctypes.byref(N))
  if rc != 0:
    raise RuntimeError(f"oracle error {rc}")
  out = out32.astype(out_dtype)
  if dims == 3:
    out = out.reshape((sz, sy, sx) if order == "C" else (sx, sy, sz), order=order)
  elif dims == 2:
    out = out.reshape((sy, sx) if order == "C" else (sx, sy), order=order)
  return (out, int(N.value)) if return_N else out


def statistics(out_labels: np.ndarray, no_slice_conversion: bool = False):
  lib = _load()
  while out_labels.ndim < 2:
    out_labels = out_labels[..., np.newaxis]
  if out_labels.dtype == bool:
    out_labels = out_labels.view(np.uint8)
  if out_labels.size == 0:
    return {"voxel_counts": None, "bounding_boxes": None, "centroids": None}
  voxels = out_labels.size
  ndim = out_labels.ndim
  N = int(np.max(out_labels))
  if N > voxels:
    raise ValueError("Statistics can only be computed on volumes containing labels with values lower than the number of voxels.")
  if np.issubdtype(out_labels.dtype, np.signedinteger):
    if np.min(out_labels) < 0:
      raise ValueError("Statistics can only be computed on volumes containing labels with values >= 0.")
    out_labels = out_labels.view(f"u{out_labels.dtype.itemsize}")
  shape3 = list(out_labels.shape) + [1] * (3 - ndim)
  bdtype = np.uint32 if max(shape3) > np.iinfo(np.uint16).max else np.uint16
  forder = out_labels.flags.f_contiguous
  lin = np.ascontiguousarray(out_labels.reshape(-1, order="F" if forder else "C"))
  mem_shape = shape3 if forder else shape3[::-1]  # (fast, mid, slow) memory axes
  counts = np.zeros(N + 1, dtype=np.uint32)
  bbox = np.zeros((N + 1, 6), dtype=np.uint32)
  bbox[:, ::2] = np.iinfo(np.uint32).max
  sums = np.zeros((N + 1, 3), dtype=np.float64)
  lib.cc3d_oracle_statistics(lin.ctypes.data, _kind(lin.dtype), mem_shape[0], mem_shape[1], mem_shape[2], N,
                             counts.ctypes.data, bbox.ctypes.data, sums.ctypes.data)
  if not forder:  # memory axes -> array axes
    sums = sums[:, ::-1]
    bbox = bbox.reshape(N + 1, 3, 2)[:, ::-1, :].reshape(N + 1, 6)
  with np.errstate(invalid="ignore", divide="ignore"):
    centroids = np.where(counts[:, None] == 0, np.nan, sums / counts[:, None].astype(np.float64))
  bb = np.where(bbox == np.iinfo(np.uint32).max, np.iinfo(bdtype).max, bbox).astype(bdtype)
  bb = np.ascontiguousarray(bb[:, : 2 * ndim])
  output = {"voxel_counts": counts, "bounding_boxes": bb, "centroids": np.ascontiguousarray(centroids[:, :ndim])}
  if no_slice_conversion:
    return output
  slices = []
  for row in bb:
    mins, maxs = row[0::2], row[1::2]
    if all(int(m) < voxels for m in mins):
      slices.append(tuple(slice(int(a), int(b) + 1) for a, b in zip(mins, maxs)))
    else:
      slices.append(None)
  output["bounding_boxes"] = slices
  return output


def dust(img, threshold, connectivity=26, in_place=False, binary_image=False, precomputed_ccl=False,
         invert=False, return_N=False):
  orig_dtype = img.dtype
  if np.issubdtype(img.dtype, np.signedinteger):
    img = img.view(f"u{img.dtype.itemsize}")
  if not in_place:
    img = np.copy(img)
  if precomputed_ccl:
    cc_labels, N = img, int(np.max(img))
  else:
    cc_labels, N = connected_components(img, connectivity=connectivity, return_N=True, binary_image=bool(binary_image))
  sizes = statistics(cc_labels, no_slice_conversion=True)["voxel_counts"]
  if isinstance(threshold, (tuple, list)):
    to_mask = [i for i in range(1, N + 1) if not (threshold[0] <= sizes[i] < threshold[1])]
  else:
    to_mask = [i for i in range(1, N + 1) if sizes[i] < threshold]
  dust_N = len(to_mask) if invert else N - len(to_mask)
  if len(to_mask) == 0:
    if invert:
      img = np.zeros(img.shape, dtype=img.dtype, order="F")
    return (img, dust_N) if return_N else img
  mask = np.isin(cc_labels, to_mask, assume_unique=True, invert=invert)
  img[mask] = 0
  img = img.view(orig_dtype)
  return (img, dust_N) if return_N else img
