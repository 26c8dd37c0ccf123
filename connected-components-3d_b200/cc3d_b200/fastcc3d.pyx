# cython: language_level=3
"""fastcc3d — the COMPILED Cython boundary of cc3d_b200 (the binding INTEGRATION.md describes, built by build()).

It is what the reference's `cc3d/fastcc3d.pyx` becomes when its C++ template externs are replaced by the C-ABI of
libcc3d_b200.so:

  reference                                                   here
  ----------------------------------------------------------  -------------------------------------------------------
  cdef extern from "cc3d.hpp" ... estimate_provisional_       cdef extern from "cc3d_b200.h": cc3d_b200_prepass
    label_count<T>(...)                 (fastcc3d.pyx:60-66)
  cdef extern from "cc3d_continuous.hpp" ...                   cc3d_b200_label_resolve + cc3d_b200_label_write (two phases:
    connected_components3d<T,U>(...)    (fastcc3d.pyx:67-74)     the out dtype is chosen from epl before the output exists)
  18 typed dispatch branches            (fastcc3d.pyx:471-608)  ONE call: the element kind is an argument of the C-ABI
  Cython statistics loops               (fastcc3d.pyx:771-938)  cc3d_b200_statistics (+ the same host finalisation)

The argument handling above the call (validation, layout normalisation, dtype views, out-dtype rule, single-row
quirk, final reshape) follows fastcc3d.pyx:245-626 step by step so that error types, messages and dtype choices are
the reference's. The GIL is released around the C-ABI calls. Host (numpy) buffers only - device tensors enter through
cc3d_b200.connected_components (ctypes on the same C-ABI). There is no CPU fallback: if libcc3d_b200.so cannot launch
its kernels the call raises.
"""
from libc.stdint cimport int64_t, uint64_t, uint32_t, uintptr_t

import numpy as np


cdef extern from "cc3d_b200.h":
  ctypedef struct cc3d_b200_session:
    pass
  ctypedef struct cc3d_b200_resolve_info:
    uint64_t N
    uint64_t epl
    int64_t first_foreground_row
    int64_t last_foreground_row
  const char* cc3d_b200_last_error() nogil
  const char* cc3d_b200_version() nogil
  int cc3d_b200_prepass(const void* inp, int in_kind, int64_t sx, int64_t sy, int64_t sz, int mem_space,
                        uint64_t* epl, int64_t* first_row, int64_t* last_row, void* vmin, void* vmax, void* stream) nogil
  int cc3d_b200_label_resolve(const void* inp, int in_kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                              const void* delta, int binary_image, int periodic_boundary, int mem_space, void* stream,
                              cc3d_b200_resolve_info* info, cc3d_b200_session** session) nogil
  int cc3d_b200_label_write(cc3d_b200_session* session, void* out, int out_kind, int mem_space, void* stream) nogil
  void cc3d_b200_session_release(cc3d_b200_session* session) nogil
  int cc3d_b200_statistics(const void* labels, int kind, int64_t sx, int64_t sy, int64_t sz, uint64_t N,
                           uint32_t* counts, uint32_t* bbox, uint64_t* sums, int mem_space, void* stream) nogil

cdef enum:
  K_U8 = 0
  K_U16 = 1
  K_U32 = 2
  K_U64 = 3
  K_F32 = 4
  K_F64 = 5
  HOST = 0


class DimensionError(Exception):
  """The array has the wrong number of dimensions."""
  pass


cdef object _raise_status(int rc):
  msg = cc3d_b200_last_error().decode("utf8", "replace")
  raise RuntimeError(f"cc3d_b200 error {rc}: {msg}")


cdef int _kind_of(dtype) except -1:
  dtype = np.dtype(dtype)
  if dtype == np.float32:
    return K_F32
  if dtype == np.float64:
    return K_F64
  if dtype == np.bool_ or np.issubdtype(dtype, np.integer):
    return {1: K_U8, 2: K_U16, 4: K_U32, 8: K_U64}[dtype.itemsize]
  raise TypeError(
    f"Type {dtype} is not currently supported. "
    f"Supported: bool, int8, int16, int32, int64, uint8, uint16, uint32, uint64, float16, float32, float64")


cdef size_t _even_ceil(size_t n):
  return n << 1 if n & 1 else n      # (sic) fastcc3d.pyx:163-166


def version():
  return cc3d_b200_version().decode()


def estimate_provisional_labels(data):
  """(epl, first foreground row, last foreground row) - fastcc3d.pyx:169-242 over cc3d_b200_prepass."""
  cdef uint64_t epl = 0
  cdef int64_t first = 0, last = 0
  if not (data.flags.f_contiguous or data.flags.c_contiguous):
    data = np.ascontiguousarray(data)
  cdef int64_t sx = data.shape[0] if data.flags.f_contiguous else data.shape[-1]
  cdef int64_t rows = data.size // sx if sx else 0
  cdef int kind = _kind_of(data.dtype)
  cdef uintptr_t ptr = data.ctypes.data
  cdef int rc
  with nogil:
    rc = cc3d_b200_prepass(<const void*>ptr, kind, sx, rows, 1, HOST, &epl, &first, &last, NULL, NULL, NULL)
  if rc:
    _raise_status(rc)
  return int(epl), int(first), int(last)


def connected_components(
  data, int64_t max_labels=-1, int64_t connectivity=26, bint return_N=False, delta=0, out_dtype=None, out_file=None,
  bint periodic_boundary=False, bint binary_image=False,
):
  """Connected components of a 1D / 2D / 3D numpy array (or CPU torch tensor); contract of
  cc3d.connected_components (fastcc3d.pyx:245-626)."""
  cdef bint is_torch = hasattr(data, "cpu") and hasattr(data, "numpy")
  if is_torch:
    data = data.cpu().numpy()
  cdef int dims = len(data.shape)
  if dims not in (1, 2, 3):
    raise DimensionError("Only 1D, 2D, and 3D arrays supported. Got: " + str(dims))
  if dims == 2 and connectivity not in (4, 8, 6, 18, 26):
    raise ValueError("Only 4, 8, and 6, 18, 26 connectivities are supported for 2D images. Got: " + str(connectivity))
  elif dims != 2 and connectivity not in (6, 18, 26):
    raise ValueError("Only 6, 18, and 26 connectivities are supported for 3D images. Got: " + str(connectivity))
  if periodic_boundary and connectivity not in (4, 8, 6):
    raise ValueError(f"periodic_boundary is not yet implemented for {connectivity}-connectivity.")
  if periodic_boundary and delta != 0:
    raise ValueError("periodic_boundary is not yet implemented continuous data.")

  if data.size == 0:
    odt = data.dtype if out_dtype is None else out_dtype
    out_labels = np.zeros(shape=(0,), dtype=odt)
    if is_torch:
      import torch
      out_labels = torch.from_numpy(out_labels)
    return (out_labels, 0) if return_N else out_labels

  order = "F" if data.flags.f_contiguous else "C"
  if not data.flags.c_contiguous and not data.flags.f_contiguous:
    data = np.copy(data, order=order)
  shape = list(data.shape)
  shape_in = tuple(shape)
  while len(shape) < 3:      # C: new leading axes, F: new trailing axes (fastcc3d.pyx:337-341)
    shape = [1] + shape if order == "C" else shape + [1]

  dtype = data.dtype
  if dtype == np.float16:
    if delta == 0:
      data = data.view(np.uint16)
      dtype = data.dtype
    else:
      raise TypeError("float16 is not supported for continuous images (delta != 0).")
  elif dtype == np.bool_:
    data = data.view(np.uint8)
  elif np.issubdtype(dtype, np.signedinteger):
    data = data.view(f"u{dtype.itemsize}")

  cdef int64_t sx, sy, sz
  if order == "C":
    sx, sy, sz = shape[2], shape[1], shape[0]
  else:
    sx, sy, sz = shape[0], shape[1], shape[2]
  cdef size_t voxels = <size_t>sx * <size_t>sy * <size_t>sz
  cdef int kind = _kind_of(dtype)

  binary_image = binary_image or dtype == np.bool_
  if np.issubdtype(dtype, np.floating):
    delta = float(delta)
    is_max_delta = delta == np.finfo(dtype).max
  else:
    delta = int(delta)
    is_max_delta = (dtype != np.bool_) and delta == np.iinfo(dtype).max
  cdef bint epl_skipped = binary_image                      # fastcc3d.pyx:381-386
  binary_image = binary_image or is_max_delta               # fastcc3d.pyx:390-395
  kdtype = data.dtype                                       # the unsigned / float view the kernels see
  with np.errstate(over="ignore"):
    if np.issubdtype(kdtype, np.floating):
      delta_arr = np.array([delta], dtype=kdtype)
    else:
      delta_arr = np.array([delta & ((1 << (8 * kdtype.itemsize)) - 1)], dtype=kdtype)

  cdef uintptr_t in_ptr = data.ctypes.data
  cdef uintptr_t delta_ptr = delta_arr.ctypes.data
  cdef cc3d_b200_resolve_info info
  cdef cc3d_b200_session* sess = NULL
  cdef int rc
  cdef int c_conn = <int>connectivity, c_bin = 1 if binary_image else 0, c_per = 1 if periodic_boundary else 0
  with nogil:
    rc = cc3d_b200_label_resolve(<const void*>in_ptr, kind, sx, sy, sz, c_conn, <const void*>delta_ptr, c_bin, c_per,
                                 HOST, NULL, &info, &sess)
  if rc:
    _raise_status(rc)

  cdef size_t epl
  cdef int64_t first_row, last_row
  cdef size_t max_lab, uf
  cdef uintptr_t out_ptr
  cdef int out_kind
  try:
    if epl_skipped:
      epl, first_row, last_row = voxels, 0, sy
    else:
      epl, first_row, last_row = info.epl, info.first_foreground_row, info.last_foreground_row
    # a single foreground row is labelled by a fast path that ignores periodic_boundary (fastcc3d.pyx:469-470, 644-679)
    if periodic_boundary and delta == 0 and first_row == last_row and first_row >= 0:
      cc3d_b200_session_release(sess)
      sess = NULL
      with nogil:
        rc = cc3d_b200_label_resolve(<const void*>in_ptr, kind, sx, sy, sz, c_conn, <const void*>delta_ptr, c_bin, 0,
                                     HOST, NULL, &info, &sess)
      if rc:
        _raise_status(rc)

    max_lab = min(epl, voxels)
    uf = _even_ceil(shape[0]) * _even_ceil(shape[1]) * _even_ceil(shape[2])
    if binary_image:
      if connectivity in (4, 6):
        max_lab = min(max_lab, (uf // 2) + 1)
      else:                                   # (sic) 8 and 18 take the 26-connected bound, fastcc3d.pyx:412
        max_lab = min(max_lab, (uf // 8) + 1)

    if out_dtype is not None:
      out_dtype = np.dtype(out_dtype)
      if out_dtype not in (np.uint16, np.uint32, np.uint64):
        raise ValueError(f"Explicitly defined out_dtype ({out_dtype}) must be one of: np.uint16, np.uint32, np.uint64")
      if np.iinfo(out_dtype).max < max_lab:
        raise ValueError(f"Explicitly defined out_dtype ({out_dtype}) is too small "
                         f"to contain the estimated maximum number of labels ({max_lab}).")
    elif max_lab < np.iinfo(np.uint16).max:
      out_dtype = np.dtype(np.uint16)
    elif max_lab < np.iinfo(np.uint32).max:
      out_dtype = np.dtype(np.uint32)
    else:
      out_dtype = np.dtype(np.uint64)

    if out_file is None:
      out_flat = np.empty((voxels,), dtype=out_dtype)
    else:
      import os
      if isinstance(out_file, str):
        with open(out_file, "wb") as f:
          os.ftruncate(f.fileno(), voxels * out_dtype.itemsize)
      out_flat = np.memmap(out_file, order="F", dtype=out_dtype, shape=(voxels,))
    out_ptr = out_flat.ctypes.data
    out_kind = {2: K_U16, 4: K_U32, 8: K_U64}[out_dtype.itemsize]
    with nogil:
      rc = cc3d_b200_label_write(sess, <void*>out_ptr, out_kind, HOST, NULL)     # releases the session
    sess = NULL
    if rc:
      _raise_status(rc)
  finally:
    if sess != NULL:
      cc3d_b200_session_release(sess)

  out_labels = out_flat.reshape(shape_in, order=order)      # _final_reshape, fastcc3d.pyx:628-642
  if is_torch:
    import torch
    out_labels = torch.from_numpy(out_labels)
  if return_N:
    return (out_labels, int(info.N))
  return out_labels


def statistics(out_labels, bint no_slice_conversion=False):
  """Voxel counts, bounding boxes and centroids per label (fastcc3d.pyx:682-938) over cc3d_b200_statistics."""
  if hasattr(out_labels, "cpu") and hasattr(out_labels, "numpy"):
    out_labels = out_labels.cpu().numpy()
  while out_labels.ndim < 2:
    out_labels = out_labels[..., np.newaxis]
  if out_labels.dtype == np.bool_:
    out_labels = out_labels.view(np.uint8)
  cdef size_t voxels = out_labels.size
  if voxels == 0:
    return {"voxel_counts": None, "bounding_boxes": None, "centroids": None}
  cdef uint64_t N = int(np.max(out_labels))
  if N > voxels:
    raise ValueError(
      f"Statistics can only be computed on volumes containing labels with values lower than the number of voxels. Max: {N}")
  if np.issubdtype(out_labels.dtype, np.signedinteger):
    n_min = int(np.min(out_labels))
    if n_min < 0:
      raise ValueError(f"Statistics can only be computed on volumes containing labels with values >= 0. Min: {n_min}")
    out_labels = out_labels.view(f"u{out_labels.dtype.itemsize}")
  cdef int ndim = out_labels.ndim
  shape3 = list(out_labels.shape) + [1] * (3 - ndim)
  bdtype = np.uint32 if max(shape3) > np.iinfo(np.uint16).max else np.uint16
  forder = out_labels.flags.f_contiguous
  if not (out_labels.flags.f_contiguous or out_labels.flags.c_contiguous):
    out_labels = np.ascontiguousarray(out_labels)
    forder = False
  mem = shape3 if forder else shape3[::-1]
  counts = np.empty(N + 1, dtype=np.uint32)
  bbox = np.empty((N + 1, 6), dtype=np.uint32)
  sums = np.empty((N + 1, 3), dtype=np.uint64)
  cdef uintptr_t lp = out_labels.ctypes.data, cp = counts.ctypes.data, bp = bbox.ctypes.data, sp = sums.ctypes.data
  cdef int kind = _kind_of(out_labels.dtype)
  cdef int64_t mx = mem[0], my = mem[1], mz = mem[2]
  cdef int rc
  with nogil:
    rc = cc3d_b200_statistics(<const void*>lp, kind, mx, my, mz, N, <uint32_t*>cp, <uint32_t*>bp, <uint64_t*>sp, HOST, NULL)
  if rc:
    _raise_status(rc)
  if not forder:      # memory axes (x fastest) -> array axes
    sums = sums[:, ::-1]
    bbox = bbox.reshape(N + 1, 3, 2)[:, ::-1, :].reshape(N + 1, 6)
  bbox, sums = bbox[:, : 2 * ndim], sums[:, :ndim]
  with np.errstate(invalid="ignore", divide="ignore"):
    centroids = sums.astype(np.float64) / counts[:, None].astype(np.float64)
  centroids[counts == 0] = np.nan
  bbxes = np.ascontiguousarray(np.where(bbox == np.iinfo(np.uint32).max, np.iinfo(bdtype).max, bbox).astype(bdtype))
  output = {"voxel_counts": counts, "bounding_boxes": bbxes, "centroids": np.ascontiguousarray(centroids)}
  if no_slice_conversion:
    return output
  slices = []
  for row in bbxes:
    mins, maxs = row[0::2], row[1::2]
    if all(int(m) < voxels for m in mins):      # fastcc3d.pyx:837, 931
      slices.append(tuple(slice(int(a), int(b) + 1) for a, b in zip(mins, maxs)))
    else:
      slices.append(None)
  output["bounding_boxes"] = slices
  return output


# The reference exposes its run helpers as cc3d.fastcc3d.* too (fastcc3d.pyx:1258-1326); they live in the Python
# layer here (the run table is extracted by cc3d_b200_runs / drawn by cc3d_b200_draw).
def runs(labels):
  from . import runs as _runs
  return _runs(labels)


def draw(label, runs, image):
  from . import draw as _draw
  return _draw(label, runs, image)


def erase(runs, image):
  from . import erase as _erase
  return _erase(runs, image)


_erase = erase
