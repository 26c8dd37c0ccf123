"""cc3d_b200 — B200-native drop-in for the hot path of seung-lab/connected-components-3d.

    import cc3d_b200 as cc3d
    labels, N = cc3d.connected_components(volume, connectivity=26, return_N=True)
    stats = cc3d.statistics(labels)
    clean = cc3d.dust(volume, threshold=100)

Host layer = the reference's Cython/Python boundary logic (argument validation, layout
normalisation, out-dtype rule; cc3d/fastcc3d.pyx:245-626, 682-938 and cc3d/__init__.py:71-155)
re-expressed over the C-ABI in include/cc3d_b200.h. All compute runs in hand-written sm_100a CUDA
kernels (csrc/); there is no CPU fallback. Inputs may be numpy arrays (staged through device
memory), CPU torch tensors, or CUDA torch tensors (zero copy, result stays on the device).
"""
from __future__ import annotations

import ctypes
from typing import Any, Optional, Tuple, Union

import numpy as np

from . import _lib
from ._lib import CC3DB200Error

__all__ = [
  "connected_components", "statistics", "dust", "estimate_provisional_labels",
  "largest_k", "voxel_connectivity_graph", "color_connectivity_graph", "contacts", "region_graph",
  "runs", "draw", "erase", "each", "connected_components_stack",
  "DimensionError", "CC3DB200Error", "last_timings", "set_timing",
]


class DimensionError(Exception):
  """The array has the wrong number of dimensions."""
  pass


_UNSIGNED = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}
_OUT_KIND = {np.dtype(np.uint16): _lib.U16, np.dtype(np.uint32): _lib.U32, np.dtype(np.uint64): _lib.U64}


def _kind_of(dtype) -> int:
  dtype = np.dtype(dtype)
  if dtype == np.float32:
    return _lib.F32
  if dtype == np.float64:
    return _lib.F64
  if dtype == bool or np.issubdtype(dtype, np.integer):
    return {1: _lib.U8, 2: _lib.U16, 4: _lib.U32, 8: _lib.U64}[dtype.itemsize]
  raise TypeError(
    f"Type {dtype} is not currently supported. "
    f"Supported: bool, int8, int16, int32, int64, uint8, uint16, uint32, uint64, float16, float32, float64"
  )


def set_timing(enabled: bool) -> None:
  """Record per-kernel CUDA-event timings for the following calls (see last_timings)."""
  _lib.lib().cc3d_b200_set_timing(int(bool(enabled)))


def last_timings():
  return _lib.last_timings()


# ----------------------------------------------------------------------------------------------
# torch interop helpers (torch is optional: only touched when a tensor is passed in)
# ----------------------------------------------------------------------------------------------
def _is_torch(x) -> bool:
  return hasattr(x, "cpu") and hasattr(x, "data_ptr")


def _adopt_device_array(x):
  """Arrays that are neither numpy arrays nor torch tensors but expose their buffer - `__cuda_array_interface__`
  (CuPy, Numba device arrays) or DLPack (`__dlpack__`) - are wrapped zero copy as torch tensors, so that device
  buffers of other libraries take the same zero-copy path as CUDA torch tensors (the result is a torch tensor on
  the same device; hand it back with `cupy.from_dlpack(out)` or the like). Anything else is returned unchanged."""
  if isinstance(x, np.ndarray) or _is_torch(x):
    return x
  if hasattr(x, "__cuda_array_interface__"):
    import torch
    return torch.as_tensor(x, device="cuda")
  if hasattr(x, "__dlpack__") and hasattr(x, "__dlpack_device__"):
    import torch
    return torch.from_dlpack(x)
  return x


def _torch_np_dtype(t):
  import torch
  table = {
    torch.bool: np.bool_, torch.uint8: np.uint8, torch.int8: np.int8, torch.int16: np.int16,
    torch.int32: np.int32, torch.int64: np.int64, torch.float16: np.float16,
    torch.float32: np.float32, torch.float64: np.float64,
  }
  for name in ("uint16", "uint32", "uint64"):
    if hasattr(torch, name):
      table[getattr(torch, name)] = getattr(np, name)
  if t.dtype not in table:
    raise TypeError(f"Type {t.dtype} is not currently supported.")
  return np.dtype(table[t.dtype])


def _torch_dtype(np_dtype):
  import torch
  return {np.dtype(np.uint16): torch.uint16, np.dtype(np.uint32): torch.uint32, np.dtype(np.uint64): torch.uint64,
          np.dtype(np.uint8): torch.uint8}[np.dtype(np_dtype)]


def _torch_order(t) -> Tuple[Any, str]:
  """Returns (tensor that is dense in memory, 'C' or 'F')."""
  if t.is_contiguous():
    return t, "C"
  if t.ndim > 1 and t.permute(*reversed(range(t.ndim))).is_contiguous():
    return t, "F"
  return t.contiguous(), "C"


# ----------------------------------------------------------------------------------------------
# estimate_provisional_labels (fastcc3d.pyx:169-242)
# ----------------------------------------------------------------------------------------------
def estimate_provisional_labels(data: np.ndarray) -> Tuple[int, int, int]:
  if _is_torch(data):
    data = data.cpu().numpy()
  sx = data.shape[0] if data.flags.f_contiguous else data.shape[-1]
  if not (data.flags.f_contiguous or data.flags.c_contiguous):
    data = np.ascontiguousarray(data)
  kind = _kind_of(data.dtype)
  epl, first, last = ctypes.c_uint64(0), ctypes.c_int64(0), ctypes.c_int64(0)
  rows = data.size // sx if sx else 0
  _lib.check(_lib.lib().cc3d_b200_prepass(
    data.ctypes.data, kind, sx, rows, 1, _lib.HOST,
    ctypes.byref(epl), ctypes.byref(first), ctypes.byref(last), None, None, None))
  return int(epl.value), int(first.value), int(last.value)


def _even_ceil(n: int) -> int:
  return n << 1 if n & 1 else n  # (sic) fastcc3d.pyx:163-166


# ----------------------------------------------------------------------------------------------
# connected_components (fastcc3d.pyx:245-626)
# ----------------------------------------------------------------------------------------------
def connected_components(
  data, max_labels: int = -1, connectivity: int = 26, return_N: bool = False,
  delta: Union[int, float] = 0, out_dtype: Optional[Any] = None, out_file=None,
  periodic_boundary: bool = False, binary_image: bool = False, out=None,
):
  """Connected components of a 1D/2D/3D image; same contract as cc3d.connected_components.

  Extension: `out` may be a preallocated contiguous numpy array (e.g. backed by pinned memory) of
  the right dtype and size for host inputs; the labels are written into it instead of a new array.

  connectivity: 6/18/26 (3D) or 4/8 (2D); delta > 0 joins values differing by <= delta;
  binary_image treats non-zero as foreground; periodic_boundary wraps 4/8/6-connected images.
  Components are numbered 1..N in order of first appearance in memory order. `max_labels` is
  accepted and ignored, as in the reference (fastcc3d.pyx:263-268, 388).
  """
  L = _lib.lib()
  data = _adopt_device_array(data)
  is_torch = _is_torch(data)
  on_device = False
  tensor = None
  if is_torch:
    if data.is_cuda:
      on_device = True
      tensor, order = _torch_order(data.detach())
      shape_in = tuple(tensor.shape)
      dtype = _torch_np_dtype(tensor)
      size = tensor.numel()
    else:
      data = data.cpu().numpy()

  if not on_device:
    shape_in = tuple(data.shape)
    dtype = data.dtype
    size = data.size

  dims = len(shape_in)
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

  if size == 0:
    odt = dtype if out_dtype is None else out_dtype
    out_labels = np.zeros(shape=(0,), dtype=odt)
    if is_torch:
      import torch
      out_labels = torch.from_numpy(out_labels)
      if on_device:
        out_labels = out_labels.to(tensor.device)
    return (out_labels, 0) if return_N else out_labels

  if not on_device:
    order = "F" if data.flags.f_contiguous else "C"
    if not data.flags.c_contiguous and not data.flags.f_contiguous:
      data = np.copy(data, order=order)

  shape3 = list(shape_in)
  while len(shape3) < 3:  # C: new leading axes, F: new trailing axes (fastcc3d.pyx:337-341)
    shape3 = [1] + shape3 if order == "C" else shape3 + [1]

  if dtype == np.float16:
    if delta == 0:
      dtype = np.dtype(np.uint16)
    else:
      raise TypeError("float16 is not supported for continuous images (delta != 0).")

  sx, sy, sz = (shape3[::-1] if order == "C" else shape3)  # x = fastest memory axis
  voxels = sx * sy * sz
  kind = _kind_of(dtype)
  if voxels >= _MAX_CALL_VOXELS:
    # One C-ABI call handles < 2^32-1 voxels (32-bit voxel / run indices on the device). The reference has no such
    # limit (int64 counts, uint64 labels), so larger volumes are split into z-slabs here and merged by the
    # sharded machinery - same labels, same numbering, same dtype rule as one monolithic call would give.
    return _connected_components_split(tensor if on_device else data, on_device, is_torch, order, shape_in, (sx, sy, sz),
                                       connectivity, return_N, delta, out_dtype, periodic_boundary, binary_image, out_file)
  orig_dtype = np.dtype(dtype)
  binary_image = bool(binary_image) or orig_dtype == bool

  if np.issubdtype(orig_dtype, np.floating):
    delta = float(delta)
    is_max_delta = (delta == np.finfo(orig_dtype).max)
  else:
    delta = int(delta)
    is_max_delta = (orig_dtype != bool) and (delta == np.iinfo(orig_dtype).max)
  epl_skipped = binary_image                      # fastcc3d.pyx:381-386
  binary_image = binary_image or is_max_delta     # fastcc3d.pyx:390-395

  # delta as one element of the kernel's element type (signed ints run as their unsigned views)
  kdtype = orig_dtype
  if orig_dtype == bool:
    kdtype = np.dtype(np.uint8)
  elif np.issubdtype(orig_dtype, np.signedinteger):
    kdtype = np.dtype(_UNSIGNED[orig_dtype.itemsize])
  with np.errstate(over="ignore"):
    if np.issubdtype(kdtype, np.floating):
      delta_arr = np.array([delta], dtype=kdtype)
    else:
      delta_arr = np.array([delta & ((1 << (8 * kdtype.itemsize)) - 1)], dtype=kdtype)

  if on_device:
    in_ptr, space = tensor.data_ptr(), _lib.DEVICE
    import torch
    stream = ctypes.c_void_p(torch.cuda.current_stream(tensor.device).cuda_stream)
    dev_guard = torch.cuda.device(tensor.device)
  else:
    in_ptr, space = data.ctypes.data, _lib.HOST
    stream = None
    dev_guard = None

  def resolve(periodic):
    info = _lib.ResolveInfo()
    sess = ctypes.c_void_p()
    _lib.check(L.cc3d_b200_label_resolve(
      in_ptr, kind, sx, sy, sz, int(connectivity), delta_arr.ctypes.data, int(binary_image),
      int(periodic), space, stream, ctypes.byref(info), ctypes.byref(sess)))
    return info, sess

  def dtype_rule(epl):
    """The reference's out-dtype rule (fastcc3d.pyx:388-434) for an epl estimate."""
    max_lab = min(epl, voxels)
    uf_voxels = _even_ceil(shape3[0]) * _even_ceil(shape3[1]) * _even_ceil(shape3[2])
    if binary_image:
      if connectivity in (4, 6):
        max_lab = min(max_lab, (uf_voxels // 2) + 1)
      else:  # (sic) 8 and 18 take the 26-connected bound, fastcc3d.pyx:412
        max_lab = min(max_lab, (uf_voxels // 8) + 1)
    if out_dtype is not None:
      odt = np.dtype(out_dtype)
      if odt not in (np.uint16, np.uint32, np.uint64):
        raise ValueError(
          f"Explicitly defined out_dtype ({odt}) must be one of: np.uint16, np.uint32, np.uint64")
      if np.iinfo(odt).max < max_lab:
        raise ValueError(
          f"Explicitly defined out_dtype ({odt}) is too small "
          f"to contain the estimated maximum number of labels ({max_lab}).")
      return odt
    if max_lab < np.iinfo(np.uint16).max:
      return np.dtype(np.uint16)
    if max_lab < np.iinfo(np.uint32).max:
      return np.dtype(np.uint32)
    return np.dtype(np.uint64)

  if dev_guard is not None:
    dev_guard.__enter__()
  try:
    # Device-resident fast path: the out dtype is guessed (uint32 unless the caller fixed it) so that both
    # phases run back to back without a host round trip; the reference's rule is checked afterwards and the
    # (small-volume) cases where it picks another width are converted.
    if on_device and not (periodic_boundary and delta == 0):
      import torch
      guess = np.dtype(np.uint32)
      if out_dtype is not None and np.dtype(out_dtype) in (np.uint16, np.uint32, np.uint64) and (
          np.dtype(out_dtype) != np.uint16 or voxels < 65535):
        guess = np.dtype(out_dtype)
      out_flat = torch.empty((voxels,), dtype=_torch_dtype(guess), device=tensor.device)
      info = _lib.ResolveInfo()
      _lib.check(L.cc3d_b200_label_with_info(
        in_ptr, kind, sx, sy, sz, int(connectivity), delta_arr.ctypes.data, int(binary_image), 0, out_flat.data_ptr(),
        _OUT_KIND[guess], _lib.DEVICE, ctypes.byref(info), stream))
      final = dtype_rule(voxels if epl_skipped else int(info.epl))
      N = int(info.N)
      if final != guess:
        signed = {2: torch.int16, 4: torch.int32, 8: torch.int64}
        out_flat = out_flat.view(signed[guess.itemsize]).to(signed[final.itemsize]).view(_torch_dtype(final))
      out_labels = out_flat.reshape(shape_in) if order == "C" else \
        out_flat.reshape(tuple(reversed(shape_in))).permute(*reversed(range(dims)))
      return (out_labels, N) if return_N else out_labels

    info, sess = resolve(bool(periodic_boundary))
    try:
      if epl_skipped:
        epl, first_row, last_row = voxels, 0, sy
      else:
        epl, first_row, last_row = int(info.epl), int(info.first_foreground_row), int(info.last_foreground_row)
      # A single foreground row is labelled by a fast path that ignores periodic_boundary
      # (fastcc3d.pyx:469-470, 644-679); reproduce that by labelling without the wrap.
      if periodic_boundary and delta == 0 and first_row == last_row and first_row >= 0:
        L.cc3d_b200_session_release(sess)
        sess = None
        info, sess = resolve(False)

      max_lab = min(epl, voxels)
      uf_voxels = _even_ceil(shape3[0]) * _even_ceil(shape3[1]) * _even_ceil(shape3[2])
      if binary_image:
        if connectivity in (4, 6):
          max_lab = min(max_lab, (uf_voxels // 2) + 1)
        else:  # (sic) 8 and 18 take the 26-connected bound, fastcc3d.pyx:412
          max_lab = min(max_lab, (uf_voxels // 8) + 1)

      if out_dtype is not None:
        out_dtype = np.dtype(out_dtype)
        if out_dtype not in (np.uint16, np.uint32, np.uint64):
          raise ValueError(
            f"Explicitly defined out_dtype ({out_dtype}) must be one of: np.uint16, np.uint32, np.uint64")
        if np.iinfo(out_dtype).max < max_lab:
          raise ValueError(
            f"Explicitly defined out_dtype ({out_dtype}) is too small "
            f"to contain the estimated maximum number of labels ({max_lab}).")
      elif max_lab < np.iinfo(np.uint16).max:
        out_dtype = np.dtype(np.uint16)
      elif max_lab < np.iinfo(np.uint32).max:
        out_dtype = np.dtype(np.uint32)
      else:
        out_dtype = np.dtype(np.uint64)

      N = int(info.N)
      if on_device:
        import torch
        out_flat = torch.empty((voxels,), dtype=_torch_dtype(out_dtype), device=tensor.device)
        s2, sess = sess, None
        _lib.check(L.cc3d_b200_label_write(s2, out_flat.data_ptr(), _OUT_KIND[out_dtype], _lib.DEVICE, stream))
      else:
        if out is not None:
          if not isinstance(out, np.ndarray) or out.dtype != out_dtype or out.size != voxels or not (
              out.flags.c_contiguous or out.flags.f_contiguous):
            raise ValueError(f"out must be a contiguous numpy array of dtype {out_dtype} and size {voxels}")
          out_flat = out.ravel(order="K")
          if not np.shares_memory(out_flat, out):
            raise ValueError("out must be contiguous")
        elif out_file is None:
          out_flat = np.empty((voxels,), dtype=out_dtype)
        else:
          import os
          if isinstance(out_file, str):
            with open(out_file, "wb") as f:
              os.ftruncate(f.fileno(), voxels * np.dtype(out_dtype).itemsize)
          out_flat = np.memmap(out_file, order="F", dtype=out_dtype, shape=(voxels,))
        s2, sess = sess, None
        _lib.check(L.cc3d_b200_label_write(s2, out_flat.ctypes.data, _OUT_KIND[out_dtype], _lib.HOST, stream))
    finally:
      if sess is not None and sess.value:
        L.cc3d_b200_session_release(sess)
  finally:
    if dev_guard is not None:
      dev_guard.__exit__(None, None, None)

  # _final_reshape (fastcc3d.pyx:628-642)
  if on_device:
    if order == "C":
      out_labels = out_flat.reshape(shape_in)
    else:
      out_labels = out_flat.reshape(tuple(reversed(shape_in))).permute(*reversed(range(dims)))
  else:
    out_labels = out_flat.reshape(shape_in, order=order)
    if is_torch:
      import torch
      out_labels = torch.from_numpy(out_labels)

  if return_N:
    return (out_labels, N)
  return out_labels


_MAX_CALL_VOXELS = 0xFFFFFFFF   # tests lower this to exercise the split path on small volumes


def _connected_components_split(src, on_device, is_torch, order, shape_in, sxyz, connectivity, return_N, delta, out_dtype,
                                periodic_boundary, binary_image, out_file):
  """connected_components for volumes with >= 2^32-1 voxels: consecutive z-slabs (z = slowest memory axis) through
  cc3d_b200.sharded.connected_components_slabs on the current device. Needs torch (device buffers) and enough device
  memory for the input, the output and the per-slab workspaces; beyond that use connected_components_stack."""
  import torch
  from . import sharded
  sx, sy, sz = sxyz
  if connectivity not in (6, 18, 26) or sz < 2 or periodic_boundary:
    raise ValueError(
      f"A volume of {sx * sy * sz} voxels (>= 2^32-1) can only be labelled in z-slabs, which needs a 3D volume, "
      f"connectivity 6/18/26 and no periodic_boundary (got shape {shape_in}, connectivity {connectivity}).")
  planes = (_MAX_CALL_VOXELS - 1) // (sx * sy)
  if planes < 1:
    raise ValueError(f"A single z-plane of {sx * sy} voxels exceeds the per-call limit of 2^32-2 voxels.")
  dims = len(shape_in)
  if on_device:
    zyx = (src if order == "C" else src.permute(*reversed(range(dims)))).reshape(sz, sy, sx)
    dev = src.device
  else:
    zyx = (src if order == "C" else src.T).reshape(sz, sy, sx)   # views: the array is dense in this order
    dev = torch.device("cuda", torch.cuda.current_device())
  cuts = list(range(0, sz, planes)) + [sz]
  with torch.cuda.device(dev):
    if on_device:
      slabs = [zyx[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    else:
      slabs = [torch.from_numpy(np.ascontiguousarray(zyx[a:b])).to(dev) for a, b in zip(cuts[:-1], cuts[1:])]
    outs, N = sharded.connected_components_slabs(slabs, connectivity=connectivity, return_N=True, delta=delta,
                                                 out_dtype=out_dtype, binary_image=binary_image, whole=on_device)
  if on_device:
    full = outs
    out_labels = full.reshape(shape_in) if order == "C" else \
      full.reshape(tuple(reversed(shape_in))).permute(*reversed(range(dims)))
    return (out_labels, N) if return_N else out_labels
  np_dt = {2: np.uint16, 4: np.uint32, 8: np.uint64}[outs[0].element_size()]
  if out_file is None:
    flat = np.empty((sz, sy, sx), dtype=np_dt)
  else:
    import os
    if isinstance(out_file, str):
      with open(out_file, "wb") as f:
        os.ftruncate(f.fileno(), sx * sy * sz * np.dtype(np_dt).itemsize)
    flat = np.memmap(out_file, order="F", dtype=np_dt, shape=(sx * sy * sz,)).reshape(sz, sy, sx)
  for (a, b), o in zip(zip(cuts[:-1], cuts[1:]), outs):
    flat[a:b] = o.cpu().numpy().view(np_dt)
  out_labels = flat.reshape(-1).reshape(shape_in, order=order)
  if is_torch:
    out_labels = torch.from_numpy(out_labels)
  return (out_labels, N) if return_N else out_labels


# ----------------------------------------------------------------------------------------------
# statistics (fastcc3d.pyx:682-938)
# ----------------------------------------------------------------------------------------------
_stat_cap = {"host": 1 << 16}     # table capacity that was enough last time (per device / host)


def _statistics_arrays(out_labels, N: Optional[int] = None):
  """Raw per-label arrays in ARRAY axes: counts u32[N+1], bbox u32[N+1, 2*ndim], sums u64[N+1, ndim], and N.
  N = None: the largest label is found by the SAME sweep that accumulates the statistics (cc3d_b200_statistics_auto;
  the reference runs np.max over the volume first, fastcc3d.pyx:713-720) - the tables are sized by the capacity that
  was enough last time and the call is repeated once if the maximum turns out larger."""
  L = _lib.lib()
  ndim = out_labels.ndim
  shape3 = list(out_labels.shape) + [1] * (3 - ndim)
  forder = out_labels.flags.f_contiguous
  if not (out_labels.flags.f_contiguous or out_labels.flags.c_contiguous):
    out_labels = np.ascontiguousarray(out_labels)
    forder = False
  mem = shape3 if forder else shape3[::-1]
  kind = _kind_of(out_labels.dtype)
  if N is None:
    cap = max(2, min(_stat_cap["host"], out_labels.size + 1))
    while True:
      counts = np.empty(cap, dtype=np.uint32)
      bbox = np.empty((cap, 6), dtype=np.uint32)
      sums = np.empty((cap, 3), dtype=np.uint64)
      mx = ctypes.c_uint64(0)
      _lib.check(L.cc3d_b200_statistics_auto(
        out_labels.ctypes.data, kind, mem[0], mem[1], mem[2], cap, ctypes.byref(mx),
        counts.ctypes.data, bbox.ctypes.data, sums.ctypes.data, _lib.HOST, None))
      N = int(mx.value)
      if N < cap or N > out_labels.size:      # N > voxels: the caller raises, nothing to repeat
        break
      cap = 1 << (N + 1).bit_length()
      _stat_cap["host"] = cap
    if N > out_labels.size:
      return None, None, None, N
    counts, bbox, sums = counts[:N + 1], bbox[:N + 1], sums[:N + 1]
  else:
    counts = np.empty(N + 1, dtype=np.uint32)
    bbox = np.empty((N + 1, 6), dtype=np.uint32)
    sums = np.empty((N + 1, 3), dtype=np.uint64)
    _lib.check(L.cc3d_b200_statistics(
      out_labels.ctypes.data, kind, mem[0], mem[1], mem[2], N,
      counts.ctypes.data, bbox.ctypes.data, sums.ctypes.data, _lib.HOST, None))
  if not forder:  # memory axes (x fastest) -> array axes
    sums = sums[:, ::-1]
    bbox = bbox.reshape(N + 1, 3, 2)[:, ::-1, :].reshape(N + 1, 6)
  return counts, bbox[:, : 2 * ndim], sums[:, :ndim], N


def _statistics_arrays_device(t, N: Optional[int] = None):
  """Same as _statistics_arrays for a CUDA tensor (labels stay on the device; only the per-label arrays come back:
  one sweep, one synchronisation for the maximum, one copy of the first N + 1 table entries)."""
  import torch
  L = _lib.lib()
  t, order = _torch_order(t.detach())
  ndim = t.ndim
  shape3 = list(t.shape) + [1] * (3 - ndim)
  mem = shape3 if order == "F" else shape3[::-1]
  kind = _kind_of(_torch_np_dtype(t))
  stream = ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
  key = ("cuda", t.device.index)
  auto = N is None
  cap = max(2, min(_stat_cap.get(key, 1 << 16), t.numel() + 1)) if auto else N + 1
  while True:
    # one device buffer for the three per-label tables (sums | boxes | counts): one copy back
    buf = torch.empty((cap * 52,), dtype=torch.uint8, device=t.device)
    p_sums, p_bbox, p_counts = buf.data_ptr(), buf.data_ptr() + cap * 24, buf.data_ptr() + cap * 48
    with torch.cuda.device(t.device):
      if auto:
        mx = ctypes.c_uint64(0)
        _lib.check(L.cc3d_b200_statistics_auto(
          t.data_ptr(), kind, mem[0], mem[1], mem[2], cap, ctypes.byref(mx), p_counts, p_bbox, p_sums, _lib.DEVICE, stream))
        N = int(mx.value)
      else:
        _lib.check(L.cc3d_b200_statistics(
          t.data_ptr(), kind, mem[0], mem[1], mem[2], N, p_counts, p_bbox, p_sums, _lib.DEVICE, stream))
    if not auto or N < cap or N > t.numel():
      break
    cap = 1 << (N + 1).bit_length()
    _stat_cap[key] = cap
  if N > t.numel():
    return None, None, None, N
  n1 = N + 1
  if n1 == cap:
    host = buf.cpu().numpy()
    sums = host[: n1 * 24].view(np.uint64).reshape(n1, 3)
    bbox = host[cap * 24: cap * 24 + n1 * 24].view(np.uint32).reshape(n1, 6)
    counts = host[cap * 48: cap * 48 + n1 * 4].view(np.uint32)
  else:   # only the first N + 1 entries of every table cross PCIe
    packed = torch.cat([buf[: n1 * 24], buf[cap * 24: cap * 24 + n1 * 24], buf[cap * 48: cap * 48 + n1 * 4]]).cpu().numpy()
    sums = packed[: n1 * 24].view(np.uint64).reshape(n1, 3)
    bbox = packed[n1 * 24: n1 * 48].view(np.uint32).reshape(n1, 6)
    counts = packed[n1 * 48:].view(np.uint32)
  if order != "F":  # memory axes (x fastest) -> array axes
    sums = sums[:, ::-1]
    bbox = bbox.reshape(N + 1, 3, 2)[:, ::-1, :].reshape(N + 1, 6)
  return counts, bbox[:, : 2 * ndim], sums[:, :ndim], N


def _torch_max(t) -> int:
  import torch
  if t.numel() == 0:
    return 0
  if t.dtype in (getattr(torch, "uint16", None), getattr(torch, "uint32", None), getattr(torch, "uint64", None)):
    signed = {2: torch.int16, 4: torch.int32, 8: torch.int64}[t.element_size()]
    lo, hi = (int(v) for v in torch.stack(torch.aminmax(t.view(signed))).cpu())     # one pass, one copy back
    if lo >= 0:
      return hi
    return int(t.cpu().numpy().max())
  return int(t.max())


def _too_large(N):
  return ValueError(
    f"Statistics can only be computed on volumes containing labels with values lower than the number of voxels. Max: {N}")


def statistics(out_labels, no_slice_conversion: bool = False) -> dict:
  """Voxel counts, bounding boxes and centroids per label; same contract as cc3d.statistics.
  CUDA tensors are processed in place on their device. The volume is swept ONCE: the largest label is found by the
  sweep that accumulates the statistics (the reference takes np.max first)."""
  out_labels = _adopt_device_array(out_labels)
  if _is_torch(out_labels):
    if out_labels.is_cuda and out_labels.ndim >= 2 and out_labels.dtype != __import__("torch").bool:
      voxels = out_labels.numel()
      if voxels == 0:
        return {"voxel_counts": None, "bounding_boxes": None, "centroids": None}
      np_dtype = _torch_np_dtype(out_labels)
      signed = np.issubdtype(np_dtype, np.signedinteger)
      counts, bbox32, sums, N = _statistics_arrays_device(out_labels, None)
      # a negative label shows as a huge unsigned one: sort the two errors out like the reference does (max first)
      if counts is None or (signed and N >= (1 << (8 * np_dtype.itemsize - 1))):
        if signed and int(out_labels.min()) < 0:
          if int(out_labels.max()) > voxels:
            raise _too_large(int(out_labels.max()))
          raise ValueError(
            f"Statistics can only be computed on volumes containing labels with values >= 0. Min: {int(out_labels.min())}")
        raise _too_large(N)
      ndim = out_labels.ndim
      shape3 = list(out_labels.shape) + [1] * (3 - ndim)
      bdtype = np.uint32 if max(shape3) > np.iinfo(np.uint16).max else np.uint16
      return _finish_statistics(counts, bbox32, sums, bdtype, voxels, no_slice_conversion)
    out_labels = out_labels.cpu().numpy()
  while out_labels.ndim < 2:
    out_labels = out_labels[..., np.newaxis]
  if out_labels.dtype == bool:
    out_labels = out_labels.view(np.uint8)
  voxels = out_labels.size
  if voxels == 0:
    return {"voxel_counts": None, "bounding_boxes": None, "centroids": None}
  signed = np.issubdtype(out_labels.dtype, np.signedinteger)
  view = out_labels.view(_UNSIGNED[out_labels.dtype.itemsize]) if signed else out_labels
  counts, bbox32, sums, N = _statistics_arrays(view, None)
  if counts is None or (signed and N >= (1 << (8 * out_labels.dtype.itemsize - 1))):
    if signed and int(np.min(out_labels)) < 0:
      if int(np.max(out_labels)) > voxels:
        raise _too_large(int(np.max(out_labels)))
      raise ValueError(
        f"Statistics can only be computed on volumes containing labels with values >= 0. Min: {int(np.min(out_labels))}")
    raise _too_large(N)
  ndim = out_labels.ndim
  shape3 = list(out_labels.shape) + [1] * (3 - ndim)
  bdtype = np.uint32 if max(shape3) > np.iinfo(np.uint16).max else np.uint16
  return _finish_statistics(counts, bbox32, sums, bdtype, voxels, no_slice_conversion)


def _finish_statistics(counts, bbox32, sums, bdtype, voxels, no_slice_conversion):
  with np.errstate(invalid="ignore", divide="ignore"):
    centroids = sums.astype(np.float64) / counts[:, None].astype(np.float64)
  centroids[counts == 0] = np.nan
  bbxes = np.where(bbox32 == np.iinfo(np.uint32).max, np.iinfo(bdtype).max, bbox32).astype(bdtype)
  bbxes = np.ascontiguousarray(bbxes)
  output = {
    "voxel_counts": counts,
    "bounding_boxes": bbxes,
    "centroids": np.ascontiguousarray(centroids),
  }
  if no_slice_conversion:
    return output
  slices = []
  for row in bbxes:
    mins, maxs = row[0::2], row[1::2]
    if all(int(m) < voxels for m in mins):  # fastcc3d.pyx:837, 931
      slices.append(tuple(slice(int(a), int(b) + 1) for a, b in zip(mins, maxs)))
    else:
      slices.append(None)
  output["bounding_boxes"] = slices
  return output


# ----------------------------------------------------------------------------------------------
# dust (cc3d/__init__.py:71-155)
# ----------------------------------------------------------------------------------------------
def _view_as_unsigned(img):
  if np.issubdtype(img.dtype, np.unsignedinteger) or img.dtype == bool:
    return img
  if np.issubdtype(img.dtype, np.signedinteger):
    return img.view(_UNSIGNED[img.dtype.itemsize])
  return img


def _validate_connectivity(dims, connectivity):
  """Same checks and messages as connected_components (fastcc3d.pyx:313-320)."""
  if dims not in (1, 2, 3):
    raise DimensionError("Only 1D, 2D, and 3D arrays supported. Got: " + str(dims))
  if dims == 2 and connectivity not in (4, 8, 6, 18, 26):
    raise ValueError("Only 4, 8, and 6, 18, 26 connectivities are supported for 2D images. Got: " + str(connectivity))
  elif dims != 2 and connectivity not in (6, 18, 26):
    raise ValueError("Only 6, 18, and 26 connectivities are supported for 3D images. Got: " + str(connectivity))


def _dust_bounds(threshold):
  """[lo, hi) of the component sizes that stay (cc3d/__init__.py:117-127); sizes are integers, so a real bound b is
  equivalent to ceil(b). Non-finite bounds are legal in the reference (it compares sizes with the threshold
  directly): +inf / -inf clamp to the ends of the range, NaN compares false with everything (nothing stays)."""
  import math
  big = (1 << 62)

  def as_int(b, default):
    if b is None:
      return default
    if isinstance(b, (float, np.floating)):
      b = float(b)
      if math.isnan(b):
        return None
      if math.isinf(b):
        return big if b > 0 else -big
      b = math.ceil(b)
    return max(-big, min(big, int(b)))

  if isinstance(threshold, (tuple, list)):
    lo, hi = as_int(threshold[0], -big), as_int(threshold[1], big)
  else:
    lo, hi = as_int(threshold, -big), big
    if lo is None:            # scalar NaN: `size < nan` is false, nothing is dust
      return -big, big
  if lo is None or hi is None:
    return big, -big          # empty range: `lo <= size < hi` is false for every size, as with a NaN bound
  return lo, hi


def _dust_fused(img, threshold, connectivity, in_place, binary_image, invert, return_N):
  """dust without a label volume: cc3d_b200_dust labels the image, takes the component sizes from the run table
  and masks the image while expanding (numpy arrays and CUDA tensors; integer images, contiguous)."""
  L = _lib.lib()
  lo, hi = _dust_bounds(threshold)
  N, nm = ctypes.c_uint64(0), ctypes.c_uint64(0)
  if _is_torch(img):
    import torch
    t, order = _torch_order(img.detach())
    out = t if (in_place and t.data_ptr() == img.data_ptr()) else torch.empty_like(t, memory_format=torch.preserve_format)
    shape3 = list(t.shape) + [1] * (3 - t.ndim) if order == "F" else [1] * (3 - t.ndim) + list(t.shape)
    sx, sy, sz = shape3 if order == "F" else shape3[::-1]
    if t.numel():
      with torch.cuda.device(t.device):
        _lib.check(L.cc3d_b200_dust(
          t.data_ptr(), out.data_ptr(), _kind_of(_torch_np_dtype(t)), sx, sy, sz, int(connectivity), int(bool(binary_image)),
          lo, hi, int(bool(invert)), _lib.DEVICE, ctypes.byref(N), ctypes.byref(nm),
          ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)))
    elif out is not t:
      out.copy_(t)
    n_total, n_mask = int(N.value), int(nm.value)
    dust_N = n_mask if invert else n_total - n_mask
    return (out, dust_N) if return_N else out
  orig_dtype = img.dtype
  src = _view_as_unsigned(img)
  if src.dtype == bool:
    src = src.view(np.uint8)
  order = "F" if (src.flags.f_contiguous and not src.flags.c_contiguous) else "C"
  out = src if in_place else np.empty(src.shape, dtype=src.dtype, order=order)
  shape3 = list(src.shape)
  while len(shape3) < 3:
    shape3 = [1] + shape3 if order == "C" else shape3 + [1]
  sx, sy, sz = (shape3[::-1] if order == "C" else shape3)
  if src.size:
    _lib.check(L.cc3d_b200_dust(
      src.ctypes.data, out.ctypes.data, _kind_of(src.dtype), sx, sy, sz, int(connectivity), int(bool(binary_image)),
      lo, hi, int(bool(invert)), _lib.HOST, ctypes.byref(N), ctypes.byref(nm), None))
  n_total, n_mask = int(N.value), int(nm.value)
  dust_N = n_mask if invert else n_total - n_mask
  if n_mask == 0 and invert:
    out = np.zeros(img.shape, dtype=src.dtype, order="F")   # (sic) cc3d/__init__.py:139-146
  out = out.view(orig_dtype)
  return (out, dust_N) if return_N else out


def dust(img, threshold, connectivity: int = 26, in_place: bool = False, binary_image: bool = False,
         precomputed_ccl: bool = False, invert: bool = False, return_N: bool = False):
  """Remove connected components smaller than threshold (or outside [lo, hi)); same contract as cc3d.dust.
  A CUDA tensor is processed entirely on its device (labelling, statistics and masking) and a tensor is returned."""
  L = _lib.lib()
  if not precomputed_ccl:
    if _is_torch(img) and img.is_cuda and _torch_np_dtype(img).kind in "biu":
      return _dust_fused(img, threshold, connectivity, in_place, binary_image, invert, return_N)
    if isinstance(img, np.ndarray) and img.dtype.kind in "biu" and img.ndim in (1, 2, 3) and (
        img.flags.c_contiguous or img.flags.f_contiguous) and (img.ndim == 3 or connectivity in (4, 8, 6, 18, 26)):
      _validate_connectivity(img.ndim, connectivity)
      return _dust_fused(img, threshold, connectivity, in_place, binary_image, invert, return_N)
  if _is_torch(img) and img.is_cuda:
    return _dust_device(img, threshold, connectivity, in_place, binary_image, precomputed_ccl, invert, return_N)
  orig_dtype = img.dtype
  img = _view_as_unsigned(img)
  if not in_place:
    img = np.copy(img)

  if precomputed_ccl:
    cc_labels = img
    N = int(np.max(cc_labels))
  else:
    cc_labels, N = connected_components(img, connectivity=connectivity, return_N=True, binary_image=bool(binary_image))

  stats = statistics(cc_labels, no_slice_conversion=True)
  mask_sizes = stats["voxel_counts"]
  del stats

  sizes = mask_sizes[1:N + 1].astype(np.int64)
  if isinstance(threshold, (tuple, list)):
    masked = ~((threshold[0] <= sizes) & (sizes < threshold[1]))
  else:
    masked = sizes < threshold
  n_mask = int(np.count_nonzero(masked))

  dust_N = n_mask if invert else N - n_mask
  if n_mask == 0:
    if invert:
      img = np.zeros(img.shape, dtype=img.dtype, order="F")
    return (img, dust_N) if return_N else img

  # np.isin(cc_labels, to_mask, invert=invert) -> zero where the mask is True (__init__.py:148-150)
  keep = np.ones(N + 1, dtype=np.uint8)
  keep[1:][masked] = 0
  if invert:
    keep = 1 - keep  # background (label 0) is "not in to_mask" -> masked when inverted
  if not (img.flags.c_contiguous or img.flags.f_contiguous):
    raise ValueError("dust requires a contiguous image")
  same_layout = (cc_labels.flags.c_contiguous and img.flags.c_contiguous) or (cc_labels.flags.f_contiguous and img.flags.f_contiguous)
  if not same_layout:
    cc_labels = np.asarray(cc_labels, order="C" if img.flags.c_contiguous else "F")
  if img.itemsize not in (1, 2, 4, 8) or img.dtype.kind not in "buif":
    raise TypeError(f"Type {img.dtype} is not currently supported.")
  _lib.check(L.cc3d_b200_mask_by_label(
    img.ctypes.data, img.itemsize, cc_labels.ctypes.data, _kind_of(cc_labels.dtype), img.size,
    keep.ctypes.data, N, _lib.HOST, None))
  img = img.view(orig_dtype)
  return (img, dust_N) if return_N else img


def _dust_device(img, threshold, connectivity, in_place, binary_image, precomputed_ccl, invert, return_N):
  import torch
  L = _lib.lib()
  t, order = _torch_order(img.detach())
  if not in_place or t.data_ptr() != img.data_ptr():
    t = t.clone(memory_format=torch.preserve_format)
  if precomputed_ccl:
    cc_labels = t
    N = _torch_max(t)
  else:
    cc_labels, N = connected_components(t, connectivity=connectivity, return_N=True, binary_image=bool(binary_image))
  counts = statistics(cc_labels, no_slice_conversion=True)["voxel_counts"] if cc_labels.ndim >= 2 else \
    statistics(cc_labels[..., None], no_slice_conversion=True)["voxel_counts"]
  sizes = counts[1:N + 1].astype(np.int64)
  if isinstance(threshold, (tuple, list)):
    masked = ~((threshold[0] <= sizes) & (sizes < threshold[1]))
  else:
    masked = sizes < threshold
  n_mask = int(np.count_nonzero(masked))
  dust_N = n_mask if invert else N - n_mask
  if n_mask == 0:
    if invert:
      t = torch.zeros_like(t)
    return (t, dust_N) if return_N else t
  keep = np.ones(N + 1, dtype=np.uint8)
  keep[1:][masked] = 0
  if invert:
    keep = 1 - keep
  keep_d = torch.from_numpy(keep).to(t.device)
  lab, lorder = _torch_order(cc_labels)
  if lorder != order:
    raise ValueError("dust requires the image and its labels in the same memory layout")
  stream = ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
  with torch.cuda.device(t.device):
    _lib.check(L.cc3d_b200_mask_by_label(
      t.data_ptr(), t.element_size(), lab.data_ptr(), _kind_of(_torch_np_dtype(lab)), t.numel(),
      keep_d.data_ptr(), N, _lib.DEVICE, stream))
  return (t, dust_N) if return_N else t


# ----------------------------------------------------------------------------------------------
# Callers either side of the labelling path: largest_k, voxel / colour connectivity graphs
# ----------------------------------------------------------------------------------------------
def _remap(labels, table, N, out_dtype):
  """out[i] = table[labels[i]] on the GPU; numpy in/out (same memory order) or CUDA tensor in/out."""
  L = _lib.lib()
  table = np.ascontiguousarray(table, dtype=np.uint32)
  okind = _kind_of(np.dtype(np.uint8) if out_dtype == bool else np.dtype(out_dtype))
  if _is_torch(labels) and labels.is_cuda:
    import torch
    t = labels.contiguous()
    tdt = torch.bool if out_dtype == bool else _torch_dtype(np.dtype(out_dtype))
    out = torch.empty(t.shape, dtype=torch.uint8 if out_dtype == bool else tdt, device=t.device)
    tab = torch.from_numpy(table.view(np.int32)).to(t.device)
    with torch.cuda.device(t.device):
      _lib.check(L.cc3d_b200_remap_labels(t.data_ptr(), _kind_of(_torch_np_dtype(t)), t.numel(), tab.data_ptr(), int(N),
                                          out.data_ptr(), okind, _lib.DEVICE,
                                          ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)))
    return out.view(torch.bool) if out_dtype == bool else out
  labels = np.asarray(labels)
  order = "F" if (labels.flags.f_contiguous and not labels.flags.c_contiguous) else "C"
  src = np.asarray(_view_as_unsigned(labels), order=order)
  out = np.empty(labels.shape, dtype=np.uint8 if out_dtype == bool else out_dtype, order=order)
  if src.size:
    _lib.check(L.cc3d_b200_remap_labels(src.ctypes.data, _kind_of(src.dtype), src.size, table.ctypes.data, int(N),
                                        out.ctypes.data, okind, _lib.HOST, None))
  return out.view(bool) if out_dtype == bool else out


def largest_k(img, k: int, connectivity: int = 26, delta=0, return_N: bool = False, binary_image: bool = False,
              precomputed_ccl: bool = False):
  """Returns the k largest connected components of the image; same contract as cc3d.largest_k
  (cc3d/__init__.py:199-279, the path without fastremap: kept components are renumbered 1..k from the
  smallest to the largest). Labelling, voxel counts and the relabelling run on the GPU."""
  assert k >= 0
  is_t = _is_torch(img)
  if k == 0:
    if is_t:
      import torch
      return torch.zeros(img.shape, dtype=torch.uint16, device=img.device)
    order = "C" if img.flags.c_contiguous else "F"
    return np.zeros(img.shape, dtype=np.uint16, order=order)

  if precomputed_ccl:
    cc_labels = img.clone() if is_t else np.copy(img, order="F")
    N = _torch_max(cc_labels) if is_t else int(np.max(cc_labels))
  else:
    cc_labels, N = connected_components(img, connectivity=connectivity, return_N=True, delta=delta,
                                        binary_image=bool(binary_image))
  if N <= k:
    return (cc_labels, N) if return_N else cc_labels

  cts = statistics(cc_labels, no_slice_conversion=True)["voxel_counts"]
  if k == 1:
    table = np.zeros(N + 1, dtype=np.uint32)
    table[int(np.argmax(cts[1:])) + 1] = 1
    cc_out = _remap(cc_labels, table, N, bool)
    return (cc_out, 1) if return_N else cc_out

  preserve = np.argpartition(cts[1:], len(cts) - k - 1)[-k:]
  preserve += 1
  preserve_list = [int(l) for l in sorted(preserve, key=lambda label: cts[label])]
  table = np.zeros(N + 1, dtype=np.uint32)
  for i, label in enumerate(preserve_list):
    table[label] = i + 1
  cc_out = _remap(cc_labels, table, N, _torch_np_dtype(cc_labels) if is_t else cc_labels.dtype)
  return (cc_out, len(preserve_list)) if return_N else cc_out


def voxel_connectivity_graph(data, connectivity: int = 26):
  """Voxel connectivity graph of a multi-label image; same contract and bit layout as
  cc3d.voxel_connectivity_graph (fastcc3d.pyx:1021-1170): uint8 for connectivity 4, 8, 6 and uint32 for 18, 26,
  array-index axes (x = axis 0), Fortran-ordered result."""
  L = _lib.lib()
  is_t = _is_torch(data)
  dims = data.ndim
  if dims not in (1, 2, 3):
    raise DimensionError("Only 1D, 2D, and 3D arrays supported. Got: " + str(dims))
  if dims == 2 and connectivity not in (4, 8, 6, 18, 26):
    raise ValueError("Only 4, 8, and 6, 18, 26 connectivities are supported for 2D images. Got: " + str(connectivity))
  elif dims != 2 and connectivity not in (6, 18, 26):
    raise ValueError("Only 6, 18, and 26 connectivities are supported for 3D images. Got: " + str(connectivity))
  out_dtype = np.uint8 if connectivity in (4, 8, 6) else np.uint32
  if is_t and data.is_cuda:
    import torch
    if data.numel() == 0:
      return torch.zeros((0,), dtype=_torch_dtype(np.dtype(out_dtype)), device=data.device)
    ndt = _torch_np_dtype(data)
    if ndt.kind not in "biu":
      raise TypeError("Type {} not currently supported.".format(ndt))
    shape = tuple(data.shape) + (1,) * (3 - dims)
    t = data.reshape(shape).permute(2, 1, 0).contiguous()     # memory order: axis 0 fastest
    g = torch.empty(t.shape, dtype=_torch_dtype(np.dtype(out_dtype)), device=t.device)
    with torch.cuda.device(t.device):
      _lib.check(L.cc3d_b200_voxel_connectivity_graph(
        t.data_ptr(), _kind_of(ndt), shape[0], shape[1], shape[2], int(connectivity), g.data_ptr(), _lib.DEVICE,
        ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)))
    return g.permute(2, 1, 0).reshape(tuple(data.shape))
  if is_t:
    data = data.numpy()
  if data.size == 0:
    return np.zeros(shape=(0,), dtype=out_dtype)
  data = np.asfortranarray(data)
  if data.dtype.kind not in "biu":
    raise TypeError("Type {} not currently supported.".format(data.dtype))
  src = _view_as_unsigned(data)
  shape = tuple(src.shape) + (1,) * (3 - dims)
  graph = np.zeros(src.shape, dtype=out_dtype, order="F")
  _lib.check(L.cc3d_b200_voxel_connectivity_graph(
    src.ctypes.data, _kind_of(src.dtype), shape[0], shape[1], shape[2], int(connectivity), graph.ctypes.data,
    _lib.HOST, None))
  return graph


def color_connectivity_graph(vcg, connectivity: int = 26, return_N: bool = False):
  """Labels the components of a voxel connectivity graph; same contract as cc3d.color_connectivity_graph
  (fastcc3d.pyx:941-1018): uint32 labels, every voxel labelled, numbered by first appearance in Fortran order."""
  L = _lib.lib()
  vcg = _adopt_device_array(vcg)
  if _is_torch(vcg):
    if vcg.is_cuda:
      return _color_connectivity_graph_device(vcg, connectivity, return_N)
    import torch
    out = color_connectivity_graph(vcg.numpy(), connectivity, return_N)
    if return_N:
      return torch.from_numpy(out[0].view(np.int32)).view(torch.uint32), out[1]
    return torch.from_numpy(out.view(np.int32)).view(torch.uint32)
  dims = len(vcg.shape)
  if dims not in (2, 3):
    raise DimensionError("Only 2D, and 3D arrays supported. Got: " + str(dims))
  if dims == 2 and connectivity not in [4, 8, 6, 26]:
    raise ValueError(f"Only 4 and 8 connectivity is supported for 2D images. Got: {connectivity}")
  elif dims != 2 and connectivity not in [6, 26]:
    raise ValueError(f"Only 6 and 26 connectivity are supported for 3D images. Got: {connectivity}")
  if vcg.dtype not in [np.uint8, np.uint32]:
    raise ValueError(f"Only uint8 and uint32 are supported. Got: {vcg.dtype}")
  if vcg.size == 0:
    return np.zeros([0] * dims, dtype=np.uint32, order="F")
  while vcg.ndim < 3:
    vcg = vcg[..., np.newaxis]
  vcg = np.asfortranarray(vcg)
  sx, sy, sz = vcg.shape
  if connectivity in [18, 26] and sz > 1 and vcg.dtype != np.uint32:
    raise ValueError(f"Only uint32 is supported for 18 and 26 connected. Got: {vcg.dtype}")
  out_labels = np.zeros((sx, sy, sz), dtype=np.uint32, order="F")
  N = ctypes.c_uint64(0)
  _lib.check(L.cc3d_b200_color_connectivity_graph(
    vcg.ctypes.data, _kind_of(vcg.dtype), sx, sy, sz, int(connectivity), out_labels.ctypes.data, ctypes.byref(N),
    _lib.HOST, None))
  while out_labels.ndim > dims:
    out_labels = out_labels[..., 0]
  if return_N:
    return (out_labels, int(N.value))
  return out_labels


def _fortran_flat_device(t):
  """(flat tensor in Fortran memory order, i.e. first axis fastest; (sx, sy, sz)) of a 2-D / 3-D CUDA tensor - a view
  when the tensor already is Fortran-ordered, else one transposing copy on the device."""
  shape = list(t.shape) + [1] * (3 - t.ndim)
  rev = t.permute(*reversed(range(t.ndim)))       # (.., y, x): C-contiguous iff t is Fortran-ordered
  return rev.contiguous().reshape(-1), tuple(shape)


def _color_connectivity_graph_device(vcg, connectivity, return_N):
  """color_connectivity_graph of a CUDA tensor without leaving the device (zero copy for Fortran-ordered graphs)."""
  import torch
  L = _lib.lib()
  dims = vcg.ndim
  if dims not in (2, 3):
    raise DimensionError("Only 2D, and 3D arrays supported. Got: " + str(dims))
  if dims == 2 and connectivity not in [4, 8, 6, 26]:
    raise ValueError(f"Only 4 and 8 connectivity is supported for 2D images. Got: {connectivity}")
  elif dims != 2 and connectivity not in [6, 26]:
    raise ValueError(f"Only 6 and 26 connectivity are supported for 3D images. Got: {connectivity}")
  np_dt = _torch_np_dtype(vcg)
  if np_dt.kind == "i" and np_dt.itemsize in (1, 4):
    np_dt = np.dtype(_UNSIGNED[np_dt.itemsize])     # torch spells uint32 graphs as int32 more often than not
  if np_dt not in (np.uint8, np.uint32):
    raise ValueError(f"Only uint8 and uint32 are supported. Got: {np_dt}")
  if vcg.numel() == 0:
    return torch.zeros([0] * dims, dtype=torch.uint32, device=vcg.device)
  flat, (sx, sy, sz) = _fortran_flat_device(vcg.detach())
  if connectivity in [18, 26] and sz > 1 and np_dt != np.uint32:
    raise ValueError(f"Only uint32 is supported for 18 and 26 connected. Got: {np_dt}")
  out = torch.empty((sx * sy * sz,), dtype=torch.uint32, device=vcg.device)
  N = ctypes.c_uint64(0)
  with torch.cuda.device(vcg.device):
    _lib.check(L.cc3d_b200_color_connectivity_graph(
      flat.data_ptr(), _kind_of(np_dt), sx, sy, sz, int(connectivity), out.data_ptr(), ctypes.byref(N), _lib.DEVICE,
      ctypes.c_void_p(torch.cuda.current_stream(vcg.device).cuda_stream)))
  out = out.reshape(tuple(reversed(vcg.shape))).permute(*reversed(range(dims)))      # Fortran-ordered like the reference's
  return (out, int(N.value)) if return_N else out


def _contacts_wide_labels(labels, connectivity, surface_area, anisotropy):
  """contacts for uint64 label VALUES >= 2^32 (the pair keys of the kernel pack two 32-bit labels): the labels are
  replaced by their ranks among the distinct values (order preserving, 0 stays 0), the ranks go through the kernel and
  the pairs are mapped back. cc3d_graphs.hpp:300-468 handles any label width."""
  uniq, inv = np.unique(labels, return_inverse=True)
  inv = np.asarray(inv).reshape(labels.shape)
  shift = 0 if (uniq.size and uniq[0] == 0) else 1          # rank 0 must mean background
  ranks = np.asfortranarray((inv + shift).astype(np.uint32))
  res = contacts(ranks, connectivity=connectivity, surface_area=surface_area, anisotropy=anisotropy)
  return {(int(uniq[a - shift]), int(uniq[b - shift])): v for (a, b), v in res.items()}


def contacts(labels, connectivity: int = 26, surface_area: bool = True, anisotropy=(1, 1, 1)) -> dict:
  """Region adjacency graph with contact areas; same contract as cc3d.contacts (fastcc3d.pyx:1196-1252):
  {(label_1, label_2): float} with label_1 < label_2. The GPU counts the contacts of every pair per direction
  class exactly; the value is count x face area (float32 like the reference, which adds areas one by one: the two
  agree whenever the reference's running sum stays exactly representable, e.g. integer areas below 2^24).
  Label values must be < 2^32."""
  L = _lib.lib()
  labels = _adopt_device_array(labels)
  device_flat = None
  if _is_torch(labels):
    if labels.is_cuda and labels.ndim in (2, 3) and _torch_np_dtype(labels).kind in "iu" and labels.numel():
      # the label volume stays on the device; only the pair table comes back
      device_flat, dshape = _fortran_flat_device(labels.detach())
      dkind = _kind_of(_torch_np_dtype(labels))
    labels = (labels[:1].cpu() if device_flat is not None else labels.cpu()).numpy()   # dtype / rank checks below
  labels = np.asarray(labels)
  while labels.ndim < 3:
    labels = labels[..., np.newaxis]
  anisotropy = tuple(anisotropy)
  while len(anisotropy) < 3:
    anisotropy = anisotropy + (1,)
  if connectivity not in (4, 8, 6, 18, 26):
    raise ValueError(f"Only (2d) 4, 8, (3d) 6, 18, and 26 connectivities are supported. Got: {connectivity}")
  if labels.dtype.kind not in "biu":
    raise TypeError("Type {} not currently supported.".format(labels.dtype))
  labels = np.asfortranarray(_view_as_unsigned(labels))
  if labels.dtype == bool:
    labels = labels.view(np.uint8)
  sx, sy, sz = labels.shape if device_flat is None else dshape
  if connectivity in (4, 8) and sz != 1:
    raise RuntimeError("z thickness must be 1 for 2d region graph extraction.")
  if labels.size == 0:
    return {}
  cap = 1 << 16
  while True:
    n = ctypes.c_uint64(0)
    try:
      if device_flat is not None:
        import torch
        dkeys = torch.empty((cap,), dtype=torch.int64, device=device_flat.device)
        dvals = torch.empty((cap, 4), dtype=torch.int32, device=device_flat.device)
        with torch.cuda.device(device_flat.device):
          _lib.check(L.cc3d_b200_contacts(device_flat.data_ptr(), dkind, sx, sy, sz, int(connectivity), dkeys.data_ptr(),
                                          dvals.data_ptr(), cap, ctypes.byref(n), _lib.DEVICE,
                                          ctypes.c_void_p(torch.cuda.current_stream(device_flat.device).cuda_stream)))
        if n.value <= cap:
          keys, vals = dkeys[: n.value].cpu().numpy().view(np.uint64), dvals[: n.value].cpu().numpy().view(np.uint32)
      else:
        keys = np.empty(cap, dtype=np.uint64)
        vals = np.empty((cap, 4), dtype=np.uint32)
        _lib.check(L.cc3d_b200_contacts(labels.ctypes.data, _kind_of(labels.dtype), sx, sy, sz, int(connectivity),
                                        keys.ctypes.data, vals.ctypes.data, cap, ctypes.byref(n), _lib.HOST, None))
    except CC3DB200Error as e:
      if e.code != -5 or "2^32" not in str(e):
        raise
      # label values >= 2^32: go through order-preserving ranks (host arrays; device tensors come to the host for it)
      wide = labels if device_flat is None else np.asfortranarray(
        device_flat.cpu().numpy().view(_UNSIGNED[device_flat.element_size()]).reshape((sx, sy, sz), order="F"))
      return _contacts_wide_labels(wide, connectivity, surface_area, anisotropy)
    if n.value <= cap:
      break
    cap = 1 << int(n.value - 1).bit_length()
  keys, vals = keys[: n.value], vals[: n.value]
  wx, wy, wz = (np.float32(a) for a in anisotropy)
  if connectivity in (4, 8):
    areas = [wy, wx, np.float32(0), np.float32(0)] if surface_area else [np.float32(1)] * 4
  else:
    areas = [wy * wz, wx * wz, wx * wy, np.float32(0)] if surface_area else [np.float32(1)] * 4
  total = np.zeros(keys.size, dtype=np.float64)
  for c in range(4):
    total += vals[:, c].astype(np.float64) * float(areas[c])
  total = total.astype(np.float32)
  lo, hi = (keys >> np.uint64(32)).tolist(), (keys & np.uint64(0xFFFFFFFF)).tolist()
  return {(a, b): float(v) for a, b, v in zip(lo, hi, total)}


def region_graph(labels, connectivity: int = 26) -> set:
  """Set of label pairs that touch; same contract as cc3d.region_graph (fastcc3d.pyx:1180-1194)."""
  return set(contacts(labels, connectivity=connectivity).keys())


# ----------------------------------------------------------------------------------------------
# runs / draw / erase / each (fastcc3d.pyx:1258-1372, cc3d_graphs.hpp:470-523)
# ----------------------------------------------------------------------------------------------
_RUN_KINDS = (np.dtype(np.uint8), np.dtype(np.uint16), np.dtype(np.uint32), np.dtype(np.uint64))


def _flat_memory_order(arr: np.ndarray) -> np.ndarray:
  """1-D view of the array in memory order (reference `_reshape(arr, (arr.size,))`, fastcc3d.pyx:129-161):
  Fortran or C contiguous arrays are viewed in place, anything else is copied to C order."""
  if arr.flags.f_contiguous:
    return arr.reshape(-1, order="F")
  if arr.flags.c_contiguous:
    return arr.reshape(-1)
  return arr.reshape((arr.size,))


def _run_dtype(dtype) -> np.dtype:
  dtype = np.dtype(dtype)
  if dtype == np.bool_:
    return np.dtype(np.uint8)
  if dtype not in _RUN_KINDS:
    raise TypeError("Unsupported type: " + str(dtype))
  return dtype


def _torch_flat(t):
  """(dense 1-D tensor in memory order, order) of a CUDA tensor."""
  t, order = _torch_order(t)
  if order == "F":
    return t.permute(*reversed(range(t.ndim))).reshape(-1), order
  return t.reshape(-1), order


def _run_table_device(flat_t, np_dtype):
  """(values, starts, ends) int64 CUDA tensors (bit patterns of uint64) of the runs of a flat CUDA tensor, in
  position order; two calls at most: the first one with a guessed capacity."""
  import torch
  L = _lib.lib()
  n = flat_t.numel()
  kind = _kind_of(_run_dtype(np_dtype))
  count = ctypes.c_uint64(0)
  cap = int(min(n, max(1 << 16, n // 16)))
  with torch.cuda.device(flat_t.device):
    stream = ctypes.c_void_p(torch.cuda.current_stream(flat_t.device).cuda_stream)
    while True:
      tab = torch.empty((3, max(cap, 1)), dtype=torch.int64, device=flat_t.device)
      _lib.check(L.cc3d_b200_runs(flat_t.data_ptr(), kind, n, tab[0].data_ptr(), tab[1].data_ptr(), tab[2].data_ptr(),
                                  cap, ctypes.byref(count), _lib.DEVICE, stream))
      if count.value <= cap:
        break
      cap = int(count.value)
  k = int(count.value)
  return tab[0, :k], tab[1, :k], tab[2, :k]


def _stable_argsort_u64(values: np.ndarray) -> np.ndarray:
  """Stable argsort of uint64 label values. numpy radix-sorts 16-bit keys only, so values that fit 16 (32) bits -
  the usual case: labels 1..N - are sorted in one (two) least-significant-digit passes over uint16 digits."""
  if values.size == 0:
    return np.zeros(0, np.int64)
  vmax = int(values.max())
  if vmax < (1 << 16):
    return np.argsort(values.astype(np.uint16), kind="stable")
  if vmax < (1 << 32):
    o1 = np.argsort(values.astype(np.uint16), kind="stable")   # astype keeps the low 16 bits
    hi = (values >> np.uint64(16)).astype(np.uint16)[o1]
    return o1[np.argsort(hi, kind="stable")]
  return np.argsort(values, kind="stable")


def _group_runs(values: np.ndarray, starts: np.ndarray, ends: np.ndarray):
  """Run table in position order -> (labels ascending, offsets, starts, ends grouped by label, position order kept):
  the iteration order of the reference's std::map<T, vector<pair>> (cc3d_graphs.hpp:472-474)."""
  order = _stable_argsort_u64(values)
  v = values[order]
  first = np.concatenate(([0], np.flatnonzero(v[1:] != v[:-1]) + 1)) if v.size else np.zeros(0, np.int64)
  labels = v[first]
  offsets = np.append(first, v.size).astype(np.int64)
  return labels, offsets, starts[order], ends[order]


def _runs_table(labels):
  """Grouped run table of a numpy array or CUDA tensor: (label values, offsets, starts, ends) as numpy uint64 /
  int64 arrays. The extraction runs on the GPU (cc3d_b200_runs); grouping by label is host bookkeeping."""
  L = _lib.lib()
  if _is_torch(labels) and labels.is_cuda:
    flat, _ = _torch_flat(labels)
    if flat.numel() == 0:
      raise IndexError("Out of bounds on buffer access (axis 0)")   # the reference indexes labels[0] (fastcc3d.pyx:1269)
    v, s, e = _run_table_device(flat, _torch_np_dtype(labels))
    values, starts, ends = (x.cpu().numpy().view(np.uint64) for x in (v, s, e))
    n, first = flat.numel(), (int(flat[0].item()) if flat.numel() == 1 else None)
  else:
    labels = np.asarray(labels.cpu().numpy() if _is_torch(labels) else labels)
    dt = _run_dtype(labels.dtype)
    flat = _flat_memory_order(labels)
    if flat.dtype != dt:
      flat = flat.view(dt)
    n = flat.size
    if n == 0:
      raise IndexError("Out of bounds on buffer access (axis 0)")   # the reference indexes labels[0] (fastcc3d.pyx:1269)
    first = int(flat[0]) if n == 1 else None
    flat = np.ascontiguousarray(flat)
    count = ctypes.c_uint64(0)
    cap = int(min(n, max(1 << 16, n // 16)))
    while True:
      tab = np.empty((3, max(cap, 1)), dtype=np.uint64)
      _lib.check(L.cc3d_b200_runs(flat.ctypes.data, _kind_of(dt), n, tab[0].ctypes.data, tab[1].ctypes.data,
                                  tab[2].ctypes.data, cap, ctypes.byref(count), _lib.HOST, None))
      if count.value <= cap:
        break
      cap = int(count.value)
    k = int(count.value)
    values, starts, ends = tab[0, :k], tab[1, :k], tab[2, :k]
  if n == 1 and first == 0:
    # a single-voxel array reports its run even when it is background (cc3d_graphs.hpp:481-484)
    values, starts, ends = np.zeros(1, np.uint64), np.zeros(1, np.uint64), np.ones(1, np.uint64)
  return _group_runs(values, starts, ends)


def runs(labels) -> dict:
  """Returns a dictionary describing where each label is located: {label: [(start, end), ...]} over the
  flattened (memory order) array, half-open voxel ranges; same contract as cc3d.fastcc3d.runs
  (fastcc3d.pyx:1258-1279 -> extract_runs, cc3d_graphs.hpp:470-503). Keys ascend, runs ascend by position."""
  import gc
  lab, off, starts, ends = _runs_table(labels)
  off = off.tolist()
  # millions of small tuples: the cyclic collector would rescan them over and over while they are being built
  gc_was_on = gc.isenabled()
  gc.disable()
  try:
    pairs = list(zip(starts.tolist(), ends.tolist()))
    return {l: pairs[off[i]:off[i + 1]] for i, l in enumerate(lab.tolist())}
  finally:
    if gc_was_on:
      gc.enable()


def _label_for_image(label, np_dtype) -> int:
  np_dtype = np.dtype(np_dtype)
  if np_dtype == np.bool_:
    return int(label != 0)
  v = int(label)
  if v < 0 or v > int(np.iinfo(np_dtype).max):
    raise OverflowError(f"value {v} does not fit {np_dtype}")   # Cython's conversion to the image type raises too
  return v


def _runs_as_arrays(rns):
  arr = np.asarray(rns, dtype=np.uint64) if len(rns) else np.zeros((0, 2), np.uint64)
  arr = arr.reshape(-1, 2)
  return np.ascontiguousarray(arr[:, 0]), np.ascontiguousarray(arr[:, 1])


def draw(label, runs, image):
  """Draws label onto the provided image according to runs (fastcc3d.pyx:1281-1314 -> set_run_voxels,
  cc3d_graphs.hpp:505-523). Returns the flattened (memory order) image like the reference. numpy images are staged
  through the device (only the window the runs span); CUDA tensors are drawn in place. An invalid run raises
  RuntimeError("Invalid run.") and leaves the image untouched."""
  L = _lib.lib()
  starts, ends = _runs_as_arrays(runs)
  if _is_torch(image) and image.is_cuda:
    import torch
    np_dt = _torch_np_dtype(image)
    flat, _ = _torch_flat(image)
    if flat.data_ptr() != image.data_ptr() and image.numel():
      raise ValueError("draw: CUDA image must be dense in memory (C or Fortran contiguous)")
    value = _label_for_image(label, np_dt)
    if starts.size:
      dev = image.device
      st = torch.from_numpy(starts.view(np.int64)).to(dev)
      en = torch.from_numpy(ends.view(np.int64)).to(dev)
      with torch.cuda.device(dev):
        _lib.check(L.cc3d_b200_draw(flat.data_ptr(), _kind_of(_run_dtype(np_dt)), flat.numel(), value, st.data_ptr(),
                                    en.data_ptr(), starts.size, _lib.DEVICE,
                                    ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return flat
  dt = _run_dtype(image.dtype)
  flat = _flat_memory_order(image)
  value = _label_for_image(label, image.dtype)
  if starts.size:
    target = flat if flat.flags.c_contiguous else np.ascontiguousarray(flat)
    _lib.check(L.cc3d_b200_draw(target.view(dt).ctypes.data, _kind_of(dt), target.size, value, starts.ctypes.data,
                                ends.ctypes.data, starts.size, _lib.HOST, None))
    if target is not flat:
      flat[...] = target
  return flat


def erase(runs, image):
  """Erases (sets to 0) part of the provided image according to runs (fastcc3d.pyx:1316-1326)."""
  return draw(0, runs, image)


_erase = erase


def each(labels, binary: bool = False, in_place: bool = False):
  """Returns an iterator that extracts each label from a dense labeling; same contract as cc3d.each
  (fastcc3d.pyx:1328-1372): yields (label, image) for every non-zero label in ascending order, image = the label's
  mask (bool when `binary`, else the label value in the labels' dtype) with the labels' shape and memory order.
  in_place: one image is reused (read-only while it is out). The run table is extracted once on the GPU and stays
  there; every image is rendered by the draw kernels. CUDA tensors give CUDA tensors (no host traffic); for numpy
  input only the window of the image that the label's runs span crosses PCIe."""
  is_t = _is_torch(labels) and labels.is_cuda
  if not is_t:
    labels = np.asarray(labels.cpu().numpy() if _is_torch(labels) else labels)
    np_dt = labels.dtype
    order = "F" if labels.flags.f_contiguous else "C"
    shape = labels.shape
  else:
    import torch
    np_dt = _torch_np_dtype(labels)
    _, order = _torch_flat(labels)
    shape = tuple(labels.shape)
  _run_dtype(np_dt)
  lab, off, starts, ends = _runs_table(labels)
  img_dt = np.dtype(np.bool_) if binary else np.dtype(np_dt)
  kind = _kind_of(_run_dtype(img_dt))
  n_vox = int(np.prod(shape)) if len(shape) else 1
  nonzero = [i for i, l in enumerate(lab.tolist()) if l != 0]
  L = _lib.lib()
  if not is_t:
    # numpy input: torch is not needed. Every image is a zeroed numpy array drawn by cc3d_b200_draw on HOST buffers
    # (the library stages only the window that the label's runs span).
    hdt = _run_dtype(img_dt)
    hkind = _kind_of(hdt)

    def host_draw(img, i, value):
      a, b = int(off[i]), int(off[i + 1])
      flat = img.reshape(-1, order=order).view(hdt)
      _lib.check(L.cc3d_b200_draw(flat.ctypes.data, hkind, n_vox, value, starts[a:b].ctypes.data, ends[a:b].ctypes.data,
                                  b - a, _lib.HOST, None))

    class HostImageIterator:
      def __len__(self):
        return len(nonzero)

      def __iter__(self):
        for i in nonzero:
          key = int(lab[i])
          img = np.zeros(shape, dtype=img_dt, order=order)
          host_draw(img, i, _label_for_image(key, img_dt))
          yield key, img

    class HostInPlaceImageIterator(HostImageIterator):
      def __iter__(self):
        img = np.zeros(shape, dtype=img_dt, order=order)
        for i in nonzero:
          key = int(lab[i])
          host_draw(img, i, _label_for_image(key, img_dt))
          img.setflags(write=0)
          yield key, img
          img.setflags(write=1)
          host_draw(img, i, 0)

    return HostInPlaceImageIterator() if in_place else HostImageIterator()

  dev = labels.device
  state = {}

  def table():
    if "st" not in state:   # grouped run table, uploaded once
      state["st"] = torch.from_numpy(starts.view(np.int64)).to(dev)
      state["en"] = torch.from_numpy(ends.view(np.int64)).to(dev)
    return state["st"], state["en"]

  def device_image():
    # zero-filled through the signed type of the same width (fill kernels exist for every signed type)
    tdt = torch.uint8 if binary else (labels.dtype if is_t else _torch_dtype(_run_dtype(np_dt)))
    signed = {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[img_dt.itemsize]
    z = torch.zeros(n_vox, dtype=signed, device=dev)
    return z if z.dtype == tdt else z.view(tdt)

  def render(flat_img, i, value):
    st, en = table()
    a, b = int(off[i]), int(off[i + 1])
    with torch.cuda.device(dev):
      _lib.check(L.cc3d_b200_draw(flat_img.data_ptr(), kind, n_vox, value, st[a:b].data_ptr(), en[a:b].data_ptr(),
                                  b - a, _lib.DEVICE, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))

  def shaped(flat_img):
    t = flat_img.view(torch.bool) if binary else flat_img
    if order == "F" and len(shape) > 1:
      return t.view(*reversed(shape)).permute(*reversed(range(len(shape))))
    return t.view(*shape)

  def window(i):
    a, b = int(off[i]), int(off[i + 1])
    return int(starts[a:b].min()), int(ends[a:b].max())

  def host_flat(img):
    f = img.reshape(-1, order=order)
    return f.view(np.uint8) if binary else f.view(_run_dtype(np_dt))

  def fetch(flat_img, host_img, lo, hi):   # device window -> the same window of the numpy image
    src = flat_img[lo:hi]
    dst = torch.from_numpy(host_flat(host_img)[lo:hi])
    if dst.dtype != src.dtype:
      src = src.view(dst.dtype)
    dst.copy_(src)

  class ImageIterator:
    def __len__(self):
      return len(nonzero)

    def __iter__(self):
      for i in nonzero:
        key = int(lab[i])
        value = _label_for_image(key, img_dt)
        dimg = device_image()
        render(dimg, i, value)
        if is_t:
          yield key, shaped(dimg)
        else:
          img = np.zeros(shape, dtype=img_dt, order=order)
          lo, hi = window(i)
          fetch(dimg, img, lo, hi)
          yield key, img

  class InPlaceImageIterator(ImageIterator):
    def __iter__(self):
      dimg = device_image()
      img = None if is_t else np.zeros(shape, dtype=img_dt, order=order)
      for i in nonzero:
        key = int(lab[i])
        render(dimg, i, _label_for_image(key, img_dt))
        if is_t:
          yield key, shaped(dimg)
          render(dimg, i, 0)
        else:
          lo, hi = window(i)
          fetch(dimg, img, lo, hi)
          img.setflags(write=0)
          yield key, img
          img.setflags(write=1)
          render(dimg, i, 0)
          fetch(dimg, img, lo, hi)

  return InPlaceImageIterator() if in_place else ImageIterator()




def connected_components_stack(stacked_images, connectivity: int = 26, return_N: bool = False,
                               binary_image: bool = False, out_dtype=None, out=None, scratch_dir=None, order=None):
  """Streaming (out-of-GPU-memory) labelling of an iterable of z-slabs; see sharded.connected_components_stack
  (counterpart of cc3d.connected_components_stack, cc3d/__init__.py:353-501)."""
  from .sharded import connected_components_stack as _stack
  return _stack(stacked_images, connectivity=connectivity, return_N=return_N, binary_image=binary_image,
                out_dtype=out_dtype, out=out, scratch_dir=scratch_dir, order=order)


try:   # the compiled Cython boundary (fastcc3d.pyx, built by build.py); also gives cc3d.fastcc3d.runs / draw like the reference
  from . import fastcc3d  # noqa: E402
except ImportError:   # not built yet: everything else works through ctypes on the same C-ABI
  fastcc3d = None
