"""oracle.py — TEST INFRASTRUCTURE ONLY. Never imported by the product package.

Python front end of the CPU oracle: a ctypes binding of oracle/cc3d_oracle.c (plain-C restatement
of the reference's two-pass union-find labelling) plus numpy restatements of the reference's
boundary logic, `statistics` and `dust`:

  connected_components  <- fastcc3d.pyx:245-626 (validation, layout normalisation, out-dtype rule A.2)
  estimate_provisional_labels <- fastcc3d.pyx:169-242 / cc3d.hpp:287-315
  statistics            <- fastcc3d.pyx:682-938
  dust                  <- cc3d/__init__.py:71-155
  largest_k             <- cc3d/__init__.py:199-279 (the path without fastremap)
  voxel_connectivity_graph <- fastcc3d.pyx:1021-1170 / cc3d_graphs.hpp:31-247 (numpy slicing)
  color_connectivity_graph <- fastcc3d.pyx:941-1018 / cc3d_graphs.hpp:583-1106 (backward-bit graph, scipy components)
  contacts / region_graph <- fastcc3d.pyx:1180-1252 / cc3d_graphs.hpp:259-468 (compute_neighborhood offsets incl. borders)
  runs / draw / erase / each <- fastcc3d.pyx:1258-1372 / cc3d_graphs.hpp:470-523 (numpy change-point restatement)

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


def reference_package():
  """The reference's Python layer (cc3d/__init__.py: dust, largest_k, ...) executed IN PLACE from
  /root/reference on top of the built extension, or None when either is absent (e.g. on the GPU box).
  Nothing is copied: the source is read where it lies and run as the package `cc3d_ref`."""
  ext = reference_module()
  src = "/root/reference/cc3d/__init__.py"
  if ext is None or not os.path.isfile(src):
    return None
  if "cc3d_ref" in sys.modules:
    return sys.modules["cc3d_ref"]
  import types
  pkg = types.ModuleType("cc3d_ref")
  pkg.__path__ = []
  pkg.__package__ = "cc3d_ref"
  pkg.__file__ = src
  sys.modules["cc3d_ref"] = pkg
  sys.modules["cc3d_ref.fastcc3d"] = ext
  try:
    exec(compile(open(src).read(), src, "exec"), pkg.__dict__)
  except Exception:
    del sys.modules["cc3d_ref"], sys.modules["cc3d_ref.fastcc3d"]
    return None
  return pkg


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


def largest_k(img, k, connectivity=26, delta=0, return_N=False, binary_image=False, precomputed_ccl=False):
  """cc3d/__init__.py:199-279, fallback branch (no fastremap): kept labels are redrawn as 1..k, smallest first."""
  assert k >= 0
  order = "C" if img.flags.c_contiguous else "F"
  if k == 0:
    return np.zeros(img.shape, dtype=np.uint16, order=order)
  if precomputed_ccl:
    cc_labels = np.copy(img, order="F")
    N = int(np.max(cc_labels))
  else:
    cc_labels, N = connected_components(img, connectivity=connectivity, return_N=True, delta=delta,
                                        binary_image=bool(binary_image))
  if N <= k:
    return (cc_labels, N) if return_N else cc_labels
  cts = statistics(cc_labels, no_slice_conversion=True)["voxel_counts"]
  if k == 1:
    cc_out = cc_labels == (np.argmax(cts[1:]) + 1)
    return (cc_out, 1) if return_N else cc_out
  preserve = np.argpartition(cts[1:], len(cts) - k - 1)[-k:]
  preserve += 1
  preserve_list = [int(l) for l in sorted(preserve, key=lambda label: cts[label])]
  table = np.zeros(N + 1, dtype=cc_labels.dtype)
  for i, label in enumerate(preserve_list):
    table[label] = i + 1
  cc_out = table[cc_labels]
  cc_out = np.asarray(cc_out, order="C" if cc_labels.flags.c_contiguous else "F")
  return (cc_out, len(preserve_list)) if return_N else cc_out


# direction tables in the reference's bit order (cc3d_graphs.hpp:91-110 and :31-76)
_VCG_DIR3 = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1),
             (1, 1, 0), (-1, 1, 0), (1, -1, 0), (-1, -1, 0), (1, 0, 1), (-1, 0, 1), (0, 1, 1), (0, -1, 1),
             (1, 0, -1), (-1, 0, -1), (0, 1, -1), (0, -1, -1),
             (1, 1, 1), (-1, 1, 1), (1, -1, 1), (-1, -1, 1), (1, 1, -1), (-1, 1, -1), (1, -1, -1), (-1, -1, -1)]
_VCG_DIR2 = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (1, 1, 0), (-1, 1, 0), (1, -1, 0), (-1, -1, 0)]


def _shifted_pairs(shape, d):
  """Slices (p, q) with q = p + d for every voxel p whose neighbour q lies inside the array."""
  sp, sq = [], []
  for n, k in zip(shape, d):
    if k == 0:
      sp.append(slice(0, n)); sq.append(slice(0, n))
    elif k > 0:
      sp.append(slice(0, n - 1)); sq.append(slice(1, n))
    else:
      sp.append(slice(1, n)); sq.append(slice(0, n - 1))
  return tuple(sp), tuple(sq)


def voxel_connectivity_graph(data, connectivity=26):
  dims = data.ndim
  if dims == 2 and connectivity not in (4, 8, 6, 18, 26):
    raise ValueError("Only 4, 8, and 6, 18, 26 connectivities are supported for 2D images. Got: " + str(connectivity))
  elif dims != 2 and connectivity not in (6, 18, 26):
    raise ValueError("Only 6, 18, and 26 connectivities are supported for 3D images. Got: " + str(connectivity))
  out_dtype = np.uint8 if connectivity in (4, 8, 6) else np.uint32
  if data.size == 0:
    return np.zeros(shape=(0,), dtype=out_dtype)
  x = np.asfortranarray(data)
  while x.ndim < 3:
    x = x[..., np.newaxis]
  if connectivity in (4, 8):
    if x.shape[2] != 1:
      raise RuntimeError("sz must be 1 for 2D connectivities.")
    dirs = _VCG_DIR2[:connectivity]
  else:
    dirs = _VCG_DIR3[:connectivity]
  graph = np.full(x.shape, (1 << len(dirs)) - 1, dtype=np.uint32, order="F")
  for b, d in enumerate(dirs):
    sp, sq = _shifted_pairs(x.shape, d)
    graph[sp] &= np.where(x[sp] != x[sq], np.uint32(~(1 << b) & 0xFFFFFFFF), np.uint32(0xFFFFFFFF))
  return np.asfortranarray(graph.astype(out_dtype)).reshape(data.shape, order="F")


def color_connectivity_graph(vcg, connectivity=26, return_N=False):
  from scipy.sparse import coo_matrix
  from scipy.sparse.csgraph import connected_components as graph_cc
  dims = vcg.ndim
  if dims == 2 and connectivity not in [4, 8, 6, 26]:
    raise ValueError(f"Only 4 and 8 connectivity is supported for 2D images. Got: {connectivity}")
  elif dims != 2 and connectivity not in [6, 26]:
    raise ValueError(f"Only 6 and 26 connectivity are supported for 3D images. Got: {connectivity}")
  if vcg.dtype not in [np.uint8, np.uint32]:
    raise ValueError(f"Only uint8 and uint32 are supported. Got: {vcg.dtype}")
  if vcg.size == 0:
    return np.zeros([0] * dims, dtype=np.uint32, order="F")
  g = np.asfortranarray(vcg)
  while g.ndim < 3:
    g = g[..., np.newaxis]
  sx, sy, sz = g.shape
  # backward edges the reference follows: (direction, bit number); 2D graphs have two layouts
  edges = [((-1, 0, 0), 2), ((0, -1, 0), 4)]
  if sz == 1:
    if connectivity in (8, 26):
      edges += [((-1, -1, 0), 8), ((1, -1, 0), 7)] if g.dtype == np.uint8 else [((-1, -1, 0), 10), ((1, -1, 0), 9)]
  else:
    edges.append(((0, 0, -1), 6))
    if connectivity == 26:
      if g.dtype != np.uint32:
        raise ValueError(f"Only uint32 is supported for 18 and 26 connected. Got: {g.dtype}")
      edges += [((-1, -1, 0), 10), ((1, -1, 0), 9), ((-1, 0, -1), 16), ((1, 0, -1), 15), ((0, -1, -1), 18),
                ((0, 1, -1), 17), ((-1, -1, -1), 26), ((1, -1, -1), 25), ((-1, 1, -1), 24), ((1, 1, -1), 23)]
  idx = np.arange(g.size, dtype=np.int64).reshape(g.shape, order="F")
  ia, ib = [], []
  for d, bit in edges:
    sp, sq = _shifted_pairs(g.shape, d)
    m = (g[sp].astype(np.uint32) & np.uint32(1 << (bit - 1))) != 0
    ia.append(idx[sp][m]); ib.append(idx[sq][m])
  ia, ib = np.concatenate(ia), np.concatenate(ib)
  n = g.size
  _, lab = graph_cc(coo_matrix((np.ones(ia.size, dtype=np.int8), (ia, ib)), shape=(n, n)), directed=False)
  # number by first appearance in Fortran (memory) order
  _, first = np.unique(lab, return_index=True)
  rank = np.empty(first.size, dtype=np.int64)
  rank[np.argsort(first)] = np.arange(1, first.size + 1)
  out = rank[lab].astype(np.uint32).reshape(g.shape, order="F")
  while out.ndim > dims:
    out = out[..., 0]
  return (out, int(first.size)) if return_N else out


def contacts(labels, connectivity=26, surface_area=True, anisotropy=(1, 1, 1)):
  """cc3d_graphs.hpp:300-468 with compute_neighborhood (:259-313) restated on index arrays (vectorised numpy);
  float32 accumulation in raster order per pair like the reference (np.add.at on float32 is sequential)."""
  labels = np.asarray(labels)
  while labels.ndim < 3:
    labels = labels[..., np.newaxis]
  anisotropy = tuple(anisotropy)
  while len(anisotropy) < 3:
    anisotropy = anisotropy + (1,)
  if connectivity not in (4, 8, 6, 18, 26):
    raise ValueError(f"Only (2d) 4, 8, (3d) 6, 18, and 26 connectivities are supported. Got: {connectivity}")
  lab = np.asfortranarray(labels)
  if np.issubdtype(lab.dtype, np.signedinteger):
    lab = lab.view(f"u{lab.dtype.itemsize}")
  sx, sy, sz = lab.shape
  if connectivity in (4, 8) and sz != 1:
    raise RuntimeError("z thickness must be 1 for 2d region graph extraction.")
  flat = lab.reshape(-1, order="F").astype(np.uint64)
  n = flat.size
  if n == 0:
    return {}
  loc = np.arange(n, dtype=np.int64)
  x, y, z = loc % sx, (loc // sx) % sy, loc // (sx * sy)
  px, mx = (x < sx - 1).astype(np.int64), -(x > 0).astype(np.int64)
  py, my = sx * (y < sy - 1).astype(np.int64), -sx * (y > 0).astype(np.int64)
  mz = -sx * sy * (z > 0).astype(np.int64)
  wx, wy, wz = (np.float32(a) for a in anisotropy)
  if connectivity in (4, 8):
    offs = [mx, my, (connectivity > 4) * (mx + my), (connectivity > 4) * (px + my)][: connectivity // 2]
    areas = [wy, wx, np.float32(0), np.float32(0)] if surface_area else [np.float32(1)] * 4
  else:
    b = lambda a: (a != 0).astype(np.int64)
    offs = [mx, my, mz,
            (mx + my) * (b(mx) & b(my)), (px + my) * (b(px) & b(my)), (mx + mz) * (b(mx) & b(mz)), (px + mz) * (b(px) & b(mz)),
            (my + mz) * (b(my) & b(mz)), (py + mz) * (b(py) & b(mz)),
            (mx + my + mz) * (b(my) & b(mz)), (px + my + mz) * (b(my) & b(mz)), (mx + py + mz) * (b(py) & b(mz)),
            (px + py + mz) * (b(py) & b(mz))][: connectivity // 2]
    areas = ([wy * wz, wx * wz, wx * wy] + [np.float32(0)] * 10) if surface_area else [np.float32(1)] * 13
  # contacts in the reference's order: voxel by voxel (raster), direction by direction
  recs = []
  for i, off in enumerate(offs):
    q = flat[loc + off]
    m = (flat != 0) & (q != 0) & (q != flat)
    idx = np.nonzero(m)[0]
    recs.append((idx, np.full(idx.size, i, dtype=np.int64), np.minimum(flat[idx], q[idx]), np.maximum(flat[idx], q[idx])))
  if not recs or sum(r[0].size for r in recs) == 0:
    return {}
  idx = np.concatenate([r[0] for r in recs]); d = np.concatenate([r[1] for r in recs])
  a = np.concatenate([r[2] for r in recs]); bb = np.concatenate([r[3] for r in recs])
  order = np.lexsort((d, idx))
  a, bb, d = a[order], bb[order], d[order]
  pairs, inv = np.unique(np.stack([a, bb], 1), axis=0, return_inverse=True)
  acc = np.zeros(pairs.shape[0], dtype=np.float32)
  np.add.at(acc, inv.reshape(-1), np.asarray(areas, dtype=np.float32)[d])
  return {(int(p[0]), int(p[1])): float(v) for p, v in zip(pairs, acc)}


def region_graph(labels, connectivity=26):
  return set(contacts(labels, connectivity=connectivity).keys())


# ---- runs / draw / erase / each (fastcc3d.pyx:1258-1372; cc3d_graphs.hpp:470-523) ----
def _flat_memory_order(arr: np.ndarray) -> np.ndarray:
  """reference `_reshape(arr, (arr.size,))` (fastcc3d.pyx:129-161)."""
  if arr.flags.f_contiguous:
    return arr.reshape(-1, order="F")
  if arr.flags.c_contiguous:
    return arr.reshape(-1)
  return arr.reshape((arr.size,))


def runs(labels: np.ndarray) -> dict:
  """extract_runs (cc3d_graphs.hpp:470-503): maximal runs of equal non-zero values of the flattened array, keyed by
  label in ascending order (std::map), each list in position order. A one-voxel array reports (0, 1) even for
  background (:481-484)."""
  labels = np.asarray(labels)
  if labels.dtype != np.bool_ and labels.dtype not in (np.uint8, np.uint16, np.uint32, np.uint64):
    raise TypeError("Unsupported type: " + str(labels.dtype))
  flat = _flat_memory_order(labels)
  if flat.dtype == np.bool_:
    flat = flat.view(np.uint8)
  n = flat.size
  if n == 0:
    raise IndexError("Out of bounds on buffer access (axis 0)")   # `&labels[0]` is bounds checked (fastcc3d.pyx:1269)
  if n == 1:
    return {int(flat[0]): [(0, 1)]}
  change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
  starts = np.concatenate(([0], change))
  ends = np.concatenate((change, [n]))
  vals = flat[starts]
  keep = vals != 0
  starts, ends, vals = starts[keep], ends[keep], vals[keep]
  out = {}
  for v in np.unique(vals).tolist():
    m = vals == v
    out[int(v)] = list(zip(starts[m].tolist(), ends[m].tolist()))
  return out


def draw(label, runs, image: np.ndarray) -> np.ndarray:
  """set_run_voxels (cc3d_graphs.hpp:505-523) through fastcc3d.pyx:1281-1314: runs are drawn one after the other and
  the first invalid one raises (the ones before it stay drawn). Returns the flattened image."""
  if image.dtype != np.bool_ and image.dtype not in (np.uint8, np.uint16, np.uint32, np.uint64):
    raise TypeError("Unsupported type: " + str(image.dtype))
  flat = _flat_memory_order(image)
  value = (label != 0) if image.dtype == np.bool_ else label
  for a, b in runs:
    if b > flat.size or a >= b:
      raise RuntimeError("Invalid run.")
    flat[a:b] = value
  return flat


def erase(runs, image: np.ndarray) -> np.ndarray:
  return draw(0, runs, image)


def each(labels: np.ndarray, binary: bool = False, in_place: bool = False):
  """fastcc3d.pyx:1328-1372 as a plain generator of (label, image) (the in_place image is reused and read-only
  while it is out; callers compare before advancing)."""
  all_runs = runs(labels)
  order = "F" if labels.flags.f_contiguous else "C"
  dtype = np.bool_ if binary else labels.dtype
  img = np.zeros(labels.shape, dtype=dtype, order=order) if in_place else None
  for key, rns in all_runs.items():
    if key == 0:
      continue
    if in_place:
      draw(key, rns, img)
      img.setflags(write=0)
      yield key, img
      img.setflags(write=1)
      erase(rns, img)
    else:
      one = np.zeros(labels.shape, dtype=dtype, order=order)
      draw(key, rns, one)
      yield key, one
