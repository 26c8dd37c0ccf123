"""crackle v0 decode on the GPU.

crackle is the on-disk format either side of the labelling path: the reference's benchmark volume
(benchmarks/connectomics.npy.ckl.gz) is stored in it and its connected_components_stack emits it
(cc3d/__init__.py:415, 476-492). A crackle decoder turns every z slice's crack codes into a 4-connected pixel graph and
colours it with cc3d.color_connectivity_graph (cc3d_graphs.hpp:583-1106, 1018-1074); here the crack codes are parsed
by one GPU thread per slice, all slices are coloured at once by the same union-find kernels as
cc3d_b200.color_connectivity_graph, and the component -> label table is applied on the device
(cc3d_b200_crackle_v0_decode). The host only reads the small header / label section.

Supported: format version 0 as written for the reference's fixture - flat labels, "impermissible" crack codes,
whole-slice grid, Markov order 0, Fortran order, 1/2/4-byte stored labels. Anything else raises NotImplementedError.
"""
from __future__ import annotations

import ctypes
import gzip
import struct

import numpy as np

from . import _lib


def _read(binary):
  if isinstance(binary, (bytes, bytearray, memoryview)):
    raw = bytes(binary)
  else:
    with open(binary, "rb") as f:
      raw = f.read()
  if raw[:2] == b"\x1f\x8b":
    raw = gzip.decompress(raw)
  return raw


def header(binary) -> dict:
  """Parsed 24-byte header of a crackle v0 stream (bytes or path; gzip is undone)."""
  raw = _read(binary)
  if raw[:4] != b"crkl":
    raise ValueError("not a crackle stream")
  version = raw[4]
  fmt, = struct.unpack_from("<H", raw, 5)
  sx, sy, sz = struct.unpack_from("<III", raw, 7)
  grid_log2 = raw[19]
  num_label_bytes, = struct.unpack_from("<I", raw, 20)
  return {
    "version": version, "format": fmt, "shape": (sx, sy, sz), "grid_log2": grid_log2, "num_label_bytes": num_label_bytes,
    "data_width": 1 << (fmt & 3), "stored_width": 1 << ((fmt >> 2) & 3), "crack_format": (fmt >> 4) & 1,
    "label_format": (fmt >> 5) & 3, "fortran_order": bool((fmt >> 7) & 1), "signed": bool((fmt >> 8) & 1),
    "markov_order": (fmt >> 9) & 15, "raw": raw,
  }


def decompress(binary, device=None):
  """Decodes a crackle v0 stream (bytes, or a path to a .ckl / .ckl.gz file) into the label volume: a Fortran-ordered
  numpy array of shape (sx, sy, sz), or - with device="cuda" / a torch device - a CUDA tensor of that shape and order
  that never leaves the GPU."""
  h = header(binary)
  raw = h["raw"]
  if h["version"] != 0:
    raise NotImplementedError(f"crackle format version {h['version']} (only version 0 is implemented)")
  if h["crack_format"] != 0 or h["label_format"] != 0 or h["markov_order"] != 0 or not h["fortran_order"] or h["signed"]:
    raise NotImplementedError("only flat labels, impermissible crack codes, Markov order 0, unsigned, Fortran order are implemented")
  sx, sy, sz = h["shape"]
  if h["grid_log2"] < 31 and (1 << h["grid_log2"]) < max(sx, sy):
    raise NotImplementedError("gridded crack codes are not implemented (the whole slice must be one grid cell)")
  if h["data_width"] > 4:
    raise NotImplementedError("64-bit labels are not implemented")
  sw = h["stored_width"]
  udt = {1: "<u1", 2: "<u2", 4: "<u4"}.get(sw)
  if udt is None:
    raise NotImplementedError("64-bit stored labels are not implemented")
  off = 24
  zindex = np.frombuffer(raw, dtype="<u4", count=sz, offset=off).astype(np.uint64)
  off += 4 * sz
  lab = raw[off:off + h["num_label_bytes"]]
  off += h["num_label_bytes"]
  num_unique, = struct.unpack_from("<Q", lab, 0)
  uniq = np.frombuffer(lab, dtype=udt, count=num_unique, offset=8).astype(np.uint32)
  cps = np.frombuffer(lab, dtype="<u4", count=sz, offset=8 + sw * num_unique)
  total = int(cps.astype(np.int64).sum())
  kdt = "<u1" if num_unique <= 0xFF else ("<u2" if num_unique <= 0xFFFF else "<u4")
  keys = np.frombuffer(lab, dtype=kdt, count=total, offset=8 + sw * num_unique + 4 * sz)
  lut = np.empty(total + 1, dtype=np.uint32)
  lut[0] = 0
  lut[1:] = uniq[keys]
  slice_off = np.zeros(sz + 1, dtype=np.uint64)
  np.cumsum(zindex, out=slice_off[1:])
  stream = np.frombuffer(raw, dtype=np.uint8, offset=off)
  if int(slice_off[-1]) > stream.size:
    raise ValueError("crackle stream is truncated")
  out_dtype = {1: np.uint8, 2: np.uint16, 4: np.uint32}[h["data_width"]]
  L = _lib.lib()
  n = ctypes.c_uint64(0)
  if device is not None:
    import torch
    dev = torch.device(device)
    flat = torch.empty((sx * sy * sz,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
      _lib.check(L.cc3d_b200_crackle_v0_decode(
        stream.ctypes.data, slice_off.ctypes.data, sx, sy, sz, lut.ctypes.data, lut.size, flat.data_ptr(), ctypes.byref(n),
        _lib.HOST, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    if int(n.value) != total:
      raise ValueError(f"crackle stream is inconsistent: {int(n.value)} components decoded, {total} keys stored")
    out = flat.view(torch.uint32).reshape(sz, sy, sx).permute(2, 1, 0)
    if out_dtype != np.uint32:
      out = out.to({np.uint8: torch.uint8, np.uint16: torch.uint16}[out_dtype])
    return out
  flat = np.empty(sx * sy * sz, dtype=np.uint32)
  _lib.check(L.cc3d_b200_crackle_v0_decode(
    stream.ctypes.data, slice_off.ctypes.data, sx, sy, sz, lut.ctypes.data, lut.size, flat.ctypes.data, ctypes.byref(n),
    _lib.HOST, None))
  if int(n.value) != total:
    raise ValueError(f"crackle stream is inconsistent: {int(n.value)} components decoded, {total} keys stored")
  return flat.reshape((sx, sy, sz), order="F").astype(out_dtype, copy=False)
