"""Decodes the reference's config-#1 benchmark volume (benchmarks/connectomics.npy.ckl.gz, crackle v0)
without crackle-codec, following SURVEY.md Appendix C, and stores it as a compressed fixture under
oracle/_ref/ (git-ignored, travels to the GPU box). TEST/BENCH INFRASTRUCTURE ONLY.

    python oracle/decode_connectomics.py   # needs /root/reference and oracle/_ref/fastcc3d*.so

Pinned facts: 512^3 uint32 Fortran order,
sha256(F-order bytes) = 4f5ad1c03f6fa0478a5c332a1ff51cf7636a83a59c87e1d086e07fb0a5fc77d6.
"""
import gzip
import hashlib
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle  # noqa: E402

SRC = "/root/reference/benchmarks/connectomics.npy.ckl.gz"
DST = os.path.join(HERE, "_ref", "connectomics_512_u32.npz")
SHA = "4f5ad1c03f6fa0478a5c332a1ff51cf7636a83a59c87e1d086e07fb0a5fc77d6"


def decode(path=SRC):
  ref = oracle.reference_module()
  raw = gzip.open(path, "rb").read()
  assert raw[:4] == b"crkl" and raw[4] == 0
  fmt, = struct.unpack_from("<H", raw, 5)
  sx, sy, sz = struct.unpack_from("<III", raw, 7)
  num_label_bytes, = struct.unpack_from("<I", raw, 20)
  assert fmt == 0x008A
  off = 24
  zindex = np.frombuffer(raw, dtype="<u4", count=sz, offset=off); off += 4 * sz
  lab = raw[off:off + num_label_bytes]; off += num_label_bytes
  num_unique, = struct.unpack_from("<Q", lab, 0)
  uniq = np.frombuffer(lab, dtype="<u4", count=num_unique, offset=8)
  cps = np.frombuffer(lab, dtype="<u4", count=sz, offset=8 + 4 * num_unique)
  keys = np.frombuffer(lab, dtype="<u2", count=int(cps.sum()), offset=8 + 4 * num_unique + 4 * sz)
  out = np.zeros((sx, sy, sz), dtype=np.uint32, order="F")
  key_off = 0
  DX = (0, 1, 0, -1); DY = (-1, 0, 1, 0)
  for z in range(sz):
    blob = raw[off:off + int(zindex[z])]; off += int(zindex[z])
    index_bytes, = struct.unpack_from("<I", blob, 0)
    idx = np.frombuffer(blob, dtype="<u2", count=index_bytes // 2, offset=4)
    starts = []
    p = 0; n_rows = int(idx[p]); p += 1; y = 0
    for _ in range(n_rows):
      y += int(idx[p]); n = int(idx[p + 1]); p += 2
      x = 0
      for _ in range(n):
        x += int(idx[p]); p += 1
        starts.append((x, y))
    codes = np.frombuffer(blob, dtype=np.uint8, offset=4 + index_bytes)
    syms = np.stack([(codes >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(-1)
    dirs = (np.cumsum(syms) & 3).tolist()
    vcg = np.full((sx, sy), 0b1111, dtype=np.uint8)

    def cut(x, y, d):
      if d == 0:   a, b, ba, bb = (x - 1, y - 1), (x, y - 1), 0, 1
      elif d == 2: a, b, ba, bb = (x - 1, y), (x, y), 0, 1
      elif d == 3: a, b, ba, bb = (x - 1, y - 1), (x - 1, y), 2, 3
      else:        a, b, ba, bb = (x, y - 1), (x, y), 2, 3
      if 0 <= a[0] < sx and 0 <= a[1] < sy: vcg[a] &= ~(1 << ba) & 0xF
      if 0 <= b[0] < sx and 0 <= b[1] < sy: vcg[b] &= ~(1 << bb) & 0xF

    i = 0
    for (x, y) in starts:
      branches = 1; last = None; stack = []
      while branches > 0:
        d = dirs[i]; i += 1
        if last is not None and ((d - last) & 3) == 2:
          if d in (0, 3):
            branches -= 1
            if stack: x, y = stack.pop()
          else:
            branches += 1; stack.append((x, y))
          last = None
          continue
        if last is not None:
          cut(x, y, last); x += DX[last]; y += DY[last]
        last = d
    cc, n = ref.color_connectivity_graph(np.asfortranarray(vcg), connectivity=4, return_N=True)
    assert n == int(cps[z]), (z, n, int(cps[z]))
    lut = np.concatenate([[0], uniq[keys[key_off:key_off + n]]]).astype(np.uint32)
    out[:, :, z] = lut[cc]
    key_off += n
  return out


def load_fixture():
  """The decoded volume if the fixture travels with the repo, else None."""
  if not os.path.exists(DST):
    return None
  with np.load(DST) as z:
    return np.asfortranarray(z["labels"].transpose(2, 1, 0))


if __name__ == "__main__":
  vol = decode()
  h = hashlib.sha256(vol.tobytes(order="F")).hexdigest()
  print("sha256", h, "ok" if h == SHA else "MISMATCH")
  assert h == SHA
  # store C-contiguous (z,y,x) so the compressor sees x-runs
  np.savez_compressed(DST, labels=np.ascontiguousarray(vol.transpose(2, 1, 0)))
  print("wrote", DST, os.path.getsize(DST) >> 20, "MiB")
