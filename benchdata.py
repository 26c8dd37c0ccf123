"""Synthetic workloads for bench.py and the GPU tests (BASELINE.json configs).
Generated with torch on the target device so that nothing large crosses PCIe before the timed region."""
from __future__ import annotations

import torch


def random_binary(shape, density=0.5, seed=1, device="cuda", dtype=torch.uint8):
  """configs[1]: random 0/1 volume (worst case for union merging)."""
  g = torch.Generator(device=device)
  g.manual_seed(seed)
  out = torch.empty(shape, dtype=dtype, device=device)
  # chunked along the slowest axis to bound temporary memory
  step = max(1, (1 << 27) // max(1, int(torch.tensor(shape[1:]).prod()))) if len(shape) > 1 else shape[0]
  for z0 in range(0, shape[0], step):
    z1 = min(shape[0], z0 + step)
    out[z0:z1] = (torch.rand((z1 - z0,) + tuple(shape[1:]), generator=g, device=device) < density).to(dtype)
  return out


def voronoi_multilabel(shape, cell=40, seed=2, device="cuda", dtype=torch.int32, zero_fraction=0.0, id_bits=31,
                       z_range=None):
  """configs[0]/[2]-like multilabel volume: jittered-grid Voronoi cells with random ids.
  shape = (sz, sy, sx), C-contiguous (x fastest). One seed per `cell`^3 coarse cell; every voxel takes
  the id of the nearest seed among the 27 surrounding coarse cells.
  z_range = (z0, z1): only that slab of the volume is generated (sharded volumes: every rank draws the
  same seeds and fills its own planes)."""
  g = torch.Generator(device=device)
  g.manual_seed(seed)
  sz, sy, sx = shape
  zlo, zhi = (0, sz) if z_range is None else z_range
  nz, ny, nx = (sz + cell - 1) // cell + 2, (sy + cell - 1) // cell + 2, (sx + cell - 1) // cell + 2
  jitter = torch.rand((3, nz, ny, nx), generator=g, device=device) * cell
  hi = (1 << id_bits) - 1
  ids = torch.randint(1, hi, (nz, ny, nx), generator=g, device=device, dtype=torch.int64)
  if zero_fraction > 0:
    ids = ids * (torch.rand((nz, ny, nx), generator=g, device=device) >= zero_fraction)
  out = torch.empty((zhi - zlo, sy, sx), dtype=dtype, device=device)
  xs = torch.arange(sx, device=device, dtype=torch.float32)
  ys = torch.arange(sy, device=device, dtype=torch.float32)
  cx = (torch.arange(sx, device=device) // cell) + 1
  cy = (torch.arange(sy, device=device) // cell) + 1
  zchunk = max(1, (1 << 24) // (sy * sx))
  for z0 in range(zlo, zhi, zchunk):
    z1 = min(zhi, z0 + zchunk)
    zs = torch.arange(z0, z1, device=device, dtype=torch.float32)
    cz = (torch.arange(z0, z1, device=device) // cell) + 1
    best = torch.full((z1 - z0, sy, sx), float("inf"), device=device)
    best_id = torch.zeros((z1 - z0, sy, sx), dtype=torch.int64, device=device)
    for dz in (-1, 0, 1):
      for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
          iz, iy, ix = (cz + dz)[:, None, None], (cy + dy)[None, :, None], (cx + dx)[None, None, :]
          pz = (iz - 1) * cell + jitter[0][iz, iy, ix]
          py = (iy - 1) * cell + jitter[1][iz, iy, ix]
          px = (ix - 1) * cell + jitter[2][iz, iy, ix]
          d = (pz - zs[:, None, None]) ** 2 + (py - ys[None, :, None]) ** 2 + (px - xs[None, None, :]) ** 2
          closer = d < best
          best = torch.where(closer, d, best)
          best_id = torch.where(closer, ids[iz, iy, ix].expand_as(best_id), best_id)
    out[z0 - zlo:z1 - zlo] = best_id.to(dtype)
  return out


def three_tone_noise(shape, cell=64, seed=3, device="cuda"):
  """configs[3]: tones {64,128,192} on a Voronoi layout plus U(-4,4) noise, float32, no zeros."""
  lab = voronoi_multilabel(shape, cell=cell, seed=seed, device=device, dtype=torch.int32)
  g = torch.Generator(device=device)
  g.manual_seed(seed + 1000)
  tone = ((lab % 3) + 1).to(torch.float32) * 64.0
  del lab
  noise = torch.rand(shape, generator=g, device=device) * 8.0 - 4.0
  return tone + noise
