"""GPU parity tests of SURVEY 8(f)4: runs / draw / erase / each (run with -m gpu on the B200 box). Every call goes
through the C-ABI (cc3d_b200_runs / cc3d_b200_draw); results are compared with
 - fixtures generated from the unmodified reference (tests/golden/runs_*.npz, make_golden_runs.py),
 - the numpy oracle (oracle/oracle.py, pinned against the reference in tests/test_oracle.py) on seeded inputs,
 - the encode -> draw round trip at BASELINE.json's 512^3 size (device resident)."""
import ctypes

import numpy as np
import pytest

from helpers import blobs
from test_oracle import _run_cases, _run_goldens, each_checksums, load_run_golden

pytestmark = pytest.mark.gpu


def _same_runs(got, want, ctx=""):
  assert list(got.keys()) == list(want.keys()), ctx
  assert got == want, ctx
  for v in got.values():
    assert all(type(a) is int and type(b) is int for a, b in v[:4])


@pytest.mark.parametrize("path", _run_goldens(), ids=[p.split("/")[-1][:-4] for p in _run_goldens()])
def test_run_goldens(cc3d, path):
  import torch
  z, x, want = load_run_golden(path)
  order = "F" if bool(z["f_order"]) else "C"
  _same_runs(cc3d.runs(x), want, path)
  _same_runs(cc3d.fastcc3d.runs(x), want, path)
  canvas = np.full(x.shape, 7 if x.dtype != np.bool_ else 0, dtype=x.dtype, order=order)
  flat = cc3d.draw(int(z["draw_value"]), want[int(z["draw_key"])], canvas)
  assert np.array_equal(canvas, z["drawn"]) and flat.ndim == 1 and np.shares_memory(flat, canvas)
  for binary in (False, True):
    for in_place in (False, True):
      assert np.array_equal(each_checksums(cc3d.each(x, binary=binary, in_place=in_place), order), z[f"each_{int(binary)}"])
  # device-resident labels: same table, images come back as CUDA tensors
  if x.dtype in (np.uint8, np.bool_):
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    _same_runs(cc3d.runs(t), cc3d.runs(np.ascontiguousarray(x)))
    for label, img in cc3d.each(t, binary=False, in_place=True):
      assert img.is_cuda and img.shape == t.shape
      assert torch.equal(img, torch.where(t == label, t, torch.zeros_like(t)))


def test_runs_draw_each_fuzz_vs_oracle(cc3d, oracle_mod):
  rng = np.random.default_rng(23)
  n = 0
  for x in _run_cases(rng, 150):
    want = oracle_mod.runs(x)
    _same_runs(cc3d.runs(x), want, (x.shape, x.dtype))
    for binary in (False, True):
      for in_place in (False, True):
        got = [(k, im.copy()) for k, im in cc3d.each(x, binary=binary, in_place=in_place)]
        exp = [(k, im.copy()) for k, im in oracle_mod.each(x, binary=binary, in_place=in_place)]
        assert len(got) == len(exp) == len(cc3d.each(x, binary=binary, in_place=in_place))
        for (ka, ia), (kb, ib) in zip(got, exp):
          assert ka == kb and ia.dtype == ib.dtype and ia.shape == ib.shape and np.array_equal(ia, ib)
          assert ia.flags.f_contiguous == ib.flags.f_contiguous and ia.flags.c_contiguous == ib.flags.c_contiguous
    if want:
      k = list(want.keys())[-1]
      c1 = np.full(x.shape, 1, dtype=x.dtype, order="F" if x.flags.f_contiguous else "C")
      c2 = c1.copy(order="K")
      cc3d.erase(want[k], c1), oracle_mod.erase(want[k], c2)
      assert np.array_equal(c1, c2)
    n += 1
  assert n == 150


def test_runs_edge_cases(cc3d, oracle_mod):
  one = np.zeros((1, 1, 1), np.uint16)
  assert cc3d.runs(one) == {0: [(0, 1)]}
  assert cc3d.runs(one + np.uint16(9)) == {9: [(0, 1)]}
  with pytest.raises(IndexError):
    cc3d.runs(np.zeros((0,), np.uint8))
  for dt in (np.float32, np.int32, np.int8):
    with pytest.raises(TypeError):
      cc3d.runs(np.zeros((3, 3), dt))
    with pytest.raises(TypeError):
      cc3d.draw(1, [(0, 1)], np.zeros((3, 3), dt))
  for bad in ([(3, 3)], [(5, 2)], [(0, 65)], [(0, 4), (70, 71)]):
    img = np.zeros((8, 8), np.uint8)
    with pytest.raises(RuntimeError, match="Invalid run"):
      cc3d.draw(1, bad, img)
    assert not img.any()   # validated before anything is drawn
  with pytest.raises(OverflowError):
    cc3d.draw(256, [(0, 1)], np.zeros((8, 8), np.uint8))
  assert cc3d.draw(5, [], np.zeros((4,), np.uint8)).tolist() == [0, 0, 0, 0]
  # non-contiguous input is walked as its C-order copy (reference _reshape fallback)
  base = np.arange(64, dtype=np.uint32).reshape(8, 8) // 3
  view = base[::2, 1::2]
  assert cc3d.runs(view) == oracle_mod.runs(view)
  # overlapping and repeated runs
  img = np.zeros(100, np.uint16)
  cc3d.draw(3, [(10, 50), (40, 60), (10, 50)], img)
  assert np.array_equal(np.flatnonzero(img), np.arange(10, 60)) and set(img[10:60]) == {3}


@pytest.mark.parametrize("dt", [np.uint8, np.uint16, np.uint32, np.uint64])
def test_runs_chunk_boundaries_and_long_runs(cc3d, oracle_mod, dt):
  """Runs that cross the 32-voxel warp steps, the 512-voxel warp segments and the 4096-voxel chunks; sizes around
  the chunk size; runs longer than the 16384-voxel split of the draw kernels."""
  import torch
  rng = np.random.default_rng(5)
  for n in (31, 32, 33, 511, 512, 513, 4095, 4096, 4097, 8192 + 17, 3 * 4096):
    cuts = np.unique(np.concatenate(([0, n], rng.integers(0, n, 12), [32, 512, 4096, 4095, 4097, 8192])))
    cuts = cuts[cuts <= n]
    x = np.zeros(n, dt)
    for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
      x[a:b] = (i * 7) % 5
    want = oracle_mod.runs(x)
    _same_runs(cc3d.runs(x), want, (n, dt))
    _same_runs(cc3d.runs(torch.from_numpy(x.view(np.uint8).copy()).cuda().view(
      {1: torch.uint8, 2: torch.uint16, 4: torch.uint32, 8: torch.uint64}[x.itemsize])), want, (n, dt, "cuda"))
    if n > 40:   # a view that starts off the 16-byte grid takes the scalar kernels
      tc = torch.from_numpy(x.view(np.uint8).copy()).cuda().view({1: torch.uint8, 2: torch.uint16, 4: torch.uint32, 8: torch.uint64}[x.itemsize])
      assert tc[1:].data_ptr() % 16 != 0
      _same_runs(cc3d.runs(tc[1:]), oracle_mod.runs(x[1:]), (n, dt, "cuda unaligned"))
  # dense random values: more runs than the first capacity guess (second call of the protocol)
  x = rng.integers(0, 3, 1 << 21).astype(dt)
  got, want = cc3d.runs(x), oracle_mod.runs(x)
  assert got == want
  # long runs: one constant volume and a half / half volume, host and device images
  x = np.ones((64, 64, 40), dt)
  x[:, :, 20:] = 2
  want = oracle_mod.runs(x)
  _same_runs(cc3d.runs(x), want)
  imgs = list(cc3d.each(np.asfortranarray(x)))
  assert [k for k, _ in imgs] == [1, 2] and all(np.array_equal(im, np.where(x == k, x, 0)) for k, im in imgs)
  canvas = np.zeros(x.size + 100, dt)
  cc3d.draw(4, [(50, x.size + 50), (3, 4)], canvas)
  assert canvas[:50].tolist() == [0, 0, 0, 4] + [0] * 46 and (canvas[50:-50] == 4).all() and not canvas[-50:].any()
  tdt = {1: torch.uint8, 2: torch.uint16, 4: torch.uint32, 8: torch.uint64}[x.itemsize]
  t = torch.zeros((x.size + 100) * x.itemsize, dtype=torch.uint8, device="cuda").view(tdt)
  cc3d.draw(4, [(50, x.size + 50), (3, 4)], t)
  assert np.array_equal(t.cpu().numpy(), canvas)
  with pytest.raises(RuntimeError, match="Invalid run"):
    cc3d.draw(4, [(50, x.size + 101)], t)
  assert np.array_equal(t.cpu().numpy(), canvas)


def test_runs_round_trip_full_size(cc3d):
  """512^3 uint32 (BASELINE configs[0] shape), device resident: the run table redrawn label by label reproduces the
  volume; erasing every run leaves zeros; the table is sorted and disjoint."""
  import torch
  from cc3d_b200 import _lib
  g = torch.Generator(device="cuda").manual_seed(11)
  coarse = torch.randint(0, 7, (32, 32, 32), generator=g, device="cuda", dtype=torch.int32)
  vol = coarse.repeat_interleave(16, 0).repeat_interleave(16, 1).repeat_interleave(16, 2).contiguous()
  vol[3::17, 5, :] = 9   # thin structures: short runs
  lab, off, starts, ends = cc3d._runs_table(vol.view(torch.uint32))
  assert lab.tolist() == sorted(set(torch.unique(vol).tolist()) - {0})
  s, e = starts.astype(np.int64), ends.astype(np.int64)
  assert (e > s).all() and int((e - s).sum()) == int((vol != 0).sum().item())
  for i in range(len(lab)):
    a, b = off[i], off[i + 1]
    assert (s[a + 1:b] >= e[a:b - 1]).all()
  flat = vol.view(-1)
  redraw = torch.zeros_like(flat)
  for i, l in enumerate(lab.tolist()):
    rns = np.stack([starts[off[i]:off[i + 1]], ends[off[i]:off[i + 1]]], axis=1)
    cc3d.draw(l, rns, redraw.view(torch.uint32))
  assert torch.equal(redraw, flat)
  for i in range(len(lab)):
    rns = np.stack([starts[off[i]:off[i + 1]], ends[off[i]:off[i + 1]]], axis=1)
    cc3d.erase(rns, redraw.view(torch.uint32))
  assert not redraw.any()
  assert _lib.lib().cc3d_b200_launch_count() > 0
