"""GPU tests of the sharded path's device side: the slab merge kernels (cc3d_b200_merge_slabs_device) against the host
merge (cc3d_b200_merge_slabs, itself pinned against the numpy merge in tests/test_sharded_cpu.py), and the
single-process fast path (slab_begin / merge on the device / slab_finish) against the monolithic call."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_merge(L, torch, _lib, rows, world, cap, rank, label_cap, small=False):
  gathered = torch.from_numpy(rows).cuda()
  ws = torch.empty((int(L.cc3d_b200_merge_workspace_bytes(label_cap)),), dtype=torch.uint8, device="cuda")
  remap_p, result_p = ctypes.c_void_p(), ctypes.c_void_p()
  st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
  fn = L.cc3d_b200_merge_slabs_device_small if small else L.cc3d_b200_merge_slabs_device
  _lib.check(fn(gathered.data_ptr(), world, rows.shape[1], rank, cap, ws.data_ptr(), label_cap,
                ctypes.byref(remap_p), ctypes.byref(result_p), st))
  torch.cuda.synchronize()
  roff = result_p.value - ws.data_ptr()
  res = ws[roff:roff + 48].view(torch.int64).cpu().numpy()
  n = int(rows[rank, 0])
  moff = remap_p.value - ws.data_ptr()
  remap = ws[moff:moff + 4 * (n + 1)].view(torch.int32).cpu().numpy().view(np.uint32)
  return res, remap


def test_device_merge_equals_host_merge(cc3d):
  import torch
  from cc3d_b200 import _lib, sharded
  L = _lib.lib()
  rng = np.random.default_rng(5)
  for trial in range(40):
    world = int(rng.integers(1, 7))
    N_r = [int(rng.integers(0, 60)) for _ in range(world)]
    if trial % 5 == 0:
      N_r = [int(rng.integers(1000, 40000)) for _ in range(world)]
    pair_lists = [np.zeros(0, np.int64)]
    for r in range(1, world):
      k = 0 if (N_r[r - 1] == 0 or N_r[r] == 0) else int(rng.integers(0, 3 * max(N_r[r], 4)))
      lo = rng.integers(1, N_r[r - 1] + 1, k) if k else np.zeros(0, np.int64)
      up = rng.integers(1, N_r[r] + 1, k) if k else np.zeros(0, np.int64)
      pr = (lo.astype(np.int64) << 32) | up.astype(np.int64)
      if k and trial % 3 == 0:
        pr = np.concatenate([pr, pr[: k // 2]])            # duplicates, as the face kernel produces them
      pair_lists.append(pr)
    cap = max(8, max(len(p) for p in pair_lists))
    rows = np.zeros((world, 4 + cap), np.int64)
    for r in range(world):
      rows[r, 0] = N_r[r]; rows[r, 1] = 7; rows[r, 2] = 3; rows[r, 3] = len(pair_lists[r])
      rows[r, 4:4 + len(pair_lists[r])] = pair_lists[r]
    label_cap = max(64, 1 << int(sum(N_r) + 2).bit_length())
    for rank in range(world):
      want_N, want = sharded._merge_native(N_r, pair_lists, rank)
      res, remap = _device_merge(L, torch, _lib, rows, world, cap, rank, label_cap)
      assert int(res[1]) == 0 and int(res[2]) == 0
      assert int(res[0]) == want_N, (trial, rank)
      assert np.array_equal(remap.astype(np.int64), want), (trial, rank)
      # the single-CTA merge: same tables when the graph is small enough, result[5] raised otherwise
      res_s, remap_s = _device_merge(L, torch, _lib, rows, world, cap, rank, label_cap, small=True)
      if sum(N_r) + 1 <= 65536:
        assert int(res_s[5]) == 0 and int(res_s[0]) == want_N and int(res_s[1]) == 0 and int(res_s[2]) == 0, (trial, rank)
        assert np.array_equal(remap_s.astype(np.int64), want), (trial, rank)
        assert np.array_equal(res_s[8:8 + 4 * world], res[8:8 + 4 * world])
      else:
        assert int(res_s[5]) == 1
  # capacity flags
  rows = np.zeros((2, 4 + 8), np.int64)
  rows[0, 0] = 100; rows[1, 0] = 100; rows[1, 3] = 20          # 20 pairs reported, room for 8
  res, _ = _device_merge(L, torch, _lib, rows, 2, 8, 0, 1024)
  assert int(res[2]) == 1 and int(res[1]) == 0
  rows[1, 3] = 0
  res, _ = _device_merge(L, torch, _lib, rows, 2, 8, 0, 128)   # 201 ids do not fit 128
  assert int(res[1]) == 1


@pytest.mark.parametrize("conn", [6, 26])
def test_single_process_fast_path_equals_monolithic(cc3d, conn):
  """connected_components_slab without a process group = one slab: slab_begin -> device merge -> slab_finish."""
  import torch
  from cc3d_b200 import sharded
  rng = np.random.default_rng(conn)
  vol = np.repeat(np.repeat(np.repeat(rng.integers(0, 5, (12, 11, 13)), 5, 0), 5, 1), 5, 2)[:57, :53, :61].astype(np.int32)
  t = torch.from_numpy(vol).cuda()
  for kw in (dict(), dict(binary_image=True), dict(out_dtype=np.uint64), dict(out_dtype=np.uint16)):
    want, Nw = cc3d.connected_components(t, connectivity=conn, return_N=True, **kw)
    got, N = sharded.connected_components_slab(t, connectivity=conn, return_N=True, **kw)
    assert N == Nw and got.dtype == want.dtype and got.shape == want.shape, kw
    assert np.array_equal(got.cpu().numpy(), want.cpu().numpy()), kw
  # label capacity smaller than the number of components: the step repeats with a larger workspace
  sharded._label_cap_seen[t.device.index] = 64
  x = torch.from_numpy((np.arange(40 * 30 * 20, dtype=np.int32) + 1).reshape(20, 30, 40)).cuda()
  got, N = sharded.connected_components_slab(x, connectivity=conn, return_N=True)
  assert N == x.numel() and np.array_equal(got.cpu().numpy().reshape(-1), np.arange(1, x.numel() + 1, dtype=np.uint32))
