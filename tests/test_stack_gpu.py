"""GPU tests of the streaming front end cc3d_b200.connected_components_stack (counterpart of the
reference's connected_components_stack, cc3d/__init__.py:353-501): slab by slab through the C-ABI, bit-identical to
the monolithic labelling of the Fortran-ordered concatenation (oracle at small sizes, the monolithic GPU call - itself
pinned by tests/test_parity_gpu.py - at a larger one)."""
import numpy as np
import pytest

from test_sharded_cpu import _stack_cases

pytestmark = pytest.mark.gpu


def test_stack_equals_monolithic_small(cc3d, oracle_mod):
  n = 0
  for vol, images, kw in _stack_cases():
    want, Nw = oracle_mod.connected_components(np.asfortranarray(vol), return_N=True, **kw)
    got, N = cc3d.connected_components_stack(iter(images), return_N=True, order="F", **kw)
    assert N == Nw and got.dtype == want.dtype and got.shape == want.shape and got.flags.f_contiguous, kw
    got_k = cc3d.connected_components_stack(iter(images), **kw)        # memory order follows the first image
    first = next(im for im in images if im.size)
    assert got_k.flags.f_contiguous == bool(first.flags.f_contiguous) and np.array_equal(got_k, want), kw
    assert np.array_equal(got, want), kw
    mono, Nm = cc3d.connected_components(np.asfortranarray(vol), return_N=True, **kw)
    assert Nm == N and np.array_equal(mono, got)
    n += 1
  assert n == 6


@pytest.mark.parametrize("conn,binary", [(26, False), (6, False), (26, True)])
def test_stack_equals_monolithic_large(cc3d, conn, binary):
  """320 x 200 x 150 volume in five uneven slabs (one of depth 1), memmap-like preallocated result."""
  from helpers import blobs
  rng = np.random.default_rng(31)
  shape = (320, 200, 150)
  vol = blobs(rng, shape, 5, 6).astype(np.uint32)
  if binary:
    vol = (rng.random(shape) < 0.35).astype(np.uint8)
  vol = np.asfortranarray(vol)
  cuts = [0, 40, 41, 90, 128, 150]
  images = (vol[:, :, a:b] for a, b in zip(cuts[:-1], cuts[1:]))
  want, Nw = cc3d.connected_components(vol, connectivity=conn, return_N=True, binary_image=binary)
  out = np.zeros(shape, dtype=want.dtype, order="F")
  got, N = cc3d.connected_components_stack(images, connectivity=conn, return_N=True, binary_image=binary,
                                           out_dtype=want.dtype, out=out)
  assert got is out and N == Nw
  assert np.array_equal(got, want)
