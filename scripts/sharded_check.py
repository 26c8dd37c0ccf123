"""Sharded (z-slab) labelling under torchrun: every rank labels its slab of ONE volume with
cc3d_b200.sharded.connected_components_slab; rank 0 checks the concatenation against the single-GPU
labelling of the whole volume (bit exact) and prints timings.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sharded_check.py [SZ_PER_RANK] [SY] [SX]
"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200, benchdata
from cc3d_b200 import sharded

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
szr = int(sys.argv[1]) if len(sys.argv) > 1 else 128
sy = int(sys.argv[2]) if len(sys.argv) > 2 else 512
sx = int(sys.argv[3]) if len(sys.argv) > 3 else 512
shape = (szr * world, sy, sx)
for name, conn, kw in (("voronoi", 26, {}), ("voronoi", 6, {}), ("binary", 26, dict(binary_image=True))):
    if name == "voronoi":
        slab = benchdata.voronoi_multilabel(shape, cell=40, seed=2, device=dev, dtype=torch.int32, z_range=(rank * szr, (rank + 1) * szr))
    else:
        g = torch.Generator(device=dev); g.manual_seed(100 + rank)
        slab = (torch.rand((szr, sy, sx), generator=g, device=dev) < 0.5).to(torch.uint8)
    out, N = sharded.connected_components_slab(slab, connectivity=conn, return_N=True, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out, N = sharded.connected_components_slab(slab, connectivity=conn, return_N=True, **kw)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    # gather and check on rank 0
    ok = None
    if out.dtype in (torch.uint32, torch.uint16, torch.uint64):   # NCCL has no unsigned 16/32/64
        out = out.view({torch.uint16: torch.int16, torch.uint32: torch.int32, torch.uint64: torch.int64}[out.dtype])
    if world > 1:
        parts = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
        dist.gather(out, parts, dst=0)
        vparts = [torch.empty_like(slab) for _ in range(world)] if rank == 0 else None
        dist.gather(slab, vparts, dst=0)
    else:
        parts, vparts = [out], [slab]
    if rank == 0:
        whole = torch.cat(vparts, 0)
        if whole.numel() < 2**32 - 1:
            ref, Nr = cc3d_b200.connected_components(whole, connectivity=conn, return_N=True, **kw)
            got = torch.cat(parts, 0)
            ok = (Nr == N) and ref.element_size() == got.element_size() and bool(torch.equal(ref.view(got.dtype), got))
        vox = whole.numel()
        print(f"{name} conn={conn} {kw} shape={tuple(whole.shape)} world={world}: N={N} identical_to_single_gpu={ok} "
              f"best {min(ts)*1e3:.3f} ms -> {vox/min(ts)/1e9:.1f} GVx/s", flush=True)
if world > 1:
    dist.destroy_process_group()
