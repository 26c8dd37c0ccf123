"""Sharded (z-slab) labelling under torchrun: every rank labels its slab of ONE volume with
cc3d_b200.sharded.connected_components_slab. Volumes below 2^32-1 voxels are checked bit for bit
against the single-GPU labelling of the whole volume on rank 0; larger ones (2048^3) are checked
through size-independent properties (same N on every rank, labels agree across every slab interface
wherever the voxel values join, label range == [1, N]).

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sharded_check.py SZ_PER_RANK SY SX [u32|u64] [cell]
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
kind = sys.argv[4] if len(sys.argv) > 4 else "u32"
cell = int(sys.argv[5]) if len(sys.argv) > 5 else 40
shape = (szr * world, sy, sx)
total_vox = shape[0] * sy * sx
small = total_vox < 2**32 - 1
cases = [("voronoi", 26, {}), ("voronoi", 6, {})] + ([("binary", 26, dict(binary_image=True))] if small else [])
signed = {2: torch.int16, 4: torch.int32, 8: torch.int64}
for name, conn, kw in cases:
    if name == "voronoi":
        slab = benchdata.voronoi_multilabel(shape, cell=cell, seed=2, device=dev, dtype=torch.int64 if kind == "u64" else torch.int32,
                                            id_bits=62 if kind == "u64" else 31, z_range=(rank * szr, (rank + 1) * szr))
    else:
        g = torch.Generator(device=dev); g.manual_seed(100 + rank)
        slab = (torch.rand((szr, sy, sx), generator=g, device=dev) < 0.5).to(torch.uint8)
    out, N = sharded.connected_components_slab(slab, connectivity=conn, return_N=True, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out, N = sharded.connected_components_slab(slab, connectivity=conn, return_N=True, **kw)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    t = torch.tensor([min(ts)], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    best = float(t.item())
    if out.dtype in (torch.uint16, torch.uint32, torch.uint64):   # NCCL has no unsigned 16/32/64
        out = out.view(signed[out.element_size()])
    ok = None
    if small:
        if world > 1:
            parts = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
            dist.gather(out, parts, dst=0)
            vparts = [torch.empty_like(slab) for _ in range(world)] if rank == 0 else None
            dist.gather(slab, vparts, dst=0)
        else:
            parts, vparts = [out], [slab]
        st = sharded.statistics_slab(out, N, no_slice_conversion=True)      # every rank: statistics of the whole volume
        if rank == 0:
            whole = torch.cat(vparts, 0)
            ref, Nr = cc3d_b200.connected_components(whole, connectivity=conn, return_N=True, **kw)
            got = torch.cat(parts, 0)
            ok = (Nr == N) and ref.element_size() == got.element_size() and bool(torch.equal(ref.view(got.dtype), got))
            rs = cc3d_b200.statistics(ref, no_slice_conversion=True)
            st_ok = all(np.array_equal(st[k], rs[k], equal_nan=(k == "centroids")) for k in rs)
        check = f"identical_to_single_gpu={ok}" + (f" sharded_statistics_equal={st_ok}" if rank == 0 else "")
    else:
        # interface property: my first plane vs the previous rank's last plane (straight neighbours, equal values)
        good = torch.ones(1, dtype=torch.int64, device=dev)
        if world > 1:
            ops = []
            if rank + 1 < world:
                ops += [dist.P2POp(dist.isend, slab[szr - 1].contiguous(), rank + 1), dist.P2POp(dist.isend, out[szr - 1].contiguous().to(torch.int64), rank + 1)]
            if rank > 0:
                pv = torch.empty_like(slab[0]); pl = torch.empty((sy, sx), dtype=torch.int64, device=dev)
                ops += [dist.P2POp(dist.irecv, pv, rank - 1), dist.P2POp(dist.irecv, pl, rank - 1)]
            for r in dist.batch_isend_irecv(ops): r.wait()
            if rank > 0:
                join = (pv == slab[0]) & (slab[0] != 0)
                good[0] = int(bool(torch.all(out[0].to(torch.int64)[join] == pl[join])))
        lo = int(out[out != 0].min()) if bool((out != 0).any()) else 1
        hi = torch.tensor([int(out.max())], dtype=torch.int64, device=dev)
        Ns = torch.tensor([N], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(good, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            Nmin = Ns.clone(); dist.all_reduce(Nmin, op=dist.ReduceOp.MIN); dist.all_reduce(Ns, op=dist.ReduceOp.MAX)
        else:
            Nmin = Ns
        check = f"interfaces_consistent={bool(good.item())} same_N_everywhere={int(Nmin.item()) == int(Ns.item())} max_label==N={int(hi.item()) == N} min_label>=1={lo >= 1}"
    if rank == 0:
        print(f"{name} {kind} conn={conn} {kw} shape={shape} world={world}: N={N} out={out.dtype} {check} "
              f"best {best*1e3:.3f} ms -> {total_vox/best/1e9:.1f} GVx/s", flush=True)
    del out, slab
if world > 1:
    dist.destroy_process_group()
