"""Per-phase host / device times of the single-sync slab path (torchrun, CC3D_SHARDED_TIMING=<rank to print>)."""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200, benchdata
from cc3d_b200 import sharded
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
szr = int(sys.argv[1]); sy = int(sys.argv[2]); sx = int(sys.argv[3]); dt = torch.int64 if (len(sys.argv) > 4 and sys.argv[4] == "u64") else torch.int32
slab = benchdata.voronoi_multilabel((szr * world, sy, sx), cell=int(sys.argv[5]) if len(sys.argv) > 5 else 40, seed=2, device=dev, dtype=dt,
                                    id_bits=62 if dt == torch.int64 else 31, z_range=(rank * szr, (rank + 1) * szr))
tim = os.environ.pop("CC3D_SHARDED_TIMING", None)
for _ in range(5):
    sharded.connected_components_slab(slab, connectivity=26, return_N=True)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    out, N = sharded.connected_components_slab(slab, connectivity=26, return_N=True)
torch.cuda.synchronize(); dist.barrier()
if rank == 0: print(f"world {world} slab {tuple(slab.shape)} {dt}: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms/step N={N}", flush=True)
if tim is not None:
    os.environ["CC3D_SHARDED_TIMING"] = tim
    for _ in range(4):
        dist.barrier()
        sharded.connected_components_slab(slab, connectivity=26, return_N=True)
dist.destroy_process_group()
