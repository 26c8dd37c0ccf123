"""A/B helper: headline workload (512^3 uint32 Voronoi, 26-connected) step time with whatever library CC3D_B200_LIB
selects; prints median / best of 40 steps and the per-kernel marks. Run once per library variant."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200, benchdata
tag = os.path.basename(os.environ.get("CC3D_B200_LIB", "libcc3d_b200.so")) + ("[no_tma]" if os.environ.get("CC3D_B200_NO_TMA") else "")
from oracle import decode_connectomics
_vol = decode_connectomics.load_fixture()
_extra = []
if _vol is not None:
    _extra.append(("connectomics512_u32_c26", torch.from_numpy(np.ascontiguousarray(_vol.transpose(2, 1, 0)).view(np.int32)).cuda(), dict(connectivity=26)))
if os.environ.get("AB_MORE"):
    _extra.append(("binary512_u8_c26", benchdata.random_binary((512, 512, 512), 0.5, 1, "cuda"), dict(connectivity=26, binary_image=True)))
    _extra.append(("tone512_f32_c26_d10", benchdata.three_tone_noise((512, 512, 512), cell=64, seed=3, device="cuda"), dict(connectivity=26, delta=10)))
for name, x, kw in _extra + [
    ("voronoi512_u32_c26", benchdata.voronoi_multilabel((512, 512, 512), cell=40, seed=2, device="cuda", dtype=torch.int32), dict(connectivity=26)),
    ("voronoi256_u32_c26", benchdata.voronoi_multilabel((256, 256, 256), cell=40, seed=2, device="cuda", dtype=torch.int32), dict(connectivity=26)),
    ("binary512_u8_c6", benchdata.random_binary((512, 512, 512), 0.5, 1, "cuda"), dict(connectivity=6, binary_image=True)),
]:
    for _ in range(5):
        out, N = cc3d_b200.connected_components(x, return_N=True, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(40):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out, N = cc3d_b200.connected_components(x, return_N=True, **kw); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    cc3d_b200.set_timing(True)
    cc3d_b200.connected_components(x, return_N=True, **kw)
    tm = cc3d_b200.last_timings()
    cc3d_b200.set_timing(False)
    print(f"{tag} {name}: N={N} median {ts[20]:.4f} ms best {ts[0]:.4f} ms | " + " ".join(f"{k}={v:.3f}" for k, v in tm), flush=True)
