"""ORS projector at the full 224x400 configuration (6 cameras x 28x50 pixels x 320 samples = 2.69 M lookups):
GPU kernel (CUDA events) next to the oracle port of the reference's CPU path."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from dualdiff_b200 import ops
from dualdiff_b200.networks import OccupancyRay
from oracle import ors_oracle as O
from test_ors import _cam_data, GOLD
g = torch.load(GOLD)
proj = OccupancyRay(image_shape=(896, 1600), sample_point=320, compress_ratio=400 / 8 / 1600, camera_data={"t": _cam_data(g)}, occ3d_idx={})
o, d = proj.rays(proj.camera_data["t"])
sem = g["semantics"].cuda()
for _ in range(3):
    ops.ors_project(o, d, sem, sample_point=320, want_rows=True)
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.ors_project(o, d, sem, sample_point=320, want_ids=False, want_rows=True); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
t_gpu = sorted(ts)[5]
nbytes = o.shape[0] * 320 * 2 + o.shape[0] * 24
t0 = time.perf_counter()
ref = O.project(g["semantics"], g["Ks"], g["Rts"], (896, 1600), 400 / 8 / 1600, 320, 0.2)
t_cpu = time.perf_counter() - t0
ids, _ = ops.ors_project(o, d, sem, sample_point=320)
print(f"ORS projector 6x28x50x320: GPU kernel {t_gpu * 1e3:.1f} us ({nbytes / t_gpu / 1e6:.0f} GB/s of output), oracle port on "
      f"{torch.get_num_threads()} CPU threads {t_cpu * 1e3:.0f} ms, mismatches vs oracle {int((ids.cpu().view(6, 28, 50, 320).long() != ref).sum())} of {ref.numel()}")
