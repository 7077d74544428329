"""K1 roofline sweep (dev tool): GB/s of calc_paths at B=2048 x EigenWorms for a few shapes."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "log-neural-cdes_b200"))
import torch
from lncde import _build
_build.build_library()
from lncde.data_dir.generate_paths import calc_paths

def run(B, L, C, s, depth, reps=10):
    x = torch.randn(B, L, C, device="cuda").cumsum(1)
    out = calc_paths(x, s, depth)
    for _ in range(3): calc_paths(x, s, depth, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): calc_paths(x, s, depth, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = x.numel() * 4 + out.numel() * 4
    print(f"B={B} L={L} C={C} s={s} d={depth}: {ms:.3f} ms  {byts/ms/1e6:.0f} GB/s  ({byts/ms/1e6/6545.3:.3f} of measured HBM)", flush=True)

if __name__ == "__main__":
    run(2048, 17984, 7, 12, 2)
    run(2048, 17984, 6, 12, 2)
    run(512, 49920, 7, 10, 2)
    run(2048, 17984, 7, 12, 1)
    run(512, 17984, 7, 12, 3, reps=3)
    run(8192, 896, 6, 16, 2)
