"""Small driver for ncu captures of the clearance-bounds kernel: a few batches through the C ABI."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from path_optimizer_2_b200 import bounds, sharedmap  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--n", type=int, default=120)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dmap = sharedmap.DistanceMap()
lines = sharedmap.make_lines(a.batch, a.n, dmap=dmap)
dev = torch.device("cuda", 0)
pbn = bounds.PathBounds(dmap.dist, dmap.res)
t = [torch.from_numpy(v).to(dev) for v in (lines.states, lines.n, lines.spline, lines.k)]
d_bounds = torch.zeros((a.batch, 6, a.n), dtype=torch.float64, device=dev)
d_nv = torch.zeros(a.batch, dtype=torch.int32, device=dev)
bi = bounds.BoundsIn(a.batch, a.n, lines.k_max, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr())
bo = bounds.BoundsOut(d_bounds.data_ptr(), d_nv.data_ptr(), None)
for _ in range(a.reps):
    pbn.compute_device(bi, bo, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("bounds kernel ms", pbn.last_kernel_ms, "states", a.batch * a.n)
