"""One dist step of a BLOCK of the configs[2] all-vs-all (queries [0, nq) x references [0, nr) of the 20000-sketch panel)
for an ncu launch list: what one rank of a q_parts x r_parts grid executes after the exchange step.
  python profiles/r02_dist_block_prof.py 10000 5000      # the block of a 2 x 4 grid on 8 GPUs
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import __graft_entry__ as g

g._paths()
import fpmash_b200 as fpm

nq, nr = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
nd, S, K = 20000, 1000, 21
ctx = fpm.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
panel = bench.gen_sketch_panel(torch, nd, S, dev, seed=3)
sizes = torch.full((nd,), S, dtype=torch.int32, device=dev)
lengths = torch.full((nd,), 5_000_000, dtype=torch.int64, device=dev)
q = (panel.data_ptr(), sizes.data_ptr(), lengths.data_ptr(), nq, S)
r = (panel[nd - nr:].data_ptr(), sizes.data_ptr(), lengths.data_ptr(), nr, S)
out = torch.empty(nq * nr * 24, dtype=torch.uint8, device=dev)
step = lambda: ctx.dist_tile_dev(r, q, S, K, 4.0 ** K, out.data_ptr())
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(5):
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("block %d x %d: step %.3f ms (min of 5; all: %s)" % (nq, nr, min(ts), " ".join("%.3f" % t for t in ts)))
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
