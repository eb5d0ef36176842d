"""One all-vs-all dist step of the configs[2] shape (20000 sketches, s=1000) for ncu: warm-up outside the profiled range.
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dist_launches.csv \
      python profiles/dist_prof.py [hits]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import __graft_entry__ as g

g._paths()
import fpmash_b200 as fpm

hits_mode = len(sys.argv) > 1 and sys.argv[1] == "hits"
dev = torch.device("cuda", 0)
nd, S, K = 20000, 1000, 21
ctx = fpm.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
panel = bench.gen_sketch_panel(torch, nd, S, dev, seed=3)
sizes = torch.full((nd,), S, dtype=torch.int32, device=dev)
lengths = torch.full((nd,), 5_000_000, dtype=torch.int64, device=dev)
ptrs = (panel.data_ptr(), sizes.data_ptr(), lengths.data_ptr(), nd, S)
if hits_mode:
    cap = 32 << 20
    out = torch.empty(cap * 32, dtype=torch.uint8, device=dev)
    step = lambda: ctx.dist_hits_dev(ptrs, ptrs, S, K, 4.0 ** K, out.data_ptr(), cap, max_distance=0.25)
else:
    out = torch.empty(nd * nd * 24, dtype=torch.uint8, device=dev)
    step = lambda: ctx.dist_tile_dev(ptrs, ptrs, S, K, 4.0 ** K, out.data_ptr())
for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
r = step()
e1.record()
torch.cuda.synchronize()
print("step %.2f ms%s" % (e0.elapsed_time(e1), "  hits %d" % r if hits_mode else ""))
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
