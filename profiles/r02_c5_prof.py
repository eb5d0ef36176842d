"""configs[4] dist on one GPU (100 000 queries x 10 000 references, s = 10 000, k = 32; panels as in bench.py): step time and,
under ncu, the launch list.  usage: python profiles/r02_c5_prof.py [n_queries]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import __graft_entry__ as g
g._paths()
import fpmash_b200 as fpm
dev = torch.device("cuda", 0)
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nr, s5 = 10000, 10000
ctx = fpm.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
qp = bench.gen_panel_rows(torch, 0, nq, s5, dev, 52, 100, core_seed=5)
rp = bench.gen_panel_rows(torch, 0, nr, s5, dev, 51, 100, core_seed=5)
qs = torch.full((nq,), s5, dtype=torch.int32, device=dev); ql = torch.full((nq,), 5_000_000, dtype=torch.int64, device=dev)
rs = torch.full((nr,), s5, dtype=torch.int32, device=dev); rl = torch.full((nr,), 5_000_000, dtype=torch.int64, device=dev)
out = torch.empty(nq * nr * 24, dtype=torch.uint8, device=dev)
step = lambda: ctx.dist_tile_dev((rp.data_ptr(), rs.data_ptr(), rl.data_ptr(), nr, s5), (qp.data_ptr(), qs.data_ptr(), ql.data_ptr(), nq, s5), s5, 32, 4.0 ** 32, out.data_ptr())
step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print("C5 dist %d x %d: %.2f ms  digest %s" % (nq, nr, e0.elapsed_time(e1), [int(x) for x in bench.pair_digest(torch, out, nq * nr).tolist()]))
torch.cuda.profiler.start(); step(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
