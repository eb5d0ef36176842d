"""Quick configs[2] dist timing on one GPU: merge-all and pruned step, plus an order-independent digest of the result matrix
(so that kernel variants -- FPMASH_B200_LIB=... -- can be compared for speed AND equality)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import __graft_entry__ as g
g._paths()
import fpmash_b200 as fpm
dev = torch.device("cuda", 0)
nd, S, K = 20000, 1000, 21
ctx = fpm.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
panel = bench.gen_sketch_panel(torch, nd, S, dev, seed=3)
sizes = torch.full((nd,), S, dtype=torch.int32, device=dev)
lengths = torch.full((nd,), 5_000_000, dtype=torch.int64, device=dev)
ptrs = (panel.data_ptr(), sizes.data_ptr(), lengths.data_ptr(), nd, S)
out = torch.empty(nd * nd * 24, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, mode in (("merge_all", {"no_prune": True}), ("pruned", {})):
    ctx.set_dist_mode(**mode)
    for _ in range(2):
        ctx.dist_tile_dev(ptrs, ptrs, S, K, 4.0 ** K, out.data_ptr())
    torch.cuda.synchronize()
    ctx.set_timing(True)
    ts = []
    for _ in range(5):
        e0.record(); ctx.dist_tile_dev(ptrs, ptrs, S, K, 4.0 ** K, out.data_ptr()); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    tile = ctx.get_timing(fpm.KERNEL_DIST_TILE); pack = ctx.get_timing(fpm.KERNEL_DIST_PACK)
    ctx.set_timing(False)
    print("%-9s step %.3f ms (min of 5)  tiles %.3f  prepass %.3f  digest %s" % (name, min(ts), tile[0] / 5, pack[0] / 5,
          [int(x) for x in bench.pair_digest(torch, out, nd * nd).tolist()]))
ctx.set_dist_mode()
# the same collection with relatives next to each other (rows sorted by family), as a taxonomy-ordered sketch file would have them:
# the default leaves such panels in their own order; group=True regroups them anyway
order = torch.argsort(torch.arange(nd, device=dev) % 20, stable=True)
panel_o = panel[order].contiguous()
ptrs_o = (panel_o.data_ptr(), sizes.data_ptr(), lengths.data_ptr(), nd, S)
for name, mode in (("ordered", {}), ("ordered, regrouped anyway", {"group": True})):
    ctx.set_dist_mode(**mode)
    for _ in range(2):
        ctx.dist_tile_dev(ptrs_o, ptrs_o, S, K, 4.0 ** K, out.data_ptr())
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0.record(); ctx.dist_tile_dev(ptrs_o, ptrs_o, S, K, 4.0 ** K, out.data_ptr()); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%-26s step %.3f ms (min of 5)  digest %s" % (name, min(ts), [int(x) for x in bench.pair_digest(torch, out, nd * nd).tolist()]))
ctx.set_dist_mode()
