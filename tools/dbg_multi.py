"""debug: row-partitioned two-loop vs compact inverse and device vs host-buffer forward apply; prints norms and differing counts"""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import torch, torch.distributed as dist
import linearoperators_jl_b200 as lo
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = lo.default_context(lr)
ctx.init_comm_from_torch(); ctx.connect_mailbox()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10**6
for mem, inverse in ((10, False), (20, True)):
    op = lo.LBFGSOperator(n, mem=mem, inverse=inverse, ctx=ctx)
    for i in range(mem):
        s = ctx.uniform(n, 1000 * rank + 100 + i); y = s + 0.1 * ctx.uniform(n, 1000 * rank + 200 + i)
        lo.push_(op, s, y)
    x = ctx.uniform(n, 1000 * rank + 7)
    r1 = op * x
    if inverse:
        op.set_option("inverse_mode", 1)
        r2 = op * x
    else:
        xh, rh = ctx.host_empty(n), ctx.host_empty(n)
        xh.copy_(x.cpu()); op.apply_host(rh, xh); r2 = rh.cuda()
    d = (r1 - r2)
    print("rank", rank, "inverse", inverse, "norm r1 %.6e" % float(r1.norm()), "norm r2 %.6e" % float(r2.norm()), "norm diff %.3e" % float(d.norm()),
          "ndiff", int((d != 0).sum()), "r1[:3]", r1[:3].tolist(), "gamma", op.data.scaling_factor, flush=True)
dist.barrier(); dist.destroy_process_group()
