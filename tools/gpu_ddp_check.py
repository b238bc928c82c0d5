"""Multi-GPU parity of the data-parallel trainer (run under torchrun, one rank per GPU, NCCL):

  * every rank steps on its own shard of a global batch (per-rank BatchNorm statistics, as DataParallel);
  * afterwards all ranks must hold bit-identical parameters;
  * rank 0 re-computes the same update on ONE GPU (each shard's gradient in turn, averaged, same fused AdamW)
    and must match the distributed result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/gpu_ddp_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import altformer_b200 as ab  # noqa: E402
from altformer_b200 import functional as AF, ops  # noqa: E402
from altformer_b200.trainer import shard_range  # noqa: E402
from oracle import altformer_oracle as O  # noqa: E402  (synthetic batch + seeded weights only)


def build(state, T, V, cls, dev):
    m = ab.ST_GCN_AltFormer(3, cls, num_frame=T, num_joints=V, style="ST", graph="graph.SHRE", graph_args={"labeling_mode": "spatial"})
    m.load_state_dict(state)
    for mod in m.modules():
        if type(mod).__name__ == "DropPath":
            mod.drop_prob = 0.0
    return m.to(dev).train()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    N, T, V, cls = 8 * world, 32, 22, 28
    state = O.random_state(O.model_spec(3, cls, T, V), 3)
    x, y = O.synthetic_batch(N, T, V, cls, 77)
    lo, hi = shard_range(N, rank, world)
    tr = ab.DataParallelTrainer(build(state, T, V, cls, dev), use_graph=False)
    loss, _ = tr.step(x[lo:hi].to(dev), y[lo:hi].to(dev))
    torch.cuda.synchronize()
    # 1. identical parameters (and summed gradients) on every rank
    mine = torch.cat([tr.flat_p, tr.flat_g])
    ref0 = mine.clone()
    dist.broadcast(ref0, src=0)
    flags = torch.tensor([int(torch.equal(mine, ref0))], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ok_same = bool(flags.item())
    ok_ref = True
    if rank == 0:
        # 2. single-GPU recomputation of the all-reduced gradient: each shard's gradient in turn, summed.
        #    (Parameters after AdamW are ill-conditioned to compare: the first update is +-lr whatever |g| is,
        #    so a sign flip of a ~0 gradient moves a weight by 2*lr; the gradient itself is the right check.)
        ref = ab.DataParallelTrainer(build(state, T, V, cls, dev), use_graph=False)
        acc = torch.zeros_like(ref.flat_g)
        bn_backup = {k: v.clone() for k, v in ref.model.state_dict().items() if "running" in k or "num_batches" in k}
        for r in range(world):
            a, b = shard_range(N, r, world)
            ref.model.load_state_dict(bn_backup, strict=False)   # every rank starts from the same BN buffers
            ref._fwd_bwd(x[a:b].to(dev), y[a:b].to(dev))
            acc += ref.flat_g
        torch.cuda.synchronize()
        d = tr.flat_g.double() - acc.double()
        rel = float(d.norm() / acc.double().norm())
        print(f"ddp-check: world={world} loss={float(loss):.4f} identical_across_ranks={ok_same} "
              f"rel_l2(all-reduced grad vs 1-GPU per-shard recompute)={rel:.3e} max_abs={float(d.abs().max()):.3e}")
        ok_ref = rel < 1e-4
    flag = torch.tensor([int(ok_ref)], device=dev)
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if not (ok_same and bool(flag.item())):
        print("ddp-check: FAIL")
        sys.exit(1)
    if rank == 0:
        print("ddp-check: PASS")


if __name__ == "__main__":
    main()
