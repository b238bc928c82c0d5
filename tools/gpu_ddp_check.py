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
    steps = 2
    for _ in range(steps):
        loss, _ = tr.step(x[lo:hi].to(dev), y[lo:hi].to(dev))
    torch.cuda.synchronize()
    # 1. identical parameters on every rank
    mine = tr.flat_p.clone()
    ref0 = mine.clone()
    dist.broadcast(ref0, src=0)
    same = bool(torch.equal(mine, ref0))
    flags = torch.tensor([int(same)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ok_same = bool(flags.item())
    ok_ref = True
    if rank == 0:
        # 2. single-GPU recomputation: gradient of each shard in turn, averaged, same AdamW kernel
        ref = ab.DataParallelTrainer(build(state, T, V, cls, dev), use_graph=False)
        # per-shard BatchNorm running stats diverge between ranks by design; compare parameters only
        for _ in range(steps):
            acc = torch.zeros_like(ref.flat_g)
            bn_backup = {k: v.clone() for k, v in ref.model.state_dict().items() if "running" in k}
            for r in range(world):
                a, b = shard_range(N, r, world)
                ref.model.load_state_dict(bn_backup, strict=False)
                ref._fwd_bwd(x[a:b].to(dev), y[a:b].to(dev))
                acc += ref.flat_g
            ref.flat_g.copy_(acc)
            h = ref.hp
            ops.adamw(ref.flat_p, ref.flat_g, ref.flat_m, ref.flat_v, ref.flat_lowp, ref.step_count, h["lr"], h["b1"], h["b2"], h["eps"],
                      h["wd"], 1.0 / world)
            AF.bump_weights_epoch()
        torch.cuda.synchronize()
        d = (mine.double() - ref.flat_p.double())
        rel = float(d.norm() / ref.flat_p.double().norm())
        # scale of the update itself (lr 2e-4 * ~1 per element): differences must be far below it
        moved = float((ref.flat_p.double() - torch.nan_to_num(ref.flat_p.double() * 0)).norm())  # noqa: F841
        print(f"ddp-check: world={world} loss={float(loss):.4f} identical_across_ranks={ok_same} "
              f"rel_l2(params vs 1-GPU recompute)={rel:.3e} max_abs={float(d.abs().max()):.3e}")
        ok_ref = rel < 1e-4   # fp32 atomics reorder sums and AdamW normalises tiny gradients: loose bound
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
