"""Batch data-parallel training: one process per GPU, replaces the reference's single-process
nn.DataParallel (SHREC/ST_TS/train_sttran.py:84) and its train step (:89-102,185-191):

    zero_grad -> forward -> CrossEntropyLoss -> backward -> [all-reduce mean over ranks] -> AdamW

All live parameters are re-homed into ONE flat fp32 buffer (with flat gradient / Adam-moment / bf16
shadow twins), so the gradient exchange is a single NCCL all-reduce over NVLink/NVSwitch and the
optimizer is a single fused kernel that also refreshes the bf16 GEMM operands.  Weight-gradient
kernels accumulate straight into the flat gradient buffer.  BatchNorm statistics stay per-rank, which
is what DataParallel does.  The forward+backward of a fixed shape can be captured in a
CUDA graph (`use_graph=True`) so the ~300 launches of a step cost one host call.
"""
import contextlib
import os

import torch
import torch.distributed as dist

from . import functional as AF
from . import ops


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of a global batch of n samples owned by `rank` (n % world may be != 0)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _round_up(n, m):
    return (n + m - 1) // m * m


class FlatBuffers:
    """Layout of the flat parameter space: name -> (offset, numel, shape); offsets 8-element aligned so every
    bf16 view is 16-byte aligned (TMA requirement).  Pure host logic (testable on CPU)."""

    ALIGN = 8

    def __init__(self, named_shapes):
        self.index = {}
        off = 0
        for name, shape in named_shapes:
            n = 1
            for s in shape:
                n *= s
            self.index[name] = (off, n, tuple(shape))
            off += _round_up(max(n, 1), self.ALIGN)
        self.total = _round_up(off, self.ALIGN)

    def view(self, flat, name):
        off, n, shape = self.index[name]
        return flat[off:off + n].view(shape)


_NVTX = os.environ.get("AFB_NVTX", "0") == "1"


@contextlib.contextmanager
def _nvtx(name):
    """NVTX range around a phase of the step (tracing), enabled by AFB_NVTX=1 for an nsys / ncu --nvtx
    timeline; a no-op otherwise (and inside CUDA-graph replay, where only the captured kernels run)."""
    if not _NVTX:
        yield
        return
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


class GradReducer:
    """Mean all-reduce of the flat gradient buffer.  With backend nccl this is one collective over
    NVLink/NVSwitch; the division by world size is folded into the optimizer kernel (grad_scale).

    Every rank's gradient is the gradient of ITS shard's mean loss.  The reference's DataParallel loss is the mean over
    the global batch (train_sttran.py:161,189), so with shards of unequal size rank r's gradient carries the weight
    local_n_r / global_n instead of 1 / world: `shard_weight` returns the factor to apply before the SUM."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1

    @property
    def grad_scale(self):
        return 1.0 / self.world

    def shard_weight(self, local_n, global_n):
        """Pre-reduction factor f with  f * grad_scale == local_n / global_n  (1.0 for equal shards)."""
        if global_n is None or local_n * self.world == global_n:
            return 1.0
        return float(local_n) * self.world / float(global_n)

    def reduce(self, flat_grad):
        if self.world > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        return flat_grad

    def broadcast(self, tensors, src=0):
        """Rank `src`'s values everywhere (parameters and BatchNorm buffers at construction: DataParallel replicates
        module 0 every forward, train_sttran.py:84; here it happens once)."""
        if self.world > 1:
            for t in tensors:
                dist.broadcast(t, src=src, group=self.group)


class DataParallelTrainer:
    def __init__(self, model, lr=2e-4, weight_decay=0.1, betas=(0.9, 0.999), eps=1e-8, use_graph=False, group=None,
                 reduce_dtype=None):
        """reduce_dtype: dtype of the gradient all-reduce (world > 1).  None = follow the precision mode: bf16 in bf16 mode
        (32 MB instead of 64.5 MB over NVLink; the rounding, 2^-9 relative per rank, is below the ~1e-2 noise the bf16
        activations already put on every gradient), fp32 in the fp32 parity mode -- the reference's DataParallel reduces
        fp32 gradients (train_sttran.py:84,189)."""
        dev = next(model.parameters()).device
        if not next(model.parameters()).is_cuda:
            raise RuntimeError("DataParallelTrainer needs the model on a CUDA device (there is no CPU fallback)")
        self.model, self.device = model, dev
        self.hp = dict(lr=lr, wd=weight_decay, b1=betas[0], b2=betas[1], eps=eps)
        live = model.live_parameters() if hasattr(model, "live_parameters") else list(model.named_parameters())
        self.layout = FlatBuffers([(n, p.shape) for n, p in live])
        T = self.layout.total
        z = lambda dt: torch.zeros(T, device=dev, dtype=dt)  # noqa: E731
        self.flat_p, self.flat_g, self.flat_m, self.flat_v = z(torch.float32), z(torch.float32), z(torch.float32), z(torch.float32)
        self.flat_lowp = z(torch.bfloat16)
        self.step_count = torch.ones((), device=dev, dtype=torch.int32)
        self.params = []
        with torch.no_grad():
            for name, p in live:
                dst = self.layout.view(self.flat_p, name)
                ops.cast(p.detach().contiguous(), torch.float32, out=dst)
                p.data = dst
                p._afb_grad = self.layout.view(self.flat_g, name)
                p._afb_shadow = self.layout.view(self.flat_lowp, name)
                p.grad = p._afb_grad
                self.params.append(p)
        self.reducer = GradReducer(group)
        if reduce_dtype is None:
            reduce_dtype = torch.bfloat16 if AF.get_precision() == "bf16" else torch.float32
        self.reduce_dtype = reduce_dtype
        self.flat_g_low = (torch.zeros(T, device=dev, dtype=reduce_dtype)
                           if self.reducer.world > 1 and reduce_dtype != torch.float32 else None)
        # ranks start from rank 0's weights and BatchNorm buffers (a model built per rank without a shared seed or
        # checkpoint would otherwise train diverging replicas: only gradients are exchanged afterwards)
        self.reducer.broadcast([self.flat_p] + [b for b in model.buffers() if b.is_cuda])
        self.sync_shadow()
        self.use_graph = use_graph
        self._graph = None
        self._static = None
        self.launches_per_step = None

    @torch.no_grad()
    def sync_shadow(self):
        """Re-cast the fp32 master weights into the bf16 GEMM operands.  Called at construction; call it (or rely on the
        per-parameter version check in functional.lowp) after editing parameters outside AdamW, e.g. load_state_dict."""
        ops.cast(self.flat_p, torch.bfloat16, out=self.flat_lowp)
        for p in self.params:
            p._afb_shadow_ver = p._version
        AF.bump_weights_epoch()

    def load_state_dict(self, state, strict=True):
        """model.load_state_dict + shadow refresh (the parameters live in the flat buffer: copy_ lands there)."""
        out = self.model.load_state_dict(state, strict=strict)
        self.sync_shadow()
        return out

    # -- one training step -------------------------------------------------------------------
    def _fwd_bwd(self, x, labels):
        self.flat_g.zero_()                      # model.zero_grad() (train_sttran.py:187); memset plumbing
        with _nvtx("afb.forward"):
            logits = self.model(x)
            loss = AF.cross_entropy(logits, labels)
        with _nvtx("afb.backward"):
            loss.backward()
        return loss.detach(), logits.detach()

    def _optimize(self, shard_weight=1.0):
        with _nvtx("afb.reduce+adamw"):
            self._optimize_impl(shard_weight)

    def _optimize_impl(self, shard_weight=1.0):
        if shard_weight != 1.0:
            ops.axpby(self.flat_g, shard_weight, self.flat_g, 0.0, out=self.flat_g)
        if self.flat_g_low is not None:      # reduced-precision exchange: cast, one all-reduce of half the bytes, cast back
            ops.cast(self.flat_g, self.reduce_dtype, out=self.flat_g_low)
            self.reducer.reduce(self.flat_g_low)
            ops.cast(self.flat_g_low, torch.float32, out=self.flat_g)
        else:
            self.reducer.reduce(self.flat_g)
        h = self.hp
        ops.adamw(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.flat_lowp, self.step_count, h["lr"], h["b1"],
                  h["b2"], h["eps"], h["wd"], self.reducer.grad_scale)
        AF.bump_weights_epoch()

    def step(self, x, labels, global_batch=None):
        """x (n, T, V, 3) float32 and labels (n,) int64 of THIS rank's shard: tensors on the trainer's device, or
        (pinned) host tensors, which are copied asynchronously straight into the step's input buffers.
        `global_batch`: total samples over all ranks; needed only when shards are unequal (shard_range allows it) so the
        reduced gradient is the global-batch mean.  Returns (loss, logits) as device tensors (no host sync); with
        use_graph they are the graph's static buffers, valid until the next step() — clone to keep them."""
        self.model.train()
        sw = self.reducer.shard_weight(x.shape[0], global_batch)
        if not self.use_graph:
            if not x.is_cuda:
                x, labels = x.to(self.device, non_blocking=True), labels.to(self.device, non_blocking=True)
            out = self._fwd_bwd(x, labels)
            self._optimize(sw)
            return out
        if self._graph is None:
            self._capture(x, labels)
        sx, sy, sloss, slogits = self._static
        if tuple(x.shape) != tuple(sx.shape):
            raise RuntimeError(f"captured step has batch shape {tuple(sx.shape)}, got {tuple(x.shape)} (use_graph needs a fixed shape)")
        sx.copy_(x, non_blocking=True)
        sy.copy_(labels, non_blocking=True)
        self._graph.replay()
        self._optimize(sw)
        return sloss, slogits

    def _capture(self, x, labels):
        sx, sy = x.to(self.device, copy=True), labels.to(self.device, copy=True)
        # the warm-up passes must leave no trace: BatchNorm running statistics / num_batches_tracked and the RNG stream
        # (DropPath masks) are restored, so step 1 under a graph equals step 1 without one
        bufs = [b for b in self.model.buffers()]
        saved = [b.detach().clone() for b in bufs]
        rng = torch.cuda.get_rng_state(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # warm-up outside capture (lazy inits, smem attributes, caches)
            for _ in range(2):
                self._fwd_bwd(sx, sy)
        torch.cuda.current_stream().wait_stream(side)
        with torch.no_grad():
            for b, s0 in zip(bufs, saved):
                b.copy_(s0)
        torch.cuda.set_rng_state(rng, self.device)
        AF.bump_weights_epoch()                   # force derived-weight kernels to be part of the graph
        self._graph = torch.cuda.CUDAGraph()
        ops.LAUNCHES[0] = 0
        with torch.cuda.graph(self._graph):
            loss, logits = self._fwd_bwd(sx, sy)
        self.launches_per_step = ops.LAUNCHES[0] + 2   # + AdamW and the step counter (outside the graph)
        self._static = (sx, sy, loss, logits)

    @torch.no_grad()
    def evaluate(self, x):
        self.model.eval()
        return self.model(x)
