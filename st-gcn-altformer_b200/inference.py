"""Fixed-shape inference with the eval forward captured in a CUDA graph.

The eval forward of one ST_GCN_AltFormer is ~95 kernel launches; at the reference's own evaluation batch (32 sequences,
BASELINE configs[0]) every kernel is far shorter than its launch, so the pass is bound by the host enqueueing work.  Capturing the
forward once and replaying it removes that: the serving-side counterpart of the trainer's captured step (trainer.py).
Semantics are the module's own eval forward (model.eval(), torch.no_grad(); reference: the test loop of
SHREC/ST_TS/train_sttran.py:196-214)."""
import torch

from . import functional as AF


class GraphedInference:
    def __init__(self, model, example):
        """model: a CUDA module of this package (or any callable of them, e.g. lambda x: streams.ensemble_forward(x, models));
        example: a CUDA tensor (or pinned host tensor) with the shape / dtype every later call will have."""
        self.model = model
        if hasattr(model, "eval"):
            model.eval()
        dev = example.device if example.is_cuda else next(model.parameters()).device
        self.device = dev
        self._x = example.to(dev, copy=True)
        with torch.no_grad():
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):          # warm-up outside capture: lazy inits, smem attributes, derived-weight caches
                for _ in range(2):
                    self.model(self._x)
            torch.cuda.current_stream(dev).wait_stream(side)
            AF.bump_weights_epoch()                # derived weights are re-made inside the graph
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._y = self.model(self._x)

    def __call__(self, x):
        """x: same shape as the example (device or pinned host tensor; copied asynchronously into the graph's input).
        Returns the graph's static output buffer -- valid until the next call; clone to keep it."""
        if tuple(x.shape) != tuple(self._x.shape):
            raise RuntimeError(f"captured forward has input shape {tuple(self._x.shape)}, got {tuple(x.shape)}")
        self._x.copy_(x, non_blocking=True)
        self._graph.replay()
        return self._y
