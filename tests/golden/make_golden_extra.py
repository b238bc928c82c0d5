"""Round-2 additions to the golden set (run here, where /root/reference exists; the GPU box never runs this):

  model_ST_22_refckpt.pt   what a checkpoint TRAINED BY THE REFERENCE sees: the reference module tree exactly as its
                           constructor leaves it -- `PA = nn.Parameter(A)` + `constant_(PA, 1e-6)` overwrite the adjacency
                           (model/unit_agcn.py:36-38; asserted below) -- with only the de-aliasing that `model.cuda()`
                           performs (the parameter gets its own storage, `self.A` stays the 1e-6 constant).  Eval-mode logits
                           for a seeded state_dict; tests load the same state with 'module.'-prefixed keys
                           (SHREC/ST_TS/emsemble.py:99-104) into the CUDA model.
  unit2d_dim3_*.pt         the reference's Unit2D(dim=3) (model/net.py:29-36), stride 1 and 2, train mode: y, dx, parameter gradients,
                           running statistics.
  streams_22.pt            Hand_Dataset.motion / Hand_Dataset.bone (data_process/Hand_Dataset.py:183-217) and the palm-centre
                           normalisation (:61) on a seeded skeleton, called on the reference class itself.

    python tests/golden/make_golden_extra.py
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import altformer_oracle as O  # noqa: E402
from oracle import refshim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = refshim.load()
    torch.manual_seed(0)
    N, T, V, cls, seed = 2, 8, 22, 14, 71
    mod = ref.ST_GCN_AltFormer(channel=3, num_class=cls, num_frame=T, num_joints=V, style="ST", graph="graph.SHRE",
                               graph_args={"labeling_mode": "spatial"})
    # the quirk, straight from the unmodified constructor: the adjacency the module holds is 1e-6 everywhere
    assert torch.all(mod.gcn0.A == 1e-6) and mod.gcn0.A.data_ptr() == mod.gcn0.PA.data_ptr()
    mod.gcn0.A = mod.gcn0.A.clone()          # what .cuda() does: PA moves to new storage, self.A stays behind as a constant
    st = O.random_state(O.model_spec(3, cls, T, V), seed)
    res = mod.load_state_dict(st, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.all(mod.gcn0.A == 1e-6)
    mod.eval()
    x, _ = O.synthetic_batch(N, T, V, cls, seed + 100)
    with refshim.cpu_cuda_noop(), torch.no_grad():
        y = mod(x)
    torch.save({"y": y.clone(), "state_seed": seed, "batch_seed": seed + 100, "shape": (N, T, V, cls)},
               os.path.join(OUT, "model_ST_22_refckpt.pt"))
    print("model_ST_22_refckpt: logits", tuple(y.shape), float(y.norm()))

    # Unit2D(dim=3): 1 x k convolution along the joints (model/net.py:29-36), stride 1 and 2, train mode, dx + parameter grads
    from tests.golden.make_golden import run_module
    for name, stride, seed in (("unit2d_dim3_train", 1, 41), ("unit2d_dim3_s2_train", 2, 42)):
        cin, cout, k, Nn, Tt, Vv = 64, 64, 3, 2, 6, 22
        spec = O.unit2d_spec("", cin, cout, k)
        spec["conv.weight"] = (cout, cin, 1, k)
        st2 = O.random_state(spec, seed)
        m2 = ref.Unit2D(cin, cout, kernel_size=k, stride=stride, dim=3)
        gx = torch.Generator().manual_seed(seed + 100)
        case = run_module(m2, st2, torch.randn(Nn, cin, Tt, Vv, generator=gx), True, True)
        case.update({"seed": seed, "shape": (cin, cout, k, Nn, Tt, Vv, stride)})
        torch.save(case, os.path.join(OUT, name + ".pt"))
        print(name, tuple(case["y"].shape))

    hd = importlib.import_module("data_process.Hand_Dataset").Hand_Dataset
    g = torch.Generator().manual_seed(5)
    sk = (0.2 * torch.randn(32, 22, 3, generator=g)).double().numpy()
    palm = sk.copy()
    palm -= palm[0][1]                                         # Hand_Dataset.py:61, verbatim semantics
    motion = hd.motion(None, palm.copy())[:32]
    bone = hd.bone(None, palm.copy())[:32]
    torch.save({"x": torch.from_numpy(sk).float(), "palm": torch.from_numpy(palm).float(),
                "motion": torch.from_numpy(np.asarray(motion)).float(), "bone": torch.from_numpy(np.asarray(bone)).float()},
               os.path.join(OUT, "streams_22.pt"))
    print("streams_22: motion", motion.shape, "bone", bone.shape)


if __name__ == "__main__":
    main()
