"""Golden vectors for the device augmentation (round 2; run here, where /root/reference exists).

  augment_22.pt   Hand_Dataset.data_aug (data_process/Hand_Dataset.py:84-157) called on the UNMODIFIED reference class, once
                  per transform, with the module's random sources (`randint`, `shuffle`, `np.random.uniform`) wrapped so the
                  draws are recorded: inputs, (kind, params) in the layout afb_augment takes, and the reference's outputs.

    python tests/golden/make_golden_aug.py
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    refshim.load()
    mod = importlib.import_module("data_process.Hand_Dataset")
    T, V = 16, 22
    ds = mod.Hand_Dataset(data=[], time_len=T, use_data_aug=True, expand=False)
    g = torch.Generator().manual_seed(11)
    xs, kinds, params, outs = [], [], [], []
    real_uniform, real_shuffle = np.random.uniform, mod.shuffle
    for case, kind in enumerate([0, 1, 2, 3, 2, 3, 0, 1]):
        sk = (0.3 * torch.randn(T, V, 3, generator=g)).double().numpy()
        draws, order = [], []

        def uniform(low, high, size=None, _d=draws):
            v = real_uniform(low, high, size)
            _d.append(np.atleast_1d(np.asarray(v, dtype=np.float64)).copy())
            return v

        def shuffle(lst, _o=order):
            real_shuffle(lst)
            _o.extend(lst)

        np.random.seed(100 + case)
        mod.random.seed(200 + case)
        mod.randint = lambda a, b, _k=kind: _k          # the transform under test (the reference draws randint(0, 3))
        mod.shuffle = shuffle
        np.random.uniform = uniform
        try:
            out = ds.data_aug(sk.copy())
        finally:
            np.random.uniform = real_uniform
            mod.shuffle = real_shuffle
        p = np.zeros(16)
        if kind == 0:
            p[0] = draws[0][0]
        elif kind == 1:
            p[:3] = draws[0]
        elif kind == 2:
            p[:4] = order[:4]
            for q in range(4):
                p[4 + 3 * q: 7 + 3 * q] = draws[q]
        else:
            p[0] = draws[0][0]
        xs.append(torch.from_numpy(sk).float())
        kinds.append(kind)
        params.append(torch.from_numpy(p).float())
        outs.append(torch.from_numpy(np.asarray(out)).float())
        assert outs[-1].shape == (T, V, 3), outs[-1].shape
    torch.save({"x": torch.stack(xs), "kind": torch.tensor(kinds, dtype=torch.int32), "params": torch.stack(params), "y": torch.stack(outs)},
               os.path.join(OUT, "augment_22.pt"))
    print("augment_22:", torch.stack(xs).shape, kinds)


if __name__ == "__main__":
    main()
