"""Host-side random policies: the pooled DropPath masks (model_ST.prefill_drop_paths): one draw per forward, queued per Block in
the order Block.forward_rows consumes them, timm semantics (0 or 1/(1-p), reference model_ST.py:84-87)."""
import torch

from altformer_b200.model.AltFormer.model_ST import Block, DropPath, prefill_drop_paths


def _blocks(rates):
    return torch.nn.ModuleList([Block(dim=32, num_heads=4, mlp_ratio=2.0, qkv_bias=True, drop_path=r) for r in rates])


def test_masks_are_queued_per_block_with_timm_values():
    torch.manual_seed(0)
    a, b = _blocks([0.0, 0.1, 0.3]).train(), _blocks([0.2]).train()
    prefill_drop_paths([(a, 4000), (b, 50)], torch.device("cpu"))
    assert not isinstance(a[0].drop_path, DropPath)            # rate 0 -> nn.Identity, as the reference builds it
    for blk, B in ((a[1], 4000), (a[2], 4000), (b[0], 50)):
        dp = blk.drop_path
        keep = 1.0 - dp.drop_prob
        assert len(dp.queue) == 2 and all(m.shape == (B,) and m.dtype == torch.float32 for m in dp.queue)
        for m in dp.queue:
            vals = set(round(v, 5) for v in m.unique().tolist())
            assert vals <= {0.0, round(1.0 / keep, 5)}
            if B >= 1000:
                assert abs(float((m > 0).float().mean()) - keep) < 0.03
        m1 = dp.row_scale(B, torch.device("cpu"))
        m2 = dp.row_scale(B, torch.device("cpu"))
        assert m1.shape == (B,) and m2.shape == (B,) and dp.queue == []
        assert dp.row_scale(B, torch.device("cpu")).shape == (B,)   # queue exhausted -> per-call draw


def test_eval_and_pinned_masks_bypass_the_pool():
    a = _blocks([0.2, 0.2])
    a[0].drop_path.queue = [torch.ones(3)]                     # stale entries are discarded by the next prefill
    a.eval()
    prefill_drop_paths([(a, 8)], torch.device("cpu"))
    assert a[0].drop_path.queue == [] and a[0].drop_path.row_scale(8, torch.device("cpu")) is None
    a.train()
    pin = torch.full((8,), 1.25)
    a[1].drop_path.pinned = pin
    prefill_drop_paths([(a, 8)], torch.device("cpu"))
    assert a[1].drop_path.queue == [] and a[1].drop_path.row_scale(8, torch.device("cpu")) is pin
    assert len(a[0].drop_path.queue) == 2


def test_augment_policy_draws_stay_in_the_reference_ranges():
    """streams.draw_augment_params (host logic of the device augmentation): the four transforms with equal probability, factor in
    [0.8, 1.2], offsets in [-0.1, 0.1], four DISTINCT joints, r in [0, 1] (data_process/Hand_Dataset.py:84-157); applying the
    oracle's explicit-parameter augmentation with them reproduces each transform's definition."""
    from altformer_b200 import streams
    from oracle import altformer_oracle as O
    g = torch.Generator().manual_seed(4)
    N, T, V = 4000, 6, 22
    kind, p = streams.draw_augment_params(N, V, torch.device("cpu"), g)
    assert kind.dtype == torch.int32 and p.shape == (N, 16) and p.dtype == torch.float32
    counts = torch.bincount(kind.long(), minlength=4).float() / N
    assert (counts - 0.25).abs().max() < 0.03
    assert ((p[kind == 0, 0] >= 0.8) & (p[kind == 0, 0] <= 1.2)).all() and (p[kind == 0, 1:] == 0).all()
    assert (p[kind == 1, :3].abs() <= 0.1).all() and (p[kind == 1, 3:] == 0).all()
    j = p[kind == 2, :4]
    assert ((j >= 0) & (j < V) & (j == j.round())).all() and all(len(set(r.tolist())) == 4 for r in j[:200])
    assert (p[kind == 2, 4:].abs() <= 0.1).all()
    assert ((p[kind == 3, 0] >= 0) & (p[kind == 3, 0] <= 1)).all() and (p[kind == 3, 1:] == 0).all()
    x = torch.randn(8, T, V, 3, generator=g)
    y = O.augment(x, kind[:8], p[:8])
    for n in range(8):
        k = int(kind[n])
        if k == 0:
            assert torch.allclose(y[n], x[n] * p[n, 0])
        elif k == 1:
            assert torch.allclose(y[n], x[n] + p[n, :3])
        elif k == 3:
            assert torch.allclose(y[n, :-1], x[n, :-1] + p[n, 0] * (x[n, 1:] - x[n, :-1])) and torch.equal(y[n, -1], y[n, -2])
