"""Host-side logic of the pooled DropPath masks (model_ST.prefill_drop_paths): one draw per forward, queued per Block in
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
