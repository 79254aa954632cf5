"""GPU parity: FusedClipAdamW (2 launches) vs torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW on identical
parameters / gradients, several steps, channels_last and contiguous layouts, clipping active and inactive."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scale", [10.0, 1e-3])
def test_fused_clip_adamw_matches_torch(scale):
    from b200seg.optim import FusedClipAdamW
    g = torch.Generator(device="cuda").manual_seed(7)
    shapes = [(64, 3, 3, 3), (128, 64, 3, 3), (32, 64, 1, 1), (128,), (1, 64, 1, 1), (300000,)]
    ref, mine = [], []
    for i, s in enumerate(shapes):
        t = torch.randn(s, device="cuda", generator=g)
        if len(s) == 4 and i % 2 == 1:
            t = t.contiguous(memory_format=torch.channels_last)
        ref.append(torch.nn.Parameter(t.clone(memory_format=torch.preserve_format)))
        mine.append(torch.nn.Parameter(t.clone(memory_format=torch.preserve_format)))
    o_ref = torch.optim.AdamW(ref, lr=1e-2, weight_decay=5e-4)
    o_mine = FusedClipAdamW(mine, lr=1e-2, weight_decay=5e-4, max_norm=1.0)
    for step in range(4):
        grads = [torch.randn(p.shape, device="cuda", generator=g) * scale for p in ref]
        for p, q, gr in zip(ref, mine, grads):
            p.grad = gr.clone()
            # the conv ops hand back channels_last-strided gradients: exercise the layout fix-up path too
            q.grad = gr.clone().contiguous(memory_format=torch.channels_last) if gr.dim() == 4 else gr.clone()
        tn = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        o_ref.step()
        o_mine.step()
        assert abs(float(o_mine.total_norm) - float(tn)) <= 1e-4 * float(tn)
        for p, q in zip(ref, mine):
            assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), (step, p.shape, float((p - q).abs().max()))
        if step == 1:
            o_ref.param_groups[0]["lr"] = o_mine.param_groups[0]["lr"] = 3e-3      # scheduler-style LR change
