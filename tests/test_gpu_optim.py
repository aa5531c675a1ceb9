"""GPU parity of the fused clip + AdamW step against torch.nn.utils.clip_grad_norm_ +
torch.optim.AdamW with the reference's three learning-rate groups (main.py:253-277, engine.py:77-81)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Toy(torch.nn.Module):
    """Parameter names chosen to hit every lr group and the skip rule; shapes cover matrices (planes), vectors
    whose numel is not a multiple of 4 (scalar tail) and a tensor larger than one 4096-element chunk."""

    def __init__(self):
        super().__init__()
        self.transformer = torch.nn.Module()
        self.transformer.reference_points = torch.nn.Linear(256, 2)          # never gets a gradient: skipped
        self.transformer.sampling_offsets = torch.nn.Linear(256, 512)        # lr * lr_linear_proj_mult
        self.transformer.linear1 = torch.nn.Linear(256, 1024)
        self.translation_head = torch.nn.Linear(256, 66)                      # bias numel 66: tail of 2
        self.backbone = torch.nn.ModuleList([torch.nn.Linear(8, 3)])         # "backbone.0": lr_backbone, numel 24 / 3


def test_fused_clip_adamw_matches_torch():
    from poet_b200 import ops
    from poet_b200.data_parallel import FlatGradReducer
    from poet_b200.optim import FusedClipAdamW
    torch.manual_seed(3)
    ours = _Toy().to(DEV)
    ref = copy.deepcopy(ours)
    red = FlatGradReducer(ours.parameters())
    kw = dict(lr=2e-3, weight_decay=1e-2, max_norm=0.1, lr_backbone=2e-4, lr_linear_proj_mult=0.1)
    opt = FusedClipAdamW(ours, red, **kw)

    def in_group(n, keys):
        return any(k in n for k in keys)
    named = list(ref.named_parameters())
    groups = [
        {"params": [p for n, p in named if not in_group(n, ("backbone.0",)) and not in_group(n, ("reference_points", "sampling_offsets"))], "lr": kw["lr"]},
        {"params": [p for n, p in named if in_group(n, ("backbone.0",))], "lr": kw["lr_backbone"]},
        {"params": [p for n, p in named if in_group(n, ("reference_points", "sampling_offsets"))], "lr": kw["lr"] * kw["lr_linear_proj_mult"]},
    ]
    topt = torch.optim.AdamW(groups, lr=kw["lr"], weight_decay=kw["weight_decay"])
    g = torch.Generator(device="cpu").manual_seed(5)
    for step in range(3):
        red.zero()
        scale = 10.0 if step == 0 else 1e-3                     # step 0 clips (norm >> 0.1), later steps do not
        for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
            if "reference_points" in n:
                q.grad = None                                   # unused on the path: the reference never touches it
                continue
            gr = (torch.randn(p.shape, generator=g) * scale).to(DEV)
            p.grad.copy_(gr)
            q.grad = gr.clone()
        total = torch.nn.utils.clip_grad_norm_([q for _, q in ref.named_parameters()], kw["max_norm"])
        topt.step()
        opt.step()
        torch.cuda.synchronize()
        assert abs(float(opt.grad_norm()) - float(total)) <= 1e-5 * float(total)
        for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
            err = float((p - q).abs().max())
            assert err <= 2e-6 * max(1.0, float(q.abs().max())), f"step {step} {n}: {err:.3e}"
    # the planes arena now holds the split of the UPDATED matrices, bit for bit what poet_split_bf16 gives
    pl = opt.planes
    assert pl is not None
    for p in pl.params:
        v = pl.lookup(p.data_ptr(), p.numel())
        hi, lo = v[0].view(p.shape), v[1].view(p.shape)
        assert torch.equal(hi, p.detach().to(torch.bfloat16))
        assert torch.equal(lo, (p.detach() - hi.float()).to(torch.bfloat16))
    # and a planes_scope right after the step does not launch another split
    before = ops.launch_count()
    with ops.planes_scope(ours):
        pass
    assert ops.launch_count() == before
