"""GPU parity of the on-device pose loss against the oracle and the reference fixture."""
import pytest
import torch

from helpers import load_golden
from oracle import poet_oracle as O
from oracle.make_golden import criterion_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_pose_criterion_matches_reference_and_oracle():
    from poet_b200.criterion import PoseCriterion
    g = load_golden("criterion/gt")
    t_all, R_all, _boxes, _labels, tgt_t, tgt_R, n_boxes = criterion_case()
    L, B, Q = t_all.shape[:3]
    T = max(n_boxes)
    pt, pR = torch.zeros(B, T, 3), torch.zeros(B, T, 3, 3)
    for b, n in enumerate(n_boxes):
        pt[b, :n], pR[b, :n] = tgt_t[b], tgt_R[b]
    td, Rd = t_all.to(DEV).requires_grad_(True), R_all.to(DEV).requires_grad_(True)
    outputs = {"pred_translation": td[-1], "pred_rotation": Rd[-1],
               "aux_outputs": [{"pred_translation": td[l], "pred_rotation": Rd[l]} for l in range(L - 1)]}
    crit = PoseCriterion(g["weights"])
    losses, total = crit(outputs, pt.to(DEV), pR.to(DEV), torch.tensor(n_boxes, dtype=torch.int32, device=DEV))
    total.backward()
    assert set(losses) == set(g["losses"])
    for k, v in g["losses"].items():                                   # vs the unmodified reference
        assert abs(float(losses[k]) - v) <= 5e-6 * max(1.0, abs(v)), k
    assert abs(float(total) - g["total"]) <= 5e-6 * abs(g["total"])
    # gradients vs the fp64 oracle (the reference fixture's fp32 acos' is noisy near the clamp)
    t64, R64 = t_all.double().requires_grad_(True), R_all.double().requires_grad_(True)
    _, tot64 = O.pose_criterion_gt(t64, R64, [t.double() for t in tgt_t], [r.double() for r in tgt_R], n_boxes,
                                   g["weights"]["loss_trans"], g["weights"]["loss_rot"])
    tot64.backward()
    assert float((td.grad.cpu().double() - t64.grad).abs().max()) < 1e-6
    assert float((Rd.grad.cpu().double() - R64.grad).abs().max()) < 1e-4 * float(R64.grad.abs().max())
    assert float((td.grad.cpu() - g["grad_t"]).abs().max()) < 1e-6
