"""GraphedTrainStep: the reference's optimisation step (train_advanced.py:322-346) captured into a CUDA graph must train exactly
like the eager step -- same loss curve, same parameters, learning-rate schedule honoured between replays."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vit_oracle as vo  # checker only


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 5e-3)])
def test_graphed_step_matches_eager_step(precision, tol):
    import vit_spoof_detection_pda_b200 as pkg
    dev = torch.device("cuda:0")
    ref = vo.OracleViTFaceAntiSpoofing(dropout=0.0, depth=2)
    vo.seeded_init_(ref, seed=5)

    def make():
        m = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=2, precision=precision)
        m.load_state_dict(ref.state_dict())
        m = m.to(dev).train()
        opt = pkg.FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-4, adamw=False, capturable=True)
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=8, eta_min=1e-5)
        return m, opt, sched

    crit = pkg.FocalLoss(0.25, 2.0)
    batches = [tuple(t.to(dev) for t in vo.synthetic_batch(4, seed=200 + i)) for i in range(6)]
    # eager (capturable optimizer, no graph)
    m_e, o_e, s_e = make()
    losses_e = []
    for x, y in batches:
        loss, _ = crit(m_e(x), y, with_metrics=True)
        loss.backward()
        pkg.clip_grad_norm_(m_e.parameters(), 1.0)
        o_e.step()
        o_e.zero_grad(set_to_none=True)
        s_e.step()
        losses_e.append(float(loss))
    # graphed
    m_g, o_g, s_g = make()
    p_before = m_g.flat_params().clone()
    step = pkg.GraphedTrainStep(m_g, crit, o_g, batches[0][0], batches[0][1], max_grad_norm=1.0)
    assert torch.equal(m_g.flat_params(), p_before) and o_g._step == 0      # construction has no training side effect
    losses_g = []
    for x, y in batches:
        loss, met = step(x, y)
        s_g.step()
        losses_g.append(float(loss))
    assert o_g._step == len(batches) and int(o_g._step_dev.item()) == len(batches)
    for a, b in zip(losses_e, losses_g):
        assert abs(a - b) <= tol * max(1.0, abs(a)), (losses_e, losses_g)
    pe, pg = m_e.flat_params(), m_g.flat_params()
    # Adam's update is sign-like for tiny gradients, and the split-K weight gradients are summed with atomics (order varies from
    # run to run): individual elements may differ by a whole step, so the comparison is in the 2-norm of the total movement
    err = float((pe - pg).norm() / (pe - p_before).norm())
    assert err <= (2e-2 if precision == "fp32" else 1e-1), err
    # the graphed model's bf16 shadow follows its masters
    if precision == "bf16":
        assert torch.equal(m_g.flat_params16(), pg.to(torch.bfloat16))
