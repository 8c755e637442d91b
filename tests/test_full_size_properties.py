"""Size-independent properties at the BASELINE sizes (ViT-B/16 depth 12, batch 64 = 12,608 token rows, and batch 256),
where the CPU oracle would take minutes per case:
  * batch invariance -- a sample's logits do not depend on what else is in the batch or where it sits (every kernel
    works per token row / per (image, head) item; ragged tile tails, partial CTA-pair tiles and the persistent item
    schedules must not leak across images);
  * data-parallel additivity -- the mean-loss gradient of a batch is the mean of the gradients of its two halves (the
    identity the bucketed all-reduce relies on), to bf16 accumulation noise;
  * bf16 path against the fp32 validation path of the same module (itself pinned to the oracle in test_gpu_model.py)."""
import pytest
import torch

from oracle import vit_oracle as vo

pytestmark = pytest.mark.gpu


def _models():
    import vit_spoof_detection_pda_b200 as pkg
    dev = torch.device("cuda:0")
    ref = vo.OracleViTFaceAntiSpoofing(dropout=0.0, depth=12)
    vo.seeded_init_(ref, seed=42)
    out = {}
    for prec in ("bf16", "fp32"):
        m = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=12, precision=prec)
        m.load_state_dict(ref.state_dict())
        out[prec] = m.to(dev)
    return pkg, dev, out


def test_batch_invariance_and_bf16_vs_fp32_at_full_size():
    pkg, dev, ms = _models()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(256, 3, 224, 224, generator=g).to(dev)
    with torch.no_grad():
        m = ms["bf16"].eval()
        full = m(x)                       # batch 256: 50,432 token rows
        b64 = m(x[:64])                   # the training shape
        perm = torch.randperm(64, generator=g).to(dev)
        shuffled = m(x[:64][perm])
        alone = torch.cat([m(x[i:i + 1]) for i in (0, 17, 63)])
        f32 = ms["fp32"].eval()(x[:64])
    assert torch.equal(full[:64], b64)                      # bit-exact: no cross-image leakage, no shape-dependent math
    assert torch.equal(shuffled, b64[perm])
    assert torch.equal(alone, b64[[0, 17, 63]])
    assert float((b64 - f32).abs().max()) <= 2e-2           # the north-star bf16 tolerance
    assert torch.equal(b64.argmax(1)[(f32[:, 1] - f32[:, 0]).abs() > 4e-2], f32.argmax(1)[(f32[:, 1] - f32[:, 0]).abs() > 4e-2])


def test_gradient_additivity_over_batch_halves_at_full_size():
    pkg, dev, ms = _models()
    m = ms["bf16"].train()
    crit = pkg.FocalLoss(0.25, 2.0)
    images, labels = vo.synthetic_batch(64, seed=77)
    images, labels = images.to(dev), labels.to(dev)

    def grads(x, y):
        for p in m.parameters():
            p.grad = None
        crit(m(x), y).backward()
        return m.flat_grads().clone()

    g_full = grads(images, labels)
    g_half = 0.5 * (grads(images[:32], labels[:32]) + grads(images[32:], labels[32:]))
    err = float((g_full - g_half).abs().max() / g_full.abs().max())
    assert err <= 2e-2, err
