"""2-rank data-parallel parity check (run under torchrun): averaged gradients of the DataParallel model on
two half-batches must equal the single-GPU gradients on the concatenated batch; both ranks must agree."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vit_spoof_detection_pda_b200 as pkg  # noqa: E402
from oracle import vit_oracle as vo  # noqa: E402  (checker)

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
depth, per = 12, 4
ref = vo.OracleViTFaceAntiSpoofing(dropout=0.0, depth=depth)
vo.seeded_init_(ref, seed=42)
images, labels = vo.synthetic_batch(per * world, seed=3)

for precision, tol in (("fp32", 2e-4), ("bf16", 5e-2)):
    m = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=depth, precision=precision)
    m.load_state_dict(ref.state_dict())
    m = m.to(dev).train()
    net = pkg.DataParallel(m, bucket_mb=50.0)
    crit = pkg.FocalLoss(0.25, 2.0)          # mean over the local batch; DataParallel averages over ranks
    sl = slice(rank * per, (rank + 1) * per)
    loss = crit(net(images[sl].to(dev)), labels[sl].to(dev))
    loss.backward()
    g_dp = m.flat_grads().clone()
    # single-GPU gradient on the full batch (same model object, hooks disabled)
    m._bucket_hook = None
    m._finish_hook = None
    for p in m.parameters():
        p.grad = None
    loss_full = crit(m(images.to(dev)), labels.to(dev))
    loss_full.backward()
    g_full = m.flat_grads().clone()
    err = float((g_dp - g_full).abs().max() / g_full.abs().max())
    other = g_dp.clone()
    dist.broadcast(other, src=0)
    same = bool(torch.equal(other, g_dp))
    print(f"rank {rank} [{precision}] buckets {net.bucketer.launched[:3]}... n={len(net.bucketer.launched)} "
          f"rel err DP-avg vs full-batch {err:.3e} ranks identical {same}", flush=True)
    assert err < tol and same
dist.barrier()
dist.destroy_process_group()
print("dp_check ok")
