"""Data-parallel parity worker (one process per rank; launched by tests/test_dp_parity.py or by hand under torchrun):

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/dp_worker.py nccl 12      # one GPU per rank, NCCL
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/dp_worker.py gloo 3       # both ranks on cuda:0, gloo

Checks, for the fp32 validation path and the bf16 path:
  1. averaged gradients of the DataParallel model on per-rank half-batches == single-GPU gradients of the concatenated
     batch, and every rank holds bit-identical gradients;
  2. gradient accumulation (two backwards without zero_grad) reduces BOTH micro-steps (the hooks used to
     be skipped while .grad aliased the flat buffer, so ranks silently diverged);
  3. the same after zero_grad(set_to_none=False).
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vit_spoof_detection_pda_b200 as pkg  # noqa: E402
from oracle import vit_oracle as vo  # noqa: E402  (checker)

backend = sys.argv[1] if len(sys.argv) > 1 else "nccl"
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]) if backend == "nccl" else 0)
torch.cuda.set_device(dev)
if backend == "nccl":
    dist.init_process_group("nccl", device_id=dev)
else:
    dist.init_process_group("gloo")
per = 4
ref = vo.OracleViTFaceAntiSpoofing(dropout=0.0, depth=depth)
vo.seeded_init_(ref, seed=42)
images, labels = vo.synthetic_batch(2 * per * world, seed=3)   # two micro-steps of per * world samples
crit = pkg.FocalLoss(0.25, 2.0)          # mean over the local batch; DataParallel averages over ranks


def micro(k, r):
    lo = (k * world + r) * per
    return images[lo:lo + per].to(dev), labels[lo:lo + per].to(dev)


def full(k):
    lo = k * world * per
    return images[lo:lo + world * per].to(dev), labels[lo:lo + world * per].to(dev)


def grads_of(m):
    return torch.cat([p.grad.reshape(-1) for p in m.parameters()])


def check(tag, got, want, tol):
    err = float((got - want).abs().max() / want.abs().max())
    other = got.clone()
    dist.broadcast(other, src=0)
    same = bool(torch.equal(other, got))
    print(f"rank {rank} [{tag}] rel err DP vs single-GPU {err:.3e}; ranks identical {same}", flush=True)
    assert err < tol and same, (tag, err, same)


for precision, tol in (("fp32", 2e-4), ("bf16", 5e-2)):
    m = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=depth, precision=precision)
    m.load_state_dict(ref.state_dict())
    m = m.to(dev).train()
    single = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=depth, precision=precision)
    single.load_state_dict(ref.state_dict())
    single = single.to(dev).train()
    net = pkg.DataParallel(m, bucket_mb=50.0 if depth >= 12 else 8.0)

    # single-GPU references: micro-step 0, micro-steps 0 + 1 accumulated
    crit(single(*full(0)[:1]), full(0)[1]).backward()
    want0 = grads_of(single).clone()
    crit(single(*full(1)[:1]), full(1)[1]).backward()
    want01 = grads_of(single).clone()

    # 1. one backward
    x, y = micro(0, rank)
    crit(net(x), y).backward()
    check(f"{precision} single backward, {len(net.bucketer.launched)} buckets", grads_of(m), want0, tol)
    # 2. accumulation: second backward without zero_grad (.grad aliases the flat buffer)
    x, y = micro(1, rank)
    crit(net(x), y).backward()
    check(f"{precision} accumulated 2 micro-steps", grads_of(m), want01, tol)
    # 3. zero_grad(set_to_none=False): .grad stays a view of the flat buffer
    for p in m.parameters():
        p.grad.zero_()
    x, y = micro(0, rank)
    crit(net(x), y).backward()
    check(f"{precision} after zero_grad(set_to_none=False)", grads_of(m), want0, tol)
    del net, m, single
dist.barrier()
dist.destroy_process_group()
print("dp_worker ok")
