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
dp_mode = sys.argv[3] if len(sys.argv) > 3 else "nccl"     # nccl: bucketed all-reduce under backward; nvls: fused step over multicast
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


def nvls_check(precision, tol, overlap):
    """NVLS mode reduces at the optimizer step, not per backward: compare PARAMETERS after three clipped Adam steps (the
    second with two accumulated micro-batches) with a single-GPU run on the concatenated batches.  eps = 1 makes the update
    ~ lr * g (linear in the gradient), so parameter differences measure gradient differences instead of sign flips of
    near-zero gradients."""
    m = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=depth, precision=precision)
    m.load_state_dict(ref.state_dict())
    m = m.to(dev).train()
    single = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=depth, precision=precision)
    single.load_state_dict(ref.state_dict())
    single = single.to(dev).train()
    net = pkg.DataParallel(m, mode="nvls", nvls_overlap=overlap, nvls_domains=3)
    assert net.mode == "nvls"
    o_dp = pkg.FusedAdam(m.parameters(), lr=1e-2, eps=1.0, weight_decay=1e-4, adamw=False)
    o_s = pkg.FusedAdam(single.parameters(), lr=1e-2, eps=1.0, weight_decay=1e-4, adamw=False)
    p0 = single.flat_params().clone()
    # overlap=True reduce-scatters the first domains under backward (one backward per step); overlap=False also accumulates
    for plan in (([0], [1], [0]) if overlap else ([0], [0, 1], [1])):
        for k in plan:      # micro-steps accumulate locally; ONE reduce at the step
            x, y = micro(k, rank)
            (crit(net(x), y) / len(plan)).backward()
            xs, ys = full(k)
            (crit(single(xs), ys) / len(plan)).backward()
        n_dp = pkg.clip_grad_norm_(m.parameters(), 0.05)
        n_s = pkg.clip_grad_norm_(single.parameters(), 0.05)
        o_dp.step(); o_s.step()
        o_dp.zero_grad(set_to_none=True); o_s.zero_grad(set_to_none=True)
        nerr = abs(float(n_dp) - float(n_s)) / float(n_s)
        assert nerr < tol, ("global norm", precision, float(n_dp), float(n_s))
    a, b = m.flat_params() - p0, single.flat_params() - p0
    err = float((a - b).abs().max() / b.abs().max())
    other = m.flat_params().clone()
    dist.broadcast(other, src=0)
    same = bool(torch.equal(other, m.flat_params()))
    same16 = True
    if precision == "bf16":
        o16 = m.flat_params16().clone()
        dist.broadcast(o16, src=0)
        same16 = bool(torch.equal(o16, m.flat_params16())) and bool(torch.equal(m.flat_params16(), m.flat_params().to(torch.bfloat16)))
    sd = o_dp.state_dict()      # moments are sharded: the state dict gathers them
    msum = sum(float(v["exp_avg"].abs().sum()) for v in sd["state"].values())
    ssum = sum(float(v["exp_avg"].abs().sum()) for v in o_s.state_dict()["state"].values())
    if overlap:      # a second backward without a step must fail loudly
        x, y = micro(0, rank)
        crit(net(x), y).backward()
        try:
            crit(net(x), y).backward()
            raised = False
        except RuntimeError:
            raised = True
        assert raised, "gradient accumulation under nvls_overlap=True was not rejected"
        o_dp.zero_grad(set_to_none=True)
        dist.barrier()
    print(f"rank {rank} [nvls {precision} overlap={overlap}] parameter-update rel err vs single GPU {err:.3e}; ranks identical {same}/{same16}; "
          f"|m| {msum:.4e} vs {ssum:.4e}", flush=True)
    assert err < tol * 5 and same and same16 and abs(msum - ssum) <= tol * 5 * ssum
    del net, m, single


for precision, tol in (("fp32", 2e-4), ("bf16", 5e-2)):
    if dp_mode == "nvls":
        nvls_check(precision, tol, overlap=False)
        nvls_check(precision, tol, overlap=True)
        continue
    m = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=depth, precision=precision)
    m.load_state_dict(ref.state_dict())
    m = m.to(dev).train()
    single = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=depth, precision=precision)
    single.load_state_dict(ref.state_dict())
    single = single.to(dev).train()
    net = pkg.DataParallel(m, bucket_mb=50.0 if depth >= 12 else 8.0, mode="nccl")

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
