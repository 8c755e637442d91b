"""Host-side logic that needs no GPU: the drop-in module's parameter container contract, the stage ->
flat-range map the data-parallel buckets are cut from, and the bucketed all-reduce itself on 2 gloo ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import __graft_entry__ as ge
from oracle import vit_oracle as vo


@pytest.fixture(scope="module")
def pkg():
    ge.build()
    import vit_spoof_detection_pda_b200 as p
    return p


def test_module_state_dict_contract_cpu(pkg):
    class Cfg:
        model_name = "vit_base_patch16_224"
        pretrained = False
        num_classes = 2
        dropout = 0.1
    m = pkg.ViTFaceAntiSpoofing(Cfg)
    spec = vo.expected_state_dict_spec(12)
    sd = m.state_dict()
    assert list(sd.keys()) == [n for n, _ in spec]
    assert all(tuple(sd[n].shape) == s for n, s in spec)
    assert sum(p.numel() for p in m.parameters()) == 86_194_946
    assert m.vit.num_features == 768 and len(m.classifier) == 6
    ref = vo.OracleViTFaceAntiSpoofing(depth=12)
    vo.seeded_init_(ref)
    res = m.load_state_dict(ref.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for (n, a), (_, b) in zip(m.state_dict().items(), ref.state_dict().items()):
        assert torch.equal(a, b), n


def test_wrong_model_name_rejected(pkg):
    class Cfg:
        model_name = "resnet50"
        pretrained = False
        num_classes = 2
        dropout = 0.1
    with pytest.raises(ValueError):
        pkg.ViTFaceAntiSpoofing(Cfg)


@pytest.mark.parametrize("depth", [1, 2, 12])
def test_stage_ranges_tile_the_flat_buffer(pkg, depth):
    m = pkg.ViTFaceAntiSpoofing(depth=depth, dropout=0.0)
    total, offs, sizes = m._layout()
    ranges = m.stage_ranges()
    assert len(ranges) == depth + 2
    # descending, adjacent, covering [0, total)
    assert ranges[0][1] == total and ranges[-1][0] == 0
    for (lo, hi), (lo2, hi2) in zip(ranges[:-1], ranges[1:]):
        assert hi2 == lo and lo2 < hi2
    names = [n for n, _ in m.named_parameters()]
    lo0, hi0 = ranges[0]
    tail = [n for n, o in zip(names, offs) if lo0 <= o < hi0]
    assert tail[0] == "vit.norm.weight" and tail[-1] == "classifier.5.bias"
    lo1, hi1 = ranges[1]
    blk = [n for n, o in zip(names, offs) if lo1 <= o < hi1]
    assert all(n.startswith(f"vit.blocks.{depth - 1}.") for n in blk) and len(blk) == 12


def test_bucketer_merges_adjacent_ranges_single_process(pkg):
    from vit_spoof_detection_pda_b200.dp import GradBucketer
    m = pkg.ViTFaceAntiSpoofing(depth=12, dropout=0.0)
    total, _, _ = m._layout()
    flat = torch.zeros(8)  # content irrelevant at world size 1
    b = GradBucketer(bucket_elems=int(50e6 / 4))
    for lo, hi in m.stage_ranges():
        b.add(flat, lo, hi)
    b.finish(flat)
    spans = b.launched
    assert spans[0][1] == total and spans[-1][0] == 0
    assert sum(hi - lo for lo, hi in spans) == total
    assert all(a[0] == c[1] for a, c in zip(spans[:-1], spans[1:]))
    assert 5 <= len(spans) <= 8          # ~86 M params in ~12.5 M-element buckets


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, ranges, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vit_spoof_detection_pda_b200.dp import GradBucketer
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(total, generator=g)
    local = flat.clone()
    b = GradBucketer(bucket_elems=total // 4)
    for lo, hi in ranges:
        b.add(flat, lo, hi)
    b.finish(flat)
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = sum(gathered) / world
    ok = torch.allclose(flat, expect, rtol=1e-6, atol=1e-7)
    out[rank] = (bool(ok), len(b.launched))
    dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2(pkg):
    total = 10_000
    ranges = [(9_000, 10_000), (6_000, 9_000), (3_000, 6_000), (500, 3_000), (0, 500)]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gloo_worker, args=(2, _free_port(), ranges, total, out), nprocs=2, join=True)
    assert out[0][0] and out[1][0]
    assert out[0][1] == out[1][1] >= 2
