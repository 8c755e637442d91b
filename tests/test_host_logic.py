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


def _gloo_worker(rank, world, port, ranges, total, out, comm_dtype=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vit_spoof_detection_pda_b200.dp import GradBucketer
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(total, generator=g)
    local = flat.clone()
    b = GradBucketer(bucket_elems=total // 4, comm_dtype=comm_dtype)
    for lo, hi in ranges:
        b.add(flat, lo, hi)
    b.finish(flat)
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    if comm_dtype is None:
        expect = sum(gathered) / world
        ok = torch.allclose(flat, expect, rtol=1e-6, atol=1e-7)
    else:   # torch DDP's bf16_compress_hook arithmetic: cast, reduce in 16 bit, cast back, average
        expect = sum(t.to(comm_dtype) for t in gathered).to(torch.float32) / world
        ok = torch.allclose(flat, expect, rtol=1e-2, atol=1e-2) and flat.dtype == torch.float32
    out[rank] = (bool(ok), len(b.launched))
    dist.destroy_process_group()


@pytest.mark.parametrize("comm_dtype", [None, torch.bfloat16], ids=["fp32", "bf16-compressed"])
def test_bucketed_allreduce_gloo_world2(pkg, comm_dtype):
    total = 10_000
    ranges = [(9_000, 10_000), (6_000, 9_000), (3_000, 6_000), (500, 3_000), (0, 500)]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gloo_worker, args=(2, _free_port(), ranges, total, out, comm_dtype), nprocs=2, join=True)
    assert out[0][0] and out[1][0]
    assert out[0][1] == out[1][1] >= 2


# ---------------------------------------------------------------------------------------------
# GEMM work decomposition (host logic of csrc/gemm_tc.cu, queried through the C-ABI; no GPU needed: without a device
# the library plans for 148 SMs).  Every (tile, k-block) unit must be produced exactly once, whatever the mode.
# ---------------------------------------------------------------------------------------------
def _plan(lib, I, J, R, accumulate, b_mn):
    import ctypes as C
    out = [C.c_int() for _ in range(7)]
    rc = lib.vitk_gemm_plan(I, J, R, accumulate, b_mn, *[C.byref(o) for o in out])
    assert rc == 0
    bn, cg, mode, ncl, tm, tn, kb = [o.value for o in out]
    items = []
    buf = (C.c_int * (3 * 4096))()
    for c in range(ncl):
        n = lib.vitk_gemm_plan_items(I, J, R, accumulate, b_mn, c, buf, 4096)
        assert 0 < n <= 4096, (c, n)
        items.append([(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(n)])
    return dict(bn=bn, cg=cg, mode=mode, clusters=ncl, tm=tm, tn=tn, kb=kb), items


M64 = 64 * 197
GEMM_SHAPES = [   # (I, J, R, accumulate, B MN-major): the twelve GEMMs of a bs-64 step + small / ragged ones
    (M64, 2304, 768, 0, 0), (M64, 768, 768, 0, 0), (M64, 3072, 768, 0, 0), (M64, 768, 3072, 0, 0),
    (M64, 3072, 768, 0, 1), (M64, 768, 3072, 0, 1), (M64, 768, 768, 0, 1), (M64, 768, 2304, 0, 1),
    (768, 3072, M64, 1, 1), (3072, 768, M64, 1, 1), (768, 768, M64, 1, 1), (2304, 768, M64, 1, 1),
    (197, 768, 768, 0, 0), (100, 128, 64, 0, 0), (768, 768, 197, 1, 1), (2304, 768, 8 * 197, 1, 1), (512, 128, 100, 1, 1),
]


@pytest.mark.parametrize("budget", [0, 140, 132, 116])
def test_gemm_plan_covers_every_tile_and_kblock_once(budget):
    from vit_spoof_detection_pda_b200 import _lib as L
    lib = L.load()
    prev = lib.vitk_set_sm_budget(budget)
    try:
        for I, J, R, acc, bmn in GEMM_SHAPES:
            plan, items = _plan(lib, I, J, R, acc, bmn)
            assert J % plan["bn"] == 0 and plan["cg"] in (1, 2) and plan["mode"] in (0, 1, 2)
            assert plan["tm"] == -(-I // (128 * plan["cg"])) and plan["tn"] == J // plan["bn"] and plan["kb"] == -(-R // 64)
            sms = budget if budget else 148
            assert plan["clusters"] * plan["cg"] <= sms, "more CTAs than SMs: a persistent CTA would wait for a slot"
            if not acc:
                assert plan["mode"] == 0
            cover = {}
            for cl in items:
                for tile, kb0, kb1 in cl:
                    assert 0 <= tile < plan["tm"] * plan["tn"] and 0 <= kb0 < kb1 <= plan["kb"], (I, J, R, tile, kb0, kb1)
                    for kb in range(kb0, kb1):
                        cover[(tile, kb)] = cover.get((tile, kb), 0) + 1
            assert len(cover) == plan["tm"] * plan["tn"] * plan["kb"], (I, J, R, plan)
            assert set(cover.values()) == {1}, (I, J, R, plan)
            if plan["mode"] == 2:      # sliced split-K: one item per cluster, slices of one tile differ by at most one k-block
                assert all(len(cl) == 1 for cl in items)
                lens = [cl[0][2] - cl[0][1] for cl in items]
                assert max(lens) - min(lens) <= 1
            if plan["mode"] == 1:      # stream-K: equal contiguous ranges (the last cluster may be short)
                lens = [sum(k1 - k0 for _, k0, k1 in cl) for cl in items]
                assert max(lens[:-1] or lens) - min(lens[:-1] or lens) == 0 and lens[-1] <= lens[0]
    finally:
        lib.vitk_set_sm_budget(prev)


def test_gemm_plan_bench_shapes_pick_documented_decompositions():
    """fc1 / fc2 / proj weight gradients run sliced split-K (36 x 2, 36 x 2, 9 x 8 items on 74 CTA pairs),
    the qkv weight gradient (27 tiles) contiguous stream-K ranges; J = 768 forward shapes take 256 x 192 pair tiles."""
    from vit_spoof_detection_pda_b200 import _lib as L
    lib = L.load()
    prev = lib.vitk_set_sm_budget(0)
    try:
        for (I, J, R), clusters in (((768, 3072, M64), 72), ((3072, 768, M64), 72), ((768, 768, M64), 72)):
            plan, _ = _plan(lib, I, J, R, 1, 1)
            assert (plan["mode"], plan["clusters"], plan["bn"], plan["cg"]) == (2, clusters, 256, 2)
        plan, _ = _plan(lib, 2304, 768, M64, 1, 1)
        assert (plan["mode"], plan["clusters"]) == (1, 74)
        plan, _ = _plan(lib, M64, 768, 3072, 0, 0)
        assert (plan["bn"], plan["cg"], plan["clusters"]) == (192, 2, 74)
        plan, _ = _plan(lib, M64, 2304, 768, 0, 0)
        assert (plan["bn"], plan["cg"]) == (256, 2)
    finally:
        lib.vitk_set_sm_budget(prev)


def _tail_plan(lib, I, J, R, b_mn, f32_out=0):
    import ctypes as C
    nw, nt, r0 = C.c_int(-1), C.c_int(-1), C.c_int(-1)
    assert lib.vitk_gemm_tail_plan(I, J, R, b_mn, f32_out, C.byref(nw), C.byref(nt), C.byref(r0)) == 0
    return nw.value, nt.value, r0.value


def test_gemm_split_tail_plan():
    """at bs 64 the deep J = 768 GEMMs (fc2 forward, fc1 dgrad, qkv dgrad: 150 tiles of 256 x 256 on 74 CTA
    pairs = two waves + a wave of two tiles) keep 148 whole tiles and deal the 2 x kb k-blocks of the last two tiles out to all
    74 clusters (mode 3); shallow reductions, well-filled last waves and reduced SM budgets keep the plain launch.  Whatever the shape, the items of
    all clusters cover every (tile, k-block) exactly once, every whole tile is ONE item, and a cluster's partial items come
    before its whole tiles."""
    from vit_spoof_detection_pda_b200 import _lib as L
    lib = L.load()
    prev = lib.vitk_set_sm_budget(0)
    try:
        assert lib.vitk_gemm_tail_scratch_floats(768) == 1024 + 512 * 768
        for (J, R, bmn, f32) in ((768, 3072, 0, 1), (768, 3072, 1, 0), (768, 2304, 1, 0)):     # fc2 forward, fc1 dgrad, qkv dgrad
            assert _tail_plan(lib, M64, J, R, bmn, f32) == (148, 2, 49 * 256)
            plan, items = _plan(lib, M64, J, R, 3 if f32 else 2, bmn)
            assert (plan["mode"], plan["bn"], plan["cg"], plan["clusters"]) == (3, 256, 2, 74)
        plan, _ = _plan(lib, M64, 768, 3072, 0, 0)                    # no scratch (eval): three waves of 256 x 192
        assert (plan["mode"], plan["bn"], plan["cg"]) == (0, 192, 2)
        assert _tail_plan(lib, M64, 768, 768, 0)[1] == 0            # proj: 12 k-blocks, the fix-up costs what the wave costs
        assert _tail_plan(lib, M64, 3072, 768, 0)[1] == 0           # fc1 forward: shallow
        assert _tail_plan(lib, 256 * 197, 768, 3072, 1)[1] == 0     # bs 256: 591 tiles = 7.99 waves
        assert _tail_plan(lib, 197, 768, 3072, 1)[1] == 0           # bs 1
        seen_mode3 = {0: 0, 1: 0}
        for batch in list(range(1, 70)) + [96, 128, 200]:
            for J, R, bmn, f32 in ((768, 3072, 0, 1), (768, 3072, 1, 0), (768, 2304, 1, 0), (1536, 4096, 1, 0), (768, 1536, 0, 1)):
                I = batch * 197
                nw, nt, r0 = _tail_plan(lib, I, J, R, bmn, f32)
                plan, items = _plan(lib, I, J, R, 3 if f32 else 2, bmn)
                tiles = plan["tm"] * plan["tn"]
                assert nw + nt == tiles
                cover = {}
                for cl in items:
                    whole_seen = False
                    for tile, kb0, kb1 in cl:
                        if tile >= nw:
                            assert not whole_seen
                        else:
                            whole_seen = True
                            assert (kb0, kb1) == (0, plan["kb"])
                        for kb in range(kb0, kb1):
                            cover[(tile, kb)] = cover.get((tile, kb), 0) + 1
                assert len(cover) == tiles * plan["kb"] and set(cover.values()) == {1}, (I, J, R, plan)
                if nt:
                    seen_mode3[f32] += 1
                    assert plan["mode"] == 3 and plan["kb"] >= 24 and nt * 4 <= 148 // plan["cg"]
                    assert nw % plan["clusters"] == 0 and r0 == (nw // plan["tn"]) * 128 * plan["cg"]
                    assert (I - r0) * J <= 512 * J        # the scratch rows the model provides
                else:
                    assert plan["mode"] == 0
        assert seen_mode3[0] >= 10 and seen_mode3[1] >= 3, seen_mode3
        lib.vitk_set_sm_budget(116)
        assert _tail_plan(lib, M64, 768, 3072, 1)[1] == 0           # 58 CTA pairs: 150 tiles = 2.6 waves, last wave well filled
    finally:
        lib.vitk_set_sm_budget(prev)


# ---------------------------------------------------------------------------------------------
# the guard-band arena of the GPU suite (tests/kernels_api.py::GuardArena) must itself detect stray writes: exercised on CPU
# ---------------------------------------------------------------------------------------------
def test_guard_arena_detects_out_of_bounds_writes():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import kernels_api as K
    arena = K.GuardArena(torch.device("cpu"), nbytes=1 << 20, guard=4096)
    a = arena.alloc((100, 7), torch.float32, zero=True)
    b = arena.alloc((33,), torch.bfloat16, zero=False)
    a.fill_(1.0)
    b.fill_(2.0)
    if True:
        assert arena.check() == (1 << 20) - a.numel() * 4 - b.numel() * 2      # every non-payload byte was looked at
        off_a = arena.spans[0][0]
        arena.buf[off_a + a.numel() * 4] = 0            # one byte past the end of `a`
        with pytest.raises(AssertionError):
            arena.check()
        arena.buf[off_a + a.numel() * 4] = K.GuardArena.BYTE
        arena.buf[arena.spans[1][0] - 1] = 7            # one byte in front of `b`
        with pytest.raises(AssertionError):
            arena.check()
        arena.buf[arena.spans[1][0] - 1] = K.GuardArena.BYTE
        arena.buf[-1] = 1                               # the unallocated remainder
        with pytest.raises(AssertionError):
            arena.check()
