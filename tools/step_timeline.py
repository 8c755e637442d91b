"""Per-kernel timeline of ONE real training step (bs 64, bf16) from the device-side tracer of the development build
(libvitk_dev.so; csrc/common.cuh trace_mark): programmatic dependent launch and the weight-gradient side stream stay
ON, nothing is serialised -- unlike ncu launch lists or CUDA-event bracketing.

    VITK_LIB=dev python tools/step_timeline.py [--batch 64] [--out gpurun_out/timeline]

Writes <out>.json (every launch: kernel, shape, first CTA start, dependencies satisfied, end, CTAs) and <out>.md (per-family
totals, the union of busy time, the idle gaps between launches of the main chain).  The trace build's kernels carry a few
extra instructions per CTA; durations are within ~1 % of the release build's (compare the step time printed at the end).
"""
import argparse
import json
import os
import sys

os.environ.setdefault("VITK_LIB", "dev")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import vit_spoof_detection_pda_b200 as pkg  # noqa: E402
from vit_spoof_detection_pda_b200 import _lib as L  # noqa: E402

KID = {1: "gemm_tc", 2: "attn_fwd", 3: "attn_bwd", 4: "ln_fwd", 5: "ln_bwd", 6: "adam", 7: "sumsq/scale", 8: "cast", 9: "head",
       10: "focal", 11: "colsum", 12: "patch_embed", 13: "embed_grads", 14: "scatter_cls", 15: "gemm_simt", 16: "attn_simt",
       17: "eval", 18: "patch_wgrad", 19: "gemm_tail_epilogue"}


DETAIL_PHASES, DETAIL_SLOTS = 24, 64


def parse(buf):
    n = min(int(buf[0]), int(buf[1]))
    rec = buf[4:4 + 3 * n].reshape(n, 3)
    t = rec[:, 0]
    tag = rec[:, 1]
    aux = rec[:, 2]
    order = torch.argsort(t)
    launches, cur = [], {}
    for i in order.tolist():
        tg = int(tag[i])
        kid, phase, smid, blk = (tg >> 48) & 0xffff, (tg >> 40) & 0xff, (tg >> 24) & 0xffff, tg & 0xffffff
        key = (kid, int(aux[i]))
        ln = cur.get(key)
        if phase == 0 and (ln is None or blk in ln["seen"]):
            ln = {"kid": kid, "aux": int(aux[i]), "seen": set(), "t0": int(t[i]), "t1": None, "t2": int(t[i]), "sms": set()}
            cur[key] = ln
            launches.append(ln)
        if ln is None:
            continue
        if phase == 0:
            ln["seen"].add(blk)
            ln["sms"].add(smid)
        elif phase == 1:
            ln["t1"] = int(t[i]) if ln["t1"] is None else min(ln["t1"], int(t[i]))
        else:
            ln["t2"] = max(ln["t2"], int(t[i]))
    return launches, n


def detail_of(buf, kid):
    """Kernel-internal marks of CTA 0 of the LAST launch of kernel `kid`: {phase: [time ns per slot]} (0 = not written)."""
    off = int(buf[2])
    if off == 0:
        return {}
    area = buf[off + kid * DETAIL_PHASES * DETAIL_SLOTS: off + (kid + 1) * DETAIL_PHASES * DETAIL_SLOTS].reshape(DETAIL_PHASES, DETAIL_SLOTS)
    return {8 + p: area[p].tolist() for p in range(DETAIL_PHASES) if int(area[p].max()) > 0}


def name_of(ln):
    nm = KID.get(ln["kid"], str(ln["kid"]))
    if ln["kid"] == 1:
        a = ln["aux"]
        nm += f" {a >> 40}x{(a >> 20) & 0xfffff}x{a & 0xfffff}"
    return nm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--out", default="gpurun_out/timeline")
    ap.add_argument("--detail", action="store_true", help="also write the kernel-internal marks of CTA 0 (<out>_detail.md)")
    a = ap.parse_args()
    lib = L.load()
    assert lib.vitk_is_dev_build() == 1, "run with VITK_LIB=dev (libvitk_dev.so)"
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = pkg.ViTFaceAntiSpoofing(dropout=0.1, depth=12, precision="bf16").to(dev).train()
    crit = pkg.FocalLoss(0.25, 2.0)
    opt = pkg.FusedAdam(model.parameters(), lr=1e-5, weight_decay=1e-4, adamw=False)
    xs = [torch.randn(a.batch, 3, 224, 224, device=dev) for _ in range(2)]
    y = torch.randint(0, 2, (a.batch,), device=dev)

    def step(i):
        loss, _ = crit(model(xs[i % 2]), y, with_metrics=True)
        loss.backward()
        pkg.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)

    for i in range(6):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms_untraced = e0.elapsed_time(e1) / 10
    buf = torch.zeros(3 * 400_000 + 4 + 32 * DETAIL_PHASES * DETAIL_SLOTS, dtype=torch.int64, device=dev)
    L.check(lib.vitk_trace_start(buf.data_ptr(), buf.numel() * 8), "trace_start")
    e0.record()
    for i in range(2):       # the second traced step is the one reported (host run-ahead has refilled the queue)
        step(i)
    e1.record()
    torch.cuda.synchronize()
    L.check(lib.vitk_trace_stop(), "trace_stop")
    ms_traced = e0.elapsed_time(e1) / 2
    launches, nmarks = parse(buf.cpu())
    # keep the launches of the second step: split at the largest kid==adam end
    adam_ends = [ln["t2"] for ln in launches if ln["kid"] == 6]
    cut = sorted(adam_ends)[0] if len(adam_ends) >= 2 else 0
    step2 = [ln for ln in launches if ln["t0"] > cut]
    t_origin = min(ln["t0"] for ln in step2)
    rows = []
    for ln in step2:
        t1 = ln["t1"] if ln["t1"] is not None else ln["t0"]
        rows.append({"kernel": name_of(ln), "start_us": (ln["t0"] - t_origin) / 1e3, "ready_us": (t1 - t_origin) / 1e3,
                     "end_us": (ln["t2"] - t_origin) / 1e3, "dur_us": (ln["t2"] - t1) / 1e3, "ctas": len(ln["seen"]),
                     "sms": len(ln["sms"])})
    rows.sort(key=lambda r: r["ready_us"])
    span = max(r["end_us"] for r in rows)
    fam = {}
    for r in rows:
        f = fam.setdefault(r["kernel"], [0, 0.0])
        f[0] += 1
        f[1] += r["dur_us"]
    # union of busy intervals (any kernel running) and the gaps in it
    iv = sorted((r["ready_us"], r["end_us"]) for r in rows)
    busy, gaps, cur_s, cur_e = 0.0, [], iv[0][0], iv[0][1]
    for s, e in iv[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append((cur_e, s - cur_e))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out + ".json", "w") as f:
        json.dump({"ms_per_step_untraced_devbuild": ms_untraced, "ms_per_step_traced": ms_traced, "marks": nmarks,
                   "span_us": span, "busy_union_us": busy, "launches": rows}, f)
    with open(a.out + ".md", "w") as f:
        f.write(f"# Step timeline (device-side tracer, bs {a.batch}, bf16; PDL and the weight-gradient side stream ON)\n\n")
        f.write(f"step: {ms_untraced:.3f} ms untraced (development build), {ms_traced:.3f} ms with the tracer recording; "
                f"{len(rows)} launches, span {span / 1e3:.3f} ms, union of busy time {busy / 1e3:.3f} ms, "
                f"idle inside the step {(span - busy):.1f} us in {len(gaps)} gaps\n\n")
        f.write("| kernel | launches | sum of durations (us) | avg (us) | share of span |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {t:.1f} | {t / c:.1f} | {100 * t / span:.1f}% |\n")
        f.write("\n(durations = dependencies satisfied -> last CTA finished; kernels on the two streams overlap, so the shares add "
                "up to more than 100 %)\n\n## launches in order\n\n| # | kernel | ready (us) | dur (us) | CTAs | CTA-start lead (us) |\n|---:|---|---:|---:|---:|---:|\n")
        for i, r in enumerate(rows):
            f.write(f"| {i} | `{r['kernel']}` | {r['ready_us']:.1f} | {r['dur_us']:.1f} | {r['ctas']} | {r['ready_us'] - r['start_us']:.1f} |\n")
    if a.detail:
        hb = buf.cpu()
        with open(a.out + "_detail.md", "w") as f:
            for kid in (2, 3, 12):
                d = detail_of(hb, kid)
                if not d:
                    continue
                ev = sorted((tt, ph, slot) for ph, arr in d.items() for slot, tt in enumerate(arr) if tt > 0)
                t0 = ev[0][0]
                f.write(f"## {KID.get(kid, kid)}: kernel-internal marks of CTA 0, last launch ({len(ev)} marks; t in us from the first)\n\n"
                        "| t (us) | phase | slot (aux & 63) |\n|---:|---:|---:|\n")
                for tt, ph, slot in ev:
                    f.write(f"| {(tt - t0) / 1e3:.2f} | {ph} | {slot} |\n")
                f.write("\n")
    print(f"step {ms_untraced:.3f} ms (dev build, untraced) / {ms_traced:.3f} ms traced; {len(rows)} launches; span {span / 1e3:.3f} ms; "
          f"busy {busy / 1e3:.3f} ms; wrote {a.out}.md/.json")


if __name__ == "__main__":
    main()
