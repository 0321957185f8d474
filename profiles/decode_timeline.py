#!/usr/bin/env python
"""In-graph timeline of one decode step (programmatic dependent launch ON, nothing serialised).

ncu replays kernels one at a time, so it cannot show where the decode step's slack sits (3.67 ms per step
against a 2.39 ms HBM bound).  This tool opens the library's trace (include/af3b200.h: af3_trace_begin), runs
generate() at the bench's shape (AF3-7B decoder, 32 sequences, 780-token prompts) so that the decode step is captured
in a CUDA graph WITH the trace slots baked in, replays it, and reads back per-CTA %globaltimer stamps of the LAST replay:
    mark 0 kernel entry, 1 after griddepcontrol.wait, 2 main loop done (GEMM: last accumulator ready), 3 exit.
Output: JSON (one record per launch + per-kind aggregates) and a markdown summary.

    python profiles/decode_timeline.py --out profiles/r02_decode_timeline --tag stages8
Environment knobs under test are simply inherited (AF3_SWAP_STAGES, AF3_SWAP_STAGES2, AF3_PDL, ...).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def kind_name(key):
    kind, a, b, c, flags = key
    if kind == "gemm":
        return f"gemm {b}x{c}" + (" +rope" if flags & 32 else "") + (" swiglu" if flags & 8 else "")
    return kind


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "profiles" / "r02_decode_timeline"))
    ap.add_argument("--tag", default="default")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--prompt", type=int, default=780)
    ap.add_argument("--new-tokens", type=int, default=24)
    ap.add_argument("--layers", type=int, default=28)
    args = ap.parse_args()

    import bench
    from audio_flamingo_b200 import AudioFlamingo3ForConditionalGeneration, _lib, ops

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cfg = bench.af3_config()
    cfg.text_config.num_hidden_layers = args.layers
    cfg.audio_config.num_hidden_layers = 1  # the audio tower is not exercised here (text-only prompts)
    model = AudioFlamingo3ForConditionalGeneration(cfg)
    model.to_empty(device=dev)
    model.to(torch.bfloat16)
    bench.init_synthetic_weights_(model, seed=0)
    rs = np.random.RandomState(0)
    ids = torch.from_numpy(rs.randint(1, 151643, size=(args.batch, args.prompt)).astype(np.int64)).to(dev)
    am = torch.ones_like(ids)

    lib = _lib.load()
    slot_words = lib.af3_trace_slot_bytes() // 8
    n_slots = 4096
    buf = torch.zeros((n_slots * slot_words,), device=dev, dtype=torch.int64)

    # warm run without tracing (configures kernels, allocator), state dropped so the traced run captures a fresh graph
    model.generate(input_ids=ids, attention_mask=am, max_new_tokens=4)
    model.release_decode_state()
    torch.cuda.synchronize()

    ops.TRACE_LOG = []
    lib.af3_trace_begin(buf.data_ptr(), buf.numel() * 8)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    model.stage_events = []
    model.generate(input_ids=ids, attention_mask=am, max_new_tokens=args.new_tokens)
    torch.cuda.synchronize()
    n_rec = lib.af3_trace_end()
    log, ops.TRACE_LOG = ops.TRACE_LOG, None
    ev = dict(model.stage_events)
    decode_ms = ev["prefill_done"].elapsed_time(ev["decode_done"])
    model.stage_events = None

    cap = [e for e in log if e[0] == "graph_capture"]
    assert cap, "the decode step was not captured (max_new_tokens too small?)"
    _, i0, i1 = cap[-1]
    entries = [e for e in log[i0:i1] if e[0] != "graph_capture"]
    raw = buf.view(n_slots, -1, 4).cpu().numpy().astype(np.int64)  # [slot][cta][mark]

    launches = []
    for key, s0, s1 in entries:
        for s in range(s0, s1):
            if s >= n_slots:
                continue
            m = raw[s]
            live = m[:, 0] > 0
            if not live.any():
                continue
            rec = {"slot": s, "kind": kind_name(key) + (f" #{s - s0}" if s1 - s0 > 1 else ""), "ctas": int(live.sum()),
                   "entry_min": int(m[live, 0].min()), "entry_max": int(m[live, 0].max()),
                   "wait_min": int(m[live, 1][m[live, 1] > 0].min()) if (m[live, 1] > 0).any() else None,
                   "wait_max": int(m[live, 1].max()) or None,
                   "mid_max": int(m[live, 2].max()) or None,
                   "exit_min": int(m[live, 3][m[live, 3] > 0].min()) if (m[live, 3] > 0).any() else None,
                   "exit_max": int(m[live, 3].max()) or None}
            launches.append(rec)
    launches.sort(key=lambda r: r["entry_min"])
    t0 = launches[0]["entry_min"]
    for r in launches:
        for k in ("entry_min", "entry_max", "wait_min", "wait_max", "mid_max", "exit_min", "exit_max"):
            if r[k] is not None:
                r[k] = (r[k] - t0) / 1e3  # us since the step's first kernel entry
    step_us = launches[-1]["exit_max"] - launches[0]["entry_min"]

    # per launch: lead = how long before its dependency resolved the kernel was already resident (pre-wait prefetch window),
    # body = dependency resolved -> last CTA exit, gap = predecessor's last exit -> this kernel's dependency resolved
    agg = {}
    prev_exit = None
    for r in launches:
        name = r["kind"]
        a = agg.setdefault(name, {"n": 0, "lead_us": 0.0, "body_us": 0.0, "stream_us": 0.0, "tail_us": 0.0, "gap_us": 0.0})
        a["n"] += 1
        if r["wait_max"] is not None and r["exit_max"] is not None:
            r["lead_us"] = r["wait_min"] - r["entry_min"]
            r["body_us"] = r["exit_max"] - r["wait_min"]
            a["lead_us"] += r["lead_us"]
            a["body_us"] += r["body_us"]
            if r["mid_max"] is not None:
                a["stream_us"] += r["mid_max"] - r["wait_min"]
                a["tail_us"] += r["exit_max"] - r["mid_max"]
            if prev_exit is not None:
                r["gap_us"] = r["wait_min"] - prev_exit
                a["gap_us"] += r["gap_us"]
        prev_exit = r["exit_max"] if r["exit_max"] is not None else prev_exit
    for a in agg.values():
        for k in list(a):
            if k != "n":
                a[k] = round(a[k], 2)

    out = {"tag": args.tag, "env": {k: v for k, v in os.environ.items() if k.startswith("AF3_")},
           "shape": {"batch": args.batch, "prompt": args.prompt, "layers": args.layers, "new_tokens": args.new_tokens},
           "decode_ms_per_step_events": decode_ms / (args.new_tokens - 1), "step_us_from_trace": step_us,
           "launches_in_graph": len(launches), "slots_recorded": n_rec, "per_kind": agg, "launches": launches}
    base = f"{args.out}_{args.tag}"
    Path(base + ".json").write_text(json.dumps(out))
    lines = [f"# decode-step timeline ({args.tag}): {step_us:.1f} us per step from the trace, "
             f"{decode_ms / (args.new_tokens - 1) * 1e3:.1f} us per step by CUDA events over the decode stage", "",
             f"env: {out['env']}", "",
             "| kernel | launches | body us (dep. resolved -> last exit) | stream us (-> accumulators ready) | tail us | resident before dep. us | gap after predecessor us |",
             "|---|---|---|---|---|---|---|"]
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["body_us"]):
        lines.append(f"| {name} | {a['n']} | {a['body_us']:.1f} | {a['stream_us']:.1f} | {a['tail_us']:.1f} | {a['lead_us']:.1f} | {a['gap_us']:.1f} |")
    lines += ["", "First layer in launch order (us since step start):", "",
              "| kernel | ctas | entry min..max | dep. resolved min..max | main loop done | exit min..max |", "|---|---|---|---|---|---|"]
    for r in launches[:12]:
        f = lambda v: "-" if v is None else f"{v:.1f}"  # noqa: E731
        lines.append(f"| {r['kind']} | {r['ctas']} | {f(r['entry_min'])}..{f(r['entry_max'])} | {f(r['wait_min'])}..{f(r['wait_max'])} | "
                     f"{f(r['mid_max'])} | {f(r['exit_min'])}..{f(r['exit_max'])} |")
    Path(base + ".md").write_text("\n".join(lines) + "\n")
    print(json.dumps({k: out[k] for k in ("tag", "decode_ms_per_step_events", "step_us_from_trace", "launches_in_graph")}))


if __name__ == "__main__":
    main()
