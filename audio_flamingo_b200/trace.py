"""Host side of the library's in-graph timeline (include/af3b200.h: af3_trace_begin / af3_trace_end).

While a trace is open every launch of a decode-step kernel stores %globaltimer stamps of its first 160 CTAs (entry, after
griddepcontrol.wait, main loop done, exit) into a slot of the caller's device buffer; the slot address is a launch parameter, so a
CUDA graph captured meanwhile keeps recording on every replay.  This module owns the buffer, maps slots back to the ops that
launched them (ops.TRACE_LOG) and turns the stamps of the LAST replay into per-launch records:

    with DecodeTrace(device) as tr:
        model.generate(...)            # captures the decode graph with the slots baked in
    launches = tr.graph_launches()     # [{"kind", "ctas", "entry_min", ..., "slot_us", "body_us", "stream_us", "tail_us", ...}]

IMPORTANT: a graph captured under a trace writes into this object's buffer on every replay -- drop the graph
(model.release_decode_state()) before the DecodeTrace is garbage collected.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops


def kind_name(key) -> str:
    kind, a, b, c, flags = key[:5]
    if kind == "gemm":
        return f"gemm {b}x{c}" + (" +rope" if flags & 32 else "") + (" swiglu" if flags & 8 else "")
    return kind


class DecodeTrace:
    def __init__(self, device, n_slots: int = 4096):
        self.lib = _lib.load()
        self.slot_words = self.lib.af3_trace_slot_bytes() // 8
        self.n_slots = n_slots
        self.buf = torch.zeros((n_slots * self.slot_words,), device=device, dtype=torch.int64)
        self.log = None
        self.n_recorded = 0

    def __enter__(self):
        ops.TRACE_LOG = []
        self.lib.af3_trace_begin(self.buf.data_ptr(), self.buf.numel() * 8)
        return self

    def __exit__(self, *exc):
        torch.cuda.synchronize(self.buf.device)
        self.n_recorded = self.lib.af3_trace_end()
        self.log, ops.TRACE_LOG = ops.TRACE_LOG, None
        return False

    def graph_launches(self):
        """Records of the launches captured in the (last) decode graph, in start order, times in us since the step's first entry.
        slot_us = last exit of this kernel - last exit of its predecessor: the kernel's exclusive share of the step (sums to the
        step time); body_us = dependency resolved -> last exit; stream_us = -> last accumulator ready; tail_us = the rest."""
        caps = [e for e in self.log if e[0] == "graph_capture"]
        if not caps:
            return []
        _, i0, i1 = caps[-1]
        entries = [e for e in self.log[i0:i1] if e[0] != "graph_capture"]
        raw = self.buf.view(self.n_slots, -1, 4).cpu().numpy().astype(np.int64)
        out = []
        for key, s0, s1 in entries:
            for s in range(s0, min(s1, self.n_slots)):
                m = raw[s]
                live = m[:, 0] > 0
                if not live.any():
                    continue
                pos = lambda col: m[live, col][m[live, col] > 0]  # noqa: E731
                rec = {"slot": s, "key": list(key[:5]), "kind": kind_name(key) + (f" #{s - s0}" if s1 - s0 > 1 else ""), "ctas": int(live.sum()),
                       "entry_min": int(pos(0).min()), "entry_max": int(pos(0).max()),
                       "wait_min": int(pos(1).min()) if len(pos(1)) else None, "wait_max": int(pos(1).max()) if len(pos(1)) else None,
                       "mid_max": int(pos(2).max()) if len(pos(2)) else None,
                       "exit_min": int(pos(3).min()) if len(pos(3)) else None, "exit_max": int(pos(3).max()) if len(pos(3)) else None}
                both = live & (m[:, 2] > 0) & (m[:, 3] > 0)
                if both.any():
                    # per-CTA time from "own accumulator ready" to "own exit" (us): the split-K kernels' non-reducing CTAs show the
                    # cost of publishing a partial (store, fence, counter), the reducing ones the whole tail
                    d = np.sort((m[both, 3] - m[both, 2]) / 1e3)
                    rec["cta_tail_us"] = {"min": float(d[0]), "p25": float(d[len(d) // 4]), "median": float(d[len(d) // 2]),
                                          "p75": float(d[(3 * len(d)) // 4]), "max": float(d[-1])}
                    rec["mid_spread_us"] = float((m[both, 2].max() - m[both, 2].min()) / 1e3)
                out.append(rec)
        if not out:
            return out
        out.sort(key=lambda r: r["entry_min"])
        t0 = out[0]["entry_min"]
        for r in out:
            for k in ("entry_min", "entry_max", "wait_min", "wait_max", "mid_max", "exit_min", "exit_max"):
                if r[k] is not None:
                    r[k] = (r[k] - t0) / 1e3
        prev_exit = out[0]["entry_min"]
        for r in out:
            if r["exit_max"] is None:
                continue
            r["slot_us"] = r["exit_max"] - prev_exit
            if r["wait_min"] is not None:
                r["lead_us"] = r["wait_min"] - r["entry_min"]
                r["body_us"] = r["exit_max"] - r["wait_min"]
                r["gap_us"] = r["wait_min"] - prev_exit
                if r["mid_max"] is not None:
                    r["stream_us"] = r["mid_max"] - r["wait_min"]
                    r["tail_us"] = r["exit_max"] - r["mid_max"]
            prev_exit = max(prev_exit, r["exit_max"])
        return out

    @staticmethod
    def aggregate(launches):
        agg = {}
        for r in launches:
            a = agg.setdefault(r["kind"], {"n": 0, "slot_us": 0.0, "body_us": 0.0, "stream_us": 0.0, "tail_us": 0.0, "lead_us": 0.0, "gap_us": 0.0})
            a["n"] += 1
            for k in ("slot_us", "body_us", "stream_us", "tail_us", "lead_us", "gap_us"):
                a[k] += r.get(k, 0.0) or 0.0
        for a in agg.values():
            for k in list(a):
                if k != "n":
                    a[k] = round(a[k], 2)
        return agg
