"""Microbenchmark of the few-token (decode) GEMM shapes: CUDA-event time per launch with rotating weight copies so
every launch streams its weights from HBM (not L2).  usage: [AF3_KSPLIT=n] python profiles/microbench_decode_gemm.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_flamingo_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
M = 32


def bench(name, N, K, mode, copies=12, iters=60):
    ws = [(torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16) for _ in range(copies)]
    if mode == "swiglu":
        ws = [ops.pack_gate_up(w, w) for w in ws]
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    bias = torch.zeros(N, device=dev, dtype=torch.bfloat16)
    res = torch.randn(M, N, device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)

    def run(i):
        w = ws[i % copies]
        if mode == "plain":
            ops.linear(x, w, out=out)
        elif mode == "bias":
            ops.linear(x, w, bias, out=out)
        elif mode == "resid":
            ops.linear(x, w, resid=res, out=out)
        elif mode == "resid_inplace":
            ops.linear(x, w, resid=out, out=out)
        elif mode == "swiglu":
            ops.swiglu_linear(x, w, N, out=out)

    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            run(i)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    nbytes = ws[0].numel() * 2
    print(f"{name:28s} N={N:6d} K={K:6d} {mode:14s} {us:8.2f} us/launch  {nbytes / us / 1e6:7.2f} TB/s")


for mode in ("plain", "bias", "resid", "resid_inplace"):
    bench("o_proj-like", 3584, 3584, mode)
bench("qkv", 4608, 3584, "bias")
bench("down", 3584, 18944, "resid_inplace")
bench("down", 3584, 18944, "plain")
bench("gate_up", 18944, 3584, "swiglu")
bench("lm_head-like", 152064 // 4, 3584, "plain", copies=4)
