"""Prefill / encoder attention at the bench's shapes, round-1 kernel (AF3_ATTN_V1=1) vs round-2 kernel with one (default) or two
(AF3_ATTN_TPR=2) softmax threads per query row:
   encoder: 32 windows x 20 heads x 64, 1500 frames, bidirectional (fused q/k/v buffer, as the tower calls it)
   prefill: 32 sequences x 28:4 GQA x 128, 780 tokens, causal over the KV cache
CUDA events around 20 back-to-back launches after 3 warm-ups; FLOPs = 4 Tq Tk D per (batch, head), halved when causal (the useful
work; the kernels execute whole 128 x 128 tiles).   python profiles/microbench_attention.py > gpurun_out/attn.json"""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_flamingo_b200 import ops  # noqa: E402

bf16 = torch.bfloat16
torch.manual_seed(0)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def encoder(W=32, H=20, D=64, T=1500):
    qkv = (torch.randn(W * T, 3 * H * D, device="cuda") * 0.5).to(bf16)
    out = torch.empty((W, T, H * D), device="cuda", dtype=bf16)
    fn = lambda: ops.attention(qkv, qkv[:, H * D:], qkv[:, 2 * H * D:], out, B=W, H=H, Hkv=H, D=D, Tq=T, Tk=T, scale=1.0, causal=False,  # noqa: E731
                               kv_layout=0, ldq=3 * H * D, ldk=3 * H * D)
    return fn, 4.0 * W * H * T * T * D, out


def prefill(B=32, H=28, Hkv=4, D=128, T=780, Tmax=1024):
    qkv = (torch.randn(B * T, (H + 2 * Hkv) * D, device="cuda") * 0.5).to(bf16)
    kc = (torch.randn(B, Hkv, Tmax, D, device="cuda") * 0.5).to(bf16)
    vc = (torch.randn(B, Hkv, Tmax, D, device="cuda") * 0.5).to(bf16)
    out = torch.empty((B, T, H * D), device="cuda", dtype=bf16)
    fn = lambda: ops.attention(qkv, kc, vc, out, B=B, H=H, Hkv=Hkv, D=D, Tq=T, Tk=T, scale=D ** -0.5, causal=True, kv_layout=1,  # noqa: E731
                               Tk_pitch=Tmax, ldq=(H + 2 * Hkv) * D, ldk=D)
    return fn, 4.0 * B * H * T * T * D / 2, out


res = {}
only_case, only_impl = os.environ.get("AF3_MB_CASE"), os.environ.get("AF3_MB_IMPL")   # ncu captures: one case, one kernel
for name, mk in (("encoder_d64_t1500", encoder), ("prefill_d128_t780_causal", prefill), ("chat_prefill_d128_t3512_causal", lambda: prefill(B=2, T=3512, Tmax=3584))):
    if only_case and only_case not in name:
        continue
    fn, flops, out = mk()
    row = {}
    outs = {}
    for impl in ("v1", "v2_tpr2", "v2") if not only_impl else (only_impl,):
        os.environ["AF3_ATTN_V1"] = "1" if impl == "v1" else "0"
        os.environ["AF3_ATTN_TPR"] = "2" if impl == "v2_tpr2" else "1"
        ms = timed(fn)
        row[impl] = {"ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1)}
        outs[impl] = out.float().clone()
    if not only_impl:
        row["speedup_v2_over_v1"] = round(row["v1"]["ms"] / row["v2"]["ms"], 3)
        row["max_abs_diff_v2_v1"] = float((outs["v1"] - outs["v2"]).abs().max())
    res[name] = row
print(json.dumps(res, indent=1))
