"""Decode-attention microbenchmark at config 2's decode shape (B=32, 28 q / 4 kv heads x 128, ctx 780..908): 56 launches on
distinct KV caches (28 layers' worth, > L2) captured in one CUDA graph, timed with CUDA events.  Prints us/launch and
the achieved fraction of HBM bandwidth for AF3_DECODE_SPLITS = auto / 1 / 2 / 8."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_flamingo_b200 import ops  # noqa: E402

B, H, Hkv, D, Tmax, L = 32, 28, 4, 128, 908, 28
bf16 = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn((B, (H + 2 * Hkv) * D), device="cuda", generator=g).to(bf16)
caches = [(torch.randn((B, Hkv, Tmax, D), device="cuda", generator=g).to(bf16), torch.randn((B, Hkv, Tmax, D), device="cuda", generator=g).to(bf16))
          for _ in range(L)]
starts = torch.zeros((B,), dtype=torch.int32, device="cuda")
out = torch.zeros((B, H * D), device="cuda", dtype=bf16)
scratch = ops.decode_attention_scratch(B, H, D, Tmax, "cuda")
res = {}
if os.environ.get("AF3_MB_EAGER"):  # for ncu: plain launches at ctx 780, auto splits, no graph
    ctx_len = torch.tensor([780], dtype=torch.int32, device="cuda")
    for _ in range(3):
        for k, v in caches:
            ops.decode_attention(qkv, k, v, out, scratch, B=B, H=H, Hkv=Hkv, D=D, ctx_len=ctx_len, kv_start=starts, scale=D ** -0.5)
    torch.cuda.synchronize()
    sys.exit(0)
for ctx in (780, 908):
    ctx_len = torch.tensor([ctx], dtype=torch.int32, device="cuda")
    for splits in (None, 1, 2, 8):
        if splits is None:
            os.environ.pop("AF3_DECODE_SPLITS", None)
        else:
            os.environ["AF3_DECODE_SPLITS"] = str(splits)

        def run():
            for k, v in caches:
                ops.decode_attention(qkv, k, v, out, scratch, B=B, H=H, Hkv=Hkv, D=D, ctx_len=ctx_len, kv_start=starts, scale=D ** -0.5)

        run()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            run()
        for _ in range(3):
            gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (10 * L)
        gbs = 2 * Hkv * D * 2 * ctx * B / us / 1e3
        res[f"ctx{ctx}_splits{splits}"] = {"us_per_launch": round(us, 2), "GB/s": round(gbs, 1)}
        del gr
print(json.dumps(res, indent=1))
