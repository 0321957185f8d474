"""Single-kernel targets for `ncu --set full` at the bench's shapes (one warm-up launch, then two launches of the case):
    ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 1 -o gpurun_out/X python profiles/ncu_targets.py prefill_gateup
cases: prefill_gateup (24960 x 2*18944 x 3584, SwiGLU, token-major), prefill_down, decode_gateup (32 tokens), decode_down (split-K),
       decode_qkv (split-K, bias), decode_o (split-K, residual)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_flamingo_b200 import ops  # noqa: E402

bf16 = torch.bfloat16
case = sys.argv[1]
torch.manual_seed(0)
g = lambda *s, sc=0.02: (torch.randn(*s, device="cuda") * sc).to(bf16)  # noqa: E731
if case in ("prefill_gateup", "decode_gateup"):
    n = 24960 if case.startswith("prefill") else 32
    x, w = g(n, 3584, sc=1.0), g(2 * 18944, 3584)
    fn = lambda: ops.swiglu_linear(x, w, 18944, concat=True)  # noqa: E731
elif case in ("prefill_down", "decode_down"):
    n = 24960 if case.startswith("prefill") else 32
    x, w, r = g(n, 18944, sc=1.0), g(3584, 18944), g(n, 3584, sc=1.0)
    fn = lambda: ops.linear(x, w, resid=r, out=r)  # noqa: E731
elif case == "decode_qkv":
    x, w, b = g(32, 3584, sc=1.0), g(4608, 3584), g(4608, sc=0.3)
    fn = lambda: ops.linear(x, w, b)  # noqa: E731
elif case == "decode_o":
    x, w, r = g(32, 3584, sc=1.0), g(3584, 3584), g(32, 3584, sc=1.0)
    fn = lambda: ops.linear(x, w, resid=r, out=r)  # noqa: E731
else:
    raise SystemExit(f"unknown case {case}")
for _ in range(3):
    fn()
torch.cuda.synchronize()
print("done", case)
