"""One launch of every kernel of libaf3b200.so at small shapes, for compute-sanitizer (memcheck + racecheck per kernel;
the kernels hand-roll mbarrier protocols, cross-CTA counters and programmatic-dependent-launch prologues):

    compute-sanitizer --tool memcheck  python profiles/sanitize_kernels.py > gpurun_out/sanitizer_memcheck.log  2>&1
    compute-sanitizer --tool racecheck python profiles/sanitize_kernels.py > gpurun_out/sanitizer_racecheck.log 2>&1
    compute-sanitizer --tool synccheck python profiles/sanitize_kernels.py > gpurun_out/sanitizer_synccheck.log 2>&1

Every call is also checked against a torch restatement with a loose tolerance, so a run that "passes" the sanitizer while computing
garbage is caught too.  Prints one line per kernel."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_flamingo_b200 import _lib, ops  # noqa: E402

bf16 = torch.bfloat16
torch.manual_seed(0)
dev = "cuda"


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(bf16)


def ok(name, got, ref, tol=3e-2):
    err = (got.float() - ref.float()).abs().max().item()
    lim = tol * max(ref.float().abs().max().item(), 1.0)
    print(f"{name:42s} max err {err:.4g} (limit {lim:.3g}) {'ok' if err <= lim else 'MISMATCH'}", flush=True)
    assert err <= lim, name


# ---- GEMM: token-major (TMA epilogue, residual), few-token, few-token split-K with residual, SwiGLU, fp32 out
x, w, b = rnd(300, 256), rnd(384, 256, scale=0.05), rnd(384, scale=0.3)
ok("gemm normal bias+gelu", ops.linear(x, w, b, gelu=True), torch.nn.functional.gelu((x.float() @ w.float().T + b.float()).to(bf16).float()))
r = rnd(300, 384)
ok("gemm normal resid (tma residual)", ops.linear(x, w, resid=r), (x.float() @ w.float().T).to(bf16).float() + r.float())
xs = rnd(32, 512)
ws = rnd(256, 512, scale=0.05)
ok("gemm few-token", ops.linear(xs, ws), xs.float() @ ws.float().T)
wk = rnd(512, 4096, scale=0.02)
xk, rk = rnd(32, 4096), rnd(32, 512)
x50, b512 = rnd(50, 4096), rnd(512, scale=0.3)
ok("gemm few-token, 50 tokens, bias + gelu", ops.linear(x50, wk, b512, gelu=True),
   torch.nn.functional.gelu((x50.float() @ wk.float().T + b512.float()).to(bf16).float()))
os.environ["AF3_KSPLIT"] = "5"
ok("gemm few-token split-K x5 + resid", ops.linear(xk, wk, resid=rk), (xk.float() @ wk.float().T).to(bf16).float() + rk.float())
os.environ.pop("AF3_KSPLIT")
ok("gemm few-token fp32 out", ops.linear(xs, ws, out_f32=True), (xs.float() @ ws.float().T).to(bf16).float())
g, u = rnd(256, 512, scale=0.05), rnd(256, 512, scale=0.05)
wp = ops.pack_gate_up(g, u)
for n_tok, tag in ((32, "few-token"), (200, "normal")):
    xx = rnd(n_tok, 512)
    gg, uu = (xx.float() @ g.float().T).to(bf16).float(), (xx.float() @ u.float().T).to(bf16).float()
    ok(f"gemm swiglu {tag}", ops.swiglu_linear(xx, wp, 256), (torch.nn.functional.silu(gg).to(bf16).float() * uu))

# ---- RMSNorm fused across two few-token GEMMs (producer partials + consumer transform), [gate; up] concat layout
hh, wo2, rr = rnd(32, 512), rnd(512, 512, scale=0.05), rnd(32, 512)
ssb = ops.sumsq_buffer(32, dev)
hres = ops.linear(hh, wo2, resid=rr.clone(), sumsq_out=ssb)
ok("fused norm: producer partials", ssb[:, :4], hres.float().pow(2).view(32, 4, 128).sum(-1), tol=1e-4)
wn = (rnd(512, scale=0.3).float() + 1.0).to(bf16)
w3, b3 = rnd(384, 512, scale=0.05), rnd(384, scale=0.3)
ok("fused norm: consumer vs rmsnorm kernel", ops.linear(hres, w3, b3, norm=(wn, ssb, 4, 1e-6)), ops.linear(ops.rmsnorm(hres, wn, 1e-6), w3, b3))
gc, uc = rnd(256, 512, scale=0.05), rnd(256, 512, scale=0.05)
ok("fused norm + swiglu [gate; up]", ops.swiglu_linear(hres, torch.cat([gc, uc]).contiguous(), 256, norm=(wn, ssb, 4, 1e-6), concat=True),
   ops.swiglu_linear(ops.rmsnorm(hres, wn, 1e-6), ops.pack_gate_up(gc, uc), 256))
# deep K (32 k blocks > ring depth): every ring stage is reused several times by its owning warp
hk_, wok, rrk = rnd(32, 256), rnd(2048, 256, scale=0.05), rnd(32, 2048)
ssk = ops.sumsq_buffer(32, dev)
hk2 = ops.linear(hk_, wok, resid=rrk.clone(), sumsq_out=ssk)
wnk = (rnd(2048, scale=0.3).float() + 1.0).to(bf16)
gk, uk = rnd(256, 2048, scale=0.03), rnd(256, 2048, scale=0.03)
ok("fused norm + swiglu, K = 2048", ops.swiglu_linear(hk2, torch.cat([gk, uk]).contiguous(), 256, norm=(wnk, ssk, 16, 1e-6), concat=True),
   ops.swiglu_linear(ops.rmsnorm(hk2, wnk, 1e-6), ops.pack_gate_up(gk, uk), 256))
w3k, b3k = rnd(384, 2048, scale=0.03), rnd(384, scale=0.3)
ok("fused norm, K = 2048, bias", ops.linear(hk2, w3k, b3k, norm=(wnk, ssk, 16, 1e-6)), ops.linear(ops.rmsnorm(hk2, wnk, 1e-6), w3k, b3k))
xx2 = rnd(200, 512)
ok("swiglu [gate; up] token-major", ops.swiglu_linear(xx2, torch.cat([gc, uc]).contiguous(), 256, concat=True), ops.swiglu_linear(xx2, ops.pack_gate_up(gc, uc), 256), tol=1e-6)
os.environ["AF3_CLUSTER_REDUCE"] = "1"
os.environ["AF3_KSPLIT"] = "4"
ok("gemm few-token cluster/DSMEM split-K x4 + resid (opt-in experiment)", ops.linear(xk, wk, resid=rk), (xk.float() @ wk.float().T).to(bf16).float() + rk.float())
os.environ.pop("AF3_CLUSTER_REDUCE")
os.environ.pop("AF3_KSPLIT")

# ---- fused q/k/v + RoPE + KV append (few-token), stand-alone RoPE / append
B, H, Hkv, D, K, Tmax = 4, 4, 2, 128, 256, 256
xq, wq, bq = rnd(B, K), rnd((H + 2 * Hkv) * D, K, scale=0.05), rnd((H + 2 * Hkv) * D, scale=0.3)
inv_freq = (1.0 / (1e6 ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))).to(dev)
pos = torch.tensor([37], dtype=torch.int32, device=dev)
kc1, vc1 = torch.zeros((B, Hkv, Tmax, D), device=dev, dtype=bf16), torch.zeros((B, Hkv, Tmax, D), device=dev, dtype=bf16)
kc2, vc2 = torch.zeros_like(kc1), torch.zeros_like(vc1)
q1 = ops.linear(xq, wq, bq)
ops.rope_kv_append(q1, kc1, vc1, B=B, T=1, H=H, Hkv=Hkv, D=D, pos0=0, inv_freq=inv_freq, pos0_dev=pos)
cs = ops.rope_table(B, D, pos, None, inv_freq)
q2 = ops.qkv_rope_linear(xq, wq, bq, kc2, vc2, H=H, Hkv=Hkv, D=D, rope_cs=cs, pos_dev=pos)
ok("qkv+rope fused vs unfused (q)", q2[:, : H * D], q1[:, : H * D], tol=1e-6)
ok("qkv+rope fused vs unfused (k cache)", kc2, kc1, tol=1e-6)
ok("qkv+rope fused vs unfused (v cache)", vc2, vc1, tol=1e-6)

# ---- norms
xn, gam, bet = rnd(70, 1280), rnd(1280), rnd(1280)
ok("layernorm", ops.layernorm(xn, gam, bet), torch.nn.functional.layer_norm(xn.float(), (1280,), gam.float(), bet.float()))
ok("avgpool+layernorm", ops.avgpool_layernorm(rnd(2 * 20, 1280), 2, 20, gam, bet).shape[0] * torch.ones(1), torch.tensor([20.0]))
for rows in (8, 2000):
    xr, wr = rnd(rows, 3584), rnd(3584)
    xf = xr.float()
    ref = wr.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).to(bf16).float()
    ok(f"rmsnorm {rows} rows", ops.rmsnorm(xr, wr), ref)


# ---- attention (both generations), decode attention (unsplit + split)
def sdpa(q, k, v, scale, mask):
    s = (q @ k.transpose(-1, -2)) * scale
    s = s.masked_fill(~mask, float("-inf"))
    return torch.nan_to_num(torch.softmax(s, -1), nan=0.0) @ v


for impl, tpr in (("0", "2"), ("0", "1"), ("1", "2")):
    os.environ["AF3_ATTN_V1"] = impl
    os.environ["AF3_ATTN_TPR"] = tpr
    Bq, Hh, Dd, T = 2, 2, 64, 300
    qkv = rnd(Bq * T, 3 * Hh * Dd)
    out = torch.zeros((Bq, T, Hh * Dd), device=dev, dtype=bf16)
    lens = torch.tensor([300, 77], dtype=torch.int32, device=dev)
    ops.attention(qkv, qkv[:, Hh * Dd:], qkv[:, 2 * Hh * Dd:], out, B=Bq, H=Hh, Hkv=Hh, D=Dd, Tq=T, Tk=T, scale=0.125, causal=False, kv_layout=0,
                  ldq=3 * Hh * Dd, ldk=3 * Hh * Dd, kv_len=lens)
    q, k, v = [t.float().view(Bq, T, Hh, Dd).transpose(1, 2) for t in qkv.split(Hh * Dd, dim=1)]
    m = torch.ones((Bq, 1, T, T), dtype=torch.bool, device=dev)
    m[1, :, :, 77:] = False
    ok(f"attention bidirectional D=64 (AF3_ATTN_V1={impl} TPR={tpr})", out.view(Bq, T, Hh, Dd).transpose(1, 2), sdpa(q, k, v, 0.125, m))
    H2, Hk2, D2, T2, Tm2 = 4, 2, 128, 300, 384
    qk2 = rnd(Bq * T2, (H2 + 2 * Hk2) * D2)
    kc, vc = torch.zeros((Bq, Hk2, Tm2, D2), device=dev, dtype=bf16), torch.zeros((Bq, Hk2, Tm2, D2), device=dev, dtype=bf16)
    kk = qk2[:, H2 * D2:(H2 + Hk2) * D2].view(Bq, T2, Hk2, D2).transpose(1, 2)
    vv = qk2[:, (H2 + Hk2) * D2:].view(Bq, T2, Hk2, D2).transpose(1, 2)
    kc[:, :, :T2], vc[:, :, :T2] = kk, vv
    o2 = torch.zeros((Bq, T2, H2 * D2), device=dev, dtype=bf16)
    st = torch.tensor([0, 130], dtype=torch.int32, device=dev)
    ops.attention(qk2, kc, vc, o2, B=Bq, H=H2, Hkv=Hk2, D=D2, Tq=T2, Tk=T2, scale=D2 ** -0.5, causal=True, kv_layout=1, Tk_pitch=Tm2,
                  ldq=(H2 + 2 * Hk2) * D2, ldk=D2, kv_start=st)
    q = qk2[:, : H2 * D2].float().view(Bq, T2, H2, D2).transpose(1, 2)
    m = torch.tril(torch.ones((T2, T2), dtype=torch.bool, device=dev))[None, None].repeat(Bq, 1, 1, 1)
    m[1, :, :, :130] = False
    ref = sdpa(q, kk.float().repeat_interleave(2, 1), vv.float().repeat_interleave(2, 1), D2 ** -0.5, m)
    got = o2.view(Bq, T2, H2, D2).transpose(1, 2)
    ok(f"attention causal GQA D=128 (AF3_ATTN_V1={impl} TPR={tpr})", got[0], ref[0])
    ok(f"attention causal GQA D=128 left-padded row (AF3_ATTN_V1={impl} TPR={tpr})", got[1, :, 130:], ref[1, :, 130:])
os.environ.pop("AF3_ATTN_V1")
os.environ.pop("AF3_ATTN_TPR")

for splits in ("1", "3"):
    os.environ["AF3_DECODE_SPLITS"] = splits
    Bd, Hd, Hkd, Dd, Tm, ctx = 3, 14, 2, 128, 512, 300
    qd = rnd(Bd, (Hd + 2 * Hkd) * Dd)
    kcd, vcd = rnd(Bd, Hkd, Tm, Dd), rnd(Bd, Hkd, Tm, Dd)
    kcd[:, :, ctx:], vcd[:, :, ctx:] = 0, 0
    od = torch.zeros((Bd, Hd * Dd), device=dev, dtype=bf16)
    scr = ops.decode_attention_scratch(Bd, Hd, Dd, Tm, dev)
    ctx_t = torch.tensor([ctx], dtype=torch.int32, device=dev)
    st = torch.tensor([0, 5, 140], dtype=torch.int32, device=dev)
    ops.decode_attention(qd, kcd, vcd, od, scr, B=Bd, H=Hd, Hkv=Hkd, D=Dd, ctx_len=ctx_t, kv_start=st, scale=Dd ** -0.5)
    q = qd[:, : Hd * Dd].float().view(Bd, Hd, 1, Dd)
    kx = kcd[:, :, :ctx].float().repeat_interleave(Hd // Hkd, 1)
    vx = vcd[:, :, :ctx].float().repeat_interleave(Hd // Hkd, 1)
    m = torch.ones((Bd, 1, 1, ctx), dtype=torch.bool, device=dev)
    for i, s0 in enumerate(st.tolist()):
        m[i, :, :, :s0] = False
    ok(f"decode attention (splits={splits})", od.view(Bd, Hd, 1, Dd), sdpa(q, kx, vx, Dd ** -0.5, m))
os.environ.pop("AF3_DECODE_SPLITS")

# ---- glue: conv-stem im2col, embedding scatter, argmax, log-mel, rotary time embedding
xi = torch.randn(2, 128, 64, device=dev)
cols = ops.im2col_conv1(xi)
ok("im2col conv1 (shape)", torch.tensor([float(cols.shape[0])]), torch.tensor([128.0]))
h1 = rnd(2 * 64, 256)
ok("im2col conv2 (shape)", torch.tensor([float(ops.im2col_conv2(h1, 2, 64).shape[0])]), torch.tensor([64.0]))
table = rnd(100, 256)
ids = torch.randint(0, 99, (40,), device=dev)
ids[5:15] = 99
aud = rnd(2 * 8, 256)
post = torch.tensor([6, 4], dtype=torch.int32, device=dev)
emb, counts = ops.embed_scatter(ids, table, 99, aud, 2, 8, post)
ref = table[ids].clone()
ref[5:11] = aud[0:6]
ref[11:15] = aud[8:12]
ok("embed + audio scatter", emb, ref, tol=1e-6)
lg = torch.randn(5, 5000, device=dev)
ok("argmax", ops.argmax(lg).float(), lg.argmax(-1).float(), tol=1e-6)
from audio_flamingo_b200 import AF3FeatureExtractor  # noqa: E402

fe = AF3FeatureExtractor(dev)
wave = torch.randn(1, 480000, device=dev) * 0.1
f = fe.from_device_waveform(wave, [480000])["input_features"]
print(f"{'logmel (finite, shape)':42s} {tuple(f.shape)} {'ok' if torch.isfinite(f).all() else 'MISMATCH'}", flush=True)
torch.cuda.synchronize()
print("ALL KERNELS LAUNCHED")
