"""Per-kernel parity on the GPU: each C-ABI kernel against a plain torch fp32 restatement of the same op with the
reference's bf16 rounding points (tolerances stated per test)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    from audio_flamingo_b200 import ops as _ops

    return _ops


def _rand(shape, scale=1.0, seed=0, dtype=bf16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).cuda()


def _close(got, ref, rtol, atol, what=""):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = (err > tol).sum().item()
    assert bad == 0, f"{what}: {bad}/{err.numel()} out of tolerance, max err {err.max().item():.4g}, ref absmax {ref.abs().max().item():.4g}"


# bf16 output: 1 ulp = 2^-8 relative; accumulation-order differences in fp32 are far below that, so results may
# differ from the torch restatement only where the fp32 value sits next to a bf16 rounding boundary (<= 1 ulp).
BF16_RTOL = 2 ** -7


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 512, 192), (1024, 1280, 1280), (257, 96, 320), (4000, 3840, 384)])
def test_gemm_plain(ops, M, N, K):
    x, w = _rand((M, K), 1.0, 1), _rand((N, K), 0.05, 2)
    out = ops.linear(x, w)
    ref = (x.float() @ w.float().T).to(bf16)
    _close(out, ref, BF16_RTOL, 1e-3, f"gemm {M}x{N}x{K}")


def test_gemm_unaligned_output_fallback(ops):
    """Output pitch not a multiple of 8 elements: the smem/TMA-store epilogue is not applicable, the generic direct-store
    epilogue (run-time flags) must still be exact."""
    M, N, K = 130, 100, 64
    x, w, b = _rand((M, K), 1.0, 17), _rand((N, K), 0.1, 18), _rand((N,), 0.5, 19)
    out = ops.linear(x, w, b, gelu=True)
    ref = torch.nn.functional.gelu((x.float() @ w.float().T + b.float()).to(bf16).float()).to(bf16)
    _close(out, ref, BF16_RTOL, 2e-3, "unaligned N")
    out32 = ops.linear(x[:20], w, b)  # few-token path with an odd feature count
    ref32 = (x[:20].float() @ w.float().T + b.float()).to(bf16)
    _close(out32, ref32, BF16_RTOL, 2e-3, "unaligned N (few tokens)")


@pytest.mark.parametrize("M", [1, 7, 32, 33, 64])
@pytest.mark.parametrize("N,K", [(256, 128), (3584, 512), (200, 64), (3584, 3584), (512, 4096), (96, 64), (4608, 1280)])
def test_gemm_swap_small_m(ops, M, N, K):
    x, w, b = _rand((M, K), 1.0, 3), _rand((N, K), 0.05, 4), _rand((N,), 0.5, 5)
    out = ops.linear(x, w, b)
    ref = (x.float() @ w.float().T + b.float()).to(bf16)
    _close(out, ref, BF16_RTOL, 1e-3, f"swap gemm {M}x{N}x{K}")


def test_gemm_splitk_deterministic_and_epilogue(ops):
    """Decode-shaped GEMM (32 tokens, 3584 x 18944 down projection + residual): split-K partials are reduced in a fixed
    order by the last CTA, so repeated launches are bit-identical; result matches the fp32 restatement."""
    M, N, K = 32, 3584, 18944
    x, w, res = _rand((M, K), 1.0, 14), _rand((N, K), 0.02, 15), _rand((M, N), 1.0, 16)
    outs = [ops.linear(x, w, resid=res) for _ in range(5)]
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    lin = (x.float() @ w.float().T).to(bf16).float()
    ref = (lin + res.float()).to(bf16)
    # one bf16 ulp of the (larger) linear output may survive the residual add even where lin + res cancels
    err = (outs[0].float() - ref.float()).abs()
    tol = 4e-3 + BF16_RTOL * (lin.abs() + res.float().abs())
    assert (err <= tol).all(), f"max err {err.max().item()}"


@pytest.mark.parametrize("M", [16, 640])
def test_gemm_epilogues(ops, M):
    N, K = 512, 256
    x, w, b = _rand((M, K), 1.0, 6), _rand((N, K), 0.06, 7), _rand((N,), 0.5, 8)
    res = _rand((M, N), 1.0, 9)
    acc = x.float() @ w.float().T
    lin = (acc + b.float()).to(bf16)
    # bias + GELU(erf)
    out = ops.linear(x, w, b, gelu=True)
    ref = torch.nn.functional.gelu(lin.float()).to(bf16)
    _close(out, ref, BF16_RTOL, 2e-3, "bias+gelu")
    # bias + residual
    out = ops.linear(x, w, b, resid=res)
    ref = (lin.float() + res.float()).to(bf16)
    _close(out, ref, BF16_RTOL, 4e-3, "bias+resid")
    # gelu + periodic residual (positional embedding add)
    period = 8
    pos = _rand((period, N), 1.0, 10)
    out = ops.linear(x, w, b, gelu=True, resid=pos, res_period=period)
    ref = (torch.nn.functional.gelu(lin.float()).to(bf16).float() + pos.float().repeat(M // period, 1)).to(bf16)
    _close(out, ref, 2 * BF16_RTOL, 4e-3, "gelu+pos")  # two chained bf16 roundings: up to 2 ulp
    # fp32 out of bf16-rounded values (lm_head)
    out = ops.linear(x, w, out_f32=True)
    ref = acc.to(bf16).float()
    _close(out, ref, BF16_RTOL, 1e-3, "f32out")
    assert out.dtype == torch.float32


@pytest.mark.parametrize("M", [1, 20, 32, 50, 64])
def test_gemm_few_token_epilogues(ops, M):
    """bias / GELU / residual epilogues of the few-token (split-K) kernel at the o-projection's shape class, against the fp32 restatement;
    repeated launches are bit-identical (no atomics, fixed reduction order)."""
    N, K = 3584, 3584
    x, w, b = _rand((M, K), 1.0, 80), _rand((N, K), 0.02, 81), _rand((N,), 0.5, 82)
    res = _rand((M, N), 1.0, 83)
    lin = (x.float() @ w.float().T + b.float()).to(bf16)
    out = ops.linear(x, w, b, resid=res)
    assert torch.equal(out, ops.linear(x, w, b, resid=res))
    err = (out.float() - (lin.float() + res.float()).to(bf16).float()).abs()
    assert (err <= 4e-3 + BF16_RTOL * (lin.float().abs() + res.float().abs())).all(), f"bias+resid max err {err.max().item()}"
    _close(ops.linear(x, w, b, gelu=True), torch.nn.functional.gelu(lin.float()).to(bf16), BF16_RTOL, 2e-3, "bias+gelu")
    _close(ops.linear(x, w), (x.float() @ w.float().T).to(bf16), BF16_RTOL, 1e-3, "plain")


@pytest.mark.parametrize("M", [8, 500])
@pytest.mark.parametrize("F", [128, 384, 200])
def test_gemm_swiglu(ops, M, F):
    K = 256
    x, g, u = _rand((M, K), 1.0, 11), _rand((F, K), 0.08, 12), _rand((F, K), 0.08, 13)
    packed = ops.pack_gate_up(g, u)
    out = ops.swiglu_linear(x, packed, F)
    gate = (x.float() @ g.float().T).to(bf16)
    up = (x.float() @ u.float().T).to(bf16)
    ref = (torch.nn.functional.silu(gate.float()).to(bf16).float() * up.float()).to(bf16)
    _close(out, ref, 2 * BF16_RTOL, 2e-3, f"swiglu M={M} F={F}")
    if F % 128 == 0:
        # [gate; up] layout (the two nn.Linear weights as views of one matrix): same tiles, same arithmetic -> bit-identical
        out_c = ops.swiglu_linear(x, torch.cat([g, u], 0).contiguous(), F, concat=True)
        assert torch.equal(out_c, out)


def test_layernorm_and_pool(ops):
    rows, dim = 333, 1280
    x, g, b = _rand((rows, dim), 2.0, 20), _rand((dim,), 1.0, 21), _rand((dim,), 0.3, 22)
    out = ops.layernorm(x, g, b)
    ref = torch.nn.functional.layer_norm(x.float(), (dim,), g.float(), b.float(), 1e-5).to(bf16)
    _close(out, ref, BF16_RTOL, 4e-3, "layernorm")
    n_win, T = 3, 50
    x = _rand((n_win * T, dim), 2.0, 23)
    out = ops.avgpool_layernorm(x, n_win, T, g, b)
    pooled = torch.nn.functional.avg_pool1d(x.view(n_win, T, dim).permute(0, 2, 1).float(), 2, 2).to(bf16).permute(0, 2, 1)
    ref = torch.nn.functional.layer_norm(pooled.float(), (dim,), g.float(), b.float(), 1e-5).to(bf16).reshape(-1, dim)
    _close(out, ref, BF16_RTOL, 4e-3, "avgpool+layernorm")


def test_rmsnorm(ops):
    rows, dim = 77, 3584
    x, w = _rand((rows, dim), 3.0, 24), _rand((dim,), 1.0, 25)
    out = ops.rmsnorm(x, w, 1e-6)
    h = x.float()
    h = h * torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + 1e-6)
    ref = w * h.to(bf16)
    _close(out, ref, BF16_RTOL, 1e-3, "rmsnorm")
    idx = torch.tensor([5, 0, 76], dtype=torch.int32, device="cuda")
    out2 = ops.rmsnorm(x, w, 1e-6, row_idx=idx)
    assert torch.equal(out2, out[idx.long()])


def test_im2col(ops):
    n_win, C, T = 2, 128, 300
    x = _rand((n_win, C, T), 1.0, 26, torch.float32)
    cols = ops.im2col_conv1(x)
    xp = torch.nn.functional.pad(x.to(bf16), (1, 1))
    ref = torch.stack([xp[:, :, kk:kk + T] for kk in range(3)], dim=1)  # [w, kk, c, t]
    ref = ref.permute(0, 3, 1, 2).reshape(n_win * T, 3 * C)
    assert torch.equal(cols, ref)
    h = _rand((n_win * T, 64), 1.0, 27)
    cols2 = ops.im2col_conv2(h, n_win, T)
    hp = torch.nn.functional.pad(h.view(n_win, T, 64), (0, 0, 1, 1))
    T_out = (T - 1) // 2 + 1
    ref2 = torch.stack([hp[:, kk:kk + 2 * T_out:2, :] for kk in range(3)], dim=2).reshape(n_win * T_out, 3 * 64)
    assert torch.equal(cols2, ref2)


def test_conv_stem_as_gemm(ops):
    """conv1d(k3,p1)+GELU and conv1d(k3,s2,p1)+GELU through im2col + tcgen05 GEMM vs F.conv1d."""
    n_win, C, T, Dm = 2, 128, 200, 256
    x = _rand((n_win, C, T), 1.0, 28, torch.float32)
    w1, b1 = _rand((Dm, C, 3), 0.05, 29), _rand((Dm,), 0.1, 30)
    w2, b2 = _rand((Dm, Dm, 3), 0.03, 31), _rand((Dm,), 0.1, 32)
    w1p = w1.permute(0, 2, 1).reshape(Dm, 3 * C).contiguous()
    w2p = w2.permute(0, 2, 1).reshape(Dm, 3 * Dm).contiguous()
    h1 = ops.linear(ops.im2col_conv1(x), w1p, b1, gelu=True)
    ref1 = torch.nn.functional.gelu(torch.nn.functional.conv1d(x.to(bf16).float(), w1.float(), b1.float(), padding=1).to(bf16).float()).to(bf16)
    _close(h1.view(n_win, T, Dm).permute(0, 2, 1), ref1, BF16_RTOL, 2e-3, "conv1")
    h2 = ops.linear(ops.im2col_conv2(h1, n_win, T), w2p, b2, gelu=True)
    ref2 = torch.nn.functional.gelu(torch.nn.functional.conv1d(h1.view(n_win, T, Dm).permute(0, 2, 1).float(), w2.float(), b2.float(), stride=2, padding=1).to(bf16).float()).to(bf16)
    _close(h2.view(n_win, T // 2, Dm).permute(0, 2, 1), ref2, BF16_RTOL, 2e-3, "conv2")


def _sdpa_ref(q, k, v, scale, mask):
    # q [B,H,Tq,D], k/v [B,H,Tk,D] fp32; mask bool [B,1,Tq,Tk] True = visible
    s = (q @ k.transpose(-1, -2)) * scale
    s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    p = torch.nan_to_num(p, nan=0.0)
    return p @ v


@pytest.fixture(params=["v2", "v2_tpr2", "v1"])
def attn_impl(request, monkeypatch):
    """All variants of the prefill / encoder attention kernel go through the same parity cases: v2 (two Q tiles per CTA, P in
    tensor memory; one softmax thread per row = the default, two per row = AF3_ATTN_TPR=2) and the round-1 kernel
    (AF3_ATTN_V1=1, kept as cross-check)."""
    monkeypatch.setenv("AF3_ATTN_V1", "1" if request.param == "v1" else "0")
    monkeypatch.setenv("AF3_ATTN_TPR", "2" if request.param == "v2_tpr2" else "1")
    return request.param


@pytest.mark.parametrize("D,H,T,lens", [(64, 4, 200, None), (64, 20, 1500, None), (64, 3, 333, [333, 100, 7]), (128, 2, 256, None),
                                        (64, 2, 1500, [1500, 385, 12]), (128, 3, 700, [700, 129, 256])])
def test_attention_bidirectional(ops, attn_impl, D, H, T, lens):
    B = 3 if lens else 2
    qkv = _rand((B * T, 3 * H * D), 1.0, 40)
    out = torch.zeros((B * T, H * D), device="cuda", dtype=bf16)
    kv_len = torch.tensor(lens, dtype=torch.int32, device="cuda") if lens else None
    ops.attention(qkv, qkv[:, H * D:], qkv[:, 2 * H * D:], out.view(B, T, H * D), B=B, H=H, Hkv=H, D=D, Tq=T, Tk=T,
                  scale=0.125 if D == 64 else 128 ** -0.5, causal=False, kv_layout=0, ldq=3 * H * D, ldk=3 * H * D, kv_len=kv_len)
    q, k, v = [t.float().view(B, T, H, D).transpose(1, 2) for t in qkv.split(H * D, dim=1)]
    mask = torch.ones((B, 1, T, T), dtype=torch.bool, device="cuda")
    if lens:
        for b, L in enumerate(lens):
            mask[b, :, :, L:] = False
    ref = _sdpa_ref(q, k, v, 0.125 if D == 64 else 128 ** -0.5, mask).transpose(1, 2).reshape(B * T, H * D)
    # P is rounded to bf16 before P.V (as in flash-style kernels) and O to bf16: ~2^-8 relative on O(1) values
    _close(out, ref, 2e-2, 2e-2, f"attention D={D} T={T}")


@pytest.mark.parametrize("T,starts", [(256, None), (300, [0, 37, 250]), (130, [5, 129]), (780, [0, 3, 530]), (1000, None)])
def test_attention_causal_gqa_cache(ops, attn_impl, T, starts):
    H, Hkv, D, Tmax = 8, 2, 128, 1024
    B = len(starts) if starts else 2
    qkv = _rand((B * T, (H + 2 * Hkv) * D), 1.0, 41)
    k_cache = torch.zeros((B, Hkv, Tmax, D), device="cuda", dtype=bf16)
    v_cache = torch.zeros_like(k_cache)
    kk = qkv[:, H * D:(H + Hkv) * D].view(B, T, Hkv, D).transpose(1, 2)
    vv = qkv[:, (H + Hkv) * D:].view(B, T, Hkv, D).transpose(1, 2)
    k_cache[:, :, :T] = kk
    v_cache[:, :, :T] = vv
    out = torch.zeros((B * T, H * D), device="cuda", dtype=bf16)
    kv_start = torch.tensor(starts, dtype=torch.int32, device="cuda") if starts else None
    ops.attention(qkv, k_cache, v_cache, out.view(B, T, H * D), B=B, H=H, Hkv=Hkv, D=D, Tq=T, Tk=T, scale=D ** -0.5, causal=True,
                  kv_layout=1, Tk_pitch=Tmax, ldq=(H + 2 * Hkv) * D, ldk=D, kv_start=kv_start)
    q = qkv[:, :H * D].float().view(B, T, H, D).transpose(1, 2)
    k = kk.float().repeat_interleave(H // Hkv, dim=1)
    v = vv.float().repeat_interleave(H // Hkv, dim=1)
    mask = torch.tril(torch.ones((T, T), dtype=torch.bool, device="cuda"))[None, None].repeat(B, 1, 1, 1)
    if starts:
        for b, s0 in enumerate(starts):
            mask[b, :, :, :s0] = False
    ref = _sdpa_ref(q, k, v, D ** -0.5, mask).transpose(1, 2).reshape(B * T, H * D)
    valid = torch.ones((B, T), dtype=torch.bool, device="cuda")
    if starts:
        for b, s0 in enumerate(starts):
            valid[b, :s0] = False
    sel = valid.reshape(-1)
    _close(out[sel], ref[sel], 2e-2, 2e-2, f"causal attention T={T}")


def test_rope_kv_append(ops):
    B, T, H, Hkv, D, Tmax = 2, 50, 4, 2, 128, 128
    theta = 1e6
    qkv = _rand((B * T, (H + 2 * Hkv) * D), 1.0, 42)
    orig = qkv.clone()
    k_cache = torch.zeros((B, Hkv, Tmax, D), device="cuda", dtype=bf16)
    v_cache = torch.zeros_like(k_cache)
    inv_freq = 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))
    starts = torch.tensor([0, 9], dtype=torch.int32, device="cuda")
    ops.rope_kv_append(qkv, k_cache, v_cache, B=B, T=T, H=H, Hkv=Hkv, D=D, pos0=0, inv_freq=inv_freq.cuda(), kv_start=starts)
    pos = (torch.arange(T, device="cuda")[None, :] - starts[:, None]).clamp(min=-1)
    pos = torch.where(pos < 0, torch.ones_like(pos), pos).float()  # padded -> 1
    freqs = pos[:, :, None] * inv_freq.cuda()[None, None, :]
    emb = torch.cat([freqs, freqs], -1)
    cos, sin = emb.cos().to(bf16), emb.sin().to(bf16)

    def rot(x):  # x [B,T,h,D] bf16, reference op order with bf16 rounding at every op
        x1, x2 = x[..., : D // 2], x[..., D // 2:]
        rh = torch.cat([-x2, x1], -1)
        return x * cos[:, :, None, :] + rh * sin[:, :, None, :]

    o = orig.view(B, T, H + 2 * Hkv, D)
    q_ref, k_ref, v_ref = rot(o[:, :, :H]), rot(o[:, :, H:H + Hkv]), o[:, :, H + Hkv:]
    got = qkv.view(B, T, H + 2 * Hkv, D)
    assert (got[:, :, :H].float() - q_ref.float()).abs().max().item() <= 2 ** -6 * q_ref.float().abs().max().item()
    frac_exact = (got[:, :, :H] == q_ref).float().mean().item()
    assert frac_exact > 0.999, frac_exact  # bit-exact except where fp32 sin/cos differ in the last ulp
    assert (k_cache[:, :, :T].transpose(1, 2) == k_ref).float().mean().item() > 0.999
    assert torch.equal(v_cache[:, :, :T].transpose(1, 2), v_ref)


@pytest.mark.parametrize(
    "B,H,Hkv,Tmax,ctx,starts,splits",
    [
        (3, 28, 4, 1024, 777, [0, 100, 776], None),   # auto: one chunk per split
        (3, 28, 4, 1024, 777, [0, 100, 776], 1),      # one CTA streams 7 chunks (ring wraps twice)
        (3, 28, 4, 1024, 777, [0, 100, 776], 3),      # 3 / 3 / 1 chunks; row 2 has a single live key
        (2, 16, 1, 2048, 1901, [5, 1300], 1),         # G = 16 (no padded heads), 15 chunks in one CTA
        (2, 8, 4, 2048, 1537, [0, 640], 5),           # G = 2, uneven splits, last chunk holds one key
        (40, 28, 4, 512, 300, None, None),            # B * Hkv > SM count -> nz = 1, more CTAs than SMs
        (1, 28, 4, 128, 1, [0], None),                # a single key
    ],
)
def test_decode_attention(ops, monkeypatch, B, H, Hkv, Tmax, ctx, starts, splits):
    D = 128
    if splits is None:
        monkeypatch.delenv("AF3_DECODE_SPLITS", raising=False)
    else:
        monkeypatch.setenv("AF3_DECODE_SPLITS", str(splits))
    qkv = _rand((B, (H + 2 * Hkv) * D), 1.0, 43)
    k_cache, v_cache = _rand((B, Hkv, Tmax, D), 1.0, 44), _rand((B, Hkv, Tmax, D), 1.0, 45)
    starts_t = torch.tensor(starts if starts is not None else [0] * B, dtype=torch.int32, device="cuda")
    ctx_len = torch.tensor([ctx], dtype=torch.int32, device="cuda")
    out = torch.zeros((B, H * D), device="cuda", dtype=bf16)
    scratch = ops.decode_attention_scratch(B, H, D, Tmax, "cuda")
    for _ in range(2):  # second call: the arrival counters must have been left at zero
        out.zero_()
        ops.decode_attention(qkv, k_cache, v_cache, out, scratch, B=B, H=H, Hkv=Hkv, D=D, ctx_len=ctx_len,
                             kv_start=starts_t if starts is not None else None, scale=D ** -0.5)
    q = qkv[:, :H * D].float().view(B, H, 1, D)
    k = k_cache[:, :, :ctx].float().repeat_interleave(H // Hkv, dim=1)
    v = v_cache[:, :, :ctx].float().repeat_interleave(H // Hkv, dim=1)
    mask = torch.ones((B, 1, 1, ctx), dtype=torch.bool, device="cuda")
    for b in range(B):
        mask[b, :, :, :starts_t[b]] = False
    ref = _sdpa_ref(q, k, v, D ** -0.5, mask).reshape(B, H * D)
    _close(out, ref, 2e-2, 1e-2, "decode attention")


def test_embed_scatter_and_argmax(ops):
    V, dim, n_win, frames = 1000, 256, 3, 20
    table = _rand((V, dim), 1.0, 46)
    audio = _rand((n_win * frames, dim), 1.0, 47)
    post = torch.tensor([20, 7, 13], dtype=torch.int32, device="cuda")
    aid = 999
    g = torch.Generator().manual_seed(48)
    ids = torch.randint(0, 900, (2, 60), generator=g)
    ids[0, 5:25] = aid          # 20 audio tokens
    ids[1, 30:50] = aid         # 7 + 13
    ids = ids.cuda()
    out, counts = ops.embed_scatter(ids.reshape(-1), table, aid, audio, n_win, frames, post)
    ref = table[ids.reshape(-1)].clone()
    valid = (torch.arange(frames, device="cuda")[None, :] < post[:, None]).reshape(-1)
    ref[(ids.reshape(-1) == aid)] = audio[valid]
    assert torch.equal(out, ref)
    assert counts.tolist() == [40, 40]
    logits = torch.randn((5, 152064), device="cuda")
    logits[2, 777] = logits[2, 90000] = 50.0  # tie -> first index
    got = ops.argmax(logits)
    assert torch.equal(got, logits.argmax(-1))
    assert got[2].item() == 777


def test_logmel_vs_fp64(ops):
    """Log-mel kernel vs an fp64 numpy restatement of WFE:135-164 (exact DFT); tolerance 1e-4 abs on the
    (x+4)/4 scale at fp32 (the reference quotes 1e-5 between its own torch and numpy paths on speech)."""
    from transformers.audio_utils import mel_filter_bank

    filt = mel_filter_bank(num_frequency_bins=201, num_mel_filters=128, min_frequency=0.0, max_frequency=8000.0,
                           sampling_rate=16000, norm="slaney", mel_scale="slaney")
    tables = ops.LogMelTables(filt, "cuda")
    n = 480000
    rs = np.random.RandomState(0)
    wave = (rs.randn(2, n) * 0.1).astype(np.float32)
    wave[1, 160000:] = 0.0  # zero-padded short clip
    got = ops.logmel(torch.from_numpy(wave).cuda(), tables).cpu().numpy()
    # fp64 reference
    win = torch.hann_window(400, dtype=torch.float64)
    st = torch.stft(torch.from_numpy(wave).double(), 400, 160, window=win, return_complex=True)
    mag = st[..., :-1].abs() ** 2
    mel = torch.from_numpy(filt).double().T @ mag
    ls = torch.clamp(mel, min=1e-10).log10()
    mx = ls.amax(dim=(1, 2), keepdim=True)
    ref = ((torch.maximum(ls, mx - 8.0) + 4.0) / 4.0).numpy()
    err = np.abs(got - ref)
    assert got.shape == (2, 128, 3000)
    assert err.max() < 1e-4, err.max()


@pytest.mark.parametrize("K", [512, 3584])
@pytest.mark.parametrize("B", [32, 5, 48, 64, 1])
def test_qkv_rope_fused_gemm_matches_unfused(ops, B, K):
    """Decode-step fusion: q/k/v projection + RoPE + KV append in the GEMM epilogue == GEMM then af3_rope_kv_append
    (up to 64 sequences).  The unfused chain runs through the same kernel, so the comparison is bit-exact: same accumulation
    order, same rounding points."""
    H, Hkv, D, Tmax, slot = 28, 4, 128, 64, 37
    x, w, b = _rand((B, K), 1.0, 60), _rand(((H + 2 * Hkv) * D, K), 0.05, 61), _rand(((H + 2 * Hkv) * D,), 0.3, 62)
    inv_freq = (1.0 / (1e6 ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))).cuda()
    starts = torch.randint(0, 30, (B,), dtype=torch.int32).cuda()
    pos = torch.tensor([slot], dtype=torch.int32, device="cuda")
    kc1, vc1 = torch.zeros((B, Hkv, Tmax, D), device="cuda", dtype=bf16), torch.zeros((B, Hkv, Tmax, D), device="cuda", dtype=bf16)
    kc2, vc2 = torch.zeros_like(kc1), torch.zeros_like(vc1)
    qkv1 = ops.linear(x, w, b)
    ops.rope_kv_append(qkv1, kc1, vc1, B=B, T=1, H=H, Hkv=Hkv, D=D, pos0=0, inv_freq=inv_freq, kv_start=starts, pos0_dev=pos)
    cs = ops.rope_table(B, D, pos, starts, inv_freq)
    qkv2 = ops.qkv_rope_linear(x, w, b, kc2, vc2, H=H, Hkv=Hkv, D=D, rope_cs=cs, pos_dev=pos)
    assert torch.equal(qkv2[:, :H * D], qkv1[:, :H * D])
    assert torch.equal(kc2, kc1) and torch.equal(vc2, vc1)


def test_qkv_rope_fused_gemm_never_writes_past_the_cache(ops):
    """with the cache full (slot == Tmax) the fused epilogue must drop the append instead of writing into the next
    (sequence, head) / past the allocation; the query heads are still produced."""
    B, H, Hkv, D, K, Tmax = 4, 28, 4, 128, 256, 16
    x, w, b = _rand((B, K), 1.0, 70), _rand(((H + 2 * Hkv) * D, K), 0.05, 71), _rand(((H + 2 * Hkv) * D,), 0.3, 72)
    inv_freq = (1.0 / (1e6 ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))).cuda()
    guard = torch.full((2, B, Hkv, Tmax + 8, D), 7.0, device="cuda", dtype=bf16)   # 8 guard rows behind every (seq, head) block
    kc, vc = guard[0, :, :, :Tmax], guard[1, :, :, :Tmax]
    kc_c, vc_c = kc.contiguous(), vc.contiguous()                                  # the kernel takes dense [B, Hkv, Tmax, D] caches
    pos = torch.tensor([Tmax], dtype=torch.int32, device="cuda")
    cs = ops.rope_table(B, D, pos, None, inv_freq)
    before_k, before_v = kc_c.clone(), vc_c.clone()
    q = ops.qkv_rope_linear(x, w, b, kc_c, vc_c, H=H, Hkv=Hkv, D=D, rope_cs=cs, pos_dev=pos)
    torch.cuda.synchronize()
    assert torch.equal(kc_c, before_k) and torch.equal(vc_c, before_v)
    assert torch.isfinite(q[:, : H * D].float()).all()
    # the stand-alone append kernel with a device-side position has the same guard
    qkv = ops.linear(x, w, b)
    ops.rope_kv_append(qkv, kc_c, vc_c, B=B, T=1, H=H, Hkv=Hkv, D=D, pos0=0, inv_freq=inv_freq, pos0_dev=pos)
    torch.cuda.synchronize()
    assert torch.equal(kc_c, before_k) and torch.equal(vc_c, before_v)


def test_attention_v2_matches_v1_on_a_live_cache_offset(ops, monkeypatch):
    """Continuation chunk (chunked prefill / next chat turn): Tq = 200 new rows against Tk = 650 cached + new keys, causal with the
    (Tk - Tq) offset, GQA 28:4, left padding on one row.  The two kernel generations must agree to within bf16 output rounding, and
    both with the fp32 restatement."""
    B, H, Hkv, D, Tq, Tk, Tmax = 2, 28, 4, 128, 200, 650, 768
    q = _rand((B * Tq, (H + 2 * Hkv) * D), 1.0, 90)
    k_cache, v_cache = _rand((B, Hkv, Tmax, D), 1.0, 91), _rand((B, Hkv, Tmax, D), 1.0, 92)
    k_cache[:, :, Tk:] = 0
    v_cache[:, :, Tk:] = 0
    kv_start = torch.tensor([0, 77], dtype=torch.int32, device="cuda")
    outs = {}
    for impl in ("0", "1"):
        monkeypatch.setenv("AF3_ATTN_V1", impl)
        out = torch.zeros((B * Tq, H * D), device="cuda", dtype=bf16)
        ops.attention(q, k_cache, v_cache, out.view(B, Tq, H * D), B=B, H=H, Hkv=Hkv, D=D, Tq=Tq, Tk=Tk, scale=D ** -0.5, causal=True,
                      kv_layout=1, Tk_pitch=Tmax, ldq=(H + 2 * Hkv) * D, ldk=D, kv_start=kv_start)
        outs[impl] = out
    qq = q[:, :H * D].float().view(B, Tq, H, D).transpose(1, 2)
    kk = k_cache[:, :, :Tk].float().repeat_interleave(H // Hkv, dim=1)
    vv = v_cache[:, :, :Tk].float().repeat_interleave(H // Hkv, dim=1)
    qi = torch.arange(Tq, device="cuda")[:, None] + (Tk - Tq)
    mask = (torch.arange(Tk, device="cuda")[None, :] <= qi)[None, None].repeat(B, 1, 1, 1)
    mask[1, :, :, :77] = False
    ref = _sdpa_ref(qq, kk, vv, D ** -0.5, mask).transpose(1, 2).reshape(B * Tq, H * D)
    _close(outs["0"], ref, 2e-2, 2e-2, "v2 vs fp32")
    _close(outs["1"], ref, 2e-2, 2e-2, "v1 vs fp32")
    _close(outs["0"], outs["1"], 2 ** -6, 2e-3, "v2 vs v1")


@pytest.mark.parametrize("B,K,N,mode", [(32, 3584, 4608, "plain"), (7, 3584, 3584, "plain"), (32, 3584, 18944, "swiglu"), (32, 256, 384, "swiglu"),
                                        (48, 512, 256, "plain")])
def test_fused_rmsnorm_across_few_token_gemms(ops, B, K, N, mode):
    """Decode-step fusion (af3_gemm_fusion): a residual GEMM emits per-row-tile sums of squares of what it stores; the next GEMM
    normalises its activation tiles in shared memory.  Producer: partials == torch's sums over each 128-feature tile of the stored
    bf16 values.  Consumer: GEMM(x, norm=...) vs GEMM(rmsnorm_kernel(x)): identical up to the last fp32 bit of rstd, i.e. at most one
    bf16 ulp on a vanishing share of the activations -> outputs within bf16 rounding of each other."""
    F = K                                  # the producer's output features are the consumer's K
    a, wo, res = _rand((B, 640), 1.0, 80), _rand((F, 640), 0.05, 81), _rand((B, F), 1.0, 82)
    ss = ops.sumsq_buffer(B, "cuda")
    n_parts = -(-F // 128)
    h = ops.linear(a, wo, resid=res.clone(), sumsq_out=ss)
    h_plain = ops.linear(a, wo, resid=res.clone())
    assert torch.equal(h, h_plain)
    ref_ss = h.float().pow(2).view(B, -1, 128).sum(-1) if F % 128 == 0 else None
    if ref_ss is not None:
        assert torch.allclose(ss[:, :n_parts], ref_ss, rtol=1e-5, atol=1e-6), (ss[:, :n_parts] - ref_ss).abs().max().item()
    wn = _rand((K,), 0.3, 83) + 1.0
    wn = wn.to(bf16)
    y = ops.rmsnorm(h, wn, 1e-6)
    if mode == "swiglu":
        g, u = _rand((N, K), 0.05, 84), _rand((N, K), 0.05, 85)
        wp = ops.pack_gate_up(g, u)
        unf = ops.swiglu_linear(y, wp, N)
        fus = ops.swiglu_linear(h, wp, N, norm=(wn, ss, n_parts, 1e-6))
    else:
        w2, b2 = _rand((N, K), 0.05, 86), _rand((N,), 0.3, 87)
        unf = ops.linear(y, w2, b2)
        fus = ops.linear(h, w2, b2, norm=(wn, ss, n_parts, 1e-6))
    d = (fus.float() - unf.float()).abs()
    tol = 2 ** -7 * unf.float().abs() + 2e-2
    assert int((d > tol).sum()) == 0, f"max diff {d.max().item()}"
    assert (d > 0).float().mean().item() < 0.05, "fused and unfused should agree bit for bit almost everywhere"


def test_token_step_is_the_references_per_token_rule(ops):
    """af3_token_step against the reference's lines restated with torch ops ([O] GEN:2797-2805): finished rows emit pad, the token is
    appended, the unfinished mask is updated with the EOS set, "every row finished" is published -- over several consecutive tokens,
    with and without EOS handling, B above one thread block's width."""
    for B, n_eos in ((5, 1), (300, 2), (64, 0)):
        g = torch.Generator(device="cpu").manual_seed(B)
        cap, steps, pad = 16, 6, 7
        eos = torch.tensor([3, 11][:max(n_eos, 1)], dtype=torch.int64).cuda()
        eos_dev = torch.zeros((8,), dtype=torch.int64, device="cuda")
        eos_dev[: eos.numel()] = eos
        ctl = torch.tensor([n_eos, pad], dtype=torch.int64).cuda()
        unfinished = torch.ones((B,), dtype=torch.int32, device="cuda")
        tok_buf = torch.full((B, cap), -1, dtype=torch.int64, device="cuda")
        gen_idx = torch.zeros((1,), dtype=torch.int32, device="cuda")
        ids_out = torch.zeros((B,), dtype=torch.int64, device="cuda")
        flags = torch.full((cap,), -1, dtype=torch.int32, device="cuda")
        ref_unf = torch.ones((B,), dtype=torch.int64)
        for i in range(steps):
            raw = torch.randint(0, 14 if i < steps - 1 else 4, (B,), generator=g, dtype=torch.int64)
            if i == steps - 1 and n_eos:
                raw[:] = 3   # everybody finishes on the last token at the latest
            ops.token_step(raw.cuda(), unfinished, eos_dev, ctl, tok_buf, gen_idx, ids_out, flags)
            nxt = raw.clone()
            if n_eos:
                nxt = nxt * ref_unf + pad * (1 - ref_unf)                                # GEN:2797
                ref_unf = ref_unf & ~torch.isin(nxt, eos.cpu())                          # GEN:2803
            assert torch.equal(tok_buf[:, i].cpu(), nxt) and torch.equal(ids_out.cpu(), nxt)
            assert int(flags[i]) == (1 if (n_eos and int(ref_unf.max()) == 0) else 0)    # GEN:2805
            if n_eos:
                assert torch.equal(unfinished.cpu().long(), ref_unf)
        assert int(gen_idx) == steps and (tok_buf[:, steps:] == -1).all() and (flags[steps:] == -1).all()
