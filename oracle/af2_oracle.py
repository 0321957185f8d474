"""ORACLE (test infrastructure only) for SURVEY.md 8-f.4 -- Audio Flamingo 2's gated cross-attention conditioning.

PARITY UNPINNED FOR AF2 ITSELF: the `audio_flamingo_2` branch ([R] /root/reference/README.md:159) is not in the mounted snapshot
and no AF2 / open_flamingo / AF-CLAP code exists in this container.  What this file pins is the OPERATOR AF2's language-model side
is built from -- Flamingo's gated xattn-dense block (arXiv 2204.14198 sec. 2.2; AF2: arXiv 2503.03983) -- against the one
executable implementation of that operator available here, transformers' IdeficsGatedCrossAttentionLayer
([O] transformers/models/idefics/modeling_idefics.py:684-806, cited as IDX):
  * hf_gated_layer(): the unmodified reference-analogue layer on seeded weights ("the analogue run here"),
  * ref_gated_layer(): a plain torch fp32 restatement, each step citing the IDX line it follows, pinned against the former in
    tests/test_oracle_cpu.py.
The AF-CLAP sliding-window encoder, the window/time positional scheme and the placement of these blocks in AF2's 3B decoder are
NOT restated: nothing in the container states them.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

XATTN_SMALL = dict(hidden_size=512, num_attention_heads=4, intermediate_size=768, media_dim=384)


def hf_gated_layer(seed: int = 0, alpha_type: str = "vector", alpha_scale: float = 0.5, **dims):
    """IdeficsGatedCrossAttentionLayer with seeded default init; alphas ~ N(0, alpha_scale) so that the gates are open
    (Flamingo initialises them at 0 = closed, which would make the block an identity and the test vacuous)."""
    from transformers import IdeficsConfig
    from transformers.models.idefics.modeling_idefics import IdeficsGatedCrossAttentionLayer

    d = dict(XATTN_SMALL, **dims)
    cfg = IdeficsConfig(hidden_size=d["hidden_size"], num_attention_heads=d["num_attention_heads"], intermediate_size=d["intermediate_size"],
                        vision_config={"embed_dim": d["media_dim"]}, alpha_initializer="ones", alpha_type=alpha_type, qk_layer_norms=False,
                        rms_norm_eps=1e-6, hidden_act="silu", num_hidden_layers=1, vocab_size=64)
    cfg._attn_implementation = "sdpa"
    torch.manual_seed(seed)
    layer = IdeficsGatedCrossAttentionLayer(cfg, layer_idx=0).eval()
    with torch.no_grad():   # (the analogue's own "normal" initialiser raises for alpha_type="float": size=(1) is not a tuple)
        layer.alpha_cross_attn.normal_(0.0, alpha_scale)
        layer.alpha_dense.normal_(0.0, alpha_scale)
    return layer


def key_padding_mask(B: int, T: int, media_len, Tm: int, dtype=torch.float32):
    """Additive [B, 1, T, Tm] mask as the analogue expects it (IDX:770-773): media rows >= media_len[b] are masked keys."""
    vis = torch.arange(Tm)[None, :] < torch.as_tensor(media_len)[:, None]
    m = torch.zeros((B, 1, T, Tm), dtype=dtype)
    return m.masked_fill(~vis[:, None, None, :], torch.finfo(dtype).min)


def _rms(x, w, eps):                       # IdeficsRMSNorm.forward (IDX:379-388)
    return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def ref_gated_layer(sd, hidden_states, media, media_len, cross_attention_gate=None, n_heads=4, eps=1e-6):
    """fp32 restatement of IdeficsGatedCrossAttentionLayer.forward (IDX:784-806)."""
    sd = {k: v.float() for k, v in sd.items()}
    x, m = hidden_states.float(), media.float()
    B, T, hid = x.shape
    Tm = m.shape[1]
    D = hid // n_heads
    h = _rms(x, sd["input_layernorm.weight"], eps)                                                     # IDX:789
    q = F.linear(h, sd["cross_attn.q_proj.weight"]).view(B, T, n_heads, D).transpose(1, 2)            # IDX:593
    k = F.linear(m, sd["cross_attn.k_proj.weight"]).view(B, Tm, n_heads, D).transpose(1, 2)           # IDX:599
    v = F.linear(m, sd["cross_attn.v_proj.weight"]).view(B, Tm, n_heads, D).transpose(1, 2)
    vis = torch.arange(Tm)[None, :] < torch.as_tensor(media_len)[:, None]
    s = (q @ k.transpose(-1, -2)) * D ** -0.5
    s = s.masked_fill(~vis[:, None, None, :], float("-inf"))
    a = torch.nan_to_num(torch.softmax(s, -1), nan=0.0) @ v                                            # IDX:626-635 (sdpa)
    a = F.linear(a.transpose(1, 2).reshape(B, T, hid), sd["cross_attn.o_proj.weight"])                # IDX:638
    if cross_attention_gate is not None:
        a = a.masked_fill((cross_attention_gate == 0)[:, :, None], 0.0)                               # IDX:797
    x = x + torch.tanh(sd["alpha_cross_attn"]) * a                                                     # IDX:798
    h = _rms(x, sd["post_attention_layernorm.weight"], eps)                                            # IDX:802
    h = F.linear(F.silu(F.linear(h, sd["mlp.gate_proj.weight"])) * F.linear(h, sd["mlp.up_proj.weight"]), sd["mlp.down_proj.weight"])
    return x + torch.tanh(sd["alpha_dense"]) * h                                                       # IDX:805
