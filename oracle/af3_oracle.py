"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under audio_flamingo_b200/ may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker.

What it is
----------
/root/reference (NVIDIA/audio-flamingo @ f4579633, `main`) contains no source code (README + images; SURVEY.md
section 0), so the reference's own implementation cannot be compiled or imported.  The executable statement of the
AF3 audio->text path is the upstream Hugging Face port that the README's checkpoint link (`nvidia/audio-flamingo-3-hf`,
/root/reference/README.md:79) loads: transformers==5.5.0 (pip-installed third-party dependency, not vendored):
    AF3M = transformers/models/audioflamingo3/modeling_audioflamingo3.py
    WFE  = transformers/models/whisper/feature_extraction_whisper.py     AU = transformers/audio_utils.py
    Q2M  = transformers/models/qwen2/modeling_qwen2.py                   GEN = transformers/generation/utils.py
This module holds
  (1) `hf_*` helpers that instantiate those classes unmodified on seeded synthetic weights (the "reference run here"),
  (2) `ref_*` functions: a plain torch-fp32 CPU restatement of the same algorithm, each citing the lines it follows,
  (3) seeded synthetic inputs: waveforms, prompts, weights.
Pinning: the reference snapshot ships no golden vectors / tests, so the restatement is pinned against
(1) executed live (tests/test_oracle_cpu.py) and against fixtures generated from (1) by tests/golden/make_golden.py
and committed under tests/golden/.  No real checkpoint exists offline: parity is on synthetic weights only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------ configs
AF3_7B = dict(
    audio=dict(num_mel_bins=128, num_hidden_layers=32, num_attention_heads=20, intermediate_size=5120, hidden_size=1280,
               max_source_positions=1500),
    # Qwen2.5-7B dimensions (published values; the checkpoint's config.json is not available offline)
    text=dict(vocab_size=152064, hidden_size=3584, intermediate_size=18944, num_hidden_layers=28, num_attention_heads=28,
              num_key_value_heads=4, max_position_embeddings=32768, rms_norm_eps=1e-6, rope_theta=1000000.0,
              tie_word_embeddings=False),
    audio_token_id=151669,
)

TINY = dict(
    audio=dict(num_mel_bins=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, hidden_size=128,
               max_source_positions=1500),
    text=dict(vocab_size=2048, hidden_size=256, intermediate_size=384, num_hidden_layers=2, num_attention_heads=2,
              num_key_value_heads=1, max_position_embeddings=4096, rms_norm_eps=1e-6, rope_theta=1000000.0,
              tie_word_embeddings=False),
    audio_token_id=2047,
)

# AF-Whisper at full width, shallow decoder at full width: exercises every full-size kernel shape cheaply
MID = dict(
    audio=dict(AF3_7B["audio"], num_hidden_layers=2),
    text=dict(AF3_7B["text"], num_hidden_layers=2),
    audio_token_id=151669,
)

PRESETS = {"tiny": TINY, "mid": MID, "af3-7b": AF3_7B}


def hf_config(preset: str | dict):
    """AudioFlamingo3Config from a preset ([O] AF3C:26-117)."""
    from transformers import AudioFlamingo3Config

    p = PRESETS[preset] if isinstance(preset, str) else preset
    text = dict(p["text"])
    theta = text.pop("rope_theta")
    cfg = AudioFlamingo3Config(audio_config=dict(p["audio"]),
                               text_config=dict(text, rope_parameters={"rope_type": "default", "rope_theta": theta}),
                               audio_token_id=p["audio_token_id"])
    cfg.text_config.pad_token_id = None
    cfg.text_config.eos_token_id = None
    cfg.text_config.bos_token_id = None
    return cfg


def hf_model(preset: str | dict, seed: int = 0, dtype=torch.float32, device="cpu", sharpen: float = 1.0):
    """The unmodified HF model on seeded synthetic weights (default init under torch.manual_seed(seed)).

    sharpen > 1 multiplies lm_head.weight: default N(0, 0.02) init gives top-1/top-2 logit margins comparable to
    bf16 rounding noise; a documented scale makes greedy-token parity a meaningful test (SURVEY.md 7 "hard parts").
    """
    from transformers import AudioFlamingo3ForConditionalGeneration

    cfg = hf_config(preset)
    torch.manual_seed(seed)
    model = AudioFlamingo3ForConditionalGeneration(cfg)
    if sharpen != 1.0:
        with torch.no_grad():
            model.language_model.lm_head.weight.mul_(sharpen)
    model = model.to(dtype=dtype, device=device).eval()
    model.generation_config.pad_token_id = 0
    model.generation_config.eos_token_id = None
    return model


def hf_restore_fp32_rotary(model):
    """Recompute the rotary inv_freq buffers in fp32 on the CPU and move them to the model's device.  The reference's normal
    loading path (from_pretrained(dtype=torch.bfloat16)) casts PARAMETERS; this non-persistent buffer is built in fp32
    ([O] Q2M:86-89) and stays fp32.  A blanket model.to(torch.bfloat16) -- convenient in tests -- would round it to 8 bits and
    run RoPE at slightly different frequencies than any real deployment of the reference; call this after such a cast (and after
    to_empty(), which leaves the buffer uninitialised)."""
    for mod in model.modules():
        if hasattr(mod, "inv_freq") and hasattr(mod, "compute_default_rope_parameters"):
            inv, _ = mod.compute_default_rope_parameters(mod.config)
            dev = mod.inv_freq.device if mod.inv_freq.device.type != "meta" else "cpu"
            mod.inv_freq = inv.to(dev)
            mod.original_inv_freq = inv.clone().to(dev)
    return model


def hf_feature_extractor():
    from transformers import WhisperFeatureExtractor

    return WhisperFeatureExtractor(feature_size=128)  # AF3 uses 128 mel bins (AF3C:56)


# ------------------------------------------------------------------------------------------------ synthetic inputs
def synth_waveforms(n_clips: int, seconds, seed: int = 0) -> list[np.ndarray]:
    """Seeded 16 kHz mono noise clips. seconds: float or list per clip."""
    secs = [seconds] * n_clips if np.isscalar(seconds) else list(seconds)
    return [(np.random.RandomState(seed + i).randn(int(round(s * 16000))) * 0.1).astype(np.float32)
            for i, s in enumerate(secs)]


def post_pool_len(n_frames: int) -> int:
    """frames -> conv2 length -> avg-pooled length ([O] AF3M:375-377, AF3P:96-101)."""
    return ((n_frames - 1) // 2 + 1 - 2) // 2 + 1


def synth_prompt(tok_counts: list[int], vocab: int, audio_token_id: int, seed: int = 0, n_pre: int = 5, n_post: int = 25,
                 pad_id: int = 0):
    """[n_pre text ids] + [audio_token_id]*tok + [n_post text ids], left padded ([O] AF3P:44-47).  Returns
    (input_ids int64 [B,S], attention_mask int64 [B,S])."""
    rs = np.random.RandomState(seed)
    hi = min(vocab, audio_token_id) - 1
    rows = []
    for tok in tok_counts:
        pre = rs.randint(1, hi, size=n_pre)
        post = rs.randint(1, hi, size=n_post)
        rows.append(np.concatenate([pre, np.full(tok, audio_token_id), post]))
    S = max(len(r) for r in rows)
    ids = np.full((len(rows), S), pad_id, dtype=np.int64)
    mask = np.zeros((len(rows), S), dtype=np.int64)
    for i, r in enumerate(rows):
        ids[i, S - len(r):] = r
        mask[i, S - len(r):] = 1
    return torch.from_numpy(ids), torch.from_numpy(mask)


def hf_features(waves: list[np.ndarray]):
    """WhisperFeatureExtractor exactly as AF3P:182-187 calls it (one <=30 s window per clip here)."""
    fe = hf_feature_extractor()
    out = fe(waves, sampling_rate=16000, return_attention_mask=True, padding="max_length", return_tensors="pt")
    return out["input_features"], out["attention_mask"]


# ------------------------------------------------------------------------------------------------ restatement
def ref_mel_filters() -> np.ndarray:
    """[201,128] slaney filterbank, built by the same numpy code path as the reference ([O] WFE:95-103 -> AU:453-544)."""
    from transformers.audio_utils import mel_filter_bank

    return mel_filter_bank(num_frequency_bins=201, num_mel_filters=128, min_frequency=0.0, max_frequency=8000.0,
                           sampling_rate=16000, norm="slaney", mel_scale="slaney")


def ref_logmel(wave: np.ndarray, dtype=torch.float64) -> np.ndarray:
    """[n, 480000] -> [n,128,3000]; WFE:135-164 restated with an explicit framed DFT in `dtype`:
    reflect-pad 200, frames of 400 hop 160, periodic Hann, |rfft|^2, drop last frame, mel, log10(clamp 1e-10),
    per-clip max-8 floor, (x+4)/4."""
    x = torch.from_numpy(np.asarray(wave)).to(dtype)
    n = x.shape[-1]
    xp = F.pad(x[:, None, :], (200, 200), mode="reflect")[:, 0]        # torch.stft(center=True, pad_mode="reflect")
    frames = xp.unfold(-1, 400, 160)                                     # [n_clips, 1 + n//160, 400]
    win = torch.hann_window(400, periodic=True, dtype=dtype)             # WFE:141
    spec = torch.fft.rfft(frames * win, dim=-1)                          # [., frames, 201]
    power = (spec.real ** 2 + spec.imag ** 2)[:, :-1].transpose(1, 2)    # WFE:150 drop last frame -> [., 201, 3000]
    mel = torch.from_numpy(ref_mel_filters()).to(dtype).T @ power        # WFE:153
    ls = torch.clamp(mel, min=1e-10).log10()                             # WFE:155
    mx = ls.amax(dim=(1, 2), keepdim=True)                               # WFE:157
    ls = torch.maximum(ls, mx - 8.0)
    return ((ls + 4.0) / 4.0).to(torch.float32).numpy()                  # WFE:161


def ref_frame_mask(n_samples: list[int], n_total: int = 480000) -> np.ndarray:
    """sample-level attention mask subsampled by hop ([O] WFE:328-337: mask[:, ::160])."""
    m = np.zeros((len(n_samples), n_total), dtype=np.int32)
    for i, n in enumerate(n_samples):
        m[i, :n] = 1
    return m[:, ::160]


def _sd(model_or_sd):
    return model_or_sd if isinstance(model_or_sd, dict) else model_or_sd.state_dict()


def ref_encoder(sd, cfg, feats: torch.Tensor, feat_mask: torch.Tensor) -> torch.Tensor:
    """AudioFlamingo3Encoder.forward ([O] AF3M:319-368) in fp32 with explicit softmax attention.
    feats [W,128,3000], feat_mask [W,3000] -> [W,750,d]."""
    sd = {k: v.float() for k, v in _sd(sd).items()}
    p = "audio_tower."
    ac = cfg.audio_config
    H, d = ac.num_attention_heads, ac.hidden_size
    hd = d // H
    x = feats.float()
    L = (feat_mask.sum(-1) - 1) // 2 + 1                                         # AF3M:338-339
    T = (x.shape[-1] - 1) // 2 + 1
    key_ok = torch.arange(T)[None, :] < L[:, None]                               # AF3M:340
    x = F.gelu(F.conv1d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1))          # AF3M:343
    x = F.gelu(F.conv1d(x, sd[p + "conv2.weight"], sd[p + "conv2.bias"], stride=2, padding=1))  # AF3M:344
    x = x.permute(0, 2, 1) + sd[p + "embed_positions.weight"]                    # AF3M:345-348
    bias = torch.zeros(x.shape[0], 1, 1, T).masked_fill(~key_ok[:, None, None, :], float("-inf"))  # MASK:1001-1087
    for i in range(ac.num_hidden_layers):
        lp = f"{p}layers.{i}."
        r = x
        h = F.layer_norm(x, (d,), sd[lp + "self_attn_layer_norm.weight"], sd[lp + "self_attn_layer_norm.bias"])
        q = (F.linear(h, sd[lp + "self_attn.q_proj.weight"], sd[lp + "self_attn.q_proj.bias"]) * hd ** -0.5)  # AF3M:142
        k = F.linear(h, sd[lp + "self_attn.k_proj.weight"])                      # no bias, AF3M:111
        v = F.linear(h, sd[lp + "self_attn.v_proj.weight"], sd[lp + "self_attn.v_proj.bias"])
        W = x.shape[0]
        q, k, v = [t.view(W, T, H, hd).transpose(1, 2) for t in (q, k, v)]
        a = torch.softmax(q @ k.transpose(-1, -2) + bias, dim=-1) @ v            # scaling=1.0, AF3M:181
        a = a.transpose(1, 2).reshape(W, T, d)
        x = r + F.linear(a, sd[lp + "self_attn.out_proj.weight"], sd[lp + "self_attn.out_proj.bias"])
        r = x
        h = F.layer_norm(x, (d,), sd[lp + "final_layer_norm.weight"], sd[lp + "final_layer_norm.bias"])
        h = F.gelu(F.linear(h, sd[lp + "fc1.weight"], sd[lp + "fc1.bias"]))
        x = r + F.linear(h, sd[lp + "fc2.weight"], sd[lp + "fc2.bias"])
    x = F.avg_pool1d(x.permute(0, 2, 1), 2, 2).permute(0, 2, 1)                  # AF3M:364-365
    return F.layer_norm(x, (d,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"])      # AF3M:366


def ref_projector(sd, x: torch.Tensor) -> torch.Tensor:
    """AF3M:398-402."""
    sd = _sd(sd)
    p = "multi_modal_projector."
    h = F.gelu(F.linear(x.float(), sd[p + "linear_1.weight"].float(), sd[p + "linear_1.bias"].float()))
    return F.linear(h, sd[p + "linear_2.weight"].float(), sd[p + "linear_2.bias"].float())


def ref_audio_embeds(sd, cfg, feats, feat_mask) -> torch.Tensor:
    """get_audio_features().pooler_output ([O] AF3M:447-475): valid frames of every window, concatenated."""
    e = ref_projector(sd, ref_encoder(sd, cfg, feats, feat_mask))
    post = torch.tensor([post_pool_len(int(n)) for n in feat_mask.sum(-1)])
    valid = torch.arange(e.shape[1])[None, :] < post[:, None]
    return e[valid]


def _rope(cfg_t, pos: torch.Tensor):
    """Qwen2RotaryEmbedding ([O] Q2M:86-113), fp32."""
    D = cfg_t.hidden_size // cfg_t.num_attention_heads
    theta = cfg_t.rope_parameters["rope_theta"]
    inv = 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))
    fr = pos[..., None].float() * inv
    emb = torch.cat([fr, fr], -1)
    return emb.cos(), emb.sin()


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], -1)


def ref_decoder_logits(sd, cfg, inputs_embeds: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """Qwen2ForCausalLM over the full (left padded) prompt ([O] Q2M:353-413, 462-475), fp32, no cache:
    returns logits [B,S,V].  Position ids = cumsum(mask)-1 with 1 at padding (GEN:719-721)."""
    sd = {k: v.float() for k, v in _sd(sd).items() if k.startswith("language_model.")}
    tc = cfg.text_config
    H, Hkv = tc.num_attention_heads, tc.num_key_value_heads
    D = tc.hidden_size // H
    B, S, _ = inputs_embeds.shape
    pos = attention_mask.long().cumsum(-1) - 1
    pos = pos.masked_fill(attention_mask == 0, 1)
    cos, sin = _rope(tc, pos)
    cos, sin = cos[:, None], sin[:, None]
    causal = torch.tril(torch.ones(S, S, dtype=torch.bool))
    vis = causal[None, None] & attention_mask.bool()[:, None, None, :]           # MASK:882
    bias = torch.zeros(B, 1, S, S).masked_fill(~vis, float("-inf"))

    def rms(x, w):                                                               # Q2M:258-263
        return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + tc.rms_norm_eps))

    x = inputs_embeds.float()
    for i in range(tc.num_hidden_layers):
        lp = f"language_model.model.layers.{i}."
        h = rms(x, sd[lp + "input_layernorm.weight"])
        q = F.linear(h, sd[lp + "self_attn.q_proj.weight"], sd[lp + "self_attn.q_proj.bias"]).view(B, S, H, D).transpose(1, 2)
        k = F.linear(h, sd[lp + "self_attn.k_proj.weight"], sd[lp + "self_attn.k_proj.bias"]).view(B, S, Hkv, D).transpose(1, 2)
        v = F.linear(h, sd[lp + "self_attn.v_proj.weight"], sd[lp + "self_attn.v_proj.bias"]).view(B, S, Hkv, D).transpose(1, 2)
        q = q * cos + _rot_half(q) * sin                                         # Q2M:143-144
        k = k * cos + _rot_half(k) * sin
        k = k.repeat_interleave(H // Hkv, dim=1)                                 # SDPA:28-37 repeat_kv
        v = v.repeat_interleave(H // Hkv, dim=1)
        s = q @ k.transpose(-1, -2) * D ** -0.5 + bias
        a = torch.nan_to_num(torch.softmax(s, dim=-1), nan=0.0) @ v              # padded query rows: all keys masked
        x = x + F.linear(a.transpose(1, 2).reshape(B, S, H * D), sd[lp + "self_attn.o_proj.weight"])
        h = rms(x, sd[lp + "post_attention_layernorm.weight"])
        g = F.silu(F.linear(h, sd[lp + "mlp.gate_proj.weight"])) * F.linear(h, sd[lp + "mlp.up_proj.weight"])  # Q2M:46-48
        x = x + F.linear(g, sd[lp + "mlp.down_proj.weight"])
    x = rms(x, sd["language_model.model.norm.weight"])
    return F.linear(x, sd["language_model.lm_head.weight"])


def ref_inputs_embeds(sd, cfg, input_ids, audio_embeds) -> torch.Tensor:
    """embed_tokens + masked_scatter of audio rows at <sound> positions ([O] AF3M:557, 563-566)."""
    sd = _sd(sd)
    e = F.embedding(input_ids, sd["language_model.model.embed_tokens.weight"].float())
    m = input_ids == cfg.audio_token_id
    if int(m.sum()) != audio_embeds.shape[0]:
        raise ValueError("number of <sound> tokens does not match number of audio feature rows")
    e[m] = audio_embeds.float()
    return e


def ref_forward_logits(sd, cfg, input_ids, attention_mask, feats, feat_mask) -> torch.Tensor:
    """AudioFlamingo3ForConditionalGeneration.forward(...).logits ([O] AF3M:479-578)."""
    ae = ref_audio_embeds(sd, cfg, feats, feat_mask)
    return ref_decoder_logits(sd, cfg, ref_inputs_embeds(sd, cfg, input_ids, ae), attention_mask)


def ref_greedy(sd, cfg, input_ids, attention_mask, feats, feat_mask, max_new_tokens: int) -> torch.Tensor:
    """Greedy decode by full re-forward each step (no cache; equals GEN:2743-2809 with do_sample=False, no EOS)."""
    ae = ref_audio_embeds(sd, cfg, feats, feat_mask)
    emb = ref_inputs_embeds(sd, cfg, input_ids, ae)
    table = _sd(sd)["language_model.model.embed_tokens.weight"].float()
    ids, mask = input_ids.clone(), attention_mask.clone()
    for _ in range(max_new_tokens):
        logits = ref_decoder_logits(sd, cfg, emb, mask)[:, -1]
        nxt = logits.argmax(-1)                                                  # GEN:2793
        ids = torch.cat([ids, nxt[:, None]], 1)
        mask = torch.cat([mask, torch.ones_like(nxt)[:, None]], 1)
        emb = torch.cat([emb, F.embedding(nxt, table)[:, None]], 1)
    return ids
