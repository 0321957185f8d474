"""Pins the oracle: (1) the torch-fp32 restatement (oracle/af3_oracle.py ref_*) against the reference executed live
(unmodified HF transformers classes) and (2) both against the committed golden vectors (tests/golden/*.npz, generated
from the reference by tests/golden/make_golden.py).  CPU only."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import af3_oracle as O

G = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def tiny():
    return O.hf_model("tiny", seed=0, sharpen=8.0)


@pytest.fixture(scope="module")
def golden_inputs():
    g = np.load(G / "logmel_golden.npz")
    waves = O.synth_waveforms(len(g["secs"]), list(g["secs"]), seed=int(g["seed"]))
    feats, fmask = O.hf_features(waves)
    return g, waves, feats, fmask


def test_golden_versions_recorded():
    g = np.load(G / "tiny_forward_golden.npz")
    assert any(str(v).startswith("transformers==") for v in g["versions"])


def test_logmel_restatement_vs_reference_and_golden(golden_inputs):
    g, waves, feats, fmask = golden_inputs
    # reference live == golden (deterministic CPU path)
    np.testing.assert_allclose(feats.numpy()[:, :, ::37], g["feats_sub"], atol=2e-6)
    assert fmask.sum(-1).tolist() == g["mask_sum"].tolist()
    # fp64 restatement of WFE:135-164 vs the reference's fp32 torch path: the reference quotes 1e-5 between its paths
    padded = np.stack([np.pad(w, (0, 480000 - len(w))) for w in waves])
    ref = O.ref_logmel(padded)
    assert np.abs(ref - feats.numpy()).max() < 2e-5
    assert O.ref_frame_mask([len(w) for w in waves]).sum(-1).tolist() == fmask.sum(-1).tolist()


def test_mel_filter_ranges_are_contiguous():
    """The CUDA kernel applies each mel filter over [first, last] non-zero bin: filters must have no interior zeros."""
    f = O.ref_mel_filters()
    assert f.shape == (201, 128)
    for m in range(128):
        nz = np.nonzero(f[:, m])[0]
        assert len(nz) > 0 and np.all(f[nz[0]:nz[-1] + 1, m] > 0)


def test_restatement_forward_and_greedy_vs_reference_and_golden(tiny, golden_inputs):
    g, waves, feats, fmask = golden_inputs
    t = np.load(G / "tiny_forward_golden.npz")
    ids, am = torch.from_numpy(t["input_ids"]), torch.from_numpy(t["attention_mask"])
    cfg = tiny.config
    toks = [O.post_pool_len(int(n)) for n in fmask.sum(-1)]
    ids2, am2 = O.synth_prompt(toks, cfg.text_config.vocab_size, cfg.audio_token_id, seed=int(t["prompt_seed"]))
    assert torch.equal(ids, ids2) and torch.equal(am, am2)
    with torch.no_grad():
        live = tiny(input_ids=ids, attention_mask=am, input_features=feats, input_features_mask=fmask).logits
        rest = O.ref_forward_logits(tiny, cfg, ids, am, feats, fmask)
        pooled = O.ref_audio_embeds(tiny, cfg, feats, fmask)
    # reference live == golden
    np.testing.assert_allclose(live[:, -1].numpy(), t["last_logits"], atol=2e-4)
    np.testing.assert_allclose(live[:, ::16, ::8].numpy(), t["logits_sub"], atol=2e-4)
    # restatement == reference (valid positions; padded rows are unconstrained)
    v = am.bool()
    assert (rest - live)[v].abs().max().item() < 1e-3
    assert pooled.shape[0] == int(t["pooler_rows"])
    np.testing.assert_allclose(pooled.numpy()[::5], t["pooler_sub"], atol=1e-4)
    # greedy ids: restatement (no cache, full re-forward) == reference generate == golden
    new = int(t["new_tokens"])
    gen = O.ref_greedy(tiny, cfg, ids, am, feats, fmask, new)
    assert np.array_equal(gen.numpy(), t["generated"])


def test_length_arithmetic_matches_reference_formulae():
    """frames -> tokens: processor (AF3P:90-97) vs model (AF3M:375-377, 470-473) vs our host logic."""
    from audio_flamingo_b200.processing import audio_token_length, split_windows, tokens_per_sample

    for n_frames in [1, 2, 3, 4, 5, 430, 999, 1000, 2999, 3000]:
        conv = (n_frames - 1) // 2 + 1
        assert audio_token_length(n_frames) == (conv - 2) // 2 + 1 == O.post_pool_len(n_frames)
    # multi-window sample: sum-before-floor (processor) equals per-window sum (model) when all but the last are full
    clips = [np.zeros(480000 * 2 + 123456, np.float32), np.zeros(16000, np.float32)]
    chunks, per = split_windows(clips)
    assert per == [3, 1] and [len(c) for c in chunks] == [480000, 480000, 123456, 16000]
    frames = [(len(c) + 159) // 160 for c in chunks]
    assert tokens_per_sample(frames, per) == [sum(O.post_pool_len(f) for f in frames[:3]), O.post_pool_len(frames[3])]


def test_af2_gated_xattn_restatement_matches_the_executable_analogue():
    """AF2 itself is unpinned (no code in the container); the gated xattn-dense OPERATOR is pinned against the
    executable analogue, transformers' IdeficsGatedCrossAttentionLayer, for vector and scalar gates, ragged media lengths and a
    token that attends to no media."""
    from oracle import af2_oracle as A

    for alpha_type in ("vector", "float"):
        layer = A.hf_gated_layer(seed=1, alpha_type=alpha_type)
        B, T, Tm = 3, 11, 9
        g = torch.Generator().manual_seed(2)
        h = torch.randn(B, T, 512, generator=g)
        m = torch.randn(B, Tm, 384, generator=g)
        media_len = [9, 4, 1]
        gate = torch.ones(B, T)
        gate[2, 3] = 0
        with torch.no_grad():
            ref = layer(h, image_hidden_states=m, image_attention_mask=A.key_padding_mask(B, T, media_len, Tm), cross_attention_gate=gate)
        ours = A.ref_gated_layer(layer.state_dict(), h, m, media_len, gate)
        assert (ours - ref).abs().max().item() < 2e-5, alpha_type
