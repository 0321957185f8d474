"""The CUDA path against the COMMITTED golden vectors (reference outputs, fp32 CPU; tests/golden/make_golden.py).
Nothing here imports transformers' model classes: weights are rebuilt from the seed through the oracle helper only to
obtain the same synthetic state_dict the goldens were generated with."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


def test_logmel_kernel_vs_golden():
    from audio_flamingo_b200 import AF3FeatureExtractor
    from oracle import af3_oracle as O

    g = np.load(G / "logmel_golden.npz")
    waves = O.synth_waveforms(len(g["secs"]), list(g["secs"]), seed=int(g["seed"]))
    out = AF3FeatureExtractor("cuda")(waves, sampling_rate=16000)
    got = out["input_features"].cpu().numpy()[:, :, ::37]
    # fp32 DFT vs the reference's fp32 FFT: 1e-4 abs on the (x+4)/4 scale (observed ~2e-5)
    assert np.abs(got - g["feats_sub"]).max() < 1e-4
    assert out["attention_mask"].sum(-1).cpu().tolist() == g["mask_sum"].tolist()


def test_tiny_model_vs_golden():
    from audio_flamingo_b200 import AF3FeatureExtractor, AudioFlamingo3ForConditionalGeneration
    from oracle import af3_oracle as O

    t = np.load(G / "tiny_forward_golden.npz")
    lg = np.load(G / "logmel_golden.npz")
    ref = O.hf_model("tiny", seed=int(t["weight_seed"]), sharpen=float(t["sharpen"]))
    ours = AudioFlamingo3ForConditionalGeneration.from_reference(ref, device="cuda")
    waves = O.synth_waveforms(len(lg["secs"]), list(lg["secs"]), seed=int(lg["seed"]))
    fo = AF3FeatureExtractor("cuda")(waves, sampling_rate=16000)
    ids, am = torch.from_numpy(t["input_ids"]).cuda(), torch.from_numpy(t["attention_mask"]).cuda()
    out = ours(input_ids=ids, attention_mask=am, input_features=fo["input_features"], input_features_mask=fo["input_features_mask"])
    last = out.logits[:, -1].float().cpu().numpy()
    std = t["last_logits"].std()
    # bf16 path vs fp32 golden: same bound as tests/test_model_gpu.py (the reference's own bf16 run shows ~0.035 std)
    assert np.abs(last - t["last_logits"]).max() < 0.06 * std
    audio = ours.get_audio_features(fo["input_features"], fo["input_features_mask"])
    assert audio.pooler_output.shape[0] == int(t["pooler_rows"])
    assert np.abs(audio.pooler_output.float().cpu().numpy()[::5] - t["pooler_sub"]).max() < 0.03 * np.abs(t["pooler_sub"]).max()
    gen = ours.generate(input_ids=ids, attention_mask=am, input_features=fo["input_features"],
                        input_features_mask=fo["input_features_mask"], max_new_tokens=int(t["new_tokens"]))
    assert np.array_equal(gen.cpu().numpy(), t["generated"]), "greedy ids differ from the reference's"
