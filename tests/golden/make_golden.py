"""Generates the committed golden vectors from the REFERENCE run here (unmodified Hugging Face transformers classes
on seeded synthetic weights/inputs, CPU fp32) -- see oracle/af3_oracle.py for why that is the reference.
Run:  python tests/golden/make_golden.py     (writes tests/golden/*.npz; records library versions)
The inputs are regenerated from seeds by the tests, only reference OUTPUTS are stored (subsampled to stay small)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import af3_oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent
SECS = [10.0, 4.3, 30.0]
NEW = 12


def main():
    import transformers

    versions = np.array([f"transformers=={transformers.__version__}", f"torch=={torch.__version__}"])
    # ---- log-mel: WhisperFeatureExtractor (WFE:135-164) on 3 seeded clips; keep every 37th frame + the mask sums
    waves = O.synth_waveforms(len(SECS), SECS, seed=1)
    feats, fmask = O.hf_features(waves)
    np.savez_compressed(OUT / "logmel_golden.npz", versions=versions, secs=np.array(SECS), seed=1,
                        feats_sub=feats.numpy()[:, :, ::37], mask_sum=fmask.sum(-1).numpy())
    # ---- tiny AF3 forward + greedy generate (AF3M:479-578, GEN:2658-2812), fp32 CPU
    model = O.hf_model("tiny", seed=0, sharpen=8.0)
    cfg = model.config
    toks = [O.post_pool_len(int(n)) for n in fmask.sum(-1)]
    ids, am = O.synth_prompt(toks, cfg.text_config.vocab_size, cfg.audio_token_id, seed=2)
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=am, input_features=feats, input_features_mask=fmask)
        audio = model.get_audio_features(feats, fmask)
        gen = model.generate(input_ids=ids, attention_mask=am, input_features=feats, input_features_mask=fmask,
                             max_new_tokens=NEW, do_sample=False)
    np.savez_compressed(OUT / "tiny_forward_golden.npz", versions=versions, preset="tiny", weight_seed=0, sharpen=8.0, prompt_seed=2,
                        input_ids=ids.numpy(), attention_mask=am.numpy(), last_logits=out.logits[:, -1].numpy(),
                        logits_sub=out.logits[:, ::16, ::8].numpy(), pooler_sub=audio.pooler_output.numpy()[::5],
                        pooler_rows=audio.pooler_output.shape[0], enc_last_sub=audio.last_hidden_state.numpy()[:, ::25],
                        generated=gen.numpy(), new_tokens=NEW)
    print("wrote", [p.name for p in OUT.glob("*.npz")])


if __name__ == "__main__":
    main()
