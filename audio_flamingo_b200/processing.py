"""Host-side front end of the AF3 path: window split / length arithmetic of the processor and the GPU log-mel
feature extractor.  Mirrors the reference surface:
    AudioFlamingo3Processor.__call__ window logic      [O] AF3P:159-179, 90-101
    WhisperFeatureExtractor.__call__ (feature_size=128) [O] WFE:189-345 (pad WFE:296, fbank WFE:135-164, mask WFE:328-337)
The tokenizer / chat template are out of scope (no vocabulary files offline): prompts are given as token ids and
`expand_audio_tokens` performs the `<sound>` expansion on id lists.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._lib import AF3Error

SAMPLING_RATE = 16000
CHUNK_LENGTH = 30
WINDOW_SAMPLES = SAMPLING_RATE * CHUNK_LENGTH  # 480000
HOP = 160
MAX_AUDIO_LEN_S = 600


def slaney_mel_filterbank(n_mels: int = 128, n_freq_bins: int = 201, f_max: float = 8000.0) -> np.ndarray:
    """[n_freq_bins, n_mels] float64 triangular mel filterbank of the Whisper front end (a2 of SURVEY.md 8-a): Slaney's
    auditory-toolbox mel scale (linear 200/3 Hz per mel below 1 kHz, 27 mels per factor 6.4 above), triangles drawn in Hz over
    the rFFT bin centres 0..f_max, each scaled by 2 / (its bandwidth) ("slaney" area normalisation).
    Own construction (one filter at a time); tests/test_host_logic_cpu.py asserts it is BIT-identical to the table the
    reference builds with transformers.audio_utils.mel_filter_bank(201, 128, 0, 8000, 16000, "slaney", "slaney")
    ([O] WFE:95-103 -> AU:453-544), so the product no longer imports the reference package for a constant."""
    knee_hz, knee_mel = 1000.0, 15.0
    mels_per_ln = 27.0 / np.log(6.4)
    top_mel = knee_mel + np.log(f_max / knee_hz) * mels_per_ln if f_max >= knee_hz else 3.0 * f_max / 200.0
    mel_pts = np.linspace(0.0, top_mel, n_mels + 2)
    edges_hz = 200.0 * mel_pts / 3.0
    upper = mel_pts >= knee_mel
    edges_hz[upper] = knee_hz * np.exp((np.log(6.4) / 27.0) * (mel_pts[upper] - knee_mel))
    bin_hz = np.linspace(0, f_max, n_freq_bins)
    bank = np.zeros((n_freq_bins, n_mels), dtype=np.float64)
    for m in range(n_mels):
        lo, mid, hi = edges_hz[m], edges_hz[m + 1], edges_hz[m + 2]
        rising = (bin_hz - lo) / (mid - lo)
        falling = (hi - bin_hz) / (hi - mid)
        bank[:, m] = np.maximum(0.0, np.minimum(rising, falling)) * (2.0 / (hi - lo))
    return bank


def split_windows(audio: list[np.ndarray], max_audio_len: int = MAX_AUDIO_LEN_S):
    """[O] AF3P:159-179.  Returns (flat_chunks, per_sample_windows)."""
    max_windows = int(max_audio_len // CHUNK_LENGTH)
    per_sample, flat = [], []
    for a in audio:
        n = int(a.shape[0])
        n_win = max(1, (n + WINDOW_SAMPLES - 1) // WINDOW_SAMPLES)
        n_win = min(n_win, max_windows)
        per_sample.append(n_win)
        cap = min(n, n_win * WINDOW_SAMPLES)
        for i in range(n_win):
            flat.append(a[i * WINDOW_SAMPLES: min((i + 1) * WINDOW_SAMPLES, cap)])
    return flat, per_sample


def audio_token_length(n_frames):
    """frames -> conv2 length -> pooled tokens ([O] AF3P:90-93, AF3M:375-377).  Works on ints, numpy and torch."""
    conv = (n_frames - 1) // 2 + 1
    return (conv - 2) // 2 + 1


def tokens_per_sample(frames_per_window: list[int], per_sample_windows: list[int]) -> list[int]:
    """[O] AF3P:95-97: frame counts are summed over a sample's windows *before* the two floor divisions."""
    out, i = [], 0
    for n_win in per_sample_windows:
        out.append(int(audio_token_length(sum(frames_per_window[i:i + n_win]))))
        i += n_win
    return out


def expand_audio_tokens(ids: list[int], audio_token_id: int, n_tokens: int) -> list[int]:
    """Replace every <sound> id by n_tokens copies ([O] AF3P:98-100 does this on the text with a regex)."""
    out = []
    for t in ids:
        out.extend([audio_token_id] * n_tokens if t == audio_token_id else [t])
    return out


def expand_audio_spans(ids: list[int], audio_token_id: int, counts: list[int]) -> list[int]:
    """Multi-audio prompts (AF3-Chat): the i-th <sound> placeholder becomes counts[i] audio tokens.
    The reference processor enforces one audio per text and one count for every placeholder ([O] AF3P:155-156, 98-100);
    the model side (masked_scatter, AF3M:563-566) already accepts any layout, so only this expansion is new."""
    out, k = [], 0
    for t in ids:
        if t == audio_token_id:
            if k >= len(counts):
                raise ValueError("more <sound> placeholders than audio clips")
            out.extend([audio_token_id] * counts[k])
            k += 1
        else:
            out.append(t)
    if k != len(counts):
        raise ValueError("fewer <sound> placeholders than audio clips")
    return out


def left_pad(rows: list[list[int]], pad_id: int = 0):
    """Tokenizer padding_side='left' ([O] AF3P:44-47) -> (input_ids, attention_mask) int64 tensors."""
    S = max(len(r) for r in rows)
    ids = torch.full((len(rows), S), pad_id, dtype=torch.int64)
    mask = torch.zeros((len(rows), S), dtype=torch.int64)
    for i, r in enumerate(rows):
        ids[i, S - len(r):] = torch.as_tensor(r, dtype=torch.int64)
        mask[i, S - len(r):] = 1
    return ids, mask


class AF3FeatureExtractor:
    """Log-mel features on the GPU with the call surface of WhisperFeatureExtractor(feature_size=128).

    __call__(raw_speech, sampling_rate=16000) -> {"input_features": fp32 [n,128,3000] (device),
                                                   "attention_mask": int32 [n,3000] (device)}
    Every clip is zero padded (or truncated) to 30 s as WFE:296 (padding='max_length', truncation=True) does; the
    frame mask is the sample mask subsampled by the hop (WFE:328-337).
    """

    def __init__(self, device="cuda", feature_size=128):
        if feature_size != 128:
            raise AF3Error("AF3 uses 128 mel bins")
        self.mel_filters = slaney_mel_filterbank(128, 201, SAMPLING_RATE / 2)
        self.device = torch.device(device)
        self.tables = ops.LogMelTables(self.mel_filters, self.device)
        self.n_samples = WINDOW_SAMPLES
        self.nb_max_frames = WINDOW_SAMPLES // HOP
        self.sampling_rate = SAMPLING_RATE
        self._pinned = None
        self._pinned_busy = None  # CUDA event recorded after the last async H2D copy out of the staging buffer

    def _stage(self, chunks: list[np.ndarray]):
        """Zero-padded [n, 480000] fp32 in pinned host memory + per-clip sample counts."""
        n = len(chunks)
        if self._pinned_busy is not None:
            # the previous call's non-blocking copy may still be reading the staging buffer: rewriting it now would corrupt
            # that batch's waveforms.  Waits only for that one copy, not for the stream.
            self._pinned_busy.synchronize()
            self._pinned_busy = None
        if self._pinned is None or self._pinned.shape[0] < n:
            self._pinned = torch.empty((n, self.n_samples), dtype=torch.float32).pin_memory()
        buf = self._pinned[:n]
        lens = []
        for i, c in enumerate(chunks):
            c = np.asarray(c, dtype=np.float32).reshape(-1)[: self.n_samples]
            lens.append(len(c))
            buf[i, : len(c)] = torch.from_numpy(c)
            buf[i, len(c):] = 0.0
        return buf, lens

    def __call__(self, raw_speech, sampling_rate=SAMPLING_RATE, **kwargs):
        if sampling_rate != self.sampling_rate:
            raise ValueError(f"sampling_rate must be {self.sampling_rate}")  # WFE:244-250
        if isinstance(raw_speech, np.ndarray) and raw_speech.ndim == 1:
            raw_speech = [raw_speech]
        host, lens = self._stage(list(raw_speech))
        with torch.cuda.device(self.device):
            wave = host.to(self.device, non_blocking=True)
            self._pinned_busy = torch.cuda.Event()
            self._pinned_busy.record()
        return self.from_device_waveform(wave, lens)

    def from_device_waveform(self, wave: torch.Tensor, n_valid_samples):
        """wave fp32 [n, 480000] already on the device; n_valid_samples: per-clip sample counts (list or tensor)."""
        feats = ops.logmel(wave, self.tables)
        lens = torch.as_tensor(n_valid_samples, device=self.device, dtype=torch.int64)
        n_frames = (lens + HOP - 1) // HOP  # = sum(mask[::160]) for a prefix mask of `lens` ones
        mask = (torch.arange(self.nb_max_frames, device=self.device)[None, :] < n_frames[:, None]).to(torch.int32)
        return {"input_features": feats, "attention_mask": mask, "input_features_mask": mask}
