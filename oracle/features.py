"""Oracle: log-mel featurizer.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates upstream `nemo/collections/asr/parts/preprocessing/features.py`
(`FilterbankFeatures`, `normalize_batch`) with the constructor values that
`AudioToMelSpectrogramPreprocessor` receives from TitaNet-L's model config
(`examples/speaker_tasks/recognition/conf/titanet-large.yaml`: sample_rate 16000,
window_size 0.025, window_stride 0.01, window hann, features 80, n_fft 512,
normalize per_feature, dither ignored in eval).  SURVEY.md section 8 row a5.
"""
import math

import numpy as np
import torch

from . import switches

CONSTANT = 1e-5


def _hz_to_mel_slaney(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore"):
        log_t = min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, log_t, mels)


def _mel_to_hz_slaney(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def librosa_mel(sr=16000, n_fft=512, n_mels=80, fmin=0.0, fmax=None):
    """`librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=)` with its defaults
    (htk=False -> Slaney mel scale, norm='slaney'), as called by
    FilterbankFeatures.__init__.  librosa is not installed; this is its published
    algorithm.  Returns float32 [n_mels, 1 + n_fft // 2]."""
    if fmax is None:
        fmax = sr / 2.0
    n_freq = 1 + n_fft // 2
    fftfreqs = np.linspace(0.0, sr / 2.0, n_freq)
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    mel_f = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, n_freq), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


def normalize_batch(x, seq_len, normalize_type="per_feature"):
    """features.normalize_batch, per_feature branch (masked mean, unbiased std + 1e-5)."""
    assert normalize_type == "per_feature"
    batch_size, _, max_time = x.shape
    time_steps = torch.arange(max_time, device=x.device).unsqueeze(0).expand(batch_size, max_time)
    valid_mask = time_steps < seq_len.unsqueeze(1)
    x_mean_numerator = torch.where(valid_mask.unsqueeze(1), x, 0.0).sum(axis=2)
    x_mean_denominator = valid_mask.sum(axis=1)
    x_mean = x_mean_numerator / x_mean_denominator.unsqueeze(1)
    x_std = torch.sqrt(
        torch.sum(torch.where(valid_mask.unsqueeze(1), x - x_mean.unsqueeze(2), 0.0) ** 2, axis=2)
        / (x_mean_denominator.unsqueeze(1) - 1.0)
    )
    x_std = x_std.masked_fill(x_std.isnan(), 0.0)
    x_std = x_std + CONSTANT
    return (x - x_mean.unsqueeze(2)) / x_std.unsqueeze(2), x_mean, x_std


class FilterbankFeatures(torch.nn.Module):
    """features.FilterbankFeatures (eval mode: no dither, no narrowband augmentation)."""

    def __init__(
        self,
        sample_rate=16000,
        n_window_size=400,
        n_window_stride=160,
        n_fft=512,
        preemph=0.97,
        nfilt=80,
        log_zero_guard_value=2 ** -24,
        mag_power=2.0,
        pad_to=16,
        pad_value=0.0,
        preemph_timemask=True,
    ):
        super().__init__()
        self.win_length = n_window_size
        self.hop_length = n_window_stride
        self.n_fft = n_fft
        self.preemph = preemph
        self.nfilt = nfilt
        self.log_zero_guard_value = log_zero_guard_value
        self.mag_power = mag_power
        self.pad_to = pad_to
        self.pad_value = pad_value
        self.preemph_timemask = preemph_timemask
        self.register_buffer("window", torch.hann_window(self.win_length, periodic=False))
        fb = torch.tensor(librosa_mel(sample_rate, n_fft, nfilt, 0.0, sample_rate / 2), dtype=torch.float).unsqueeze(0)
        self.register_buffer("fb", fb)

    def get_seq_len(self, seq_len):
        pad_amount = self.n_fft // 2 * 2
        seq_len = torch.floor_divide((seq_len + pad_amount - self.n_fft), self.hop_length)
        if switches.SEQ_LEN_PLUS_ONE:
            seq_len = seq_len + 1
        return seq_len.to(dtype=torch.long)

    def stft(self, x):
        return torch.stft(
            x,
            n_fft=self.n_fft,
            hop_length=self.hop_length,
            win_length=self.win_length,
            center=True,
            window=self.window.to(dtype=torch.float),
            return_complex=True,
            pad_mode=switches.STFT_PAD_MODE,
        )

    @torch.no_grad()
    def forward(self, x, seq_len):
        seq_len_time = seq_len
        seq_len_unfixed = self.get_seq_len(seq_len)
        seq_len = torch.where(seq_len == 0, torch.zeros_like(seq_len_unfixed), seq_len_unfixed)
        if self.preemph is not None:
            x = torch.cat((x[:, 0].unsqueeze(1), x[:, 1:] - self.preemph * x[:, :-1]), dim=1)
            if self.preemph_timemask:
                timemask = torch.arange(x.shape[1], device=x.device).unsqueeze(0) < seq_len_time.unsqueeze(1)
                x = x.masked_fill(~timemask, 0.0)
        x = self.stft(x)
        x = torch.view_as_real(x)
        x = torch.sqrt(x.pow(2).sum(-1))
        if self.mag_power != 1.0:
            x = x.pow(self.mag_power)
        x = torch.matmul(self.fb.to(x.dtype), x)
        x = torch.log(x + self.log_zero_guard_value)
        x, _, _ = normalize_batch(x, seq_len, normalize_type="per_feature")
        max_len = x.size(-1)
        mask = torch.arange(max_len, device=x.device)
        mask = mask.repeat(x.size(0), 1) >= seq_len.unsqueeze(1)
        x = x.masked_fill(mask.unsqueeze(1), self.pad_value)
        if self.pad_to > 0:
            pad_amt = x.size(-1) % self.pad_to
            if pad_amt != 0:
                x = torch.nn.functional.pad(x, (0, self.pad_to - pad_amt), value=self.pad_value)
        return x, seq_len


def num_frames(num_samples, hop=160):
    return int(math.floor(num_samples / hop)) + 1
