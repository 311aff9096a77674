"""Model-independent constant tables the reference builds at module construction time.

They are computed here with the same torch CPU arithmetic the reference uses and handed to the
CUDA library as "__const__.*" tensors, so that sin/cos/log rounding cannot differ between the two
implementations.
"""
from __future__ import annotations

import math

import torch


def rel_pos_table(T: int, d_model: int) -> torch.Tensor:
    """RelPositionalEncoding.extend_pe/forward (fairseq/fairseq/modules/positional_encoding.py:82-129):
    [2T-1, d_model]; row k holds relative position (T-1-k)."""
    position = torch.arange(0, T, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pos = torch.zeros(T, d_model)
    neg = torch.zeros(T, d_model)
    pos[:, 0::2] = torch.sin(position * div)
    pos[:, 1::2] = torch.cos(position * div)
    neg[:, 0::2] = torch.sin(-1 * position * div)
    neg[:, 1::2] = torch.cos(-1 * position * div)
    return torch.cat([torch.flip(pos, [0]), neg[1:]], dim=0).contiguous()


def sinusoidal_table(n: int, dim: int, padding_idx: int = 1) -> torch.Tensor:
    """SinusoidalPositionalEmbedding.get_embedding (fairseq/modules/sinusoidal_positional_embedding.py:43-63)."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float) * -e)
    e = torch.arange(n, dtype=torch.float).unsqueeze(1) * e.unsqueeze(0)
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1).view(n, -1)
    e[padding_idx, :] = 0
    return e.contiguous()


def povey_window(n: int = 400) -> torch.Tensor:
    """torchaudio/compliance/kaldi.py _feature_window_function(POVEY)."""
    return torch.hann_window(n, periodic=False, dtype=torch.float32).pow(0.85)


def mel_bank(num_bins: int = 80, padded: int = 512, sample_freq: float = 16000.0, low_freq: float = 20.0,
             high_freq: float = 0.0) -> torch.Tensor:
    """torchaudio/compliance/kaldi.py get_mel_banks (vtln_warp = 1) + the zero Nyquist column: [num_bins, padded/2+1]."""
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    fft_bin_width = sample_freq / padded

    def mel(f):
        return 1127.0 * math.log(1.0 + f / 700.0)

    mel_lo, mel_hi = mel(low_freq), mel(high_freq)
    delta = (mel_hi - mel_lo) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left = mel_lo + b * delta
    center = mel_lo + (b + 1.0) * delta
    right = mel_lo + (b + 2.0) * delta
    melf = (1127.0 * (1.0 + fft_bin_width * torch.arange(padded // 2) / 700.0).log()).unsqueeze(0)
    up = (melf - left) / (center - left)
    down = (right - melf) / (right - center)
    bins = torch.max(torch.zeros(1), torch.min(up, down))
    return torch.nn.functional.pad(bins, (0, 1)).contiguous()


def resample_kernel_3to1(lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """Windowed-sinc decimation filter 48 kHz -> 16 kHz (orig / gcd = 3, new / gcd = 1): the kernel of
    torchaudio.functional.resample's default method (sinc_interp_hann; torchaudio/functional/functional.py
    _get_sinc_resample_kernel), restated.  Returns (kernel [2 * width + 3], width): out[i] = sum_k kernel[k] * x[3 i + k - width].

    The reference resamples with sox's `rate` effect (fairseq/fairseq/data/audio/audio_utils.py:53-62 through
    torchaudio.sox_effects, which this image does not have): a different (longer) filter; see DESIGN.md."""
    orig, new = 3, 1
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float64)[:, None, None] / new + idx
    t = t * base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t)
    kernels = kernels * window * scale
    return kernels.to(torch.float32).view(-1).contiguous(), width
