"""Prompt / source mel front-end: ``mel_spectrogram`` with the reference's signature
(modules/audio.py:45-82) on the CUDA kernels.

reflect pad -> STFT -> magnitude -> mel filterbank -> log clamp.  The STFT is not an FFT here: frame t is
rows t..t+n_fft/hop-1 of the padded audio viewed as rows of ``hop`` samples, so the whole transform is ONE
segmented fp32 GEMM (``svc_gemm``, n_fft/hop segments with row shifts) against the Hann-windowed DFT matrix
``[w cos | -w sin]``; no frame matrix is materialised.  The mel projection is a second fp32 GEMM.  fp32 operands
throughout (audio samples do not survive bf16).

The Slaney mel filterbank is librosa's published algorithm (``librosa.filters.mel`` with htk=False,
norm='slaney'), restated in ``mel_filterbank`` because librosa is not a dependency of this package.
"""
from __future__ import annotations

import numpy as np
import torch

from .ops import Ops

_cache = {}


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr, n_fft, n_mels, fmin=0.0, fmax=None):
    """(n_mels, 1 + n_fft//2) float32 Slaney-normalised triangular filters."""
    fmax = float(sr) / 2 if fmax is None else fmax
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


def _tables(n_fft, num_mels, sampling_rate, win_size, fmin, fmax, dev):
    key = (n_fft, num_mels, sampling_rate, win_size, fmin, fmax, str(dev))
    if key not in _cache:
        nb = n_fft // 2 + 1
        n = np.arange(n_fft)
        w = np.zeros(n_fft)
        off = (n_fft - win_size) // 2                     # torch.stft centres a shorter window
        w[off:off + win_size] = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(win_size) / win_size)   # hann, periodic
        ang = 2 * np.pi * np.outer(np.arange(nb), n) / n_fft
        rows = 2 * nb + (-2 * nb) % 4                     # pad the output width to a multiple of 4
        dft = np.zeros((rows, n_fft), dtype=np.float32)
        dft[:nb] = (np.cos(ang) * w).astype(np.float32)
        dft[nb:2 * nb] = (-np.sin(ang) * w).astype(np.float32)
        basis = mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax)
        _cache[key] = (torch.from_numpy(dft).to(dev), torch.from_numpy(basis).to(dev).contiguous())
    return _cache[key]


@torch.no_grad()
def mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False):
    """y (B, L) float in [-1, 1] on a CUDA device -> log-mel (B, num_mels, L // hop_size)."""
    if center:
        raise NotImplementedError("center=False only (the reference never passes True)")
    if y.device.type != "cuda":
        raise RuntimeError("seedvc_b200 mel_spectrogram runs on a CUDA (sm_100a) device only")
    if n_fft % hop_size:
        raise NotImplementedError("n_fft must be a multiple of hop_size")
    ops = Ops("fp32")
    dev = y.device
    y = y.to(torch.float32).contiguous()
    B, L = y.shape
    pad = int((n_fft - hop_size) / 2)
    n_frames = 1 + (L + 2 * pad - n_fft) // hop_size
    nseg = n_fft // hop_size
    R = n_frames + nseg - 1                                  # rows of `hop` samples the frames touch
    dft, basis = _tables(n_fft, num_mels, sampling_rate, win_size, fmin, fmax, dev)
    nb = n_fft // 2 + 1
    yp = torch.empty(B, max(R * hop_size, L + 2 * pad), dtype=torch.float32, device=dev)
    ops.reflect_pad1d(y, pad, yp)
    rows = yp[:, :R * hop_size].view(B, R, hop_size) if yp.shape[1] == R * hop_size else \
        yp.as_strided((B, R, hop_size), (yp.stride(0), hop_size, 1))
    spec = torch.empty(B, n_frames, dft.shape[0], dtype=torch.float32, device=dev)
    ops.gemm([(rows, s, dft[:, s * hop_size:(s + 1) * hop_size]) for s in range(nseg)], dft.shape[0],
             B=B, T=n_frames, out_f32=spec, f32=True)
    mag = torch.empty(B, n_frames, nb, dtype=torch.float32, device=dev)
    ops.stft_mag(spec, nb, mag, eps=1e-9)
    mel = torch.empty(B, n_frames, num_mels, dtype=torch.float32, device=dev)
    ops.gemm([(mag, 0, basis)], num_mels, B=B, T=n_frames, out_f32=mel, f32=True)
    ops.log_clamp(mel, 1e-5)
    out = torch.empty(B, num_mels, n_frames, dtype=torch.float32, device=dev)
    ops.btc_to_bct(mel, out)
    return out
