"""TEST INFRASTRUCTURE ONLY - golden log-mels from the REAL reference ``modules.audio.mel_spectrogram``
(torch.stft path).  librosa is not installed here, so the reference's ``librosa_mel_fn`` import is served
by ``seedvc_oracle.slaney_mel_filterbank`` (a restatement of librosa's algorithm, see its docstring):
reflect padding, STFT, magnitude, projection and log-clamp are pinned to the reference's own code, the
filterbank values are pinned only to that restatement.

    python oracle/gen_golden_mel.py      (needs /root/reference; writes tests/golden/mel_kat.npz)
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

# name -> mel_fn_args (inference.py:315-324) and (B, L)
CASES = {
    "mel_22k": (dict(n_fft=1024, win_size=1024, hop_size=256, num_mels=80, sampling_rate=22050, fmin=0, fmax=None,
                     center=False), 2, 256 * 37),
    "mel_22k_odd": (dict(n_fft=1024, win_size=1024, hop_size=256, num_mels=80, sampling_rate=22050, fmin=0,
                         fmax=None, center=False), 1, 256 * 20 + 131),
    "mel_44k": (dict(n_fft=2048, win_size=2048, hop_size=512, num_mels=128, sampling_rate=44100, fmin=0, fmax=None,
                     center=False), 1, 512 * 23),
    "mel_22k_fmax8k": (dict(n_fft=1024, win_size=1024, hop_size=256, num_mels=80, sampling_rate=22050, fmin=0,
                            fmax=8000, center=False), 1, 256 * 16),
}


def audio(name, B, L):
    g = torch.Generator().manual_seed(11 + list(CASES).index(name))
    t = torch.arange(L, dtype=torch.float32) / 22050.0
    y = 0.3 * torch.sin(2 * torch.pi * 220.0 * t)[None, :] + 0.2 * torch.sin(2 * torch.pi * 3100.0 * t)[None, :]
    y = y + 0.05 * torch.randn(B, L, generator=g)
    return y.clamp(-1, 1)


def main():
    ns = ref_import.load()
    out, meta = {}, {}
    for name, (kw, B, L) in CASES.items():
        y = audio(name, B, L)
        with torch.no_grad():
            m = ns.mel_spectrogram(y, **kw)
        out[name] = m.numpy()
        meta[name] = dict(kw=kw, B=B, L=L)
        print(name, tuple(m.shape), "mean =", float(m.mean()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mel_kat.npz"), meta=json.dumps(meta), **out)


if __name__ == "__main__":
    main()
