"""Generates tests/golden/tacotron_stft.npz by running the UNMODIFIED reference TacotronSTFT
(/root/reference/python/fastpitch1_1/common/layers.py:102-138 + common/stft.py:51-114: the mel extractor of the
FastPitch dataset) on a seeded waveform. Build container only; the fixture is committed.

    python tests/golden/make_golden_tacotron_stft.py

librosa is not installed: `librosa.filters.mel` is torchaudio.functional.melscale_fbanks (Slaney scale and norm), an
implementation independent of oracle/hifigan.py::mel_filterbank.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import _ref_import  # noqa: E402
from make_golden_hifigan import torchaudio_mel  # noqa: E402  (installs the shims as a side effect)


def main():
    _ref_import.install()
    sys.modules["librosa.filters"].mel = torchaudio_mel
    sys.modules["librosa"].filters.mel = torchaudio_mel
    import python.fastpitch1_1.common.layers as layers

    layers.librosa_mel_fn = lambda sr, n_fft, n_mels, fmin, fmax: torchaudio_mel(sr, n_fft, n_mels, fmin, fmax)
    stft = layers.TacotronSTFT(1024, 256, 1024, 80, 22050, 0.0, 8000.0)
    g = torch.Generator().manual_seed(11)
    audio = 0.9 * torch.tanh(torch.randn(2, 5120, generator=g) * 0.4)
    mel = stft.mel_spectrogram(audio)
    out = {"audio": audio.numpy(), "mel": mel.numpy()}
    np.savez_compressed(os.path.join(HERE, "tacotron_stft.npz"), **out)
    print("wrote", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
