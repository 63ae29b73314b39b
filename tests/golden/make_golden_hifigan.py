"""Generates tests/golden/hifigan_small.npz and hifigan_keys.json by running the UNMODIFIED reference HiFi-GAN modules
(/root/reference/python/hifigan/{models,meldataset}.py) on seeded inputs. Build container only (the GPU box has no
/root/reference); the fixtures it writes are committed.

    python tests/golden/make_golden_hifigan.py

librosa is not installed: `librosa.filters.mel` is provided by torchaudio.functional.melscale_fbanks (Slaney scale and
norm), an implementation INDEPENDENT of oracle/hifigan.py::mel_filterbank, so the mel fixture also pins that restatement.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import _ref_import  # noqa: E402


def torchaudio_mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **_):
    import torchaudio

    if fmax is None:
        fmax = sr / 2.0
    return torchaudio.functional.melscale_fbanks(1 + n_fft // 2, float(fmin), float(fmax), n_mels, sr, norm="slaney",
                                                 mel_scale="slaney").T.contiguous().numpy().astype(np.float32)


_ref_import.install()
sys.modules["librosa.filters"].mel = torchaudio_mel
sys.modules["librosa"].filters.mel = torchaudio_mel

from oracle import hifigan as ohg  # noqa: E402
from python.hifigan import meldataset as ref_mel  # noqa: E402
from python.hifigan.models import AttrDict, Generator  # noqa: E402


def sample_idx(key, numel, n=16):
    h = int(hashlib.sha256(key.encode()).hexdigest()[:8], 16)
    return np.random.RandomState(h).randint(0, numel, size=n)


def main():
    torch.manual_seed(0)
    h = AttrDict(json.load(open(os.path.join(_ref_import.REFERENCE_ROOT, "python/hifigan/config_v1.json"))))
    h.USE_EMB_CONDITIONING = False
    gen = Generator(h)
    ref_sd = gen.state_dict()
    keys = [[k, list(v.shape)] for k, v in ref_sd.items()]
    json.dump(keys, open(os.path.join(HERE, "hifigan_keys.json"), "w"))

    out = {}
    sd = ohg.make_generator_state(1234, scale=1.0)
    gen.load_state_dict(sd)
    g = torch.Generator().manual_seed(5)
    mel = torch.randn(2, 80, 6, generator=g)
    out["gen/mel"] = mel.numpy()
    params = dict(gen.named_parameters())
    y = gen(mel)
    out["gen/y"] = y.detach().numpy()
    w = torch.randn(y.shape, generator=g)
    out["gen/w"] = w.numpy()
    (y * w).sum().backward()
    for k, p in params.items():
        gr = p.grad.detach().double()
        out[f"gen/grad/{k}/norm"] = np.float64(gr.norm())
        out[f"gen/grad/{k}/samples"] = gr.reshape(-1)[torch.from_numpy(sample_idx(k, gr.numel()))].numpy()

    # mel spectrogram, both filterbanks the trainer uses (config_v1.json: fmax 8000 for inputs, null for the loss)
    audio = 0.95 * torch.tanh(torch.randn(2, 4096, generator=g) * 0.3)
    out["mel/audio"] = audio.numpy()
    for tag, fmax in (("8000", 8000), ("none", None)):
        ref_mel.mel_basis.clear()
        m = ref_mel.mel_spectrogram(audio, h.n_fft, h.num_mels, h.sampling_rate, h.hop_size, h.win_size, h.fmin, fmax)
        out[f"mel/{tag}"] = m.numpy()
    out["mel/basis8000"] = torchaudio_mel(22050, 1024, 80, 0, 8000)
    np.savez_compressed(os.path.join(HERE, "hifigan_small.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()


def disc_goldens():
    """Scores and feature-map norms of the reference MPD / MSD (training mode: one spectral-norm power iteration per
    forward) on a fixed pair of waveforms, with weights from oracle.hifigan.make_disc_state."""
    from python.hifigan.models import MultiPeriodDiscriminator, MultiScaleDiscriminator

    out = {}
    g = torch.Generator().manual_seed(9)
    y = 0.9 * torch.tanh(torch.randn(2, 1, 2048, generator=g))
    yh = 0.9 * torch.tanh(torch.randn(2, 1, 2048, generator=g))
    out["disc/y"], out["disc/y_hat"] = y.numpy(), yh.numpy()
    for name, cls, spec, seed in (("mpd", MultiPeriodDiscriminator, ohg.mpd_spec(), 21), ("msd", MultiScaleDiscriminator, ohg.msd_spec(), 22)):
        torch.manual_seed(0)
        m = cls()
        keys = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        json.dump(keys, open(os.path.join(HERE, f"hifigan_{name}_keys.json"), "w"))
        m.load_state_dict(ohg.make_disc_state(spec, seed))
        m.train()
        rs, gs, frs, fgs = m(y, yh)
        for i, (r, g_) in enumerate(zip(rs, gs)):
            out[f"disc/{name}/r{i}"], out[f"disc/{name}/g{i}"] = r.detach().numpy(), g_.detach().numpy()
            out[f"disc/{name}/fnorm_r{i}"] = np.array([float(f.detach().double().norm()) for f in frs[i]])
            out[f"disc/{name}/fnorm_g{i}"] = np.array([float(f.detach().double().norm()) for f in fgs[i]])
    return out


if __name__ == "__main__":
    extra = disc_goldens()
    path = os.path.join(HERE, "hifigan_small.npz")
    base = dict(np.load(path))
    base.update(extra)
    np.savez_compressed(path, **base)
    print("added", len(extra), "discriminator arrays")
