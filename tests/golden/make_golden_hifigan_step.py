"""Generates tests/golden/hifigan_step.npz: TWO consecutive training steps of the UNMODIFIED reference HiFi-GAN modules
(/root/reference/python/hifigan/models.py Generator / MultiPeriodDiscriminator / MultiScaleDiscriminator, meldataset.py
mel_spectrogram, torch.optim.AdamW as built at xva_train.py:298-300) driven through the exact sequence of
HiFiTrainer.iteration, hifigan/xva_train.py:467-515 (G forward, loss mel, D step with optim_d.step, G step on the updated
discriminators with optim_g.step). Build container only; the fixture is committed.

    python tests/golden/make_golden_hifigan_step.py

Recorded per step: every loss term, the norm + 16 sampled entries of every parameter gradient of the D step (MPD + MSD)
and of the G step (generator), the same summary of every parameter and spectral-norm buffer (u, v) after the step.
tests/test_oracle_golden.py::test_hifigan_train_step_matches_reference pins oracle.hifigan.train_step to it, which closes
the chain CUDA step -> oracle step -> reference step for SURVEY rows a18 / a21.
"""
import hashlib
import itertools
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

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

from oracle import hifigan as ohg  # noqa: E402  (seeded weights / inputs shared with the test)
from python.hifigan import meldataset as ref_mel  # noqa: E402
from python.hifigan.models import (AttrDict, Generator, MultiPeriodDiscriminator, MultiScaleDiscriminator,  # noqa: E402
                                    discriminator_loss, feature_loss, generator_loss)

B, FRAMES, STEPS = 2, 16, 2


def sample_idx(key, numel, n=16):
    h = int(hashlib.sha256(key.encode()).hexdigest()[:8], 16)
    return np.random.RandomState(h).randint(0, numel, size=n)


def summarize(prefix, named, out):
    for k, v in named.items():
        if v is None:
            continue
        v = v.detach().reshape(-1).double()
        out[f"{prefix}/{k}/norm"] = np.float64(v.norm().item())
        out[f"{prefix}/{k}/samples"] = v[torch.from_numpy(sample_idx(k, v.numel()))].numpy()


def main():
    h = AttrDict(json.load(open(os.path.join(_ref_import.REFERENCE_ROOT, "python/hifigan/config_v1.json"))))
    h.USE_EMB_CONDITIONING = False
    torch.manual_seed(0)
    gen, mpd, msd = Generator(h), MultiPeriodDiscriminator(), MultiScaleDiscriminator()
    gen.load_state_dict(ohg.make_generator_state(1234))
    mpd.load_state_dict(ohg.make_disc_state(ohg.mpd_spec(), 21))
    msd.load_state_dict(ohg.make_disc_state(ohg.msd_spec(), 22))
    for m in (gen, mpd, msd):
        m.train()
    optim_g = torch.optim.AdamW(gen.parameters(), h.learning_rate, betas=[h.adam_b1, h.adam_b2])
    optim_d = torch.optim.AdamW(itertools.chain(msd.parameters(), mpd.parameters()), h.learning_rate,
                                betas=[h.adam_b1, h.adam_b2])
    out = {"meta/B": np.int64(B), "meta/frames": np.int64(FRAMES), "meta/steps": np.int64(STEPS)}
    for s in range(STEPS):
        x, y, y_mel = ohg.synthetic_batch(B, FRAMES, seed=100 + s)
        out[f"s{s}/in/x"], out[f"s{s}/in/y"], out[f"s{s}/in/y_mel"] = x.numpy(), y.numpy(), y_mel.numpy()
        # ---- hifigan/xva_train.py:467-515
        gen.zero_grad(set_to_none=True)
        mpd.zero_grad(set_to_none=True)
        msd.zero_grad(set_to_none=True)
        yy = y.unsqueeze(1)
        y_g_hat = gen(x)
        ref_mel.mel_basis.clear()
        y_g_hat_mel = ref_mel.mel_spectrogram(y_g_hat.squeeze(1), h.n_fft, h.num_mels, h.sampling_rate, h.hop_size,
                                              h.win_size, h.fmin, h.fmax_for_loss)
        optim_d.zero_grad()
        y_df_hat_r, y_df_hat_g, _, _ = mpd(yy, y_g_hat.detach())
        loss_disc_f, _, _ = discriminator_loss(y_df_hat_r, y_df_hat_g)
        y_ds_hat_r, y_ds_hat_g, _, _ = msd(yy, y_g_hat.detach())
        loss_disc_s, _, _ = discriminator_loss(y_ds_hat_r, y_ds_hat_g)
        loss_disc_all = loss_disc_s + loss_disc_f
        loss_disc_all.backward()
        summarize(f"s{s}/dgrad/mpd", {k: p.grad for k, p in mpd.named_parameters()}, out)
        summarize(f"s{s}/dgrad/msd", {k: p.grad for k, p in msd.named_parameters()}, out)
        optim_d.step()
        optim_g.zero_grad()
        loss_mel = F.l1_loss(y_mel, y_g_hat_mel) * 45
        y_df_hat_r, y_df_hat_g, fmap_f_r, fmap_f_g = mpd(yy, y_g_hat)
        y_ds_hat_r, y_ds_hat_g, fmap_s_r, fmap_s_g = msd(yy, y_g_hat)
        loss_fm_f = feature_loss(fmap_f_r, fmap_f_g)
        loss_fm_s = feature_loss(fmap_s_r, fmap_s_g)
        loss_gen_f, _ = generator_loss(y_df_hat_g)
        loss_gen_s, _ = generator_loss(y_ds_hat_g)
        loss_gen_all = loss_gen_s + loss_gen_f + loss_fm_s + loss_fm_f + loss_mel
        loss_gen_all.backward()
        summarize(f"s{s}/ggrad", {k: p.grad for k, p in gen.named_parameters()}, out)
        optim_g.step()
        # ----
        for name, v in (("loss_disc_f", loss_disc_f), ("loss_disc_s", loss_disc_s), ("loss_disc_all", loss_disc_all),
                        ("loss_mel", loss_mel), ("loss_fm_f", loss_fm_f), ("loss_fm_s", loss_fm_s),
                        ("loss_gen_f", loss_gen_f), ("loss_gen_s", loss_gen_s), ("loss_gen_all", loss_gen_all)):
            out[f"s{s}/loss/{name}"] = np.float64(float(v))
        out[f"s{s}/y_g_hat"] = y_g_hat.detach().numpy()
        summarize(f"s{s}/after/gen", dict(gen.state_dict()), out)
        summarize(f"s{s}/after/mpd", dict(mpd.state_dict()), out)
        summarize(f"s{s}/after/msd", dict(msd.state_dict()), out)
    np.savez_compressed(os.path.join(HERE, "hifigan_step.npz"), **out)
    print("wrote", len(out), "arrays;", {k.split("/")[-1]: float(v) for k, v in out.items() if "/loss/" in k and k.startswith("s1")})


if __name__ == "__main__":
    main()
