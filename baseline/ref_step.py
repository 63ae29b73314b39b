"""Drives the UNMODIFIED reference modules (a verbatim copy under baseline/_ref/, made by baseline/make_ref.sh and
git-ignored) through the two trainer loop bodies this repository replaces:

    FastPitchTrainer.iteration   python/fastpitch1_1/xva_train.py:784-862   (forward, FastPitchLoss, /gam, [scaled] backward,
                                                                             unscale + clip_grad_norm_(1000) + Lamb.step)
    HiFiTrainer.iteration        python/hifigan/xva_train.py:467-515        (G forward, mel, D step, G step, two AdamW)

None of this repository's kernels, models or engine is on that path. Two users:

  * bench.py --impl reference : the reference arm, on the box's host cores (`kind: "reference"`);
  * bench.py (native arm) and `python baseline/ref_step.py` : the same steps under stock PyTorch eager ON THE B200 --
    the kernel-for-kernel bar (`"eager_b200"`), in the trainer's own modes: fp16 autocast + GradScaler (its default,
    xva_train.py:698), fp32 with TF32 convolutions (torch's default), and strict fp32 (TF32 off). The probe also records how
    far the reference's own AMP / TF32 results are from its strict-fp32 results on the same inputs (the yardstick for this
    repository's tf32 parity table, profiles/r02_parity_table.txt).

Import shims only (no edits to reference code): matplotlib is stubbed; librosa.filters.mel comes from
torchaudio.functional.melscale_fbanks (Slaney scale + norm = librosa 0.8.1's default); FastPitchLoss hard-codes
`torch.device('cuda:N')` placeholders (loss_function.py:92-129), which is redirected to the CPU for the CPU run.
"""
import json
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
TT, TM = 160, 880


def available():
    return os.path.exists(os.path.join(REF, "python", "fastpitch1_1", "fastpitch", "model.py"))


def _mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **_):
    import torchaudio

    if fmax is None:
        fmax = sr / 2.0
    return torchaudio.functional.melscale_fbanks(1 + n_fft // 2, float(fmin), float(fmax), n_mels, sr, norm="slaney",
                                                 mel_scale="slaney").T.contiguous().numpy().astype(np.float32)


def install():
    """Make `python.fastpitch1_1...` / `python.hifigan...` importable from baseline/_ref."""
    if not available():
        raise RuntimeError("baseline/_ref is missing: run baseline/make_ref.sh in the build container")

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "librosa" not in sys.modules:
        mpl = stub("matplotlib", use=lambda *a, **k: None)
        mpl.pylab = stub("matplotlib.pylab")
        lib = stub("librosa")
        lib.filters = stub("librosa.filters", mel=_mel)
        lib.util = stub("librosa.util", pad_center=lambda d, size, **k: d, tiny=lambda x: 1e-30,
                        normalize=lambda x, **k: x / (np.abs(x).max() + 1e-12))
    if REF not in sys.path:
        sys.path.insert(0, REF)


class _CpuDevice:
    """torch.device('cuda:N') -> cpu while the reference FastPitchLoss runs on the host (see module docstring)."""

    def __enter__(self):
        self.real = torch.device
        real = self.real

        class Fake:
            def __new__(cls, *a, **k):
                return real("cpu")

        torch.device = Fake
        return self

    def __exit__(self, *a):
        torch.device = self.real


# ---------------------------------------------------------------------------------------------- synthetic inputs
def fastpitch_batch(B, Tt=TT, Tm=TM, seed=1234):
    """SURVEY 8(d) cfg-2 inputs as the 12-list of data_function.py:737-738 (full-length utterances): tokens U{1..147},
    durations = 1 + multinomial(Tm - Tt extra frames), mel N(0,1), pitch N(0,1) with 30 % zeros, energy U(0,10)."""
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(1, 148, (B, Tt), generator=g)
    durs = torch.ones(B, Tt)
    extra = torch.multinomial(torch.ones(B, Tt), Tm - Tt, replacement=True, generator=g)
    durs.scatter_add_(1, extra, torch.ones(B, Tm - Tt))
    mel = torch.randn(B, 80, Tm, generator=g)
    pitch = torch.randn(B, 1, Tm, generator=g) * (torch.rand(B, 1, Tm, generator=g) > 0.3)
    energy = torch.rand(B, Tm, generator=g) * 10
    in_lens = torch.full((B,), Tt, dtype=torch.int64)
    mel_lens = torch.full((B,), Tm, dtype=torch.int64)
    x = [text, in_lens, mel, mel_lens, pitch, energy, None, None, durs, torch.full((B,), float(Tt)),
         torch.full((B,), float(Tm)), ["synthetic"] * B]
    return x


def _to(x, dev):
    return [t.to(dev) if torch.is_tensor(t) else t for t in x]


# ---------------------------------------------------------------------------------------------- FastPitch step
class FastPitchRef:
    """The reference model + criterion + Lamb + GradScaler wired as FastPitchTrainer.init / iteration do."""

    def __init__(self, device, stage=3, mode="fp32", state=None, seed=1234, dropout=True):
        install()
        from python.fastpitch1_1.fastpitch.loss_function import FastPitchLoss
        from python.fastpitch1_1.fastpitch.model import FastPitch
        from python.fastpitch1_1.lamb import Lamb

        self.dev = torch.device(device)
        self.mode, self.stage = mode, stage
        self.amp = mode == "amp_fp16"
        if self.dev.type == "cuda":
            tf32 = mode != "fp32_strict"
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False  # torch default; the trainer never touches it
        torch.manual_seed(seed)
        self.model = FastPitch(logger=None).to(self.dev)
        if state is not None:
            self.model.load_state_dict(state)
        self.model.training_stage = stage
        self.model.train(dropout)
        gpus = [self.dev.index or 0]
        self.criterion = FastPitchLoss(dur_predictor_loss_scale=0.1, pitch_predictor_loss_scale=0.1, attn_loss_scale=1.0,
                                       gpus=gpus)
        self.opt = Lamb(self.model.parameters(), lr=0.1, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
        self.scaler = torch.amp.GradScaler("cuda", enabled=self.amp)
        self.it = 50000

    def lr(self):
        # adjust_learning_rate, xva_train.py:1252-1261 (noam, warm-up 1000)
        self.it += 1
        scale = 1.0 / (self.it ** 0.5) if self.it > 1000 else self.it / (1000 ** 1.5)
        for g in self.opt.param_groups:
            g["lr"] = 0.1 * scale

    def fwd_bwd(self, x):
        y = [x[2], x[1], x[3], x[9]]
        self.model.zero_grad(set_to_none=True)
        ctx = _CpuDevice() if self.dev.type == "cpu" else _Null()
        with torch.autocast("cuda", dtype=torch.float16, enabled=self.amp):
            y_pred = self.model(x)
            with ctx:
                loss, meta, parts = self.criterion(y_pred, y, training_stage=self.stage)
        if self.amp:
            self.scaler.scale(loss).backward()
        else:
            loss.backward()
        return y_pred, loss, meta

    def step(self, x):
        self.lr()
        y_pred, loss, meta = self.fwd_bwd(x)
        if self.amp:
            self.scaler.unscale_(self.opt)
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), 1000)
            self.scaler.step(self.opt)
            self.scaler.update()
        else:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), 1000)
            self.opt.step()
        return loss


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


# ---------------------------------------------------------------------------------------------- HiFi-GAN step
def hifigan_batch(B, frames=32, seed=1):
    """SURVEY 8(d) cfg-3 inputs: audio 0.95 tanh(N(0, 0.3)), input mel fmax 8000, loss mel fmax None."""
    install()
    from python.hifigan import meldataset as md

    g = torch.Generator().manual_seed(seed)
    y = 0.95 * torch.tanh(torch.randn(B, frames * 256, generator=g) * 0.3)
    x = md.mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, 8000)
    y_mel = md.mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, None)
    return x, y, y_mel


class HiFiGANRef:
    def __init__(self, device, mode="fp32", seed=1234):
        install()
        import itertools

        from python.hifigan import meldataset as md
        from python.hifigan.models import (AttrDict, Generator, MultiPeriodDiscriminator, MultiScaleDiscriminator,
                                            discriminator_loss, feature_loss, generator_loss)

        self.dev = torch.device(device)
        if self.dev.type == "cuda":
            tf32 = mode != "fp32_strict"
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False
        h = AttrDict(json.load(open(os.path.join(REF, "python", "hifigan", "config_v1.json"))))
        h.USE_EMB_CONDITIONING = False
        self.h = h
        torch.manual_seed(seed)
        self.G = Generator(h).to(self.dev)
        self.mpd = MultiPeriodDiscriminator().to(self.dev)
        self.msd = MultiScaleDiscriminator().to(self.dev)
        self.optim_g = torch.optim.AdamW(self.G.parameters(), h.learning_rate, betas=[h.adam_b1, h.adam_b2])
        self.optim_d = torch.optim.AdamW(itertools.chain(self.msd.parameters(), self.mpd.parameters()), h.learning_rate,
                                         betas=[h.adam_b1, h.adam_b2])
        self.md, self.dl, self.fl, self.gl = md, discriminator_loss, feature_loss, generator_loss
        for m in (self.G, self.mpd, self.msd):
            m.train()

    def step(self, x, y, y_mel):
        """hifigan/xva_train.py:467-515."""
        import torch.nn.functional as F

        h = self.h
        self.G.zero_grad(set_to_none=True)
        self.mpd.zero_grad(set_to_none=True)
        self.msd.zero_grad(set_to_none=True)
        y = y.unsqueeze(1)
        y_g_hat = self.G(x)
        y_g_hat_mel = self.md.mel_spectrogram(y_g_hat.squeeze(1), h.n_fft, h.num_mels, h.sampling_rate, h.hop_size,
                                              h.win_size, h.fmin, h.fmax_for_loss)
        self.optim_d.zero_grad()
        y_df_hat_r, y_df_hat_g, _, _ = self.mpd(y, y_g_hat.detach())
        loss_disc_f, _, _ = self.dl(y_df_hat_r, y_df_hat_g)
        y_ds_hat_r, y_ds_hat_g, _, _ = self.msd(y, y_g_hat.detach())
        loss_disc_s, _, _ = self.dl(y_ds_hat_r, y_ds_hat_g)
        loss_disc_all = loss_disc_s + loss_disc_f
        loss_disc_all.backward()
        self.optim_d.step()
        self.optim_g.zero_grad()
        loss_mel = F.l1_loss(y_mel, y_g_hat_mel) * 45
        y_df_hat_r, y_df_hat_g, fmap_f_r, fmap_f_g = self.mpd(y, y_g_hat)
        y_ds_hat_r, y_ds_hat_g, fmap_s_r, fmap_s_g = self.msd(y, y_g_hat)
        loss_fm_f = self.fl(fmap_f_r, fmap_f_g)
        loss_fm_s = self.fl(fmap_s_r, fmap_s_g)
        loss_gen_f, _ = self.gl(y_df_hat_g)
        loss_gen_s, _ = self.gl(y_ds_hat_g)
        loss_gen_all = loss_gen_s + loss_gen_f + loss_fm_s + loss_fm_f + loss_mel
        loss_gen_all.backward()
        self.optim_g.step()
        return {"loss_disc_all": loss_disc_all.detach(), "loss_gen_all": loss_gen_all.detach(), "loss_mel": loss_mel.detach()}


# ---------------------------------------------------------------------------------------------- xVAPitch --hifi_only
def xvapitch_available():
    return os.path.exists(os.path.join(REF, "python", "xvapitch", "model.py"))


def install_xvapitch():
    """install() plus what python/xvapitch/model.py and losses.py import at module level and this image lacks: the text
    front end (out of scope) and soundfile are stubs; np.bool is the alias numpy >= 1.24 removed
    (xvapitch/util.py:28); scipy.signal is imported first because numpy.ma must not see the alias while it initialises."""
    import scipy.signal  # noqa: F401

    np.bool = np.bool_
    install()
    if not xvapitch_available():
        raise RuntimeError("baseline/_ref/python/xvapitch is missing: run baseline/make_ref.sh in the build container")
    text = types.ModuleType("python.xvapitch.text")
    text.get_text_preprocessor, text.ALL_SYMBOLS, text.lang_names = None, list(range(200)), {}
    sys.modules["python.xvapitch.text"] = text
    sys.modules.setdefault("soundfile", types.ModuleType("soundfile"))


def vits_batch(B, T=256, seed=1):
    """Synthetic --hifi_only batch: linear spectrogram [B, 513, T] (|N(0, 0.5)|), ragged lengths in [3T/4, T] (first = T),
    waveform [B, 1, 256 T] = 0.9 tanh(N(0, 0.3)), d-vectors [B, 512]."""
    g = torch.Generator().manual_seed(seed)
    linear = torch.randn(B, 513, T, generator=g).abs() * 0.5
    waveform = 0.9 * torch.tanh(torch.randn(B, 1, T * 256, generator=g) * 0.3)
    d_vectors = torch.randn(B, 512, generator=g)
    lens = torch.randint(3 * T // 4, T + 1, (B,), generator=g)
    lens[0] = T
    return linear, lens, waveform, d_vectors


class VitsHifiOnlyRef:
    """The reference's --hifi_only iteration from its own modules: xVAPitch.train_hifi_only (xvapitch/model.py:650-678),
    xVAPitch.forward (model.py:271-340, 385-399), the trainer's loop body (xvapitch/xva_train.py:651-736) and optimizers
    (training_util.py:31-32, 66-67). The full xVAPitch model is not instantiated (its text front end needs packages this
    image lacks); the composition below is the one tests/golden/make_golden_vits_hifi_only.py recorded the golden step with."""

    def __init__(self, device, mode="fp32", seed=1234):
        install_xvapitch()
        from python.xvapitch.hifigan import HifiganGenerator
        from python.xvapitch.losses import VitsDiscriminatorLoss, VitsGeneratorLoss
        from python.xvapitch.model import PosteriorEncoder, VitsDiscriminator
        import python.xvapitch.util as ref_util

        self.dev = torch.device(device)
        self.amp = mode == "amp_fp16"
        if self.dev.type == "cuda":
            tf32 = mode != "fp32_strict"
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False
        torch.manual_seed(seed)
        self.enc = PosteriorEncoder(513, 192, 192, kernel_size=5, dilation_rate=1, num_layers=16, cond_channels=512).to(self.dev)
        self.dec = HifiganGenerator(192, 1, "1", [[1, 3, 5]] * 3, [3, 7, 11], [16, 16, 4, 4], 512, [8, 8, 2, 2],
                                    inference_padding=0, cond_channels=512, conv_pre_weight_norm=False,
                                    conv_post_weight_norm=False, conv_post_bias=False).to(self.dev)
        self.disc = VitsDiscriminator().to(self.dev)
        args = types.SimpleNamespace(hifi_only=True, analyze_loss=False, pitch=False, energy=False, mltts_rc=False)
        self.crit_g, self.crit_d = VitsGeneratorLoss(args).to(self.dev), VitsDiscriminatorLoss().to(self.dev)
        gen_params = list(self.enc.parameters()) + list(self.dec.parameters())
        self.opt0 = torch.optim.AdamW(gen_params, lr=0.000175, betas=[0.8, 0.99], eps=1e-09, weight_decay=0.01)
        self.opt1 = torch.optim.AdamW(self.disc.parameters(), lr=0.0002, betas=[0.8, 0.99], eps=1e-09, weight_decay=0.01)
        self.scaler = torch.cuda.amp.GradScaler(enabled=self.amp)
        self.util = ref_util
        for m in (self.enc, self.dec, self.disc):
            m.train()

    def step(self, linear, lens, waveform, d_vectors):
        import torch.nn.functional as F

        ac = torch.autocast("cuda", dtype=torch.float16, enabled=self.amp) if self.dev.type == "cuda" else _Null()
        self.opt0.zero_grad()
        with ac:
            g = F.normalize(d_vectors).unsqueeze(-1)
            z, m_q, logs_q, y_mask = self.enc(linear, lens, g=g)
            z_slice, slice_ids = self.util.rand_segments(z, lens, 32)
            o = self.dec(z_slice, g=g)
            wav_seg = self.util.segment(waveform, slice_ids * 256, 32 * 256)
            scores_fake, feats_fake, _, feats_real = self.disc(o, wav_seg)
            loss_dict = self.crit_g(waveform_hat=o.float(), waveform=wav_seg.float(), z_p=None, logs_q=None, m_p=None,
                                    logs_p=None, z_mask=None, scores_disc_fake=scores_fake, feats_disc_fake=feats_fake,
                                    feats_disc_real=feats_real, loss_duration=None)
        self.scaler.scale(loss_dict["loss"].mean()).backward()
        self.opt1.zero_grad()
        with ac:
            s_fake, _, s_real, _ = self.disc(o.detach(), wav_seg)
            loss_d = self.crit_d(s_real, s_fake)
        self.scaler.scale(loss_d["loss"].mean()).backward()
        for opt in (self.opt0, self.opt1):                                  # xva_train.py:729-736
            if self.amp:
                self.scaler.step(opt)
                self.scaler.update()
            else:
                opt.step()
        return {"loss": loss_dict["loss"].detach(), "loss_disc": loss_d["loss"].detach()}


# ---------------------------------------------------------------------------------------------- timing helpers
def _median(v):
    v = sorted(v)
    return v[len(v) // 2]


def time_cuda(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return _median(ts)


def time_cpu(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return _median(ts), ts


def eager_b200(device="cuda:0", batch=32, stage=3, steps=10, warmup=3, hifigan=True):
    """ms/step of the unmodified reference step under PyTorch eager on this GPU, in the trainer's modes."""
    out = {"fastpitch": {}, "hifigan": {}, "torch": torch.__version__, "median_of": steps, "warmup": warmup}
    x = _to(fastpitch_batch(batch), device)
    frames = batch * TM
    for mode in ("amp_fp16", "fp32", "fp32_strict"):
        r = FastPitchRef(device, stage, mode)
        ms = time_cuda(lambda: r.step(x), warmup, steps)
        out["fastpitch"][mode] = {"ms_per_step": ms, "frames_per_s": frames / (ms * 1e-3)}
        del r
        torch.cuda.empty_cache()
    if hifigan:
        hb = [t.to(device) for t in hifigan_batch(16)]
        for mode in ("fp32", "fp32_strict"):
            r = HiFiGANRef(device, mode)
            ms = time_cuda(lambda: r.step(*hb), warmup, steps)
            out["hifigan"][mode] = {"ms_per_step": ms, "samples_per_s": 16 * 8192 / (ms * 1e-3)}
            del r
            torch.cuda.empty_cache()
    if hifigan and xvapitch_available():
        out["xvapitch_hifi_only"] = {}
        vb = [t.to(device) for t in vits_batch(16)]
        for mode in ("amp_fp16", "fp32", "fp32_strict"):
            r = VitsHifiOnlyRef(device, mode)
            ms = time_cuda(lambda: r.step(*vb), warmup, steps)
            out["xvapitch_hifi_only"][mode] = {"ms_per_step": ms, "samples_per_s": 16 * 8192 / (ms * 1e-3),
                                               "spec_frames_per_s": float(vb[1].sum()) / (ms * 1e-3)}
            del r
            torch.cuda.empty_cache()
    out["modes"] = {"amp_fp16": "torch.autocast(fp16) + GradScaler: the trainer's default (xva_train.py:698,787)",
                    "fp32": "fp32 parameters, cuDNN TF32 convolutions allowed (torch default), fp32 matmul",
                    "fp32_strict": "TF32 off everywhere"}
    return out


def amp_error_table(device="cuda:0", B=4, Tt=40, Tm=150, stage=3):
    """How far the reference's OWN reduced-precision modes are from its strict-fp32 result on identical weights and
    inputs (dropout off): relative L2 error of mel_out, the losses and every parameter gradient."""
    x = fastpitch_batch(B, Tt, Tm, seed=11)
    x = _to(x, device)
    base = FastPitchRef(device, stage, "fp32_strict", dropout=False)
    state = {k: v.clone() for k, v in base.model.state_dict().items()}
    res = {}
    ref_out = None
    for mode in ("fp32_strict", "fp32", "amp_fp16"):
        r = FastPitchRef(device, stage, mode, state=state, dropout=False)
        y_pred, loss, meta = r.fwd_bwd(x)
        inv = 1.0 / r.scaler.get_scale() if r.amp else 1.0
        grads = {k: (p.grad.detach().double() * inv) for k, p in r.model.named_parameters() if p.grad is not None}
        d = lambda t: None if t is None else t.detach().double()
        cur = {"mel_out": d(y_pred[0]), "loss": d(loss), "pitch_pred": d(y_pred[4]), "energy_pred": d(y_pred[6]),
               "grads": grads}
        if ref_out is None:
            ref_out = cur
            continue
        rel = lambda a, b: float("nan") if a is None or b is None else float((a - b).norm() / b.norm().clamp_min(1e-30))
        ge = {k: rel(v, ref_out["grads"][k]) for k, v in grads.items() if float(ref_out["grads"][k].norm()) > 0}
        num = sum(float((v - ref_out["grads"][k]).pow(2).sum()) for k, v in grads.items())
        den = sum(float(v.pow(2).sum()) for v in ref_out["grads"].values())
        worst = max(ge.items(), key=lambda kv: kv[1])
        srt = sorted(ge.values())
        res[mode] = {"mel_out": rel(cur["mel_out"], ref_out["mel_out"]), "loss": rel(cur["loss"], ref_out["loss"]),
                     "pitch_pred": rel(cur["pitch_pred"], ref_out["pitch_pred"]),
                     "energy_pred": rel(cur["energy_pred"], ref_out["energy_pred"]),
                     "grad_global": (num / den) ** 0.5, "grad_median": srt[len(srt) // 2], "grad_worst": list(worst)}
    return res


def main():
    os.makedirs(os.path.join(os.path.dirname(HERE), "gpurun_out"), exist_ok=True)
    out = {"eager_b200": eager_b200(), "reference_vs_its_own_strict_fp32": {}}
    for shape in ((4, 40, 150), (8, 160, 880)):
        out["reference_vs_its_own_strict_fp32"]["B%d_Tt%d_Tm%d" % shape] = amp_error_table("cuda:0", *shape)
    path = os.path.join(os.path.dirname(HERE), "gpurun_out", "ref_eager_probe.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
