"""Oracle (test infrastructure, see oracle/__init__.py): functional fp32 restatement of the HiFi-GAN v1 training path of
the reference, written against plain ``{state_dict key: tensor}`` mappings with the reference's keys
(``weight_g`` / ``weight_v`` / ``bias`` per weight-normed conv).

Reference files restated (paths relative to the reference tree, python/hifigan/...):
  models.py       ResBlock1 :17-54, Generator :81-137, DiscriminatorP :140-173, MultiPeriodDiscriminator :176-200,
                  DiscriminatorS :203-228, MultiScaleDiscriminator :231-260, feature_loss :263-269,
                  discriminator_loss :272-283, generator_loss :286-294
  meldataset.py   mel_spectrogram :217-240 (librosa 0.8.1 Slaney mel filterbank restated in mel_filterbank())
  utils.py        get_padding :35-36, init_weights :23-26
  config_v1.json  upsample rates 8,8,2,2 / kernels 16,16,4,4 / initial channel 512 / resblock kernels 3,7,11 /
                  dilations (1,3,5) / segment 8192 / n_fft 1024 / hop 256 / win 1024 / 80 mels / 22050 Hz
  xva_train.py    HiFiTrainer.iteration :467-515 (D step then G step, mel loss x45, AdamW lr 2e-4 betas .8/.99)

Pinned against outputs of the imported reference by tests/golden/make_golden_hifigan.py -> tests/test_oracle_golden.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1
UP_RATES, UP_KERNELS, UP_INIT = (8, 8, 2, 2), (16, 16, 4, 4), 512
RB_KERNELS, RB_DILATIONS = (3, 7, 11), ((1, 3, 5), (1, 3, 5), (1, 3, 5))
N_FFT, HOP, WIN, N_MELS, SR, FMIN, FMAX, FMAX_LOSS = 1024, 256, 1024, 80, 22050, 0, 8000, None
PERIODS = (2, 3, 5, 7, 11)


def get_padding(kernel_size, dilation=1):
    return int((kernel_size * dilation - dilation) / 2)


# ------------------------------------------------------------------------------------------------ parameters
def generator_spec():
    """(key, shape) of Generator(h).state_dict() for config_v1 (models.py:82-108): 234 keys, 13 936 130 parameters."""
    spec = []

    def wn(prefix, shape, transposed=False):
        g_shape = (shape[0], 1, 1)
        return [(f"{prefix}.bias", (shape[1] if transposed else shape[0],)), (f"{prefix}.weight_g", g_shape),
                (f"{prefix}.weight_v", shape)]

    spec += wn("conv_pre", (UP_INIT, 80, 7))
    for i, (u, k) in enumerate(zip(UP_RATES, UP_KERNELS)):
        spec += wn(f"ups.{i}", (UP_INIT // 2 ** i, UP_INIT // 2 ** (i + 1), k), transposed=True)
    for i in range(len(UP_RATES)):
        ch = UP_INIT // 2 ** (i + 1)
        for j, k in enumerate(RB_KERNELS):
            for name in ("convs1", "convs2"):
                for m in range(3):
                    spec += wn(f"resblocks.{i * 3 + j}.{name}.{m}", (ch, ch, k))
    spec += wn("conv_post", (1, UP_INIT // 2 ** len(UP_RATES), 7))
    return spec


def make_generator_state(seed=1234, scale=1.0):
    """Seeded weights in the reference's parameterisation. ``scale`` > 1 moves the N(0, 0.01) init of the reference
    (utils.py:23-26) to magnitudes where every layer contributes visibly to the output, which makes parity tests
    sensitive to each of them."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in generator_spec():
        if key.endswith("weight_v"):
            fan_in = shape[1] * shape[2]
            std = scale / math.sqrt(fan_in) if scale != 1.0 else 0.01
            if key.startswith("conv_pre") and scale == 1.0:
                std = 1.0 / math.sqrt(3 * fan_in)
            sd[key] = torch.randn(shape, generator=g) * std
        elif key.endswith("weight_g"):
            v = sd.get(key[:-1] + "v")
            sd[key] = None  # filled below (the state_dict order has g before v)
        else:
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    for key in list(sd):
        if key.endswith("weight_g"):
            v = sd[key[:-1] + "v"]
            sd[key] = v.flatten(1).norm(dim=1).view(-1, 1, 1) * (1.0 + 0.1 * torch.rand(v.shape[0], 1, 1, generator=g))
    return {k: sd[k] for k, _ in generator_spec()}


def wn_weight(sd, prefix):
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v|| with the norm over all dims but the first."""
    v, g = sd[f"{prefix}.weight_v"], sd[f"{prefix}.weight_g"]
    return v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1))


# ------------------------------------------------------------------------------------------------ generator
def resblock1(x, sd, prefix, k, dilations=(1, 3, 5)):
    """models.py:41-48"""
    for m, d in enumerate(dilations):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, wn_weight(sd, f"{prefix}.convs1.{m}"), sd[f"{prefix}.convs1.{m}.bias"], dilation=d,
                      padding=get_padding(k, d))
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, wn_weight(sd, f"{prefix}.convs2.{m}"), sd[f"{prefix}.convs2.{m}.bias"], padding=get_padding(k, 1))
        x = xt + x
    return x


def generator(sd, mel):
    """Generator.forward, models.py:110-128. mel [B, 80, T] -> [B, 1, 256 T]."""
    x = F.conv1d(mel, wn_weight(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(UP_RATES, UP_KERNELS)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, wn_weight(sd, f"ups.{i}"), sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, rk in enumerate(RB_KERNELS):
            r = resblock1(x, sd, f"resblocks.{i * 3 + j}", rk, RB_DILATIONS[j])
            xs = r if xs is None else xs + r
        x = xs / len(RB_KERNELS)
    x = F.leaky_relu(x)  # default slope 0.01 (models.py:124)
    x = F.conv1d(x, wn_weight(sd, "conv_post"), sd["conv_post.bias"], padding=3)
    return torch.tanh(x)


def generator_vits(sd, z, g=None):
    """HifiganGenerator.forward of the xVAPitch model (python/xvapitch/hifigan.py:234-262, configured at
    xvapitch/model.py:134-149: 192 latent channels in, plain -- not weight-normed -- conv_pre and conv_post, no conv_post
    bias, speaker conditioning cond_layer 512 -> 512 added after conv_pre). Groundwork for SURVEY.md section 8f rank 1:
    apart from the input width and the per-utterance conditioning vector it is Generator.forward above -- same
    transposed convolutions, same ResBlock1 stacks (xvapitch/hifigan.py:29-103), same MRF mean, same slopes -- so the
    generator kernels carry over. z [B, 192, T], g [B, 512, 1] or None -> [B, 1, 256 T]."""
    x = F.conv1d(z, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    if g is not None:
        x = x + F.conv1d(g, sd["cond_layer.weight"], sd["cond_layer.bias"])
    for i, (u, k) in enumerate(zip(UP_RATES, UP_KERNELS)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, wn_weight(sd, f"ups.{i}"), sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, rk in enumerate(RB_KERNELS):
            r = resblock1(x, sd, f"resblocks.{i * 3 + j}", rk, RB_DILATIONS[j])
            xs = r if xs is None else xs + r
        x = xs / len(RB_KERNELS)
    x = F.leaky_relu(x)
    return torch.tanh(F.conv1d(x, sd["conv_post.weight"], None, padding=3))


# ------------------------------------------------------------------------------------------------ mel spectrogram
def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr=SR, n_fft=N_FFT, n_mels=N_MELS, fmin=FMIN, fmax=FMAX):
    """librosa.filters.mel of librosa 0.8.1 (htk=False, norm='slaney'): the call at meldataset.py:225. librosa is not
    installed in the build image; this restates its published algorithm (triangles on the Slaney mel scale, each
    normalised by 2 / bandwidth). float32 like librosa's return value."""
    if fmax is None:
        fmax = sr / 2.0
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
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
    return torch.from_numpy(weights.astype(np.float32))


def mel_spectrogram(y, fmax=FMAX, n_fft=N_FFT, hop=HOP, win=WIN, n_mels=N_MELS, sr=SR, fmin=FMIN):
    """meldataset.py:217-240. y [B, N] -> [B, 80, N / 256]."""
    basis = mel_filterbank(sr, n_fft, n_mels, fmin, fmax).to(y.device)
    pad = int((n_fft - hop) / 2)
    yp = F.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = torch.stft(yp, n_fft, hop_length=hop, win_length=win, window=torch.hann_window(win, device=y.device), center=False,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    mag = torch.sqrt(spec.real ** 2 + spec.imag ** 2 + 1e-9)
    return torch.log(torch.clamp(torch.matmul(basis, mag), min=1e-5))


def tacotron_mel(y, n_fft=N_FFT, hop=HOP, n_mels=N_MELS, sr=SR, fmin=FMIN, fmax=FMAX):
    """TacotronSTFT.mel_spectrogram, fastpitch1_1/common/layers.py:121-138 with STFT.transform, common/stft.py:86-114
    (the FastPitch dataset's mel extractor): reflect-pad n_fft/2, windowed-DFT-basis conv1d with stride
    hop, sqrt(re^2 + im^2) (no epsilon), Slaney mel basis, log(clamp(., 1e-5)). y [B, N] in [-1, 1] -> [B, 80, N/256 + 1]."""
    assert float(y.min()) >= -1 and float(y.max()) <= 1
    nb = n_fft // 2 + 1
    fb = np.fft.fft(np.eye(n_fft))
    basis = torch.from_numpy(np.vstack([np.real(fb[:nb]), np.imag(fb[:nb])])).float()            # stft.py:63-69
    basis = basis[:, None, :] * torch.hann_window(n_fft, periodic=True).float()                 # stft.py:75-81 (fftbins=True)
    yp = F.pad(y[:, None, None, :], (n_fft // 2, n_fft // 2, 0, 0), mode="reflect").squeeze(1)  # stft.py:93-98
    ft = F.conv1d(yp, basis, stride=hop)
    mag = torch.sqrt(ft[:, :nb] ** 2 + ft[:, nb:] ** 2)
    mel = torch.matmul(mel_filterbank(sr, n_fft, n_mels, fmin, fmax), mag)
    return torch.log(torch.clamp(mel, min=1e-5))                                                # audio_processing: C = 1


# ------------------------------------------------------------------------------------------------ synthetic inputs
def synthetic_batch(B, frames, seed=1234):
    """audio = 0.95 tanh(N(0, 0.3)), x = mel(audio, fmax 8000), y_mel = mel(audio, fmax None)."""
    g = torch.Generator().manual_seed(seed)
    y = 0.95 * torch.tanh(torch.randn(B, frames * HOP, generator=g) * 0.3)
    return mel_spectrogram(y, FMAX), y, mel_spectrogram(y, FMAX_LOSS)


# ------------------------------------------------------------------------------------------------ discriminators
P_SPECS = [(1, 32, 5, 3, 2, 1), (32, 128, 5, 3, 2, 1), (128, 512, 5, 3, 2, 1), (512, 1024, 5, 3, 2, 1), (1024, 1024, 5, 1, 2, 1)]
S_SPECS = [(1, 128, 15, 1, 7, 1), (128, 128, 41, 2, 20, 4), (128, 256, 41, 2, 20, 16), (256, 512, 41, 4, 20, 16),
           (512, 1024, 41, 4, 20, 16), (1024, 1024, 41, 1, 20, 16), (1024, 1024, 5, 1, 2, 1)]
POST = (1024, 1, 3, 1, 1, 1)


def mpd_spec():
    """(key, shape) of MultiPeriodDiscriminator().state_dict() (models.py:176-186): 90 keys."""
    spec = []
    for d in range(len(PERIODS)):
        for name, (cin, cout, k, s, p, g) in [(f"convs.{i}", sp) for i, sp in enumerate(P_SPECS)] + [("conv_post", POST)]:
            pre = f"discriminators.{d}.{name}"
            spec += [(f"{pre}.bias", (cout,)), (f"{pre}.weight_g", (cout, 1, 1, 1)), (f"{pre}.weight_v", (cout, cin // g, k, 1))]
    return spec


def msd_spec():
    """(key, shape) of MultiScaleDiscriminator().state_dict() (models.py:231-243): 80 keys; disc 0 is spectral-normed."""
    spec = []
    for d in range(3):
        for name, (cin, cout, k, s, p, g) in [(f"convs.{i}", sp) for i, sp in enumerate(S_SPECS)] + [("conv_post", POST)]:
            pre = f"discriminators.{d}.{name}"
            if d == 0:
                spec += [(f"{pre}.bias", (cout,)), (f"{pre}.weight_orig", (cout, cin // g, k)), (f"{pre}.weight_u", (cout,)),
                         (f"{pre}.weight_v", (cin // g * k,))]
            else:
                spec += [(f"{pre}.bias", (cout,)), (f"{pre}.weight_g", (cout, 1, 1)), (f"{pre}.weight_v", (cout, cin // g, k))]
    return spec


def make_disc_state(spec, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in spec:
        if key.endswith("weight_u") or (key.endswith("weight_v") and len(shape) == 1):
            sd[key] = F.normalize(torch.randn(shape, generator=g), dim=0)
        elif key.endswith("weight_v") or key.endswith("weight_orig"):
            fan_in = int(np.prod(shape[1:]))
            sd[key] = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        elif key.endswith("weight_g"):
            sd[key] = None
        else:
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    for key in list(sd):
        if key.endswith("weight_g"):
            v = sd[key[:-1] + "v"]
            sd[key] = v.flatten(1).norm(dim=1).view(sd_shape(spec, key)) * (1.0 + 0.1 * torch.rand(v.shape[0], generator=g).view(sd_shape(spec, key)))
    return {k: sd[k] for k, _ in spec}


def sd_shape(spec, key):
    return dict(spec)[key]


def _disc_weight(sd, pre, training, state):
    """weight_norm (dim 0) or spectral_norm (one power iteration per training forward, torch.nn.utils.spectral_norm).
    ``state`` carries the evolving u vectors across calls within one oracle evaluation."""
    if f"{pre}.weight_g" in sd:
        v, g = sd[f"{pre}.weight_v"], sd[f"{pre}.weight_g"]
        return v * (g / v.flatten(1).norm(dim=1).view((-1,) + (1,) * (v.dim() - 1)))
    wo = sd[f"{pre}.weight_orig"]
    mat = wo.reshape(wo.shape[0], -1)
    u = state.setdefault(f"{pre}.u", sd[f"{pre}.weight_u"].clone())
    v = state.setdefault(f"{pre}.v", sd[f"{pre}.weight_v"].clone())
    if training:
        with torch.no_grad():
            v = F.normalize(torch.mv(mat.t(), u), dim=0, eps=1e-12)
            u = F.normalize(torch.mv(mat, v), dim=0, eps=1e-12)
            state[f"{pre}.u"], state[f"{pre}.v"] = u, v
    return wo / torch.dot(u, torch.mv(mat, v))


def discriminator_p(sd, pre, x, period, training=True, state=None):
    """DiscriminatorP.forward, models.py:154-173. x [B, 1, T] -> (flattened score, fmaps)"""
    state = {} if state is None else state
    fmap = []
    b, c, t = x.shape
    if t % period != 0:
        n_pad = period - (t % period)
        x = F.pad(x, (0, n_pad), "reflect")
        t = t + n_pad
    x = x.view(b, c, t // period, period)
    for i, (cin, cout, k, s, p, g) in enumerate(P_SPECS):
        x = F.conv2d(x, _disc_weight(sd, f"{pre}.convs.{i}", training, state), sd[f"{pre}.convs.{i}.bias"], stride=(s, 1), padding=(p, 0))
        x = F.leaky_relu(x, LRELU_SLOPE)
        fmap.append(x)
    x = F.conv2d(x, _disc_weight(sd, f"{pre}.conv_post", training, state), sd[f"{pre}.conv_post.bias"], padding=(1, 0))
    fmap.append(x)
    return torch.flatten(x, 1, -1), fmap


def discriminator_s(sd, pre, x, training=True, state=None, specs=None):
    """DiscriminatorS.forward, models.py:218-228 (specs: another layer table, e.g. VITS_S_SPECS)."""
    state = {} if state is None else state
    fmap = []
    for i, (cin, cout, k, s, p, g) in enumerate(S_SPECS if specs is None else specs):
        x = F.conv1d(x, _disc_weight(sd, f"{pre}.convs.{i}", training, state), sd[f"{pre}.convs.{i}.bias"], stride=s, padding=p, groups=g)
        x = F.leaky_relu(x, LRELU_SLOPE)
        fmap.append(x)
    x = F.conv1d(x, _disc_weight(sd, f"{pre}.conv_post", training, state), sd[f"{pre}.conv_post.bias"], padding=1)
    fmap.append(x)
    return torch.flatten(x, 1, -1), fmap


def mpd(sd, y, y_hat, training=True, state=None):
    """MultiPeriodDiscriminator.forward, models.py:187-200."""
    state = {} if state is None else state
    rs, gs, frs, fgs = [], [], [], []
    for i, p in enumerate(PERIODS):
        r, fr = discriminator_p(sd, f"discriminators.{i}", y, p, training, state)
        g_, fg = discriminator_p(sd, f"discriminators.{i}", y_hat, p, training, state)
        rs.append(r); gs.append(g_); frs.append(fr); fgs.append(fg)
    return rs, gs, frs, fgs


def msd(sd, y, y_hat, training=True, state=None):
    """MultiScaleDiscriminator.forward, models.py:244-260."""
    state = {} if state is None else state
    rs, gs, frs, fgs = [], [], [], []
    for i in range(3):
        if i != 0:
            y = F.avg_pool1d(y, 4, 2, padding=2)
            y_hat = F.avg_pool1d(y_hat, 4, 2, padding=2)
        r, fr = discriminator_s(sd, f"discriminators.{i}", y, training, state)
        g_, fg = discriminator_s(sd, f"discriminators.{i}", y_hat, training, state)
        rs.append(r); gs.append(g_); frs.append(fr); fgs.append(fg)
    return rs, gs, frs, fgs


# ---- xVAPitch discriminator
VITS_S_SPECS = [(1, 16, 15, 1, 7, 1), (16, 64, 41, 4, 20, 4), (64, 256, 41, 4, 20, 16), (256, 1024, 41, 4, 20, 64),
                (1024, 1024, 41, 4, 20, 256), (1024, 1024, 5, 1, 2, 1)]


def vits_disc_spec():
    """(key, shape) of VitsDiscriminator().state_dict() (python/xvapitch/model.py:1601-1607): nets.0 = the scale
    discriminator with VITS widths (model.py:1558-1568), nets.1..5 = DiscriminatorP(2, 3, 5, 7, 11)
    (xvapitch/hifigan.py:320-335, the layer table of hifigan/models.py's): 111 keys."""
    spec = []
    for name, (cin, cout, k, s, p, g) in [(f"convs.{i}", sp) for i, sp in enumerate(VITS_S_SPECS)] + [("conv_post", POST)]:
        pre = f"nets.0.{name}"
        spec += [(f"{pre}.bias", (cout,)), (f"{pre}.weight_g", (cout, 1, 1)), (f"{pre}.weight_v", (cout, cin // g, k))]
    for d in range(len(PERIODS)):
        for name, (cin, cout, k, s, p, g) in [(f"convs.{i}", sp) for i, sp in enumerate(P_SPECS)] + [("conv_post", POST)]:
            pre = f"nets.{d + 1}.{name}"
            spec += [(f"{pre}.bias", (cout,)), (f"{pre}.weight_g", (cout, 1, 1, 1)), (f"{pre}.weight_v", (cout, cin // g, k, 1))]
    return spec


def vits_discriminator(sd, x, x_hat=None):
    """VitsDiscriminator.forward, python/xvapitch/model.py:1609-1631: x, x_hat [B, 1, T] -> (x_scores, x_feats,
    x_hat_scores, x_hat_feats), every net on x and then on x_hat. Its losses are the LSGAN / feature-matching forms
    below (xvapitch/losses.py:65-85, 329-342 are hifigan/models.py:263-294 again)."""
    nets = [lambda w: discriminator_s(sd, "nets.0", w, specs=VITS_S_SPECS)]
    nets += [lambda w, i=i, p=p: discriminator_p(sd, f"nets.{i + 1}", w, p) for i, p in enumerate(PERIODS)]
    xs, xf = [], []
    hs, hf = ([], []) if x_hat is not None else (None, None)
    for net in nets:
        s_, f_ = net(x)
        xs.append(s_); xf.append(f_)
        if x_hat is not None:
            s_, f_ = net(x_hat)
            hs.append(s_); hf.append(f_)
    return xs, xf, hs, hf


def feature_loss(fmap_r, fmap_g):
    """models.py:263-269"""
    loss = 0
    for dr, dg in zip(fmap_r, fmap_g):
        for rl, gl in zip(dr, dg):
            loss = loss + torch.mean(torch.abs(rl - gl))
    return loss * 2


def discriminator_loss(rs, gs):
    """models.py:272-283"""
    loss = 0
    for dr, dg in zip(rs, gs):
        loss = loss + torch.mean((1 - dr) ** 2) + torch.mean(dg ** 2)
    return loss


def generator_loss(gs):
    """models.py:286-294"""
    loss = 0
    for dg in gs:
        loss = loss + torch.mean((1 - dg) ** 2)
    return loss


# ------------------------------------------------------------------------------------------------ training step
def adamw_step(params, grads, state, lr=2e-4, betas=(0.8, 0.99), eps=1e-8, weight_decay=0.01):
    """torch.optim.AdamW (decoupled weight decay, bias correction), the optimizer of xva_train.py:298-300."""
    b1, b2 = betas
    for k, g in grads.items():
        if g is None:
            continue
        p = params[k]
        st = state.setdefault(k, {"m": torch.zeros_like(p), "v": torch.zeros_like(p), "t": 0})
        st["t"] += 1
        p.mul_(1 - lr * weight_decay)
        st["m"].mul_(b1).add_(g, alpha=1 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (st["v"].sqrt() / math.sqrt(1 - b2 ** st["t"])).add_(eps)
        p.addcdiv_(st["m"], denom, value=-lr / (1 - b1 ** st["t"]))


def _leaves(sd):
    is_buf = lambda k, v: k.endswith("weight_u") or (k.endswith("weight_v") and v.dim() == 1)
    return {k: (v if is_buf(k, v) else v.detach().requires_grad_(True)) for k, v in sd.items()}


def train_step(sd_g, sd_mpd, sd_msd, x, y, y_mel, opt_state, lr=2e-4, detail=None):
    """HiFiTrainer.iteration body, xva_train.py:467-515. Updates the three state dicts in place (including the
    spectral-norm u / v buffers: four power iterations per step, two per msd call). Returns the loss terms. ``detail``
    (a dict) additionally receives every loss term, the generated waveform and the D-step parameter gradients
    ({"mpd": {key: grad}, "msd": {key: grad}}) -- what tests/golden/hifigan_step.npz records from the reference."""
    y = y.unsqueeze(1) if y.dim() == 2 else y
    lg, lp, ls = _leaves(sd_g), _leaves(sd_mpd), _leaves(sd_msd)
    sn_state = {}
    y_g_hat = generator(lg, x)
    y_g_hat_mel = mel_spectrogram(y_g_hat.squeeze(1), FMAX_LOSS)
    # D step
    rs, gs, _, _ = mpd(lp, y, y_g_hat.detach(), True, sn_state)
    loss_disc_f = discriminator_loss(rs, gs)
    rs, gs, _, _ = msd(ls, y, y_g_hat.detach(), True, sn_state)
    loss_disc_s = discriminator_loss(rs, gs)
    loss_disc_all = loss_disc_s + loss_disc_f
    d_keys = [(ls, k) for k, v in ls.items() if v.requires_grad] + [(lp, k) for k, v in lp.items() if v.requires_grad]
    d_grads = torch.autograd.grad(loss_disc_all, [s[k] for s, k in d_keys])
    with torch.no_grad():
        adamw_step({("s" if s is ls else "p", k): (sd_msd if s is ls else sd_mpd)[k] for s, k in d_keys},
                   {("s" if s is ls else "p", k): g for (s, k), g in zip(d_keys, d_grads)}, opt_state.setdefault("d", {}), lr)
    if detail is not None:
        detail["dgrad"] = {"msd": {}, "mpd": {}}
        for (s_, k), g_ in zip(d_keys, d_grads):
            detail["dgrad"]["msd" if s_ is ls else "mpd"][k] = g_
        detail["y_g_hat"] = y_g_hat.detach()
    # G step (discriminators now hold the updated weights; fresh leaves)
    lp, ls = _leaves(sd_mpd), _leaves(sd_msd)
    loss_mel = F.l1_loss(y_mel, y_g_hat_mel) * 45
    rs, gs, frs, fgs = mpd(lp, y, y_g_hat, True, sn_state)
    loss_fm_f, loss_gen_f = feature_loss(frs, fgs), generator_loss(gs)
    rs, gs, frs, fgs = msd(ls, y, y_g_hat, True, sn_state)
    loss_fm_s, loss_gen_s = feature_loss(frs, fgs), generator_loss(gs)
    loss_gen_all = loss_gen_s + loss_gen_f + loss_fm_s + loss_fm_f + loss_mel
    g_keys = [k for k, v in lg.items()]
    g_grads = torch.autograd.grad(loss_gen_all, [lg[k] for k in g_keys])
    if detail is not None:
        detail["loss"] = {"loss_disc_f": loss_disc_f, "loss_disc_s": loss_disc_s, "loss_disc_all": loss_disc_all,
                          "loss_mel": loss_mel, "loss_fm_f": loss_fm_f, "loss_fm_s": loss_fm_s, "loss_gen_f": loss_gen_f,
                          "loss_gen_s": loss_gen_s, "loss_gen_all": loss_gen_all}
        detail["loss"] = {k: float(v.detach()) for k, v in detail["loss"].items()}
    with torch.no_grad():
        adamw_step({k: sd_g[k] for k in g_keys}, dict(zip(g_keys, g_grads)), opt_state.setdefault("g", {}), lr)
        for k in sd_msd:        # persist the power-iteration vectors like the module buffers do
            if k.endswith("weight_u") and f"{k[:-9]}.u" in sn_state:
                sd_msd[k] = sn_state[f"{k[:-9]}.u"].clone()
                sd_msd[k[:-1] + "v"] = sn_state[f"{k[:-9]}.v"].clone()
    return {"loss_gen_all": loss_gen_all.detach(), "loss_disc_all": loss_disc_all.detach(), "loss_mel": loss_mel.detach(),
            "loss_fm": (loss_fm_s + loss_fm_f).detach(), "loss_gen": (loss_gen_s + loss_gen_f).detach()}, dict(zip(g_keys, g_grads))
