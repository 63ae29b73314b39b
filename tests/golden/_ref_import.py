"""Import shims that let the UNMODIFIED reference modules under /root/reference import in the build container.

Only used by tests/golden/make_golden.py (fixture generation, build container only). The reference needs
`librosa` and `matplotlib`, which are not installed; the two functions of librosa that the hot path really calls
(`filters.mel`, `util.normalize`) are provided by independent restatements (Slaney mel filterbank as in librosa
0.8.1), everything else is a stub.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("XVA_REFERENCE_ROOT", "/root/reference")


def slaney_mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **_):
    """librosa.filters.mel (htk=False, norm='slaney'), restated from the published algorithm."""
    if fmax is None:
        fmax = sr / 2.0

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        f_sp = 200.0 / 3
        mels = f / f_sp
        min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
        min_log_mel = min_log_hz / f_sp
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        f_sp = 200.0 / 3
        min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
        min_log_mel = min_log_hz / f_sp
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
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


def install():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mpl = stub("matplotlib", use=lambda *a, **k: None)
    mpl.pylab = stub("matplotlib.pylab")
    lib = stub("librosa")
    lib.filters = stub("librosa.filters", mel=slaney_mel)
    lib.util = stub("librosa.util", pad_center=lambda d, size, **k: d, tiny=lambda x: 1e-30,
                    normalize=lambda x, **k: x / (np.abs(x).max() + 1e-12))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def install_xvapitch():
    """install() plus what python/xvapitch/model.py and losses.py import at module level and the build container lacks:
    the text front end (unidecode, g2pc ... -- out of scope) and soundfile are stubs; np.bool is the
    alias numpy >= 1.24 removed (xvapitch/util.py:28). scipy.signal is imported first because numpy.ma must not see the
    alias while it initialises."""
    import scipy.signal  # noqa: F401

    np.bool = np.bool_
    install()
    text = types.ModuleType("python.xvapitch.text")
    text.get_text_preprocessor, text.ALL_SYMBOLS, text.lang_names = None, list(range(200)), {}
    sys.modules["python.xvapitch.text"] = text
    sys.modules.setdefault("soundfile", types.ModuleType("soundfile"))
