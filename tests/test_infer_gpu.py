"""FastPitch.infer (free-running synthesis, model.py:426-482) on the B200 engine vs the outputs recorded from the
unmodified reference (tests/golden/make_golden_infer.py) and the CPU oracle. Product path (tf32 tensor cores): predictor
outputs within 2e-3 (dur_pred = exp(x) - 1: 6e-3, as in tests/test_fastpitch_gpu.py), mel within 3e-3. The number of
frames per token is floor(dur * pace + 0.5): it may differ from the reference only for a token whose dur * pace + 0.5 is
within the predictor tolerance of an integer; the mel comparison therefore also runs with the reference's own predicted
durations handed in as dur_tgt, which makes the frame layout identical by construction."""
import os

import numpy as np
import pytest
import torch

from oracle import fastpitch as ofp
from test_oracle_golden import infer_kwargs, infer_state

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _model():
    from xva_trainer_b200 import fastpitch as fp

    m = fp.FastPitch(device="cuda:0")
    m.load_state_dict(infer_state())
    m.eval()
    return m


@pytest.mark.parametrize("case", ["free", "pace", "forced"])
def test_infer_matches_reference_golden(lib, case):
    g = np.load(os.path.join(GOLD, "infer.npz"))
    t = lambda k: torch.from_numpy(g[k])
    kw = infer_kwargs(g, case)
    ckw = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items()}
    m = _model()
    text = t("text").cuda()
    mel, dec_lens, dur_pred, pitch_pred, energy_pred = m.infer(text, **ckw)
    torch.cuda.synchronize()
    assert rel(dur_pred, t(f"{case}/dur_pred")) < 6e-3, rel(dur_pred, t(f"{case}/dur_pred"))
    assert rel(pitch_pred, t(f"{case}/pitch_pred")) < 2e-3 and rel(energy_pred, t(f"{case}/energy_pred")) < 2e-3
    assert tuple(pitch_pred.shape) == tuple(g[f"{case}/pitch_pred"].shape) and dec_lens.dtype == torch.int64
    pace = kw["pace"]
    want_d = kw.get("dur_tgt", t(f"{case}/dur_pred"))
    got_d = kw.get("dur_tgt", dur_pred.cpu())
    reps_w, reps_g = torch.floor(want_d * pace + 0.5), torch.floor(got_d * pace + 0.5)
    frac = (want_d * pace + 0.5) % 1.0
    near_boundary = (frac < 2e-2) | (frac > 1 - 2e-2)
    assert bool(((reps_w == reps_g) | near_boundary).all()), "frame counts differ away from a rounding boundary"
    if bool((reps_w == reps_g).all()):
        assert torch.equal(dec_lens.cpu(), t(f"{case}/dec_lens"))
        assert tuple(mel.shape) == tuple(g[f"{case}/mel"].shape)
        assert rel(mel, t(f"{case}/mel")) < 3e-3, rel(mel, t(f"{case}/mel"))
    # identical frame layout by construction: the reference's predicted durations as dur_tgt
    ckw2 = dict(ckw)
    ckw2["dur_tgt"] = want_d.cuda()
    mel2, dec2, _, _, _ = m.infer(text, **ckw2)
    assert torch.equal(dec2.cpu(), t(f"{case}/dec_lens")) and tuple(mel2.shape) == tuple(g[f"{case}/mel"].shape)
    assert rel(mel2, t(f"{case}/mel")) < 3e-3, rel(mel2, t(f"{case}/mel"))


def test_infer_energy_target_and_pitch_transform(lib):
    """The two arguments the golden cannot cover: energy_tgt (the reference dies on an unbound energy_pred, model.py:482)
    and pitch_transform (a callable applied to the predicted pitch, :444-452) -- against the oracle."""
    g = np.load(os.path.join(GOLD, "infer.npz"))
    text = torch.from_numpy(g["text"])
    gen = torch.Generator().manual_seed(5)
    energy = torch.rand(text.shape[0], 1, text.shape[1], generator=gen) * (text != 0).unsqueeze(1)
    durs = torch.from_numpy(g["forced/dur_tgt"])
    sd = infer_state()
    with torch.no_grad():
        want_mel, want_lens, _, want_pitch, _ = ofp.infer(sd, text, dur_tgt=durs, energy_tgt=energy)
    m = _model()
    mel, lens, _, pitch, energy_pred = m.infer(text.cuda(), dur_tgt=durs.cuda(), energy_tgt=energy.cuda())
    assert energy_pred is None and torch.equal(lens.cpu(), want_lens)
    assert rel(mel, want_mel) < 3e-3 and rel(pitch, want_pitch) < 2e-3
    seen = {}

    def shift(p, lens_, mean, std):
        seen["args"] = (tuple(p.shape), lens_.cpu().tolist(), float(mean), float(std))
        return p + 0.5

    with torch.no_grad():
        want_mel2, _, _, _, _ = ofp.infer(sd, text, dur_tgt=durs, pitch_tgt=want_pitch + 0.5)
    mel2, _, _, pitch2, _ = m.infer(text.cuda(), dur_tgt=durs.cuda(), pitch_transform=shift)
    assert seen["args"] == ((3, 1, 21), (text != 0).sum(1).tolist(), 218.14, 67.24)     # pitch_std == 0: LJSpeech defaults
    assert rel(pitch2, want_pitch + 0.5) < 2e-3 and rel(mel2, want_mel2) < 3e-3
