"""Generates tests/golden/infer.npz by running the UNMODIFIED reference FastPitch.infer (model.py:426-482, imported from
/root/reference) on seeded token batches. Build container only:   python tests/golden/make_golden_infer.py

Weights = oracle.fastpitch.make_state(1234) with the duration predictor's output bias raised by 1.5 (so the random-init
model predicts a few frames per token instead of ~0). Cases: "free" (everything predicted, pace 1), "pace" (pace 0.85),
"forced" (durations and pitch targets supplied: the path the UI's editor sliders drive; energy_tgt cannot be recorded --
the reference returns an unbound `energy_pred` when it is given, model.py:462-467,482).
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import _ref_import  # noqa: E402

_ref_import.install()
from oracle import fastpitch as ofp  # noqa: E402  (make_state only: shared seeded weights)


def infer_state():
    sd = ofp.make_state(1234)
    sd["duration_predictor.fc.bias"] = sd["duration_predictor.fc.bias"] + 1.5
    return sd


def main():
    from python.fastpitch1_1.fastpitch.model import FastPitch

    torch.manual_seed(0)
    m = FastPitch()
    m.load_state_dict(infer_state(), strict=True)
    m.eval()
    g = torch.Generator().manual_seed(17)
    B, Tt = 3, 21
    text = torch.randint(1, ofp.N_SYMBOLS, (B, Tt), generator=g)
    text[1, 15:] = 0
    text[2, 9:] = 0
    out = {"text": text.numpy()}
    durs = torch.randint(1, 6, (B, Tt), generator=g).float() * (text != 0)
    pitch = torch.randn(B, 1, Tt, generator=g) * (text != 0).unsqueeze(1)
    out["forced/dur_tgt"], out["forced/pitch_tgt"] = durs.numpy(), pitch.numpy()
    cases = {"free": dict(pace=1.0), "pace": dict(pace=0.85),
             "forced": dict(pace=1.0, dur_tgt=durs, pitch_tgt=pitch)}
    with torch.no_grad():
        for name, kw in cases.items():
            mel, dec_lens, dur_pred, pitch_pred, energy_pred = m.infer(text, **kw)
            out[f"{name}/pace"] = np.float64(kw["pace"])
            out[f"{name}/mel"], out[f"{name}/dec_lens"] = mel.numpy(), dec_lens.numpy()
            out[f"{name}/dur_pred"], out[f"{name}/pitch_pred"] = dur_pred.numpy(), pitch_pred.numpy()
            out[f"{name}/energy_pred"] = energy_pred.numpy()
            print(name, tuple(mel.shape), dec_lens.tolist())
    np.savez_compressed(os.path.join(HERE, "infer.npz"), **out)


if __name__ == "__main__":
    main()
