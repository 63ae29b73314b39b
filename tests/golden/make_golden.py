"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on seeded inputs.

Run in the build container only:   python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read the committed .npz files instead.

What is recorded (all fp32 CPU, dropout off via model.eval(), weights = oracle.fastpitch.make_state(1234)):
  fastpitch_small.npz   B=3 ragged batch (Tt=14, Tm=50): for stages 2/3/4 the reference's forward outputs, its
                        FastPitchLoss terms, per-parameter gradient norms + 16 sampled gradient values, and the
                        parameters after one reference Lamb.step (norm + samples)
  regulate_len.npz      regulate_len / average_pitch of the reference on durations with zeros, pace != 1, truncation
The only deviation from "unmodified": FastPitchLoss builds its zero placeholders on torch.device('cuda:N')
(loss_function.py:92-129); while it runs, torch.device is redirected to the CPU device.
"""
import hashlib
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

from oracle import fastpitch as ofp  # noqa: E402  (only for make_state / synthetic_batch: shared seeded inputs)


def sample_idx(key, numel, n=16):
    h = int(hashlib.sha256(key.encode()).hexdigest()[:8], 16)
    g = np.random.RandomState(h)
    return g.randint(0, numel, size=n)


def summarize(prefix, named, out):
    for k, v in named.items():
        if v is None:
            continue
        v = v.detach().reshape(-1).double()
        out[f"{prefix}/{k}/norm"] = np.float64(v.norm().item())
        out[f"{prefix}/{k}/samples"] = v[torch.from_numpy(sample_idx(k, v.numel()))].numpy()


class _CpuDevice:
    """Redirect torch.device('cuda:N') to CPU while the reference loss runs (see module docstring)."""

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


def fastpitch_small():
    from python.fastpitch1_1.fastpitch.model import FastPitch
    from python.fastpitch1_1.fastpitch.loss_function import FastPitchLoss
    from python.fastpitch1_1.lamb import Lamb

    out = {}
    sd0 = ofp.make_state(1234)
    x, y = ofp.synthetic_batch(3, 14, 50, seed=7, ragged=True)
    out["in/text"], out["in/in_lens"] = x[0].numpy(), x[1].numpy()
    out["in/mel"], out["in/mel_lens"] = x[2].numpy(), x[3].numpy()
    out["in/pitch"], out["in/energy"], out["in/durs"] = x[4].numpy(), x[5].numpy(), x[8].numpy()
    for stage in (2, 3, 4):
        torch.manual_seed(0)
        m = FastPitch()
        missing = m.load_state_dict(sd0, strict=True)
        m.training_stage = stage
        m.eval()
        trainable = set(ofp.trainable_keys(stage))
        for k, p in m.named_parameters():
            p.requires_grad = k in trainable
        opt = Lamb([p for p in m.parameters()], lr=ofp.noam_lr(50000), betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
        crit = FastPitchLoss(dur_predictor_loss_scale=0.1, pitch_predictor_loss_scale=0.1, attn_loss_scale=1.0, gpus=[0])
        yp = m([t.clone() if torch.is_tensor(t) else t for t in x])
        with _CpuDevice():
            loss, meta, _ = crit(yp, [t.clone() if torch.is_tensor(t) else t for t in y], training_stage=stage)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1000)
        names = ["mel_out", "dec_mask", "dur_pred", "log_dur_pred", "pitch_pred", "pitch_tgt", "energy_pred", "energy_tgt"]
        for n, v in zip(names, yp[:8]):
            if v is not None:
                out[f"s{stage}/out/{n}"] = v.detach().float().numpy()
        for k in ("loss", "mel_loss", "duration_predictor_loss", "pitch_loss", "energy_loss"):
            if k in meta:
                out[f"s{stage}/loss/{k}"] = np.float64(meta[k].item())
        summarize(f"s{stage}/grad", {k: p.grad for k, p in m.named_parameters()}, out)
        opt.step()
        summarize(f"s{stage}/after_lamb", {k: p for k, p in m.named_parameters() if p.grad is not None}, out)
        print(f"stage {stage}: loss {loss.item():.6f}")
    np.savez_compressed(os.path.join(HERE, "fastpitch_small.npz"), **out)
    import json
    json.dump([[k, list(v.shape)] for k, v in FastPitch().state_dict().items()],
              open(os.path.join(HERE, "fastpitch_keys.json"), "w"))


def regulate_cases():
    from python.fastpitch1_1.fastpitch.model import regulate_len, average_pitch

    out = {}
    g = torch.Generator().manual_seed(11)
    cases = {"plain": (1.0, None), "pace": (0.87, None), "trunc": (1.0, 37), "pace_trunc": (1.3, 41)}
    durs = torch.randint(0, 7, (4, 19), generator=g).float()
    durs[1, 12:] = 0          # padded tokens
    durs[2, :3] = 0           # leading zero-length tokens
    durs[3] = torch.rand(19, generator=g) * 5   # fractional (predicted) durations
    enc = torch.randn(4, 19, 8, generator=g)
    out["durs"], out["enc"] = durs.numpy(), enc.numpy()
    for name, (pace, mx) in cases.items():
        rep, lens = regulate_len(durs, enc, pace, mx)
        out[f"{name}/enc_rep"], out[f"{name}/dec_lens"] = rep.numpy(), lens.numpy()
        out[f"{name}/pace"], out[f"{name}/mel_max_len"] = np.float64(pace), np.int64(-1 if mx is None else mx)
    idurs = torch.randint(0, 6, (3, 11), generator=g).float()
    Tm = int(idurs.sum(1).max())
    pitch = torch.randn(3, 1, Tm, generator=g) * (torch.rand(3, 1, Tm, generator=g) > 0.3)
    out["avg/durs"], out["avg/pitch"] = idurs.numpy(), pitch.numpy()
    out["avg/out"] = average_pitch(pitch, idurs).numpy()
    np.savez_compressed(os.path.join(HERE, "regulate_len.npz"), **out)


if __name__ == "__main__":
    regulate_cases()
    fastpitch_small()
    print("golden fixtures written to", HERE)
