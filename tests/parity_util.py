"""Measurements shared by the parity tests (-m gpu) and scripts/parity_table.py: one training step of the PRODUCT path
(tcgen05 tf32 tap-GEMMs through the C ABI) against the CPU oracle on identical weights and inputs, dropout off.
Every function returns plain dicts of relative L2 errors; the tests put bounds on them, the script prints them
(profiles/r02_parity_table.txt). Test infrastructure: this is the only place besides tests/ that imports oracle/."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import fastpitch as ofp  # noqa: E402
from oracle import hifigan as ohg  # noqa: E402

FWD_NAMES = ["mel_out", "dec_mask", "dur_pred", "log_dur_pred", "pitch_pred", "pitch_tgt", "energy_pred", "energy_tgt"]
LOSS_NAMES = ("loss", "mel_loss", "duration_predictor_loss", "pitch_loss", "energy_loss")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def grad_summary(got, want):
    """got / want: {key: tensor}. -> global relative error of the concatenated vector, median and worst per tensor."""
    num = den = 0.0
    per = {}
    for k, w in want.items():
        if w is None or float(w.norm()) == 0.0:
            continue
        g = got[k].detach().cpu().double()
        w = w.detach().double()
        num += float((g - w).pow(2).sum())
        den += float(w.pow(2).sum())
        per[k] = float((g - w).norm() / w.norm())
    srt = sorted(per.values())
    worst = max(per.items(), key=lambda kv: kv[1])
    return {"global": (num / den) ** 0.5, "median": srt[len(srt) // 2], "worst": worst[1], "worst_key": worst[0],
            "n": len(per), "per": per}


def cuda_batch(x, y):
    cx = [t.cuda() if torch.is_tensor(t) else t for t in x]
    cy = [t.cuda() if torch.is_tensor(t) else t for t in y]
    return cx, cy


def fastpitch_model(sd, stage, training=True):
    from xva_trainer_b200 import fastpitch as fp

    m = fp.FastPitch(device="cuda:0")
    m.load_state_dict(sd)
    m.training_stage = stage
    m.train(training)
    m.p_drop = 0.0
    return fp, m


def fastpitch_step_errors(stage, B, Tt, Tm, ragged, seed=11, state_seed=1234, threads=None):
    """Forward tensors, losses and every parameter gradient of one step at the given shape."""
    if threads:
        torch.set_num_threads(threads)
    x, y = ofp.synthetic_batch(B, Tt, Tm, seed=seed, ragged=ragged)
    sd = ofp.make_state(state_seed)
    fp, m = fastpitch_model({k: v.clone() for k, v in sd.items()}, stage)
    crit = fp.FastPitchLoss()
    crit.training_stage = stage
    cx, cy = cuda_batch(x, y)
    out = m(cx)
    loss, meta = crit(out, cy)
    m.zero_grad()
    m.backward(crit, 1.0)
    torch.cuda.synchronize()
    res = {"fwd": {}, "loss": {}, "exact": {}}
    want = ofp.forward(sd, x, stage)
    for n, g_, w_ in zip(FWD_NAMES, out[:8], want[:8]):
        if w_ is None:
            assert g_ is None, n
            continue
        assert g_.shape == w_.shape, (n, g_.shape, w_.shape)
        if w_.dtype == torch.bool:
            res["exact"][n] = bool(torch.equal(g_.cpu(), w_))
        else:
            res["fwd"][n] = rel(g_, w_)
    wmeta, wgrads = ofp.train_step({k: v.clone() for k, v in sd.items()}, x, y, stage, 1e-3, {}, drop=0.0, training=False)
    for k in LOSS_NAMES:
        a, b = float(meta[k]), float(wmeta[k])
        res["loss"][k] = abs(a - b) / abs(b) if b != 0 else abs(a)
    keys = fp.trainable_keys(stage)
    res["grad"] = grad_summary(m.grads(keys), wgrads)
    res["frozen_zero"] = all(float(m.grads([k])[k].abs().max()) == 0.0 for k in keys if wgrads[k] is None)
    return res


def fastpitch_trajectory(stage, B, Tt, Tm, ragged, steps=5, seed=11):
    """`steps` consecutive optimizer steps (forward, loss, backward, clip, LAMB at the noam learning rate of iteration
    50 000 + i) on both sides from the same state: relative loss difference per step."""
    x, y = ofp.synthetic_batch(B, Tt, Tm, seed=seed, ragged=ragged)
    sd = ofp.make_state(1234)
    fp, m = fastpitch_model({k: v.clone() for k, v in sd.items()}, stage)
    crit = fp.FastPitchLoss()
    crit.training_stage = stage
    opt = fp.Lamb(m, lr=0.1, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
    cx, cy = cuda_batch(x, y)
    osd = {k: v.clone() for k, v in sd.items()}
    ostate = {}
    out = []
    for i in range(steps):
        it = 50000 + i
        fp.adjust_learning_rate(it, opt, 0.1, 1000)
        m.zero_grad()
        o = m(cx)
        loss, meta = crit(o, cy)
        m.backward(crit, 1.0)
        opt.step()
        wmeta, _ = ofp.train_step(osd, x, y, stage, ofp.noam_lr(it), ostate, drop=0.0, training=False)
        a, b = float(loss), float(wmeta["loss"])
        out.append({"step": i, "loss": a, "oracle": b, "rel": abs(a - b) / abs(b)})
    after = m.state_dict()
    werr = grad_summary({k: after[k] for k in fp.trainable_keys(stage)}, {k: osd[k] for k in fp.trainable_keys(stage)})
    return out, werr


class _H(dict):
    __getattr__ = dict.__getitem__


def hifigan_config():
    return _H(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
              resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
              learning_rate=2e-4, adam_b1=0.8, adam_b2=0.99, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256,
              win_size=1024, fmin=0, fmax=8000, fmax_for_loss=None)


def hifigan_models(seed_g=5, scale=0.7):
    from xva_trainer_b200 import hifigan as hg

    h = hifigan_config()
    sd_g = ohg.make_generator_state(seed_g, scale=scale)
    sd_p = ohg.make_disc_state(ohg.mpd_spec(), 21)
    sd_s = ohg.make_disc_state(ohg.msd_spec(), 22)
    G = hg.Generator(h, device="cuda:0")
    G.load_state_dict({k: v for k, v in sd_g.items()})
    G.train()
    mpd = hg.MultiPeriodDiscriminator(device="cuda:0")
    mpd.load_state_dict(sd_p)
    mpd.train()
    msd = hg.MultiScaleDiscriminator(device="cuda:0")
    msd.load_state_dict(sd_s)
    msd.train()
    return hg, h, (G, mpd, msd), (sd_g, sd_p, sd_s)


def hifigan_step_errors(B, frames, steps=1, seed=3):
    """`steps` consecutive HiFiTrainer.iteration bodies on both sides: loss terms per step; D-step and G-step parameter
    gradients of the FIRST step; every weight after the last step."""
    hg, h, (G, mpd, msd), (sd_g, sd_p, sd_s) = hifigan_models()
    step = hg.HiFiGANStep(G, mpd, msd, h)
    og, op, os_ = ({k: v.clone() for k, v in d.items()} for d in (sd_g, sd_p, sd_s))
    ostate = {}
    res = {"loss": [], "steps": steps}
    for s in range(steps):
        x, y, y_mel = ohg.synthetic_batch(B, frames, seed=seed + s)
        losses = step.step(x.cuda(), y.cuda(), y_mel.cuda())
        torch.cuda.synchronize()
        detail = {}
        want, ggrads = ohg.train_step(og, op, os_, x, y, y_mel, ostate, detail=detail)
        res["loss"].append({k: abs(float(losses[k]) - float(want[k])) / abs(float(want[k]))
                            for k in ("loss_disc_all", "loss_mel", "loss_fm", "loss_gen", "loss_gen_all")})
        if s == 0:
            res["y_g_hat"] = None
            got_d = {("mpd", k): p.grad for k, p in mpd.named_parameters()}
            got_d.update({("msd", k): p.grad for k, p in msd.named_parameters()})
            want_d = {("mpd", k): v for k, v in detail["dgrad"]["mpd"].items()}
            want_d.update({("msd", k): v for k, v in detail["dgrad"]["msd"].items()})
            res["dgrad"] = grad_summary(got_d, want_d)
            res["ggrad"] = grad_summary({k: p.grad for k, p in G.named_parameters()}, ggrads)
    res["weights"] = {}
    for name, model, ref in (("G", G, og), ("mpd", mpd, op), ("msd", msd, os_)):
        after = model.state_dict()
        res["weights"][name] = grad_summary({k: after[k] for k in ref}, ref)
    return res
