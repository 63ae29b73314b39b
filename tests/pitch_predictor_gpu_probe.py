"""Runs the xVAPitch pitch predictor (textenc.RelativePositioningPitchEnergyEncoder) on cuda:0 against the oracle and the
reference recording and prints ONE JSON line of measured errors. Executed in a child process by
tests/test_vits_zz_pitch_predictor_gpu.py: these shapes have not had a hardware run yet, and a fault or a hang of a first
run must not take the rest of the GPU suite with it (the parent kills the child after a timeout)."""
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from textenc_util import fill_pitch, pitch_ref_spec, rel  # noqa: E402

CASES = [(3, 196, 13, [13, 8], 71), (2, 268, 40, [40, 17, 33], 5)]


def case(layers, hidden, T, lens, seed):
    from oracle import vits as ov

    gen = torch.Generator().manual_seed(seed)
    sd = fill_pitch(pitch_ref_spec(layers=layers, hidden=hidden), gen)
    B = len(lens)
    x = torch.randn(B, T, hidden, generator=gen)
    spk = torch.nn.functional.normalize(torch.randn(B, 512, 1, generator=gen), dim=1)
    r = torch.randn(B, 1, T, generator=gen)
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = ov.pitch_predictor(p, x, lens, spk, num_layers=layers)
    (want * r).sum().backward()
    return sd, x, spk, r, want.detach(), {k: v.grad for k, v in p.items() if v.grad is not None}


def run(sd, layers, hidden, x, lens, spk, r, p=0.0):
    from xva_trainer_b200 import textenc

    m = textenc.RelativePositioningPitchEnergyEncoder(1, hidden, 768, 2, layers, 3, p, conditioning_emb_dim=512)
    m.load_state_dict(sd)
    m.train()
    m.zero_grad()
    pred = m(x.cuda(), lens, speaker_emb=spk.cuda())
    m.backward(r.cuda())
    torch.cuda.synchronize()
    return m, pred, m.grads()


def errors(pred, want, got, wgrads, lens, T, floor_frac):
    floor = floor_frac * max(float(v.norm()) for v in wgrads.values())
    per = {k: float((got[k].cpu() - w).norm()) / max(float(w.norm()), floor) for k, w in wgrads.items()}
    num = sum(float((got[k].cpu() - w).norm()) ** 2 for k, w in wgrads.items())
    den = sum(float(w.norm()) ** 2 for w in wgrads.values())
    worst = max(per.items(), key=lambda kv: kv[1])
    pad = max([float(pred[b, :, n:].abs().max()) for b, n in enumerate(lens) if n < T] or [0.0])
    return {"fwd": rel(pred, want), "grad_global": math.sqrt(num / den), "grad_worst": worst[1], "grad_worst_key": worst[0],
            "same_keys": set(got) == set(wgrads), "pad_max": pad}


def main():
    from xva_trainer_b200 import capi, hifigan, ops, textenc

    out = {"exact": [], "product": []}
    for (layers, hidden, T, lens, seed) in CASES:
        sd, x, spk, r, want, wgrads = case(layers, hidden, T, lens, seed)
        orig = ops.gemm_launch
        ops.gemm_launch = lambda args, ref=False: orig(args, True)
        capi.call("xva_set_operand_rounding", 0)
        try:
            _, pred, got = run(sd, layers, hidden, x, lens, spk, r)
        finally:
            capi.call("xva_set_operand_rounding", 1)
            ops.gemm_launch = orig
        out["exact"].append(errors(pred, want, got, wgrads, lens, T, 1e-4))
        _, pred, got = run(sd, layers, hidden, x, lens, spk, r)
        out["product"].append(errors(pred, want, got, wgrads, lens, T, 1e-2))
    # the reference recording, then one AdamW step: the six untrained tensors must be bit-identical afterwards
    g = np.load(os.path.join(HERE, "golden", "vits_pitch_predictor.npz"))
    sd = fill_pitch(pitch_ref_spec(), torch.Generator().manual_seed(71))
    x, spk, r = torch.from_numpy(g["x"]), torch.from_numpy(g["spk"]), torch.from_numpy(g["r"])
    m = textenc.RelativePositioningPitchEnergyEncoder(1, 196, 768, 2, 3, 3, 0.1, conditioning_emb_dim=512)
    m.load_state_dict(sd)
    m.eval()
    out["golden_fwd"] = rel(m(x.cuda(), [13, 8], speaker_emb=spk.cuda()), torch.from_numpy(g["pitch_pred"]))
    opt = hifigan.AdamW([m.flat], lr=1.75e-4, betas=(0.8, 0.99), eps=1e-9, weight_decay=0.01)
    opt.zero_grad()
    m.train()
    m(x.cuda(), [13, 8], speaker_emb=spk.cuda())
    m.backward(r.cuda())
    opt.step()
    torch.cuda.synchronize()
    after = m.state_dict()
    out["keys_in_reference_order"] = list(after) == list(sd)
    out["dead_untouched"] = all(torch.equal(after[k].cpu(), sd[k]) for k in m.dead_keys())
    out["moved"] = sum(1 for k in sd if k not in m.dead_keys() and not torch.equal(after[k].cpu(), sd[k]))
    out["trainable"] = len(sd) - 6
    print("PITCH_PREDICTOR_PROBE " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
