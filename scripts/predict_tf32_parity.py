"""Predicts the product-path (tf32 operand) parity of the xVAPitch text encoder, the pitch predictor and the FastPitch step ON THE CPU: the modules'
host code through tests/cabi_emu.py with its tf32 operand model on (xva_gemm truncates operands as the tensor cores do,
producers round to nearest) against the fp32 oracle -- the same comparison tests/test_vits_text_encoder_gpu.py makes on the
device. For the text encoder the prediction can be held against the B200 measurement (profiles/r02_textenc.txt); for the
pitch predictor, whose GPU tests have not run yet, it is the evidence the test bounds are set from.
    python scripts/predict_tf32_parity.py  ->  profiles/r02_tf32_parity_predicted.txt"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cabi_emu  # noqa: E402
from textenc_util import TE_PATCHES, oracle_grads, rel  # noqa: E402


def summary(got, want, floor_frac=1e-2):
    floor = floor_frac * max(float(w.norm()) for w in want.values())
    per = sorted(((float((got[k] - w).norm()) / max(float(w.norm()), floor), k) for k, w in want.items()), reverse=True)
    num = sum(float((got[k] - w).norm()) ** 2 for k, w in want.items())
    den = sum(float(w.norm()) ** 2 for w in want.values())
    return {"grad_global": math.sqrt(num / den), "grad_worst": per[0][0], "grad_worst_key": per[0][1], "grad_median": per[len(per) // 2][0]}


def main():
    import test_vits_text_encoder_gpu as T
    import pitch_predictor_gpu_probe as P

    out = {"text_encoder": [], "pitch_predictor": []}
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    with cabi_emu.installed():
        import xva_trainer_b200
        te = cabi_emu.load_module("textenc", TE_PATCHES)
        sys.modules["xva_trainer_b200.textenc"] = te
        xva_trainer_b200.textenc = te
        cabi_emu.TF32 = True
        try:
            for (Tn, lens, layers, cfg) in T.CASES:
                sd, tokens, lang, seeds = T._case(Tn, lens, layers, cfg, 100 + Tn)
                want_out, want = oracle_grads(sd, tokens, lens, lang, layers, *seeds)
                m = T._build(sd, layers, **cfg)
                o, got, dlang = T._run(m, tokens, lens, lang, *seeds)
                row = {"case": f"T={Tn} lens={lens} layers={layers} cfg={cfg}", "x": rel(o["x"], want_out["x"]),
                       "m_p": rel(o["m_p"], want_out["m_p"]), "dlang": rel(dlang, want["lang"])}
                row.update(summary(got, {k: want[k] for k in sd}))
                out["text_encoder"].append(row)
                print(row, flush=True)
            for (layers, hidden, Tn, lens, seed) in P.CASES:
                sd, x, spk, r, want, wgrads = P.case(layers, hidden, Tn, lens, seed)
                _, pred, got = P.run(sd, layers, hidden, x, lens, spk, r)
                row = {"case": f"layers={layers} channels={hidden}+512 T={Tn} lens={lens}", "pitch_pred": rel(pred, want)}
                row.update(summary(got, wgrads))
                out["pitch_predictor"].append(row)
                print(row, flush=True)
        finally:
            cabi_emu.TF32 = False
    # ---- the FastPitch step at the toy shape of profiles/r02_parity_table.txt (4 x 40 x 150 ragged, seeds of tests/parity_util.py)
    from oracle import fastpitch as ofp
    from parity_util import FWD_NAMES, grad_summary
    from test_cabi_emu_cpu import FP_PATCHES
    out["fastpitch_toy"] = []
    with cabi_emu.installed():
        fp = cabi_emu.load_module("fastpitch", FP_PATCHES)
        for stage in (3, 4):
            x, y = ofp.synthetic_batch(4, 40, 150, seed=11, ragged=True)
            sd = ofp.make_state(1234)
            m = fp.FastPitch(device="cpu")
            m.training_stage = stage
            m.train()
            m.p_drop = 0.0
            crit = fp.FastPitchLoss()
            crit.training_stage = stage
            cabi_emu.TF32 = True
            try:
                m.load_state_dict({k: v.clone() for k, v in sd.items()})       # (the operand copy is rounded at load time)
                o = m(x)
                loss, meta = crit(o, y)
                m.zero_grad()
                m.backward(crit, 1.0)
            finally:
                cabi_emu.TF32 = False
            want = ofp.forward(sd, x, stage)
            wmeta, wgrads = ofp.train_step({k: v.clone() for k, v in sd.items()}, x, y, stage, 1e-3, {}, drop=0.0, training=False)
            gs = grad_summary(m.grads(fp.trainable_keys(stage)), wgrads)
            row = {"case": f"stage {stage}, 4 x 40 x 150 ragged"}
            row.update({n: rel(g_.float(), w_.float()) for n, g_, w_ in zip(FWD_NAMES, o[:8], want[:8])
                        if w_ is not None and w_.dtype != torch.bool and n in ("mel_out", "pitch_pred", "energy_pred")})
            row.update(loss=abs(float(meta["loss"]) - float(wmeta["loss"])) / abs(float(wmeta["loss"])), grad_global=gs["global"],
                       grad_median=gs["median"], grad_worst=gs["worst"], grad_worst_key=gs["worst_key"])
            out["fastpitch_toy"].append(row)
            print(row, flush=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_tf32_parity_predicted.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
