"""Generates tests/golden/stage1.npz by running the UNMODIFIED reference stage-1 aligner (imported from /root/reference)
on seeded inputs. Build container only:   python tests/golden/make_golden_stage1.py

Recorded (fp32 CPU, weights = oracle.fastpitch.make_state(1234) loaded into the reference FastPitch):
  the reference's ConvAttention outputs (attn_soft, attn_logprob) on the token embedding / mel target / beta-binomial
  prior of a ragged B=3 batch, the hard alignment of its numba b_mas and the durations, AttentionCTCLoss and
  AttentionBinarizationLoss values, and -- for loss = ctc + 0.5 * binarization, what FastPitchTrainer.iteration
  back-propagates in stage 1 (xva_train.py:790-798) -- per-parameter gradient norms + 16 sampled values, plus the list
  of parameters whose grad stays None.
The glue between the reference pieces is the body of FastPitch.get_alignment_durations (model.py:298-323) minus its
`.to(attn.get_device())`, which cannot run without a GPU (get_device() is -1 on the CPU).
A second case ("sharp") scales the mel target so the scores are far from uniform, and a third ("short") has an
utterance with fewer frames than tokens (impossible alignment -> zero_infinity).
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

from make_golden import summarize  # noqa: E402
from oracle import fastpitch as ofp  # noqa: E402  (make_state / synthetic_batch only: shared seeded inputs)


def run_case(name, x, out, kl_weight=0.5):
    from python.fastpitch1_1.fastpitch.model import FastPitch
    from python.fastpitch1_1.fastpitch.alignment import b_mas
    from python.fastpitch1_1.fastpitch.attn_loss_function import AttentionBinarizationLoss, AttentionCTCLoss
    from python.fastpitch1_1.fastpitch.transformer import mask_from_lens

    (inputs, input_lens, mel_tgt, mel_lens, _, _, _, attn_prior, _, max_inp_lengths, _, _) = x
    torch.manual_seed(0)
    m = FastPitch()
    m.load_state_dict(ofp.make_state(1234), strict=True)
    m.eval()
    m.zero_grad()
    text_emb = m.encoder.word_emb(inputs)
    attn_mask = mask_from_lens(input_lens, max_inp_lengths[0])[..., None] == 0
    attn_soft, attn_logprob = m.attention(mel_tgt, text_emb.permute(0, 2, 1), mel_lens, attn_mask, key_lens=input_lens,
                                          keys_encoded=None, attn_prior=attn_prior)
    attn_hard = torch.from_numpy(b_mas(attn_soft.data.cpu().numpy(), input_lens.cpu().numpy(), mel_lens.cpu().numpy(),
                                       width=1))
    durs = attn_hard.sum(2)[:, 0, :]
    ctc = AttentionCTCLoss()(attn_logprob, input_lens, mel_lens)
    kl = AttentionBinarizationLoss()(attn_hard, attn_soft)
    (ctc * 1.0 + kl_weight * kl).backward()
    for k, v in (("text", inputs), ("in_lens", input_lens), ("mel", mel_tgt), ("mel_lens", mel_lens), ("prior", attn_prior)):
        out[f"{name}/in/{k}"] = v.numpy()
    out[f"{name}/attn_soft"] = attn_soft.detach().numpy()
    out[f"{name}/attn_logprob"] = attn_logprob.detach().numpy()
    out[f"{name}/attn_hard"] = attn_hard.numpy()
    out[f"{name}/durs"] = durs.numpy()
    out[f"{name}/ctc"], out[f"{name}/kl"] = np.float64(ctc.item()), np.float64(kl.item())
    out[f"{name}/kl_weight"] = np.float64(kl_weight)
    summarize(f"{name}/grad", {k: p.grad for k, p in m.named_parameters()}, out)
    out[f"{name}/no_grad"] = np.array(sorted(k for k, p in m.named_parameters() if p.grad is None))
    print(f"{name}: ctc {ctc.item():.6f} kl {kl.item():.6f}, {sum(p.grad is not None for p in m.parameters())} tensors with grad")


def main():
    out = {}
    x, _ = ofp.synthetic_batch(3, 14, 50, seed=7, ragged=True, prior=True)
    run_case("small", x, out)
    x, _ = ofp.synthetic_batch(3, 20, 64, seed=9, ragged=True, prior=True)
    x[2] = x[2] * 12.0                                           # sharper scores: |q - k|^2 * 0.0005 of order 1..10
    run_case("sharp", x, out)
    x, _ = ofp.synthetic_batch(2, 12, 30, seed=11, ragged=True, prior=True)
    x[3] = x[3].clone()
    x[3][1] = int(x[1][1]) - 2                                   # fewer frames than tokens: CTC cost is infinite -> 0
    run_case("short", x, out)
    np.savez_compressed(os.path.join(HERE, "stage1.npz"), **out)


if __name__ == "__main__":
    main()
