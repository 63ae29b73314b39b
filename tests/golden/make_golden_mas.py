"""Generates tests/golden/mas.npz by running the UNMODIFIED reference b_mas (numba,
/root/reference/python/fastpitch1_1/fastpitch/alignment.py:110-118) on seeded soft alignments. Build container only.

    python tests/golden/make_golden_mas.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import  # noqa: E402

_ref_import.install()
from python.fastpitch1_1.fastpitch.alignment import b_mas  # noqa: E402


def case(seed, B, Tm, Tt, sharp):
    r = np.random.RandomState(seed)
    in_lens = r.randint(max(2, Tt // 2), Tt + 1, size=B)
    out_lens = np.maximum(r.randint(max(2, Tm // 2), Tm + 1, size=B), in_lens)   # at least one frame per token
    in_lens[0], out_lens[0] = Tt, Tm
    # a noisy diagonal: softmax over text of -sharp * (j - i * Tt / Tm)^2 + noise, like a converging aligner
    i = np.arange(Tm)[:, None] / Tm
    j = np.arange(Tt)[None, :] / Tt
    logits = -sharp * (j - i) ** 2 * Tt + r.randn(B, 1, Tm, Tt) * 1.5
    p = np.exp(logits - logits.max(-1, keepdims=True))
    p = (p / p.sum(-1, keepdims=True)).astype(np.float32)
    return p, in_lens.astype(np.int64), out_lens.astype(np.int64)


def main():
    out = {}
    for name, args in {"small": (1, 3, 24, 9, 4.0), "mid": (2, 4, 160, 40, 2.0), "flat": (3, 2, 60, 20, 0.0),
                       "short_mel": (4, 2, 12, 12, 3.0)}.items():
        p, il, ol = case(*args)
        if name == "short_mel":
            ol[1] = il[1] - 3 if il[1] > 4 else ol[1]           # fewer frames than tokens: the double mark in row 0
        out[f"{name}/attn"], out[f"{name}/in_lens"], out[f"{name}/out_lens"] = p, il, ol
        out[f"{name}/hard"] = b_mas(p, il, ol, width=1)
    np.savez_compressed(os.path.join(HERE, "mas.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("hard")})


if __name__ == "__main__":
    main()
