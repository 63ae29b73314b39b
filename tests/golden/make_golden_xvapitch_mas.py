"""Generates tests/golden/xvapitch_mas.npz: the UNMODIFIED xVAPitch monotonic alignment search (python/xvapitch/util.py:14-53,
numpy on the CPU, called once per training step after a device->host copy) on seeded inputs. Groundwork for SURVEY.md
section 8f rank 1: is it the search the engine already has (xva_mas_width1)?  Build container only."""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
np.bool = np.bool_                              # xvapitch/util.py:28 uses the alias numpy >= 1.24 removed
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import  # noqa: E402

_ref_import.install()
from python.xvapitch.util import maximum_path  # noqa: E402


def main():
    out = {}
    for name, (seed, B, Tx, Ty, quant) in {"small": (0, 3, 7, 19, 0), "mid": (1, 4, 40, 150, 0), "ties": (2, 3, 9, 30, 1)}.items():
        g = torch.Generator().manual_seed(seed)
        val = torch.randn(B, Tx, Ty, generator=g) * 2
        if quant:
            val = torch.round(val)              # many exact ties between "stay" and "advance"
        x_lens = torch.randint(max(2, Tx // 2), Tx + 1, (B,), generator=g)
        y_lens = torch.maximum(torch.randint(Ty // 2, Ty + 1, (B,), generator=g), x_lens)
        x_lens[0], y_lens[0] = Tx, Ty
        mask = ((torch.arange(Tx)[None, :, None] < x_lens[:, None, None])
                & (torch.arange(Ty)[None, None, :] < y_lens[:, None, None])).float()
        out[f"{name}/value"], out[f"{name}/x_lens"], out[f"{name}/y_lens"] = val.numpy(), x_lens.numpy(), y_lens.numpy()
        out[f"{name}/path"] = maximum_path(val, mask).numpy()
    np.savez_compressed(os.path.join(HERE, "xvapitch_mas.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("path")})


if __name__ == "__main__":
    main()
