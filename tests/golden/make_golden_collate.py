"""Generates tests/golden/collate.npz by running the UNMODIFIED reference TTSCollate.__call__ and batch_to_gpu
(fastpitch/data_function.py:560-741, imported from /root/reference) on seeded items shaped like TTSDataset.__getitem__'s
(:300-352: text LongTensor, mel / pitch / energy / prior / durs numpy arrays). Build container only:

    python tests/golden/make_golden_collate.py

Import shims (fixture tooling only): `parselmouth` and the text front end (`common.text.text_processing`, which pulls
inflect / unidecode / CMUdict) are stubbed -- neither is used by the collate.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import  # noqa: E402

_ref_import.install()
sys.modules["parselmouth"] = types.ModuleType("parselmouth")
import python.fastpitch1_1.common  # noqa: E402,F401

_tx = types.ModuleType("python.fastpitch1_1.common.text")
_tx.__path__ = []
_tp = types.ModuleType("python.fastpitch1_1.common.text.text_processing")
_tp.TextProcessing = object
sys.modules["python.fastpitch1_1.common.text"] = _tx
sys.modules["python.fastpitch1_1.common.text.text_processing"] = _tp
from python.fastpitch1_1.fastpitch.data_function import TTSCollate, batch_to_gpu  # noqa: E402


def make_items(stage, seed=0):
    """Items as TTSDataset.__getitem__ returns them for a training stage."""
    r = np.random.RandomState(seed)
    items = []
    for k, (n_txt, n_mel) in enumerate(((5, 20), (8, 31), (3, 12), (8, 25))):
        text = torch.from_numpy(r.randint(1, 148, size=n_txt)).long()
        mel = r.randn(80, n_mel).astype(np.float32)
        if stage in (1, 2):
            pitch, energy = [0], [0]
        else:
            pitch = (r.randn(1, n_mel) * 2.5).astype(np.float32) * (r.rand(1, n_mel) > 0.3)
            energy = np.linalg.norm(mel, ord=2, axis=0)
        prior = r.rand(n_mel, n_txt).astype(np.float32) if stage == 1 else None
        durs = None
        if stage != 1:
            durs = np.ones(n_txt, dtype=np.float32)
            durs[: n_mel % n_txt] += 1
            durs[0] += n_mel - durs.sum()
        items.append((text, mel, n_txt, pitch, energy, None, prior, durs, f"D:/voice/wavs/{k:04d}.wav"))
    return items


def main():
    out = {}
    for stage in (1, 2, 3, 4):
        c = TTSCollate()
        c.training_stage = stage
        batch = c(make_items(stage))
        names = ["text_padded", "input_lengths", "mel_padded", "output_lengths", "len_x", "pitch_padded", "energy_padded",
                 "speaker", "attn_prior_padded", "durs_padded", "max_inp_lengths", "max_mel_lengths", "audiopaths"]
        for n, v in zip(names, batch):
            if torch.is_tensor(v):
                out[f"s{stage}/collate/{n}"] = v.numpy()
                out[f"s{stage}/collate/{n}/dtype"] = np.array(str(v.dtype))
            elif v is None:
                out[f"s{stage}/collate/{n}/none"] = np.array(1)
            else:
                out[f"s{stage}/collate/{n}"] = np.array(v)
        x, y, len_x = batch_to_gpu(batch, training_stage=stage, device=0)
        for i, v in enumerate(x):
            if torch.is_tensor(v):
                out[f"s{stage}/x/{i}"] = v.numpy()
                out[f"s{stage}/x/{i}/dtype"] = np.array(str(v.dtype))
            elif v is None:
                out[f"s{stage}/x/{i}/none"] = np.array(1)
            else:
                out[f"s{stage}/x/{i}"] = np.array(v)
        for i, v in enumerate(y):
            out[f"s{stage}/y/{i}"] = v.numpy()
        out[f"s{stage}/len_x"] = np.int64(int(len_x))
        print(f"stage {stage}: ok, frames {int(len_x)}")
    np.savez_compressed(os.path.join(HERE, "collate.npz"), **out)


if __name__ == "__main__":
    main()
