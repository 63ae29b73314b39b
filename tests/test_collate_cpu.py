"""xva-trainer_b200/data.py (TTSCollate, batch_to_gpu) against the outputs of the unmodified reference collate and
batch_to_gpu recorded by tests/golden/make_golden_collate.py -- bit for bit, dtypes included, for training stages 1-4,
including the reference's truncation of pitch / energy / durations toward zero. CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["text_padded", "input_lengths", "mel_padded", "output_lengths", "len_x", "pitch_padded", "energy_padded", "speaker",
         "attn_prior_padded", "durs_padded", "max_inp_lengths", "max_mel_lengths", "audiopaths"]


def make_items(stage, seed=0):
    """The items of tests/golden/make_golden_collate.py (same generator, same order of draws)."""
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


def _same(got, g, key):
    if key + "/none" in g.files:
        assert got is None, key
    elif torch.is_tensor(got):
        want = g[key]
        assert str(got.dtype) == str(g[key + "/dtype"]), (key, got.dtype, g[key + "/dtype"])
        assert tuple(got.shape) == tuple(want.shape), (key, got.shape, want.shape)
        assert np.array_equal(got.numpy(), want), key
    else:
        assert list(got) == list(g[key]), key


@pytest.mark.parametrize("stage", [1, 2, 3, 4])
def test_collate_and_batch_to_gpu_match_the_reference(stage):
    from xva_trainer_b200 import data

    g = np.load(os.path.join(GOLD, "collate.npz"))
    batch = data.TTSCollate(training_stage=stage)(make_items(stage))
    assert len(batch) == 13
    for n, v in zip(NAMES, batch):
        _same(v, g, f"s{stage}/collate/{n}")
    x, y, frames = data.batch_to_gpu(batch, training_stage=stage, device="cpu")
    assert len(x) == 12 and len(y) == 3 and int(frames) == int(g[f"s{stage}/len_x"])
    for i, v in enumerate(x):
        _same(v, g, f"s{stage}/x/{i}")
    for i, v in enumerate(y):
        assert np.array_equal(v.numpy(), g[f"s{stage}/y/{i}"])


def test_reference_truncation_is_reproduced_and_can_be_switched_off():
    from xva_trainer_b200 import data

    items = make_items(3)
    ref_like = data.TTSCollate(training_stage=3)(items)
    exact = data.TTSCollate(training_stage=3, exact_targets=True)(items)
    order = [1, 3, 0, 2]                                             # text lengths 8, 8, 5, 3
    pitch0 = items[order[0]][3]
    assert ref_like[5].dtype == torch.int64 and exact[5].dtype == torch.float32
    n = pitch0.shape[1]
    assert np.array_equal(ref_like[5][0, :, :n].numpy(), np.trunc(pitch0).astype(np.int64))     # toward zero, not floor
    assert np.array_equal(exact[5][0, :, :n].numpy(), pitch0)
    assert (np.trunc(pitch0) != np.floor(pitch0)).any()              # the case that tells trunc from floor is present
    assert np.array_equal(exact[6][0, :n].numpy(), items[order[0]][4])
    x, _, _ = data.batch_to_gpu(exact, training_stage=3, device="cpu")
    assert x[4].dtype == torch.float32 and x[8].dtype == torch.float32


def test_collated_batch_feeds_the_engine_layout():
    """The 12-list has the layout FastPitch.forward unpacks (model.py:325-330): padding is trailing, lengths sorted."""
    from xva_trainer_b200 import data

    for stage in (1, 3):
        x, y, _ = data.batch_to_gpu(data.TTSCollate(training_stage=stage)(make_items(stage)), stage, "cpu")
        text, in_lens, mel, mel_lens = x[:4]
        assert bool((in_lens[:-1] >= in_lens[1:]).all())
        for b in range(text.shape[0]):
            assert bool((text[b, :in_lens[b]] != 0).all()) and not bool(text[b, in_lens[b]:].any())
            assert not bool(mel[b, :, mel_lens[b]:].any())
        assert float(x[9][0]) == text.shape[1] and float(x[10][0]) == mel.shape[2]
        if stage == 1:
            assert x[7].shape == (4, mel.shape[2], text.shape[1]) and x[8] is None
        else:
            assert torch.equal(x[8].sum(1).long(), mel_lens)
