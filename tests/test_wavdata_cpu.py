"""Host side of the HiFi-GAN voice-folder loader (xva-trainer_b200/wavdata.py, mirror of python/hifigan/meldataset.py:268-379):
file list and data multiplier, decode + peak normalisation, random crop / zero pad, epoch batching with the last incomplete
batch dropped. The mel extractors are the device kernels; a stand-in callable takes their place here. CPU only."""
import os

import numpy as np
import pytest
import torch
from scipy.io.wavfile import write as write_wav


@pytest.fixture()
def voice(tmp_path):
    wavs = tmp_path / "wavs"
    wavs.mkdir()
    rng = np.random.RandomState(0)
    lengths = {"a": 30000, "b": 9000, "c": 4000, "d": 8192}
    for name, n in lengths.items():
        x = (rng.randn(n) * 3000).astype(np.int16)
        write_wav(str(wavs / f"{name}.wav"), 22050, x)
    (tmp_path / "metadata.csv").write_text("a|hello there\nwavs/b.wav|general\nc|kenobi\nd.wav|x\nmissing|not on disk\n\n")
    return tmp_path, lengths


def test_filelist_and_data_multiplier(voice):
    from xva_trainer_b200 import wavdata

    root, lengths = voice
    files, not_found, dm = wavdata.get_dataset_filelist(str(root / "metadata.csv"), str(root / "wavs"))
    assert not_found == 1 and dm == round(1000 / (4 - 1))                     # meldataset.py:296-298
    assert len(files) == 4 * dm
    assert sorted(set(os.path.basename(f) for f in files)) == ["a.wav", "b.wav", "c.wav", "d.wav"]
    files2, _, dm2 = wavdata.get_dataset_filelist(str(root / "metadata.csv"), str(root / "wavs"), dm=2)
    assert dm2 == 2 and len(files2) == 8
    (root / "empty.csv").write_text("nothing|here\n")
    with pytest.raises(FileNotFoundError):
        wavdata.get_dataset_filelist(str(root / "empty.csv"), str(root / "wavs"))


def test_items_are_normalised_cropped_or_padded(voice):
    from xva_trainer_b200 import wavdata

    root, lengths = voice
    files = [str(root / "wavs" / f"{n}.wav") for n in lengths]
    ws = wavdata.WavSegments(files, 8192, 22050, shuffle=False)
    for i, n in enumerate(lengths.values()):
        seg = ws.item(i)
        assert seg.dtype == np.float32 and seg.shape == (8192,)
        full = ws.load(files[i])
        assert abs(float(np.abs(full).max()) - 0.95) < 1e-6                   # /32768, peak-normalise, * 0.95 (:347-349)
        if n < 8192:
            assert np.array_equal(seg[:n], full) and not seg[n:].any()        # zero pad on the right (:363-364)
        else:
            # a contiguous window of the normalised waveform
            starts = [s for s in range(n - 8192 + 1) if full[s] == seg[0] and np.array_equal(full[s:s + 8192], seg)]
            assert starts
    # same seed -> same shuffle and the same crops; another seed -> different crops of the long file
    a = wavdata.WavSegments(files, 8192, 22050, seed=5)
    b = wavdata.WavSegments(files, 8192, 22050, seed=5)
    c = wavdata.WavSegments(files, 8192, 22050, seed=6)
    assert a.audio_files == b.audio_files
    ia = a.audio_files.index(files[0])
    assert np.array_equal(a.item(ia), b.item(ia))
    assert not np.array_equal(a.item(ia), c.item(c.audio_files.index(files[0])))
    with pytest.raises(ValueError):
        wavdata.WavSegments(files, 8192, 16000, shuffle=False).item(0)        # wrong sampling rate is an error, not a resample


def test_epoch_batches_layout_and_drop_last(voice):
    from xva_trainer_b200 import wavdata

    root, lengths = voice
    files, _, _ = wavdata.get_dataset_filelist(str(root / "metadata.csv"), str(root / "wavs"), dm=2)   # 8 items
    ws = wavdata.WavSegments(files, 8192, 22050)
    calls = []

    def fake_mel(tag):
        def f(y):
            calls.append((tag, tuple(y.shape)))
            return y.reshape(y.shape[0], 1, -1)[:, :, ::256].repeat(1, 80, 1)      # [B, 80, T] like MelSpectrogram
        return f

    batches = list(ws.batches(3, "cpu", fake_mel("in"), fake_mel("loss")))
    assert len(batches) == 2                                                       # 8 items, batch 3, drop_last (:321)
    for x, y, y_mel in batches:
        assert tuple(y.shape) == (3, 8192) and y.dtype == torch.float32
        assert tuple(x.shape) == (3, 32, 80) and tuple(y_mel.shape) == (3, 32, 80) and x.is_contiguous()
        assert float(y.abs().max()) <= 0.95 + 1e-6
    assert batches[0][1].data_ptr() != batches[1][1].data_ptr()                    # staging buffer is not aliased
    assert not torch.equal(batches[0][1], batches[1][1])
    assert [c[0] for c in calls] == ["in", "loss"] * 2 and all(c[1] == (3, 8192) for c in calls)
    order1, order2 = ws.epoch_order(), ws.epoch_order()
    assert sorted(order1) == list(range(8)) and order1 != order2                   # a fresh permutation per epoch
