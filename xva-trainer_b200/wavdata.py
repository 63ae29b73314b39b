"""HiFi-GAN training data from a voice folder (``metadata.csv`` + ``wavs/*.wav``, 22 050 Hz mono int16): the host side of
SURVEY.md section 8f rank 3 for the vocoder. Mirrors python/hifigan/meldataset.py:

    get_dataset_filelist   meldataset.py:268-308   file list from metadata.csv, repeated `dm` times (~1000 items / epoch)
    MelDataset             meldataset.py:311-379   decode, /32768, peak-normalise * 0.95, random 8192-sample crop (or pad)

What differs by design: the reference computes two STFT mel spectrograms PER ITEM on the CPU inside ``__getitem__`` (on the
training thread: ``num_workers=0``, hifigan/xva_train.py:321) and rebuilds the librosa filterbank for every call. Here a
batch of cropped waveforms is assembled in one pinned host buffer, copied once, and both mels (fmax 8000 for the generator
input, fmax None for the loss) are ONE batched call each of the mel kernels on the device (hifigan.MelSpectrogram) -- the
same operator the training step back-propagates through. The product has no CPU path for the mel: ``batches`` needs the two
extractor callables.
"""
import os
import random

import numpy as np
import torch

MAX_WAV_VALUE = 32768.0


def load_wav(full_path):
    """meldataset.py:16-18 (scipy.io.wavfile.read) -> (int16 / float ndarray, sampling_rate)."""
    from scipy.io.wavfile import read

    sampling_rate, data = read(full_path)
    return data, sampling_rate


def peak_normalize(audio):
    """librosa.util.normalize (norm=inf, axis 0) as meldataset.py:349 uses it: divide by max |x|; an all-zero (below tiny)
    signal is returned unchanged."""
    audio = np.asarray(audio, dtype=np.float32)
    peak = float(np.abs(audio).max()) if audio.size else 0.0
    if peak < np.finfo(np.float32).tiny:
        return audio
    return audio / peak


def get_dataset_filelist(input_training_file, input_wavs_dir, dm=None):
    """meldataset.py:268-308 -> (file list repeated dm times and shuffled per repetition, number not found, dm).
    Lines of metadata.csv are ``fname|text``; only the file name is used. dm defaults to round(1000 / n_files), >= 1."""
    files, not_found = [], 0
    with open(input_training_file, "r", encoding="utf-8") as fi:
        for line in fi.read().split("\n"):
            if not line:
                continue
            fname = line.split("|")[0].split("/")[-1]
            if not fname.strip():
                continue
            if ".wav" not in fname:
                fname += ".wav"
            path = f"{input_wavs_dir}/{fname}"
            if os.path.exists(path):
                files.append(path)
            else:
                not_found += 1
    if not files:
        raise FileNotFoundError(f"no wav file of {input_training_file} exists under {input_wavs_dir}")
    if dm is None:
        dm = max(1, round(1000 / max(1, len(files) - not_found)))
    total = []
    for _ in range(dm):
        random.shuffle(files)
        total += files
    return total, int(not_found), dm


class WavSegments:
    """MelDataset minus its per-item spectrograms: item(i) -> float32 [segment_size] waveform in [-0.95, 0.95]."""

    MAX_CACHE_ITEMS = 5000                                   # meldataset.py:337

    def __init__(self, training_files, segment_size, sampling_rate=22050, split=True, shuffle=True, seed=1234):
        self.audio_files = list(training_files)
        self.rng = random.Random(seed)                       # meldataset.py:318 seeds the global generator with 1234
        if shuffle:
            self.rng.shuffle(self.audio_files)
        self.segment_size, self.sampling_rate, self.split = int(segment_size), int(sampling_rate), split
        self.audio_cache = {}

    def __len__(self):
        return len(self.audio_files)

    def load(self, filename):
        filename = filename if filename.endswith(".wav") else filename + ".wav"
        audio = self.audio_cache.get(filename)
        if audio is None:
            data, sr = load_wav(filename)
            if sr != self.sampling_rate:
                raise ValueError(f"{filename}: {sr} Hz, the vocoder is trained on {self.sampling_rate} Hz audio")
            if data.ndim > 1:
                data = data[:, 0]
            audio = peak_normalize(np.asarray(data, dtype=np.float32) / MAX_WAV_VALUE) * 0.95   # meldataset.py:347-349
            if len(self.audio_cache) < self.MAX_CACHE_ITEMS:
                self.audio_cache[filename] = audio
        return audio

    def item(self, index):
        audio = self.load(self.audio_files[index])
        if not self.split:
            return audio
        n = self.segment_size
        if audio.shape[0] >= n:                              # meldataset.py:358-362
            start = self.rng.randint(0, audio.shape[0] - n)
            return audio[start:start + n]
        out = np.zeros(n, dtype=np.float32)                  # :363-364
        out[:audio.shape[0]] = audio
        return out

    def epoch_order(self):
        """DataLoader(shuffle=True, drop_last=True) of hifigan/xva_train.py:321: a fresh permutation per epoch."""
        order = list(range(len(self)))
        self.rng.shuffle(order)
        return order

    def batches(self, batch_size, device, mel_in, mel_loss):
        """One epoch of (x, y, y_mel) in the layout HiFiGANStep.step takes: x [B, T, 80] generator-input mel (channels
        last), y [B, segment] waveform, y_mel [B, T, 80] loss mel. mel_in / mel_loss: callables [B, N] -> [B, 80, T] on the
        device (hifigan.MelSpectrogram). The last incomplete batch is dropped. Uses one pinned staging buffer."""
        order = self.epoch_order()
        pin = torch.cuda.is_available() and torch.device(device).type == "cuda"
        stage = torch.empty(batch_size, self.segment_size, dtype=torch.float32, pin_memory=pin)
        for b0 in range(0, len(order) - batch_size + 1, batch_size):
            for r, idx in enumerate(order[b0:b0 + batch_size]):
                stage[r].copy_(torch.from_numpy(np.ascontiguousarray(self.item(idx))))
            y = stage.to(device, non_blocking=False).clone() if not pin else stage.to(device, non_blocking=True)
            if pin:
                torch.cuda.current_stream().synchronize()    # the staging buffer is reused for the next batch
            yield mel_in(y).transpose(1, 2).contiguous(), y, mel_loss(y).transpose(1, 2).contiguous()
