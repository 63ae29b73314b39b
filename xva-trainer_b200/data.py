"""FastPitch batch assembly: the step between the dataset and ``FastPitch.forward`` (the
12-list is the contract of section 8b). Mirrors python/fastpitch1_1/fastpitch/data_function.py:

    TTSCollate.__call__   :560-695   sort by text length, right-pad text / mel / pitch / energy / prior / durations
    batch_to_gpu          :706-741   casts, host -> device copies, the (x, y, num_frames) triple the trainer consumes

Items are what ``TTSDataset.__getitem__`` returns (:300-352): ``(text LongTensor [n_text], mel [80, n_mel], len(text),
pitch [n_formants, n_mel] | [0], energy [n_mel] | [0], speaker | None, attn_prior [n_mel, n_text] | None,
durs [n_text] | None, audiopath)`` with numpy arrays for everything but the text.

Faithful to the reference where it matters for results: the reference allocates its pitch / energy / duration buffers with
the TEXT dtype (int64, :596-598 and :641-644) and adds the float arrays into them, so what reaches the model is every
value TRUNCATED TOWARD ZERO (a normalised pitch of -0.9 trains as 0, an energy of 8.7 as 8) before ``batch_to_gpu`` casts
back to float. ``TTSCollate(exact_targets=False)`` (the default) reproduces that bit for bit -- the golden test compares
against the reference's own output; ``exact_targets=True`` keeps the float values, which is presumably what was meant.
Host buffers can be pinned (``pin_memory=True``) so ``batch_to_gpu`` overlaps its copies with the previous step.
"""
import numpy as np
import torch


def _as_tensor(a):
    return a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))


class TTSCollate:
    """Zero-pads model inputs and targets; ``training_stage`` selects which of pitch / energy / durations / prior exist
    (the reference sets it as an attribute after construction, xva_train.py:449-450)."""

    def __init__(self, training_stage=1, exact_targets=False, pin_memory=False):
        self.training_stage = training_stage
        self.exact_targets = exact_targets
        self.pin_memory = pin_memory

    def _target(self, values):
        """A float target as the reference's int64 buffers hold it (see module docstring)."""
        t = _as_tensor(values).to(torch.float32)
        return t if self.exact_targets else t.to(torch.float64).trunc().to(torch.int64)

    def _pin(self, t):
        return t.pin_memory() if (self.pin_memory and torch.is_tensor(t) and torch.cuda.is_available()) else t

    def __call__(self, batch):
        stage = self.training_stage
        B = len(batch)
        input_lengths, order = torch.sort(torch.LongTensor([len(x[0]) for x in batch]), dim=0, descending=True)
        order = [int(i) for i in order]
        max_input_len = int(input_lengths[0])
        target_dtype = torch.float32 if self.exact_targets else torch.int64

        text_padded = torch.zeros(B, max_input_len, dtype=torch.int64)
        num_mels = batch[0][1].shape[0]
        max_target_len = max(x[1].shape[1] for x in batch)
        mel_padded = torch.zeros(B, num_mels, max_target_len, dtype=torch.float32)
        output_lengths = torch.zeros(B, dtype=torch.int64)
        for row, i in enumerate(order):
            text, mel = batch[i][0], _as_tensor(batch[i][1]).to(torch.float32)
            text_padded[row, :text.shape[0]] = text
            mel_padded[row, :, :mel.shape[1]] = mel
            output_lengths[row] = mel.shape[1]

        if stage in (3, 4, -1):                                    # :593-611
            n_formants = batch[0][3].shape[0]
            pitch_padded = torch.zeros(B, n_formants, max_target_len, dtype=target_dtype)
            energy_padded = torch.zeros(B, max_target_len, dtype=target_dtype)
            for row, i in enumerate(order):
                pitch, energy = self._target(batch[i][3]), self._target(batch[i][4])
                pitch_padded[row, :, :pitch.shape[1]] = pitch
                energy_padded[row, :energy.shape[0]] = energy
        else:
            pitch_padded, energy_padded = torch.tensor([0]), torch.tensor([0])

        speaker = None
        if batch[0][5] is not None:                                # :613-620 (256-d resemblyzer embedding)
            speaker = torch.zeros(B, 256)
            for row, i in enumerate(order):
                speaker[row] = _as_tensor(batch[i][5]).to(torch.float32)

        if stage not in (1, -1):                                   # :624-650 extracted durations
            max_dur = max(batch[i][7].shape[0] for i in order)
            durs_padded = torch.zeros(B, max_dur, dtype=target_dtype)
            for row, i in enumerate(order):
                durs = self._target(batch[i][7])
                durs_padded[row, :durs.shape[0]] = durs
            attn_prior_padded = torch.tensor([0])
        else:                                                      # :652-676 alignment priors
            durs_padded = torch.tensor([0])
            rows = max(max_target_len, max(batch[i][6].shape[0] for i in order))
            cols = max(max_input_len, max(x[6].shape[1] for x in batch))
            attn_prior_padded = torch.zeros(B, rows, cols)
            for row, i in enumerate(order):
                prior = _as_tensor(batch[i][6]).to(torch.float32)
                attn_prior_padded[row, :prior.shape[0], :prior.shape[1]] = prior

        len_x = torch.Tensor([x[2] for x in batch])                # :681-682 (batch order, not the sorted order)
        max_inp_lengths = torch.full((B,), max_input_len, dtype=torch.int64)
        max_mel_lengths = torch.full((B,), max_target_len, dtype=torch.int64)
        audiopaths = [batch[i][8] for i in order]
        out = (text_padded, input_lengths, mel_padded, output_lengths, len_x, pitch_padded, energy_padded, speaker,
               attn_prior_padded, durs_padded, max_inp_lengths, max_mel_lengths, audiopaths)
        return tuple(self._pin(t) for t in out)


def batch_to_gpu(batch, training_stage=1, device=0):
    """data_function.py:706-741 -> (x, y, num_frames): x is the 12-list FastPitch.forward takes, y = [mel, input_lengths,
    output_lengths] (the trainer appends max_inp_lengths, xva_train.py:785), num_frames = sum of the mel lengths (the
    trainer's frames/s counter). ``device``: an index like the reference, a torch.device, or "cpu" (tests)."""
    (text_padded, input_lengths, mel_padded, output_lengths, len_x, pitch_padded, energy_padded, speaker, attn_prior,
     durs_padded, max_inp_lengths, max_mel_lengths, audiopaths) = batch
    dev = torch.device(f"cuda:{device}") if isinstance(device, int) else torch.device(device)

    def put(t, dtype):
        return t.contiguous().to(dev, non_blocking=True).to(dtype)

    text_padded, input_lengths = put(text_padded, torch.int64), put(input_lengths, torch.int64)
    mel_padded, output_lengths = put(mel_padded, torch.float32), put(output_lengths, torch.int64)
    if training_stage != 1:
        pitch_padded, energy_padded = put(pitch_padded, torch.float32), put(energy_padded, torch.float32)
        durs_padded, attn_prior = put(durs_padded, torch.float32), None
    else:
        pitch_padded = energy_padded = durs_padded = None
        attn_prior = put(attn_prior, torch.float32)
    if speaker is not None:
        speaker = put(speaker, torch.float32)
    max_inp_lengths, max_mel_lengths = put(max_inp_lengths, torch.float32), put(max_mel_lengths, torch.float32)
    x = [text_padded, input_lengths, mel_padded, output_lengths, pitch_padded, energy_padded, speaker, attn_prior,
         durs_padded, max_inp_lengths, max_mel_lengths, audiopaths]
    y = [mel_padded, input_lengths, output_lengths]
    return x, y, torch.sum(output_lengths)
