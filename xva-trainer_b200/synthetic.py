"""Seeded synthetic inputs of SURVEY.md section 8(d) for the benchmark and the trainer facades' ``synthetic:`` datasets
(there is no network for datasets; BASELINE.json's configs are quoted on synthetic utterances).

fastpitch_batch: cfg-2 -- tokens U{1..147}, durations = 1 + multinomial(extra frames) so that every row sums to its mel
length, mel N(0, 1) (never exactly zero inside an utterance: the loss mask is ``mel != 0``, loss_function.py:104-106),
pitch N(0, 1) with 30 % exact zeros (unvoiced), energy U(0, 10); layout of batch_to_gpu (data_function.py:706-741).
hifigan_batch: cfg-3 -- audio 0.95 tanh(N(0, 0.3)), input mel (fmax 8000) and loss mel (fmax None) from the engine's own
mel-spectrogram kernels."""
import torch

N_SYMBOLS, N_MEL = 148, 80


def fastpitch_batch(B, Tt, Tm, seed=1234, ragged=False):
    """-> (x, y): the 12-list model input and the 4-list criterion target, CPU tensors."""
    g = torch.Generator().manual_seed(seed)
    if ragged:
        in_lens = torch.randint(max(1, (3 * Tt) // 5), Tt + 1, (B,), generator=g)
        in_lens[0] = Tt
    else:
        in_lens = torch.full((B,), Tt, dtype=torch.long)
    text = torch.randint(1, N_SYMBOLS, (B, Tt), generator=g)
    durs = torch.zeros(B, Tt)
    for b in range(B):
        n = int(in_lens[b])
        text[b, n:] = 0
        total = Tm if b == 0 or not ragged else int(torch.randint(max(n, (3 * Tm) // 5), Tm + 1, (1,), generator=g))
        durs[b, :n] = 1
        extra = torch.randint(0, n, (total - n,), generator=g)
        durs[b, :n] += torch.bincount(extra, minlength=n).float()
    mel_lens = durs.sum(1).long()
    mel = torch.randn(B, N_MEL, Tm, generator=g)
    pitch = torch.randn(B, 1, Tm, generator=g) * (torch.rand(B, 1, Tm, generator=g) > 0.3)
    energy = torch.rand(B, Tm, generator=g) * 10
    for b in range(B):
        mel[b, :, int(mel_lens[b]):] = 0
        pitch[b, :, int(mel_lens[b]):] = 0
        energy[b, int(mel_lens[b]):] = 0
    x = [text, in_lens, mel, mel_lens, pitch, energy, None, None, durs, torch.full((B,), float(Tt)),
         torch.full((B,), float(Tm)), ["synthetic"] * B]
    return x, [mel, in_lens, mel_lens, x[9]]


def hifigan_batch(B, frames, device, seed=1234, hop=256, fmax=8000, fmax_for_loss=None):
    """-> (x [B, 80, frames], y [B, frames * hop], y_mel [B, 80, frames]) on ``device``."""
    from . import hifigan as hg

    g = torch.Generator().manual_seed(seed)
    y = (0.95 * torch.tanh(torch.randn(B, frames * hop, generator=g) * 0.3)).to(device)
    mel_in = hg.MelSpectrogram(fmax=fmax, device=device)
    mel_loss = hg.MelSpectrogram(fmax=fmax_for_loss, device=device)
    return mel_in(y).transpose(1, 2).contiguous(), y, mel_loss(y).transpose(1, 2).contiguous()
