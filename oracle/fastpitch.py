"""Oracle (test infrastructure, see oracle/__init__.py): functional fp32 restatement of the FastPitch 1.1 training
path of the reference, written against a plain ``{state_dict key: tensor}`` mapping.

Reference files restated (paths relative to the reference tree, python/fastpitch1_1/...):
  fastpitch/transformer.py   PositionalEmbedding :21-35, PositionwiseConvFF :38-77, MultiHeadAttn :80-152,
                             TransformerLayer :155-171, FFTransformer.forward :212-243, mask_from_lens :299-304
  fastpitch/model.py         regulate_len :59-79, average_pitch :82-100, TemporalPredictor :103-122,
                             FastPitch.forward :325-390, get_pitch_energy :394-423, FastPitch.infer :426-482
  common/layers.py           ConvReLUNorm :85-97
  fastpitch/loss_function.py FastPitchLoss.forward :63-154 (the reference hard-codes cuda: placeholders; same math here)
  lamb.py                    Lamb.step :40-106
  fastpitch/attention.py     ConvAttention.forward :171-220 ('3xconv' query encoder; ConvNorm common/layers.py:60-82)
  fastpitch/attn_loss_function.py  AttentionCTCLoss :20-44, AttentionBinarizationLoss :47-54
  fastpitch/model.py         get_alignment_durations :298-323 and the stage-1 return of forward :346-360
  fastpitch/alignment.py     mas_width1 / b_mas :79-118

Pinned against outputs of the imported reference by tests/golden/make_golden.py (stages 2-4), make_golden_mas.py and
make_golden_stage1.py (the aligner: ConvAttention, both attention losses, their gradients) -> tests/test_oracle_golden.py.
"""
import math

import torch
import torch.nn.functional as F

D_MODEL, D_HEAD, D_INNER, N_LAYERS, N_MEL, N_SYMBOLS, D_PRED = 384, 64, 1536, 6, 80, 148, 256


# ------------------------------------------------------------------------------------------------ parameters
def state_spec():
    """(key, shape) of every entry of FastPitch().state_dict() in the reference (model.py:125-265), stage-1
    attention module included (185 keys, 46 266 420 parameters)."""
    spec = [("pitch_mean", (1,)), ("pitch_std", (1,))]

    def fft(prefix, embed):
        out = []
        if embed:
            out.append((f"{prefix}.word_emb.weight", (N_SYMBOLS, D_MODEL)))
        out.append((f"{prefix}.pos_emb.inv_freq", (D_MODEL // 2,)))
        for i in range(N_LAYERS):
            p = f"{prefix}.layers.{i}"
            out += [(f"{p}.dec_attn.qkv_net.weight", (3 * D_HEAD, D_MODEL)), (f"{p}.dec_attn.qkv_net.bias", (3 * D_HEAD,)),
                    (f"{p}.dec_attn.o_net.weight", (D_MODEL, D_HEAD)),
                    (f"{p}.dec_attn.layer_norm.weight", (D_MODEL,)), (f"{p}.dec_attn.layer_norm.bias", (D_MODEL,)),
                    (f"{p}.pos_ff.CoreNet.0.weight", (D_INNER, D_MODEL, 3)), (f"{p}.pos_ff.CoreNet.0.bias", (D_INNER,)),
                    (f"{p}.pos_ff.CoreNet.2.weight", (D_MODEL, D_INNER, 3)), (f"{p}.pos_ff.CoreNet.2.bias", (D_MODEL,)),
                    (f"{p}.pos_ff.layer_norm.weight", (D_MODEL,)), (f"{p}.pos_ff.layer_norm.bias", (D_MODEL,))]
        return out

    def predictor(prefix):
        out = []
        for i, cin in enumerate((D_MODEL, D_PRED)):
            p = f"{prefix}.layers.{i}"
            out += [(f"{p}.conv.weight", (D_PRED, cin, 3)), (f"{p}.conv.bias", (D_PRED,)),
                    (f"{p}.norm.weight", (D_PRED,)), (f"{p}.norm.bias", (D_PRED,))]
        out += [(f"{prefix}.fc.weight", (1, D_PRED)), (f"{prefix}.fc.bias", (1,))]
        return out

    spec += fft("encoder", True)
    spec += predictor("duration_predictor")
    spec += fft("decoder", False)
    spec += predictor("pitch_predictor")
    spec += [("pitch_emb.weight", (D_MODEL, 1, 3)), ("pitch_emb.bias", (D_MODEL,))]
    spec += predictor("energy_predictor")
    spec += [("energy_emb.weight", (D_MODEL, 1, 3)), ("energy_emb.bias", (D_MODEL,))]
    spec += [("proj.weight", (N_MEL, D_MODEL)), ("proj.bias", (N_MEL,))]
    spec += [("attention.query_proj.0.conv.weight", (160, 80, 3)), ("attention.query_proj.0.conv.bias", (160,)),
             ("attention.query_proj.2.conv.weight", (80, 160, 1)), ("attention.query_proj.2.conv.bias", (80,)),
             ("attention.query_proj.4.conv.weight", (80, 80, 1)), ("attention.query_proj.4.conv.bias", (80,)),
             ("attention.attn_proj.weight", (1, 80, 1, 1)), ("attention.attn_proj.bias", (1,)),
             ("attention.key_proj.0.conv.weight", (768, 384, 3)), ("attention.key_proj.0.conv.bias", (768,)),
             ("attention.key_proj.2.conv.weight", (80, 768, 1)), ("attention.key_proj.2.conv.bias", (80,))]
    return spec


BUFFERS = ("pitch_mean", "pitch_std", "encoder.pos_emb.inv_freq", "decoder.pos_emb.inv_freq")


def make_state(seed=1234, perturb=True):
    """Deterministic (CPU generator) weights with the reference's shapes and torch-default-like scales.
    ``perturb`` makes LayerNorm gains/biases and every bias non-trivial so those paths are exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in state_spec():
        if key.endswith("inv_freq"):
            sd[key] = 1.0 / (10000 ** (torch.arange(0.0, D_MODEL, 2.0) / D_MODEL))
        elif key in ("pitch_mean", "pitch_std"):
            sd[key] = torch.zeros(1)
        elif ".layer_norm." in key or ".norm." in key:
            base = 1.0 if key.endswith("weight") else 0.0
            sd[key] = base + (0.1 * torch.randn(shape, generator=g) if perturb else torch.zeros(shape))
        elif key.endswith("word_emb.weight"):
            w = torch.randn(shape, generator=g)
            w[0] = 0.0  # padding_idx
            sd[key] = w
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if key.endswith(".bias"):
                fan_in = None
            if fan_in is None:
                # bias bound uses the fan_in of its weight; the exact value is irrelevant for parity
                sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
            else:
                bound = 1.0 / math.sqrt(fan_in)
                sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


def trainable_keys(stage):
    """Parameter keys that receive gradients in a stage (xva_train.py:607-669)."""
    keys = [k for k, _ in state_spec() if k not in BUFFERS]

    def under(*prefixes):
        return [k for k in keys if any(k.startswith(p + ".") for p in prefixes)]

    if stage == 1:
        # xva_train.py:607-669 leaves encoder + attention trainable, but the stage-1 loss only reaches the token
        # embedding (the keys are encoder.word_emb(inputs), model.py:299 -- enc_out is unused) and the two projection
        # stacks; every other parameter keeps grad = None, which Lamb.step skips (lamb.py:52-53: no decay either)
        return ["encoder.word_emb.weight"] + [k for k in under("attention") if ".attn_proj." not in k]
    if stage == 2:
        return under("encoder", "duration_predictor")
    if stage == 3:
        return [k for k in keys if not (k.startswith("attention.") or k.startswith("duration_predictor."))]
    if stage == 4:
        return under("encoder", "decoder", "energy_emb", "proj")
    raise ValueError(f"stage {stage} not covered by the oracle")


# ------------------------------------------------------------------------------------------------ index paths
def mask_from_lens(lens, max_len=None):
    """transformer.py:299-304"""
    if max_len is None:
        max_len = int(lens.max())
    ids = torch.arange(0, int(max_len), device=lens.device, dtype=lens.dtype)
    return ids < lens.unsqueeze(1)


def regulate_indices(durations, pace=1.0, mel_max_len=None):
    """Integer part of regulate_len (model.py:61-70,76-78): frame -> token index (-1 past the end) and lengths."""
    reps = (durations.float() * pace + 0.5).long()
    dec_lens = reps.sum(dim=1)
    max_len = int(dec_lens.max())
    cum = torch.cumsum(reps, dim=1)
    t = torch.arange(max_len, device=durations.device)[None, :].expand(durations.size(0), -1).contiguous()
    idx = torch.searchsorted(cum, t, right=True)
    idx = torch.where(t < dec_lens[:, None], idx, torch.full_like(idx, -1))
    if mel_max_len is not None:
        idx = idx[:, :mel_max_len]
        dec_lens = torch.clamp_max(dec_lens, mel_max_len)
    return idx, dec_lens


def regulate_len(durations, enc_out, pace=1.0, mel_max_len=None):
    """model.py:59-79. The reference multiplies by a one-hot matrix; selecting rows is the same arithmetic
    (each output row is 1*x + 0*...), bit for bit in fp32."""
    idx, dec_lens = regulate_indices(durations, pace, mel_max_len)
    valid = (idx >= 0).unsqueeze(-1).to(enc_out.dtype)
    rows = torch.gather(enc_out, 1, idx.clamp_min(0).unsqueeze(-1).expand(-1, -1, enc_out.size(2)))
    return rows * valid, dec_lens


def average_pitch(pitch, durs):
    """model.py:82-100, kept in the reference's cumsum-difference form so rounding matches it."""
    ends = torch.cumsum(durs, dim=1).long()
    starts = F.pad(ends[:, :-1], (1, 0))
    nz_cums = F.pad(torch.cumsum(pitch != 0.0, dim=2), (1, 0))
    p_cums = F.pad(torch.cumsum(pitch, dim=2), (1, 0))
    bs, l = ends.size()
    nf = pitch.size(1)
    dcs = starts[:, None, :].expand(bs, nf, l)
    dce = ends[:, None, :].expand(bs, nf, l)
    sums = (torch.gather(p_cums, 2, dce) - torch.gather(p_cums, 2, dcs)).float()
    nel = (torch.gather(nz_cums, 2, dce) - torch.gather(nz_cums, 2, dcs)).float()
    return torch.where(nel == 0.0, nel, sums / nel)


# ------------------------------------------------------------------------------------------------ blocks
def positional_embedding(T, inv_freq):
    """transformer.py:28-35 -> [1,T,d]"""
    pos = torch.arange(T, device=inv_freq.device, dtype=inv_freq.dtype)
    s = torch.outer(pos, inv_freq)
    return torch.cat([s.sin(), s.cos()], dim=1)[None]


def multi_head_attn(x, pad_mask, sd, p, drop=0.0, dropatt=0.0, training=False):
    """transformer.py:100-152 (n_head = 1, post-LN). pad_mask [B,T] True at padded KEY positions."""
    qkv = F.linear(x, sd[f"{p}.qkv_net.weight"], sd[f"{p}.qkv_net.bias"])
    q, k, v = torch.chunk(qkv, 3, dim=2)
    score = torch.bmm(q, k.transpose(1, 2)) * (1.0 / math.sqrt(D_HEAD))
    score = score.masked_fill(pad_mask[:, None, :], float("-inf"))
    prob = F.dropout(F.softmax(score, dim=2), dropatt, training)
    vec = torch.bmm(prob, v)
    out = F.dropout(F.linear(vec, sd[f"{p}.o_net.weight"]), drop, training)
    return F.layer_norm(x + out, (D_MODEL,), sd[f"{p}.layer_norm.weight"], sd[f"{p}.layer_norm.bias"], 1e-5)


def conv_ff(x, sd, p, drop=0.0, training=False):
    """transformer.py:59-77 (post-LN branch)."""
    h = F.conv1d(x.transpose(1, 2), sd[f"{p}.CoreNet.0.weight"], sd[f"{p}.CoreNet.0.bias"], padding=1)
    h = F.conv1d(F.relu(h), sd[f"{p}.CoreNet.2.weight"], sd[f"{p}.CoreNet.2.bias"], padding=1)
    h = F.dropout(h, drop, training).transpose(1, 2)
    return F.layer_norm(x + h, (D_MODEL,), sd[f"{p}.layer_norm.weight"], sd[f"{p}.layer_norm.bias"], 1e-5)


def fftransformer(sd, prefix, dec_inp, seq_lens=None, conditioning=0, drop=0.0, dropatt=0.0, training=False):
    """transformer.py:212-243 -> (out [B,T,d], mask [B,T,1] bool)."""
    if f"{prefix}.word_emb.weight" in sd and dec_inp.dtype in (torch.int64, torch.int32):
        inp = F.embedding(dec_inp, sd[f"{prefix}.word_emb.weight"], padding_idx=0)
        mask = (dec_inp != 0).unsqueeze(2)
    else:
        inp = dec_inp
        mask = mask_from_lens(seq_lens, inp.size(1)).unsqueeze(2)
    pos = positional_embedding(inp.size(1), sd[f"{prefix}.pos_emb.inv_freq"]) * mask
    out = inp + pos + conditioning
    for i in range(N_LAYERS):
        p = f"{prefix}.layers.{i}"
        out = multi_head_attn(out, ~mask.squeeze(2), sd, f"{p}.dec_attn", drop, dropatt, training) * mask
        out = conv_ff(out, sd, f"{p}.pos_ff", drop, training) * mask
    return out, mask


def temporal_predictor(enc_out, mask, sd, prefix, drop=0.0, training=False):
    """model.py:118-122 with ConvReLUNorm common/layers.py:94-97."""
    out = (enc_out * mask).transpose(1, 2)
    for i in range(2):
        p = f"{prefix}.layers.{i}"
        out = F.relu(F.conv1d(out, sd[f"{p}.conv.weight"], sd[f"{p}.conv.bias"], padding=1))
        out = F.layer_norm(out.transpose(1, 2), (D_PRED,), sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-5)
        out = F.dropout(out.transpose(1, 2), drop, training)
    out = F.linear(out.transpose(1, 2), sd[f"{prefix}.fc.weight"], sd[f"{prefix}.fc.bias"])
    return out * mask


# ------------------------------------------------------------------------------------------------ model
def forward(sd, inputs_x, stage, use_gt_pitch=True, pace=1.0, max_duration=75, drop=0.0, training=False):
    """FastPitch.forward, model.py:325-390. Returns the 13-list."""
    (inputs, input_lens, mel_tgt, mel_lens, pitch_dense, energy_dense, speaker, attn_prior, durs_padded,
     max_inp_lengths, max_mel_lengths, audiopaths) = inputs_x
    if stage == 1:
        return forward_stage1(sd, inputs_x)
    mel_max_len = int(max_mel_lengths[0].item())
    enc_out, enc_mask = fftransformer(sd, "encoder", inputs, conditioning=0, drop=drop, dropatt=drop, training=training)
    dur_tgt = durs_padded
    if stage == 2:
        log_dur_pred = temporal_predictor(enc_out, enc_mask, sd, "duration_predictor", drop, training).squeeze(-1)
        dur_pred = torch.clamp(torch.exp(log_dur_pred) - 1, 0, max_duration)
        return [None, None, dur_pred, log_dur_pred, None, None, None, None, None, None, dur_tgt, None, input_lens]
    # get_pitch_energy, model.py:394-423
    pitch_pred = temporal_predictor(enc_out, enc_mask, sd, "pitch_predictor", drop, training).permute(0, 2, 1)
    pitch_tgt = average_pitch(pitch_dense, dur_tgt)
    src = pitch_tgt if use_gt_pitch else pitch_pred
    enc_out = enc_out + F.conv1d(src, sd["pitch_emb.weight"], sd["pitch_emb.bias"], padding=1).transpose(1, 2)
    energy_pred = temporal_predictor(enc_out, enc_mask, sd, "energy_predictor", drop, training).squeeze(-1)
    energy_tgt = torch.log(1.0 + average_pitch(energy_dense.unsqueeze(1), dur_tgt))
    enc_out = enc_out + F.conv1d(energy_tgt, sd["energy_emb.weight"], sd["energy_emb.bias"], padding=1).transpose(1, 2)
    energy_tgt = energy_tgt.squeeze(1)
    len_regulated, dec_lens = regulate_len(dur_tgt, enc_out, pace, mel_max_len)
    dec_out, dec_mask = fftransformer(sd, "decoder", len_regulated, seq_lens=dec_lens, drop=drop, dropatt=drop,
                                      training=training)
    mel_out = F.linear(dec_out, sd["proj.weight"], sd["proj.bias"])
    return [mel_out, dec_mask, None, None, pitch_pred, pitch_tgt, energy_pred, energy_tgt, None, None, dur_tgt, None,
            input_lens]


def infer(sd, inputs, pace=1.0, dur_tgt=None, pitch_tgt=None, energy_tgt=None, max_duration=75):
    """FastPitch.infer, model.py:426-482 (no speaker embedding, no pitch_transform): free-running synthesis -- predicted
    durations, pitch and energy feed the length regulator and the decoder.
    -> (mel_out [B, 80, T], dec_lens, dur_pred, pitch_pred [B, 1, Tt], energy_pred [B, Tt])."""
    enc_out, enc_mask = fftransformer(sd, "encoder", inputs, conditioning=0)
    log_dur_pred = temporal_predictor(enc_out, enc_mask, sd, "duration_predictor").squeeze(-1)
    dur_pred = torch.clamp(torch.exp(log_dur_pred) - 1, 0, max_duration)
    pitch_pred = temporal_predictor(enc_out, enc_mask, sd, "pitch_predictor").permute(0, 2, 1)
    src = pitch_pred if pitch_tgt is None else pitch_tgt
    enc_out = enc_out + F.conv1d(src, sd["pitch_emb.weight"], sd["pitch_emb.bias"], padding=1).transpose(1, 2)
    energy_pred = temporal_predictor(enc_out, enc_mask, sd, "energy_predictor").squeeze(-1)
    esrc = energy_pred.unsqueeze(1) if energy_tgt is None else energy_tgt
    enc_out = enc_out + F.conv1d(esrc, sd["energy_emb.weight"], sd["energy_emb.bias"], padding=1).transpose(1, 2)
    len_regulated, dec_lens = regulate_len(dur_pred if dur_tgt is None else dur_tgt, enc_out, pace, None)
    dec_out, _ = fftransformer(sd, "decoder", len_regulated, seq_lens=dec_lens)
    mel_out = F.linear(dec_out, sd["proj.weight"], sd["proj.bias"])
    return mel_out.permute(0, 2, 1), dec_lens, dur_pred, pitch_pred, energy_pred


def loss(model_out, targets, stage, dur_scale=0.1, pitch_scale=0.1, energy_scale=0.1, attn_scale=1.0, kl_weight=0.0):
    """FastPitchLoss.forward, loss_function.py:63-154 (+ for stage 1 the binarization term the trainer adds,
    xva_train.py:792-798, weighted by kl_weight). -> (loss, dict of detached terms)"""
    (mel_out, dec_mask, dur_pred, log_dur_pred, pitch_pred, pitch_tgt, energy_pred, energy_tgt, attn_soft, attn_hard,
     attn_dur, attn_logprob, input_lens) = model_out
    mel_tgt, in_lens, out_lens, max_inp_lengths = targets
    if stage == 1:
        attn_loss = attention_ctc_loss(attn_logprob, input_lens, out_lens)            # loss_function.py:74-82
        total = attn_loss * attn_scale
        meta = {"loss": total.detach(), "attn_loss": attn_loss.detach(), "kl_loss": torch.zeros(())}
        if kl_weight:
            kl = attention_binarization_loss(attn_hard, attn_soft)
            meta["kl_loss"] = kl.detach() * kl_weight
            total = total + kl_weight * kl
        return total, meta
    dur_mask = mask_from_lens(input_lens, max_len=int(max_inp_lengths[0]))
    zero = torch.zeros(1, device=input_lens.device)
    mel_loss = pitch_loss = energy_loss = dur_loss = zero
    if stage == 2:
        log_dur_tgt = torch.log(attn_dur.float() + 1)
        dur_loss = (F.mse_loss(log_dur_pred, log_dur_tgt, reduction="none") * dur_mask).sum() / dur_mask.sum()
    else:
        tgt = mel_tgt.transpose(1, 2)
        ldiff = tgt.size(1) - mel_out.size(1)
        mel_p = F.pad(mel_out, (0, 0, 0, ldiff, 0, 0), value=0.0)
        mel_mask = tgt.ne(0).float()
        mel_loss = (F.mse_loss(mel_p, tgt, reduction="none") * mel_mask).sum() / mel_mask.sum()
        if stage == 3:
            ld = pitch_tgt.size(2) - pitch_pred.size(2)
            pp = F.pad(pitch_pred, (0, ld, 0, 0, 0, 0), value=0.0)
            pitch_loss = (F.mse_loss(pitch_tgt, pp, reduction="none") * dur_mask.unsqueeze(1)).sum() / dur_mask.sum()
            ep = F.pad(energy_pred, (0, ld, 0, 0), value=0.0)
            energy_loss = (F.mse_loss(energy_tgt, ep, reduction="none") * dur_mask).sum() / dur_mask.sum()
    total = mel_loss + dur_loss * dur_scale + pitch_loss * pitch_scale + energy_loss * energy_scale
    meta = {"loss": total.detach(), "mel_loss": mel_loss.detach(), "duration_predictor_loss": dur_loss.detach(),
            "pitch_loss": pitch_loss.detach(), "energy_loss": energy_loss.detach()}
    return total, meta


def lamb_step(params, grads, state, lr, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6):
    """Lamb.step, lamb.py:40-106, over dicts keyed like the state_dict. Tensors in `params` are updated in place."""
    b1, b2 = betas
    for k, g in grads.items():
        if g is None:
            continue
        p = params[k]
        st = state.setdefault(k, {"exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p), "step": 0})
        st["step"] += 1
        st["exp_avg"].mul_(b1).add_(g, alpha=1 - b1)
        st["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1 - b2)
        w_norm = p.pow(2).sum().sqrt().clamp(0, 10)
        adam = st["exp_avg"] / st["exp_avg_sq"].sqrt().add(eps)
        if weight_decay != 0:
            adam = adam + weight_decay * p
        a_norm = adam.pow(2).sum().sqrt()
        trust = 1.0 if (w_norm == 0 or a_norm == 0) else (w_norm / a_norm)
        p.add_(adam, alpha=-lr * float(trust))
    return params


def noam_lr(iteration, base_lr=0.1, warmup=1000):
    """adjust_learning_rate, xva_train.py:1252-1261."""
    if warmup == 0:
        scale = 1.0
    elif iteration > warmup:
        scale = 1.0 / (iteration ** 0.5)
    else:
        scale = iteration / (warmup ** 1.5)
    return base_lr * scale


def train_step(sd, batch_x, batch_y, stage, lr, opt_state, drop=0.0, training=True, clip=1000.0, kl_weight=0.0):
    """One micro-batch with gam = 1 of FastPitchTrainer.iteration, xva_train.py:784-862 (fp32, no GradScaler):
    forward, loss, backward, clip_grad_norm_(1000), LAMB. Returns (loss terms, grads)."""
    keys = trainable_keys(stage)
    leaves = {k: sd[k].detach().requires_grad_(True) for k in keys}
    work = dict(sd)
    work.update(leaves)
    out = forward(work, batch_x, stage, drop=drop, training=training)
    total, meta = loss(out, batch_y, stage, kl_weight=kl_weight)
    grads = dict(zip(keys, torch.autograd.grad(total, [leaves[k] for k in keys], allow_unused=True)))
    gl = [g for g in grads.values() if g is not None]
    norm = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in gl]))
    coef = torch.clamp(clip / (norm + 1e-6), max=1.0)
    grads = {k: (None if g is None else g * coef) for k, g in grads.items()}
    with torch.no_grad():
        lamb_step(sd, grads, opt_state, lr)
    return meta, grads


# ------------------------------------------------------------------------------------------------ synthetic batches
def synthetic_batch(B, Tt, Tm, seed=1234, ragged=False, device="cpu", prior=False):
    """SURVEY.md section 8(d) synthetic inputs, in the layout of batch_to_gpu (data_function.py:706-741).
    prior=True fills x[7] with a beta-binomial alignment prior per utterance (what stage 1 reads)."""
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
    t = lambda x: x.to(device)
    attn_prior = None
    if prior:
        attn_prior = torch.zeros(B, Tm, Tt)
        for b in range(B):
            n, m = int(in_lens[b]), int(mel_lens[b])
            attn_prior[b, :m, :n] = beta_binomial_prior(n, m)
        attn_prior = t(attn_prior)
    x = [t(text), t(in_lens), t(mel), t(mel_lens), t(pitch), t(energy), None, attn_prior, t(durs),
         t(torch.full((B,), float(Tt))), t(torch.full((B,), float(Tm))), ["synthetic"] * B]
    y = [t(mel), t(in_lens), t(mel_lens), x[9]]
    return x, y


# ------------------------------------------------------------------------------------------------ stage-1 MAS
def mas_width1(attn_map, is_log=False):
    """mas_width1, fastpitch/alignment.py:79-108, restated in numpy (the reference JIT-compiles it with numba).
    attn_map [Tm, Tt] fp32 probabilities (or log-probabilities with is_log) -> 0/1 matrix of the same shape."""
    import numpy as np

    a = np.asarray(attn_map, dtype=np.float32)
    opt = np.zeros_like(a)
    with np.errstate(divide="ignore"):
        la = a.copy() if is_log else np.log(a)                       # alignment.py:85
    la[0, 1:] = -np.inf                                                # :86
    log_p = np.zeros_like(la)
    log_p[0, :] = la[0, :]
    prev_ind = np.zeros(la.shape, dtype=np.int64)
    Tm, Tt = la.shape
    for i in range(1, Tm):                                             # :90-100, one mel row at a time (vectorised over j)
        stay = log_p[i - 1]
        adv = np.concatenate([np.array([-np.inf], dtype=np.float32), log_p[i - 1, :-1]])
        take = adv >= stay
        take[0] = False
        log_p[i] = la[i] + np.where(take, adv, stay)
        prev_ind[i] = np.arange(Tt) - take.astype(np.int64)
    cur = Tt - 1
    for i in range(Tm - 1, -1, -1):                                    # :103-107
        opt[i, cur] = 1
        cur = prev_ind[i, cur]
    opt[0, cur] = 1                                                    # :108
    return opt


def b_mas(b_attn_map, in_lens, out_lens, is_log=False):
    """b_mas, fastpitch/alignment.py:110-118. b_attn_map [B, 1, Tm, Tt] -> hard alignment of the same shape."""
    import numpy as np

    a = np.asarray(b_attn_map, dtype=np.float32)
    out = np.zeros_like(a)
    for b in range(a.shape[0]):
        out[b, 0, :out_lens[b], :in_lens[b]] = mas_width1(a[b, 0, :out_lens[b], :in_lens[b]], is_log)
    return out


# ------------------------------------------------------------------------------------------------ stage-1 aligner
def beta_binomial_prior(n_text, n_mel, scaling=1.0):
    """The alignment prior the dataset caches per utterance (data_function.py:85-99: scipy.stats.betabinom(n_text, a, b)
    with a = scaling * i, b = scaling * (n_mel + 1 - i) for mel frame i = 1..n_mel, its pmf at text positions
    0..n_text-1). Restated with lgamma so the oracle needs no scipy; only used to give synthetic batches a realistic prior."""
    i = torch.arange(1, n_mel + 1, dtype=torch.float64)[:, None]
    k = torch.arange(0, n_text, dtype=torch.float64)[None, :]
    n = float(n_text)
    a, b = scaling * i, scaling * (n_mel + 1 - i)
    lg = torch.lgamma
    log_comb = lg(torch.tensor(n + 1.0, dtype=torch.float64)) - lg(k + 1) - lg(n - k + 1)
    log_beta = lambda x, y: lg(x) + lg(y) - lg(x + y)
    return torch.exp(log_comb + log_beta(k + a, n - k + b) - log_beta(a, b)).float()


def conv_attention(sd, queries, keys, key_pad_mask, attn_prior):
    """ConvAttention.forward, attention.py:171-220, with the reference's configuration (model.py:262-264: '3xconv'
    query encoder, 80 attention channels). queries [B, 80, Tm] (the mel target), keys [B, 384, Tt] (token embedding),
    key_pad_mask [B, Tt] True on padded tokens, attn_prior [B, Tm, Tt] or None.
    -> (attn_soft, attn_logprob), both [B, 1, Tm, Tt]."""
    p = "attention"
    k = F.relu(F.conv1d(keys, sd[f"{p}.key_proj.0.conv.weight"], sd[f"{p}.key_proj.0.conv.bias"], padding=1))
    k = F.conv1d(k, sd[f"{p}.key_proj.2.conv.weight"], sd[f"{p}.key_proj.2.conv.bias"])               # [B, 80, Tt]
    q = F.relu(F.conv1d(queries, sd[f"{p}.query_proj.0.conv.weight"], sd[f"{p}.query_proj.0.conv.bias"], padding=1))
    q = F.relu(F.conv1d(q, sd[f"{p}.query_proj.2.conv.weight"], sd[f"{p}.query_proj.2.conv.bias"]))
    q = F.conv1d(q, sd[f"{p}.query_proj.4.conv.weight"], sd[f"{p}.query_proj.4.conv.bias"])           # [B, 80, Tm]
    # isotropic Gaussian log-likelihood of every (frame, token) pair, :207-209
    sq = (q.transpose(1, 2).unsqueeze(2) - k.transpose(1, 2).unsqueeze(1)).pow(2).sum(-1)             # [B, Tm, Tt]
    score = (-0.0005 * sq).unsqueeze(1)
    if attn_prior is not None:
        score = F.log_softmax(score, dim=3) + torch.log(attn_prior.unsqueeze(1) + 1e-8)               # :210-211
    attn_logprob = score.clone()
    if key_pad_mask is not None:
        # the reference fills score.data in place (:215-217); out of place here, same values and same gradient
        score = score.masked_fill(key_pad_mask[:, None, None, :], float("-inf"))
    return F.softmax(score, dim=3), attn_logprob


def attention_ctc_loss(attn_logprob, in_lens, out_lens, blank_logprob=-1.0):
    """AttentionCTCLoss.forward, attn_loss_function.py:20-44: one CTC per utterance over [blank, its own keys], target
    = the token positions in order, nn.CTCLoss(zero_infinity=True) with its default 'mean' reduction, mean over B."""
    B = attn_logprob.shape[0]
    with_blank = F.pad(attn_logprob, (1, 0), value=blank_logprob)                                      # [B,1,Tm,Tt+1]
    total = 0.0
    for b in range(B):
        n_key, n_query = int(in_lens[b]), int(out_lens[b])
        lp = with_blank[b, 0, :n_query, :n_key + 1].log_softmax(dim=-1).unsqueeze(1)                   # [T, 1, L+1]
        total = total + F.ctc_loss(lp, torch.arange(1, n_key + 1).unsqueeze(0), input_lengths=[n_query],
                                   target_lengths=[n_key], blank=0, reduction="mean", zero_infinity=True)
    return total / B


def attention_binarization_loss(hard_attention, soft_attention, eps=1e-12):
    """AttentionBinarizationLoss.forward, attn_loss_function.py:47-54."""
    picked = soft_attention[hard_attention == 1].clamp(min=eps)
    return -picked.log().sum() / hard_attention.sum()


def forward_stage1(sd, inputs_x):
    """Training stage 1 of FastPitch.forward: get_alignment_durations, model.py:298-323, and the return at :356-360.
    The encoder FFT stack also runs in the reference (:345) but nothing returned depends on it."""
    (inputs, input_lens, mel_tgt, mel_lens, _, _, _, attn_prior, _, max_inp_lengths, _, _) = inputs_x
    text_emb = F.embedding(inputs, sd["encoder.word_emb.weight"], padding_idx=0)
    key_pad = ~mask_from_lens(input_lens, int(max_inp_lengths[0]))
    attn_soft, attn_logprob = conv_attention(sd, mel_tgt, text_emb.transpose(1, 2), key_pad, attn_prior)
    hard = b_mas(attn_soft.detach().cpu().numpy(), input_lens.cpu().numpy(), mel_lens.cpu().numpy())
    attn_hard = torch.from_numpy(hard).to(attn_soft.device)
    attn_hard_dur = attn_hard.sum(2)[:, 0, :]
    return [None, None, None, None, None, None, None, None, attn_soft, attn_hard, attn_hard_dur, attn_logprob,
            input_lens]


def ctc_recursion(lp, n_key, n_query, blank_logprob=-1.0):
    """The CTC forward-backward of one utterance written out (numpy fp64), as csrc/align.cu evaluates it: independent of
    torch's ctc_loss, which attention_ctc_loss above (and the reference) call. lp [Tm, Tt] = attn_logprob of the
    utterance. Extended states s = 0..2L: blank for even s, key (s-1)/2 for odd s; the keys are all distinct, so the
    skip s-2 -> s is open for every odd s >= 3. -> (cost = -log p / max(L, 1) or 0 if impossible, d cost / d lp)."""
    import numpy as np

    L, T = int(n_key), int(n_query)
    S = 2 * L + 1
    grad = np.zeros_like(lp, dtype=np.float64)
    if T == 0:
        return 0.0, grad
    z = np.concatenate([np.full((T, 1), blank_logprob), np.asarray(lp, dtype=np.float64)[:T, :L]], axis=1)
    n = z - np.log(np.exp(z - z.max(1, keepdims=True)).sum(1, keepdims=True)) - z.max(1, keepdims=True)
    col = np.array([0 if s % 2 == 0 else (s + 1) // 2 for s in range(S)])
    odd = (np.arange(S) % 2) == 1

    def lse(*terms):
        m = np.max(terms, axis=0)
        safe = np.where(np.isfinite(m), m, 0.0)
        with np.errstate(divide="ignore"):
            return np.where(np.isfinite(m), safe + np.log(sum(np.exp(t - safe) for t in terms)), -np.inf)

    ninf = np.full(S + 4, -np.inf)
    alpha = np.full((T, S), -np.inf)
    alpha[0, :min(S, 2)] = n[0, col[:min(S, 2)]]
    for t in range(1, T):
        p = ninf.copy()
        p[2:S + 2] = alpha[t - 1]
        alpha[t] = n[t, col] + lse(p[2:S + 2], p[1:S + 1], np.where(odd, p[0:S], -np.inf))
    tail = alpha[T - 1, max(S - 2, 0):]
    nll = -float(lse(*[np.array(v) for v in tail]))
    if not np.isfinite(nll):
        return 0.0, grad
    beta = np.full((T, S), -np.inf)
    beta[T - 1, max(S - 2, 0):] = n[T - 1, col[max(S - 2, 0):]]
    for t in range(T - 2, -1, -1):
        p = ninf.copy()
        p[2:S + 2] = beta[t + 1]
        beta[t] = n[t, col] + lse(p[2:S + 2], p[3:S + 3], np.where(odd, p[4:S + 4], -np.inf))
    s_of_key = 2 * np.arange(L) + 1
    post = np.exp(alpha[:, s_of_key] + beta[:, s_of_key] - n[:, 1:] + nll)
    grad[:T, :L] = (np.exp(n[:, 1:]) - post) / max(L, 1)
    return nll / max(L, 1), grad
