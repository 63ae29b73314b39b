"""CPU oracle for the xVAPitch ``--hifi_only`` training step: posterior encoder (WaveNet
stack) + waveform decoder against the VITS discriminator. TEST INFRASTRUCTURE ONLY -- nothing in the product path may
import this module; tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg are its only callers.

A restatement in plain PyTorch (fp32, CPU, autograd) of

    WN                         python/xvapitch/wavenet.py:6-106
    PosteriorEncoder           python/xvapitch/model.py:1422-1475   (configured at model.py:93-101)
    rand_segments / segment    python/xvapitch/util.py:145-178
    TorchSTFT (log-mel)        python/xvapitch/audio.py:138-181, 194-195 (configured at xvapitch/losses.py:29-46)
    hifi_only generator loss   python/xvapitch/losses.py:170-215  (feature_loss is CALLED with (fake, real) at :196 and
                               detaches its first argument, :69: the feature-matching term reaches the discriminator's
                               real-branch activations only, never the generator -- restated as the reference runs it)
    step composition           python/xvapitch/model.py:650-678, 271-340, 385-399; xva_train.py:651-736;
                               optimizers python/xvapitch/training_util.py:31-32, 66-67

with the waveform decoder and the discriminator from oracle/hifigan.py (generator_vits, vits_discriminator). Parity is
PINNED: tests/test_oracle_golden.py checks every function here against tests/golden/vits_hifi_only.npz, recorded from
the unmodified reference modules by tests/golden/make_golden_vits_hifi_only.py."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import hifigan as ohg

HIDDEN, LATENT, SPEC_BINS, COND, WN_LAYERS, WN_KERNEL = 192, 192, 513, 512, 16, 5
SEGMENT, HOP = 32, 256
MEL_ALPHA = 45.0
LR_GEN, LR_DISC, ADAM_EPS = 0.000175, 0.0002, 1e-9


def posterior_encoder_spec():
    """(key, shape) of PosteriorEncoder(513, 192, 192, 5, 1, 16, cond_channels=512).named_parameters()."""
    H = HIDDEN
    spec = [("pre.weight", (H, SPEC_BINS, 1)), ("pre.bias", (H,))]
    for i in range(WN_LAYERS):
        spec += [(f"enc.in_layers.{i}.bias", (2 * H,)), (f"enc.in_layers.{i}.weight_g", (2 * H, 1, 1)),
                 (f"enc.in_layers.{i}.weight_v", (2 * H, H, WN_KERNEL))]
    for i in range(WN_LAYERS):
        c = 2 * H if i < WN_LAYERS - 1 else H
        spec += [(f"enc.res_skip_layers.{i}.bias", (c,)), (f"enc.res_skip_layers.{i}.weight_g", (c, 1, 1)),
                 (f"enc.res_skip_layers.{i}.weight_v", (c, H, 1))]
    spec += [("enc.cond_layer.bias", (2 * H * WN_LAYERS,)), ("enc.cond_layer.weight_g", (2 * H * WN_LAYERS, 1, 1)),
             ("enc.cond_layer.weight_v", (2 * H * WN_LAYERS, COND, 1))]
    spec += [("proj.weight", (2 * LATENT, H, 1)), ("proj.bias", (2 * LATENT,))]
    return spec


def decoder_spec():
    """(key, shape) of the waveform decoder's named_parameters() (HifiganGenerator as configured at model.py:134-149)."""
    mid = [(k, sh) for k, sh in ohg.generator_spec() if k.startswith(("ups.", "resblocks."))]
    return ([("conv_pre.bias", (512,)), ("conv_pre.weight", (512, LATENT, 7))] + mid +
            [("conv_post.weight", (1, 32, 7)), ("cond_layer.weight", (512, COND, 1)), ("cond_layer.bias", (512,))])


def sequence_mask(lengths, max_len):
    """util.py:180-197: [B, max_len] bool, True where t < length."""
    return torch.arange(max_len)[None, :] < torch.as_tensor(lengths)[:, None]


def wn(sd, pre, x, x_mask, g, num_layers=WN_LAYERS, hidden=HIDDEN, kernel=WN_KERNEL, dilation_rate=1):
    """WN.forward, wavenet.py:87-106 (dropout_p = 0). x [B, H, T], x_mask [B, 1, T], g [B, C, 1] or None."""
    output = torch.zeros_like(x)
    if g is not None:
        g = F.conv1d(g, ohg.wn_weight(sd, f"{pre}.cond_layer"), sd[f"{pre}.cond_layer.bias"])
    for i in range(num_layers):
        d = dilation_rate ** i
        x_in = F.conv1d(x, ohg.wn_weight(sd, f"{pre}.in_layers.{i}"), sd[f"{pre}.in_layers.{i}.bias"], dilation=d,
                        padding=(kernel * d - d) // 2)
        if g is not None:
            x_in = x_in + g[:, i * 2 * hidden:(i + 1) * 2 * hidden, :]
        acts = torch.tanh(x_in[:, :hidden]) * torch.sigmoid(x_in[:, hidden:])          # wavenet.py:6-13
        rs = F.conv1d(acts, ohg.wn_weight(sd, f"{pre}.res_skip_layers.{i}"), sd[f"{pre}.res_skip_layers.{i}.bias"])
        if i < num_layers - 1:
            x = (x + rs[:, :hidden]) * x_mask
            output = output + rs[:, hidden:]
        else:
            output = output + rs
    return output * x_mask


def posterior_encoder(sd, y, y_lengths, g, eps):
    """PosteriorEncoder.forward, model.py:1462-1475, with the N(0, 1) draw of :1474 passed in. y [B, 513, T] ->
    (z, mean, log_scale, mask [B, 1, T])."""
    mask = sequence_mask(y_lengths, y.shape[2])[:, None, :].to(y.dtype)
    x = F.conv1d(y, sd["pre.weight"], sd["pre.bias"]) * mask
    x = wn(sd, "enc", x, mask, g)
    stats = F.conv1d(x, sd["proj.weight"], sd["proj.bias"]) * mask
    mean, log_scale = torch.split(stats, LATENT, dim=1)
    z = (mean + eps * torch.exp(log_scale)) * mask
    return z, mean, log_scale, mask


FLOWS, FLOW_LAYERS = 4, 4


def flow_spec():
    """(key, shape) of ResidualCouplingBlocks(192, 192, 5, 1, 4, cond_channels=512).named_parameters() (model.py:103-112)."""
    H, half = HIDDEN, LATENT // 2
    spec = []
    for f in range(FLOWS):
        p = f"flows.{f}"
        spec += [(f"{p}.pre.weight", (H, half, 1)), (f"{p}.pre.bias", (H,))]
        for i in range(FLOW_LAYERS):
            spec += [(f"{p}.enc.in_layers.{i}.bias", (2 * H,)), (f"{p}.enc.in_layers.{i}.weight_g", (2 * H, 1, 1)),
                     (f"{p}.enc.in_layers.{i}.weight_v", (2 * H, H, WN_KERNEL))]
        for i in range(FLOW_LAYERS):
            c = 2 * H if i < FLOW_LAYERS - 1 else H
            spec += [(f"{p}.enc.res_skip_layers.{i}.bias", (c,)), (f"{p}.enc.res_skip_layers.{i}.weight_g", (c, 1, 1)),
                     (f"{p}.enc.res_skip_layers.{i}.weight_v", (c, H, 1))]
        spec += [(f"{p}.enc.cond_layer.bias", (2 * H * FLOW_LAYERS,)), (f"{p}.enc.cond_layer.weight_g", (2 * H * FLOW_LAYERS, 1, 1)),
                 (f"{p}.enc.cond_layer.weight_v", (2 * H * FLOW_LAYERS, COND, 1))]
        spec += [(f"{p}.post.weight", (half, H, 1)), (f"{p}.post.bias", (half,))]
    return spec


def residual_coupling_blocks(sd, x, x_mask, g=None, reverse=False):
    """ResidualCouplingBlocks.forward (model.py:1406-1421) over ResidualCouplingBlock.forward (:1516-1546, mean_only, no
    projector): x [B, 192, T] -> the same shape; forward maps the posterior latent to the prior space, reverse inverts it."""
    half = LATENT // 2

    def block(f, x):
        x0, x1 = x[:, :half], x[:, half:]
        h = F.conv1d(x0, sd[f"flows.{f}.pre.weight"], sd[f"flows.{f}.pre.bias"]) * x_mask
        h = wn(sd, f"flows.{f}.enc", h, x_mask, g, num_layers=FLOW_LAYERS)
        m = F.conv1d(h, sd[f"flows.{f}.post.weight"], sd[f"flows.{f}.post.bias"]) * x_mask
        x1 = (m + x1 * x_mask) if not reverse else (x1 - m) * x_mask
        return torch.cat([x0, x1], 1)

    if not reverse:
        for f in range(FLOWS):
            x = torch.flip(block(f, x), [1])
    else:
        for f in reversed(range(FLOWS)):
            x = block(f, torch.flip(x, [1]))
    return x


# ------------------------------------------------------------------------------------------------ text encoder (oracle only)
REL_WINDOW, TEXT_HEADS = 4, 2


def relative_attention(sd, pre, x, mask, heads=TEXT_HEADS, window=REL_WINDOW):
    """RelativePositionMultiHeadAttention.forward (self-attention, shared relative embeddings, no proximal bias),
    python/xvapitch/glow_tts.py:159-214, written with explicit relative indices instead of the reference's pad / reshape
    skewing (:260-292): with d = j - i,
        score[i, j] = (q_i . k_j + [|d| <= w] q_i . E_k[d + w]) / sqrt(d_k),   masked_fill(-1e4) outside the mask,
        out_i = sum_j p[i, j] (v_j + [|d| <= w] E_v[d + w]).
    x [B, C, T], mask [B, 1, T] -> [B, C_out, T]."""
    B, C, T = x.shape
    dk = C // heads
    proj = lambda name: F.conv1d(x, sd[f"{pre}.{name}.weight"], sd[f"{pre}.{name}.bias"]).view(B, heads, dk, T).transpose(2, 3)
    q, k, v = proj("conv_q"), proj("conv_k"), proj("conv_v")                       # [B, H, T, dk]
    e_k, e_v = sd[f"{pre}.emb_rel_k"][0], sd[f"{pre}.emb_rel_v"][0]                # [2w + 1, dk]
    d = torch.arange(T)[None, :] - torch.arange(T)[:, None]                        # d[i, j] = j - i
    near = d.abs() <= window
    idx = (d + window).clamp(0, 2 * window)
    rel_k = torch.einsum("bhid, ijd -> bhij", q, e_k[idx]) * near                  # q_i . E_k[j - i + w]
    scores = (torch.matmul(q, k.transpose(-2, -1)) + rel_k) / math.sqrt(dk)
    amask = mask.unsqueeze(2) * mask.unsqueeze(-1)                                 # [B, 1, T, T]
    scores = scores.masked_fill(amask == 0, -1e4)
    p = F.softmax(scores, dim=-1)
    out = torch.matmul(p, v) + torch.einsum("bhij, ijd -> bhid", p * near, e_v[idx])
    out = out.transpose(2, 3).contiguous().view(B, C, T)
    return F.conv1d(out, sd[f"{pre}.conv_o.weight"], sd[f"{pre}.conv_o.bias"])


def _layer_norm2(sd, pre, x):
    """LayerNorm2 (glow_tts.py:34-56): layer norm over the channel dimension of [B, C, T]."""
    return F.layer_norm(x.transpose(1, -1), (x.shape[1],), sd[f"{pre}.gamma"], sd[f"{pre}.beta"], 1e-5).transpose(1, -1)


def relative_position_transformer(sd, pre, x, mask, num_layers, kernel=3):
    """RelativePositionTransformer.forward, glow_tts.py:463-485 (dropout off; in = hidden = out channels, so no proj)."""
    for i in range(num_layers):
        x = x * mask
        x = _layer_norm2(sd, f"{pre}.norm_layers_1.{i}", x + relative_attention(sd, f"{pre}.attn_layers.{i}", x, mask))
        pad = ((kernel - 1) // 2, kernel // 2)
        h = F.conv1d(F.pad(x * mask, pad), sd[f"{pre}.ffn_layers.{i}.conv_1.weight"], sd[f"{pre}.ffn_layers.{i}.conv_1.bias"])
        h = torch.relu(h)
        y = F.conv1d(F.pad(h * mask, pad), sd[f"{pre}.ffn_layers.{i}.conv_2.weight"], sd[f"{pre}.ffn_layers.{i}.conv_2.bias"]) * mask
        x = _layer_norm2(sd, f"{pre}.norm_layers_2.{i}", x + y)
    return x * mask


def text_encoder(sd, tokens, x_lengths, lang_emb, num_layers):
    """TextEncoder.forward(stats=False), xvapitch/model.py:1152-1168: scaled embedding, language embedding concatenated to
    every token, transformer. tokens [B, T] -> (x [B, C + L, T], x_emb [B, T, C], mask [B, 1, T])."""
    C = sd["emb.weight"].shape[1]
    x_emb = F.embedding(tokens, sd["emb.weight"]) * math.sqrt(C)
    x = torch.cat((x_emb, lang_emb.transpose(2, 1).expand(x_emb.size(0), x_emb.size(1), -1)), dim=-1).transpose(1, -1)
    mask = sequence_mask(x_lengths, x.shape[2])[:, None, :].to(x.dtype)
    return relative_position_transformer(sd, "encoder", x * mask, mask, num_layers), x_emb, mask


def text_encoder_stats(sd, x, mask):
    """TextEncoder.forward(stats=True), model.py:1147-1150: the prior's mean and log-scale per token."""
    stats = F.conv1d(x, sd["proj.weight"], sd["proj.bias"]) * mask
    return torch.split(stats, stats.shape[1] // 2, dim=1)


def pitch_predictor(sd, x, x_lengths, speaker_emb, num_layers, kernel=3):
    """RelativePositioningPitchEnergyEncoder.forward, xvapitch/model.py:1310-1356 (built at :154-168 with out_channels = 1)
    over RelativePositionTransformer.forward, glow_tts.py:463-485: x [B, T, hidden] (the text encoder's output, detached by
    the caller, model.py:835), speaker_emb [B, 512, 1] -> pitch_pred [B, 1, T]. With out_channels = 1 the LAST layer keeps
    its attention half only: its FFN is evaluated and thrown away (:479-483: `x = self.proj(x)` replaces `norm(x + y)`), so
    ffn_layers[-1] and norm_layers_2[-1] never receive a gradient -- restated as the reference runs it, minus that dead
    evaluation."""
    B, T, _ = x.shape
    h = torch.cat((x, speaker_emb.transpose(2, 1).expand(B, T, -1)), dim=-1).transpose(1, -1)      # [B, C, T]
    mask = sequence_mask(x_lengths, T)[:, None, :].to(h.dtype)
    h = h * mask
    pre = "encoder"
    pad = ((kernel - 1) // 2, kernel // 2)
    for i in range(num_layers):
        h = h * mask
        h = _layer_norm2(sd, f"{pre}.norm_layers_1.{i}", h + relative_attention(sd, f"{pre}.attn_layers.{i}", h, mask))
        if i + 1 == num_layers:
            h = F.conv1d(h, sd[f"{pre}.proj.weight"], sd[f"{pre}.proj.bias"])
        else:
            f = F.conv1d(F.pad(h * mask, pad), sd[f"{pre}.ffn_layers.{i}.conv_1.weight"], sd[f"{pre}.ffn_layers.{i}.conv_1.bias"])
            f = torch.relu(f)
            y = F.conv1d(F.pad(f * mask, pad), sd[f"{pre}.ffn_layers.{i}.conv_2.weight"], sd[f"{pre}.ffn_layers.{i}.conv_2.bias"]) * mask
            h = _layer_norm2(sd, f"{pre}.norm_layers_2.{i}", h + y)
    return h * mask


# ------------------------------------------------------------------------------------------------ stochastic duration predictor
# (oracle only: the engine-side module is not built. Groundwork for SURVEY.md section 8f rank 1, pinned to the reference
# module by tests/golden/make_golden_vits_sdp.py.)
SDP_BINS, SDP_TAIL = 10, 5.0


def dds_conv(sd, pre, x, x_mask, num_layers=3, kernel=3, g=None):
    """DilatedDepthSeparableConv.forward, python/xvapitch/sdp.py:77-93 (dropout off): per layer a depth-wise convolution with
    dilation kernel^i on x * mask, LayerNorm over channels, GELU, 1x1 convolution, LayerNorm, GELU, residual add."""
    if g is not None:
        x = x + g
    C = x.shape[1]
    for i in range(num_layers):
        d = kernel ** i
        y = F.conv1d(x * x_mask, sd[f"{pre}.convs_sep.{i}.weight"], sd[f"{pre}.convs_sep.{i}.bias"], groups=C, dilation=d,
                     padding=(kernel * d - d) // 2)
        y = F.gelu(_layer_norm2(sd, f"{pre}.norms_1.{i}", y))
        y = F.conv1d(y, sd[f"{pre}.convs_1x1.{i}.weight"], sd[f"{pre}.convs_1x1.{i}.bias"])
        y = F.gelu(_layer_norm2(sd, f"{pre}.norms_2.{i}", y))
        x = x + y
    return x * x_mask


def rq_spline_forward(x, w_un, h_un, d_un, tail=SDP_TAIL, min_w=1e-3, min_h=1e-3, min_d=1e-3):
    """Forward direction of the monotonic rational-quadratic spline with linear tails (Durkan et al. 2019) as the reference
    evaluates it (python/xvapitch/util.py:244-399, inverse=False): x [...], w_un / h_un [..., K], d_un [..., K - 1] ->
    (y, log|dy/dx|). Outside [-tail, tail] the map is the identity. Written without boolean indexing: every element goes
    through the spline arithmetic on a clamped copy and the tails are selected at the end."""
    K = w_un.shape[-1]
    inside = (x >= -tail) & (x <= tail)
    xc = x.clamp(-tail, tail)
    edge = math.log(math.exp(1 - min_d) - 1)                        # softplus^-1(1 - min_d): derivative 1 at both ends
    d_un = F.pad(d_un, (1, 1), value=edge)

    def knots(un, lo, hi, min_size):
        size = min_size + (1 - min_size * K) * F.softmax(un, dim=-1)
        cum = F.pad(torch.cumsum(size, dim=-1), (1, 0))
        cum = (hi - lo) * cum + lo
        cum = torch.cat([torch.full_like(cum[..., :1], lo), cum[..., 1:-1], torch.full_like(cum[..., :1], hi)], dim=-1)
        return cum, cum[..., 1:] - cum[..., :-1]

    cw, widths = knots(w_un, -tail, tail, min_w)
    ch, heights = knots(h_un, -tail, tail, min_h)
    deriv = min_d + F.softplus(d_un)
    search = cw.clone()
    search[..., -1] = search[..., -1] + 1e-6                        # util.py:239-241: the right edge belongs to the last bin
    idx = (torch.sum(xc[..., None] >= search, dim=-1) - 1).clamp(0, K - 1)[..., None]
    take = lambda t: t.gather(-1, idx)[..., 0]
    x0, bw, y0, bh = take(cw), take(widths), take(ch), take(heights)
    delta = take(heights / widths)
    d0, d1 = take(deriv), take(deriv[..., 1:])
    th = (xc - x0) / bw
    tt = th * (1 - th)
    den = delta + (d0 + d1 - 2 * delta) * tt
    y = y0 + bh * (delta * th ** 2 + d0 * tt) / den
    logdet = torch.log(delta ** 2 * (d1 * th ** 2 + 2 * delta * tt + d0 * (1 - th) ** 2)) - 2 * torch.log(den)
    return torch.where(inside, y, x), torch.where(inside, logdet, torch.zeros_like(logdet))


def conv_flow(sd, pre, x, x_mask, g, hidden=HIDDEN):
    """ConvFlow.forward (forward direction), sdp.py:149-176: the first channel conditions a spline on the second."""
    x0, x1 = x[:, :1], x[:, 1:]
    h = F.conv1d(x0, sd[f"{pre}.pre.weight"], sd[f"{pre}.pre.bias"])
    h = dds_conv(sd, f"{pre}.convs", h, x_mask, g=g)
    h = F.conv1d(h, sd[f"{pre}.proj.weight"], sd[f"{pre}.proj.bias"]) * x_mask            # [B, 3 K - 1, T]
    h = h.permute(0, 2, 1)[:, None]                                                        # [B, 1, T, 3 K - 1]
    K = SDP_BINS
    y1, logabsdet = rq_spline_forward(x1, h[..., :K] / math.sqrt(hidden), h[..., K:2 * K] / math.sqrt(hidden), h[..., 2 * K:])
    return torch.cat([x0, y1], 1) * x_mask, torch.sum(logabsdet * x_mask, [1, 2])


def _sdp_flows(sd, pre, z, x_mask, g, num_flows=4):
    """ElementwiseAffine then num_flows x [ConvFlow, channel flip] (sdp.py:268-276, 292-297) -> (z, summed log-determinant)."""
    z = (z * torch.exp(sd[f"{pre}.0.log_scale"]) + sd[f"{pre}.0.translation"]) * x_mask
    logdet = torch.sum(sd[f"{pre}.0.log_scale"] * x_mask, [1, 2])
    for i in range(1, num_flows + 1):
        z, ld = conv_flow(sd, f"{pre}.{i}", z, x_mask, g)
        logdet = logdet + ld
        z = torch.flip(z, [1])
    return z, logdet


def sdp_nll(sd, x, x_mask, dr, g, lang_emb, noise):
    """StochasticDurationPredictor.forward (training direction), sdp.py:241-300, with the N(0, 1) draw of :264 passed in:
    x [B, C + L, T] (the text encoder's output, detached by the caller), dr [B, 1, T] durations, g [B, 512, 1],
    lang_emb [B, L, 1], noise [B, 2, T] -> negative log-likelihood per utterance [B] (variational dequantisation + data
    augmentation: a posterior flow conditioned on the durations, then the main flow on log(d - u))."""
    x = F.conv1d(x, sd["pre.weight"], sd["pre.bias"])
    if g is not None:
        x = x + F.conv1d(g, sd["cond.weight"], sd["cond.bias"])
    if lang_emb is not None:
        x = x + F.conv1d(lang_emb, sd["cond_lang.weight"], sd["cond_lang.bias"])
    x = dds_conv(sd, "convs", x, x_mask)
    x = F.conv1d(x, sd["proj.weight"], sd["proj.bias"]) * x_mask
    h = F.conv1d(dr, sd["post_pre.weight"], sd["post_pre.bias"])
    h = dds_conv(sd, "post_convs", h, x_mask)
    h = F.conv1d(h, sd["post_proj.weight"], sd["post_proj.bias"]) * x_mask
    noise = noise * x_mask
    z_q, logdet_q = _sdp_flows(sd, "post_flows", noise, x_mask, x + h)
    z_u, z_v = z_q[:, :1], z_q[:, 1:]
    u = torch.sigmoid(z_u) * x_mask
    z0 = (dr - u) * x_mask
    logdet_q = logdet_q + torch.sum((F.logsigmoid(z_u) + F.logsigmoid(-z_u)) * x_mask, [1, 2])
    nll_post = torch.sum(-0.5 * (math.log(2 * math.pi) + noise ** 2) * x_mask, [1, 2]) - logdet_q
    z0 = torch.log(torch.clamp_min(z0, 1e-5)) * x_mask
    logdet = torch.sum(-z0, [1, 2])
    z, ld = _sdp_flows(sd, "flows", torch.cat([z0, z_v], 1), x_mask, x)
    return torch.sum(0.5 * (math.log(2 * math.pi) + z ** 2) * x_mask, [1, 2]) - (logdet + ld) + nll_post


# ------------------------------------------------------------------------------------------------ alignment, prior, KL
def maximum_path(value, x_lens, y_lens):
    """xVAPitch's monotonic alignment search, python/xvapitch/util.py:14-53, restated: value [B, t_x, t_y] (masked with
    mask[b, i, j] = i < x_lens[b] and j < y_lens[b]); dynamic programme over the frames j with, for every token i, the
    better of "stay on i" and "come from i - 1" -- ties stay (util.py:35: v1 >= v0) --, tokens i > j unreachable, then a
    backtrack from token x_len - 1 at the last frame; fp32 scores as in the reference. -> path [B, t_x, t_y] of 0 / 1."""
    val = value.detach().cpu().numpy().astype(np.float32)
    B, tx, ty = val.shape
    xl, yl = np.asarray(x_lens).astype(np.int64), np.asarray(y_lens).astype(np.int64)
    mask = (np.arange(tx)[None, :, None] < xl[:, None, None]) & (np.arange(ty)[None, None, :] < yl[:, None, None])
    val = val * mask
    stay_dir = np.zeros((B, tx, ty), dtype=np.int64)
    score = np.zeros((B, tx), dtype=np.float32)
    for j in range(ty):
        from_prev = np.concatenate([np.full((B, 1), -np.inf, dtype=np.float32), score[:, :-1]], axis=1)
        keep = score >= from_prev
        best = np.where(keep, score, from_prev)
        stay_dir[:, :, j] = keep
        score = np.where(np.arange(tx)[None, :] <= j, best + val[:, :, j], -np.inf).astype(np.float32)
    stay_dir = np.where(mask, stay_dir, 1)
    path = np.zeros((B, tx, ty), dtype=np.float32)
    tok = mask[:, :, 0].sum(1).astype(np.int64) - 1
    rows = np.arange(B)
    for j in reversed(range(ty)):
        path[rows, tok, j] = 1.0
        tok = tok + stay_dir[rows, tok, j] - 1
    return torch.from_numpy(path * mask)


def alignment_logp(z_p, m_p, logs_p):
    """Log-density of every latent frame under every token's diagonal Gaussian prior, in the four-term expansion the
    reference evaluates (xvapitch/model.py:766-771). z_p [B, C, t_y], m_p / logs_p [B, C, t_x] -> [B, t_x, t_y]."""
    o_scale = torch.exp(-2 * logs_p)
    logp1 = torch.sum(-0.5 * math.log(2 * math.pi) - logs_p, [1]).unsqueeze(-1)
    logp2 = torch.einsum("bct, bcs -> bts", o_scale, -0.5 * (z_p ** 2))
    logp3 = torch.einsum("bct, bcs -> bts", m_p * o_scale, z_p)
    logp4 = torch.sum(-0.5 * (m_p ** 2) * o_scale, [1]).unsqueeze(-1)
    return logp2 + logp3 + logp1 + logp4


def prior_alignment(z_p, m_p, logs_p, x_lens, y_lens):
    """model.py:763-777 + :855-856: -> (logp, path [B, t_x, t_y], durations [B, t_x], m_p and logs_p expanded to the
    frames [B, C, t_y])."""
    with torch.no_grad():
        logp = alignment_logp(z_p, m_p, logs_p)
        path = maximum_path(logp, x_lens, y_lens)
    expand = lambda t: torch.einsum("bts, bct -> bcs", path, t)
    return logp, path, path.sum(2), expand(m_p), expand(logs_p)


def kl_loss(z_p, logs_q, m_p, logs_p, z_mask):
    """VitsGeneratorLoss.kl_loss, xvapitch/losses.py:86-103: KL(q || p) of diagonal Gaussians evaluated at the flow output,
    summed over channels and valid frames, divided by the number of valid frames. z_mask [B, 1, t_y]."""
    kl = logs_p - logs_q - 0.5 + 0.5 * (z_p - m_p) ** 2 * torch.exp(-2.0 * logs_p)
    return torch.sum(kl * z_mask) / torch.sum(z_mask)


def segment_starts(u, lengths, segment_size=SEGMENT):
    """rand_segments' index rule, util.py:160-162: floor(u * (length - segment + 1)) with u ~ U[0, 1) per utterance."""
    max_idxs = torch.as_tensor(lengths) - segment_size + 1
    assert bool((max_idxs > 0).all()), "at least one sample is shorter than the segment size"
    return (u.to(torch.float32) * max_idxs).long()


def segment(x, starts, size):
    """util.py:165-178: x [B, C, T] -> [B, C, size]."""
    return torch.stack([x[i, :, int(s):int(s) + size] for i, s in enumerate(starts)])


def torch_stft_mel(x, n_fft=1024, hop=256, win=1024, sr=22050, fmin=0, fmax=8000, n_mels=80):
    """TorchSTFT(1024, 256, 1024, sample_rate=22050, mel_fmin=0, mel_fmax=8000, n_mels=80, use_mel=True,
    do_amp_to_db=True).__call__, audio.py:138-181: centred reflect-padded STFT, sqrt(clamp(re^2 + im^2, 1e-8)), Slaney
    mel, log(clamp(., 1e-5)). x [B, 1, T] or [B, T] -> [B, 80, T / hop + 1]."""
    if x.dim() == 3:
        x = x.squeeze(1)
    o = torch.stft(x, n_fft, hop, win, torch.hann_window(win), center=True, pad_mode="reflect", normalized=False,
                   onesided=True, return_complex=True)
    s = torch.sqrt(torch.clamp(o.real ** 2 + o.imag ** 2, min=1e-8))
    s = torch.matmul(ohg.mel_filterbank(sr, n_fft, n_mels, fmin, fmax), s)
    return torch.log(torch.clamp(s, min=1e-5))


def generator_losses(wav_hat, wav, scores_fake, feats_fake, feats_real):
    """VitsGeneratorLoss.forward, hifi_only branch, losses.py:184-215 -> dict(loss, loss_gen, loss_feat, loss_mel)."""
    loss_mel = F.l1_loss(torch_stft_mel(wav), torch_stft_mel(wav_hat)) * MEL_ALPHA
    loss_gen = ohg.generator_loss(scores_fake)
    loss_feat = ohg.feature_loss([[f.detach() for f in fs] for fs in feats_fake], feats_real)      # losses.py:196 + :69
    return {"loss": loss_feat + loss_mel + loss_gen, "loss_gen": loss_gen, "loss_feat": loss_feat, "loss_mel": loss_mel}


def hifi_only_step(sd_enc, sd_dec, sd_disc, linear, waveform, d_vectors, y_lengths, eps, u, opt_state, detail=None):
    """One --hifi_only iteration. Updates the three state dicts in place (AdamW, eps 1e-9; generator-side lr 1.75e-4,
    discriminator 2e-4) and returns the loss dict. linear [B, 513, T], waveform [B, 1, T * 256], d_vectors [B, 512]."""
    enc = {k: v.clone().requires_grad_(True) for k, v in sd_enc.items()}
    dec = {k: v.clone().requires_grad_(True) for k, v in sd_dec.items()}
    disc = {k: v.clone().requires_grad_(True) for k, v in sd_disc.items()}
    g = F.normalize(d_vectors).unsqueeze(-1)
    z, m_q, logs_q, _ = posterior_encoder(enc, linear, y_lengths, g, eps)
    starts = segment_starts(u, y_lengths)
    o = ohg.generator_vits(dec, segment(z, starts, SEGMENT), g)
    wav_seg = segment(waveform, starts * HOP, SEGMENT * HOP)
    s_fake, f_fake, _, f_real = ohg.vits_discriminator(disc, o, wav_seg)
    losses = generator_losses(o, wav_seg, s_fake, f_fake, f_real)
    gen_leaves = list(enc.items()) + list(dec.items())
    grads = torch.autograd.grad(losses["loss"], [v for _, v in gen_leaves])
    g_gen = {("enc" if i < len(enc) else "dec", k): gr for i, ((k, _), gr) in enumerate(zip(gen_leaves, grads))}
    # discriminator pass on the detached output
    s_fake, _, s_real, _ = ohg.vits_discriminator(disc, o.detach(), wav_seg)
    loss_disc = ohg.discriminator_loss(s_real, s_fake)
    d_grads = torch.autograd.grad(loss_disc, list(disc.values()))
    if detail is not None:
        detail.update(z=z.detach(), m_q=m_q.detach(), logs_q=logs_q.detach(), o=o.detach(), wav_seg=wav_seg, starts=starts,
                      grads_gen=g_gen, grads_disc=dict(zip(disc.keys(), d_grads)))
    with torch.no_grad():
        st = opt_state.setdefault("gen", {})
        params = {("enc", k): sd_enc[k] for k in sd_enc}
        params.update({("dec", k): sd_dec[k] for k in sd_dec})
        ohg.adamw_step(params, g_gen, st, lr=LR_GEN, eps=ADAM_EPS)
        ohg.adamw_step(sd_disc, dict(zip(disc.keys(), d_grads)), opt_state.setdefault("disc", {}), lr=LR_DISC, eps=ADAM_EPS)
    out = {k: float(v.detach()) for k, v in losses.items()}
    out["loss_disc"] = float(loss_disc.detach())
    return out
