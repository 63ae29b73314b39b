"""Host-orchestration fingerprint: the sequence of C-ABI calls (entry point, every scalar argument, the complete GEMM
argument block except device pointers) that one FastPitch step of each training stage (3, 2, 4, 1), one HiFi-GAN step, one
xVAPitch --hifi_only step and one forward + backward of the xVAPitch text encoder emit, with the kernels stubbed out, on
tiny seeded inputs. CPU only.

tests/golden/launch_sequence.json holds the fingerprint of a tree whose GPU parity suite was green
(`python tests/launch_sequence.py --write` after such a run). tests/test_launch_sequence.py recomputes it: a refactor
of the Python host code between two GPU runs either leaves every launch and every argument unchanged -- then the
parity evidence still applies -- or shows up here, with the first differing call, before any GPU time is spent.
"""
import ctypes as C
import hashlib
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "launch_sequence.json")


class _FakeStream:
    """Stand-in for torch.cuda.Stream in the host-only dry run of the stream experiments."""
    waits = 0

    def __init__(self, *a, **k):
        pass

    cuda_stream = 0

    def wait_stream(self, other):
        _FakeStream.waits += 1

    def wait_event(self, event):
        _FakeStream.waits += 1

    def synchronize(self):
        pass


class _FakeEvent:
    def __init__(self, *a, **k):
        pass

    def record(self, stream=None):
        pass


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


def record(streams=False):
    """streams=True: the same steps with the (off by default) multi-stream experiments switched on -- XVA_BWD_STREAMS,
    XVA_DISC_STREAMS, XVA_GEN_STREAMS -- and torch's CUDA stream API replaced by stand-ins: Python issues the launches in
    the same order whatever stream they go to, so the sequence must not change."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import torch

    from xva_trainer_b200 import capi, ops
    from oracle import fastpitch as ofp, hifigan as ohg          # seeded synthetic batches only

    seq = []

    def ser(a):
        if isinstance(a, (int, float)):
            return a
        if isinstance(a, C.c_float):
            return a.value
        g = getattr(a, "_obj", None)
        if isinstance(g, capi.GemmArgs):                           # byref(GemmArgs): every field but the pointers
            out = {}
            for name, typ in g._fields_:
                v = getattr(g, name)
                if typ is C.c_void_p:
                    out[name] = None if v is None else "ptr"
                elif hasattr(v, "__len__"):
                    out[name] = list(v)
                else:
                    out[name] = v
            return out
        return None if a is None else "ptr"

    saved = (capi.load, capi.call, ops._stream, ops._check3, ops.duration_scan)
    saved_cuda = (torch.cuda.Stream, torch.cuda.stream, torch.cuda.current_stream, getattr(torch.Tensor, "record_stream", None),
                  torch.cuda.Event)
    saved_env = {k: os.environ.get(k) for k in ("XVA_BWD_STREAMS", "XVA_DISC_STREAMS", "XVA_GEN_STREAMS")}
    Tm = 40
    try:
        if streams:
            os.environ.update(XVA_BWD_STREAMS="1", XVA_DISC_STREAMS="4", XVA_GEN_STREAMS="1")
            torch.cuda.Stream = _FakeStream
            torch.cuda.Event = _FakeEvent
            torch.cuda.stream = lambda s: _NullCtx()
            torch.cuda.current_stream = lambda *a, **k: _FakeStream()
            torch.Tensor.record_stream = lambda self, s: None
            _FakeStream.waits = 0
        capi.load = lambda: types.SimpleNamespace(xva_attn_ctc_workspace_bytes=lambda B, T, Tt: 2 * B * T * (2 * Tt + 1) * 8 + B * 8 + (B * T * 4 + 7) // 8 * 8)
        capi.call = lambda name, *a: seq.append([name, [ser(x) for x in a]])
        ops._stream = lambda: None
        ops._check3 = lambda t, name: None
        # the stubbed scan leaves its outputs unwritten: hand back legal lengths so the host code can size its tensors
        ops.duration_scan = lambda d, pace=1.0, mx=None: (torch.zeros(d.shape[0], d.shape[1] + 1, dtype=torch.int32),
                                                          torch.full((d.shape[0],), Tm, dtype=torch.int32))

        def load(modname, patches):
            """The product modules refuse a non-CUDA device (there is no CPU path); for this host-only dry run the
            check is patched out of a private copy of the module source."""
            src = open(os.path.join(ROOT, "xva-trainer_b200", modname + ".py")).read()
            for a, b in patches:
                assert a in src, f"{modname}: patch anchor not found"
                src = src.replace(a, b)
            m = types.ModuleType(f"xva_trainer_b200.{modname}_dry")
            m.__package__ = "xva_trainer_b200"
            exec(compile(src, modname + "_dry", "exec"), m.__dict__)
            return m

        fp = load("fastpitch", [('if self.device_.type != "cuda":', "if False:"),
                                ('self.device_ = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")',
                                 'self.device_ = torch.device("cpu")')])
        hg = load("hifigan", [('if dev.type != "cuda":', "if False:")])
        torch.manual_seed(0)
        x, y = ofp.synthetic_batch(2, 16, Tm, seed=7, prior=True)
        m = fp.FastPitch(device="cpu")
        m.train()
        crit, opt = fp.FastPitchLoss(), fp.Lamb(m, lr=0.1)
        kl = fp.AttentionBinarizationLoss()
        for stage in (3, 2, 4, 1):
            m.training_stage = crit.training_stage = stage
            m.zero_grad()
            out = m(x) if stage == 1 else m(x, host_lens=(Tm, Tm))
            crit(out, y)
            if stage == 1:
                kl(out[9], out[8])
                m.backward(crit, 0.5, kl=(kl, 0.25))
            else:
                m.backward(crit, 1.0)
            opt.step()
            m.step_dropout()
        m.eval()
        m.infer(x[0], pace=0.9)

        class H(dict):
            __getattr__ = dict.__getitem__

        h = H(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
              resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3, learning_rate=2e-4, adam_b1=0.8,
              adam_b2=0.99, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256, win_size=1024, fmin=0, fmax=8000,
              fmax_for_loss=None)
        G = hg.Generator(h, device="cpu")
        G.train()
        mpd, msd = hg.MultiPeriodDiscriminator(device="cpu"), hg.MultiScaleDiscriminator(device="cpu")
        mpd.train()
        msd.train()
        xx, yy, y_mel = ohg.synthetic_batch(2, 32, seed=1)
        hg.HiFiGANStep(G, mpd, msd, h).step(xx, yy, y_mel)

        # the xVAPitch --hifi_only step (vits.py imports its building blocks from hifigan: hand it the dry-run copy)
        real_hg = sys.modules.get("xva_trainer_b200.hifigan")
        sys.modules["xva_trainer_b200.hifigan"] = hg
        try:
            vt = load("vits", [('if dev.type != "cuda":', "if False:"),
                               ('dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")',
                                'dev = torch.device("cpu")'),
                               ('capi.call("xva_device_check", dev.index or 0)', "pass")])
        finally:
            if real_hg is not None:
                sys.modules["xva_trainer_b200.hifigan"] = real_hg
            else:
                sys.modules.pop("xva_trainer_b200.hifigan", None)
        enc = vt.PosteriorEncoder(513, 192, 192, kernel_size=5, dilation_rate=1, num_layers=16, cond_channels=512, device="cpu")
        dec = hg.HifiganGenerator(192, 1, "1", [[1, 3, 5]] * 3, [3, 7, 11], [16, 16, 4, 4], 512, [8, 8, 2, 2],
                                  inference_padding=0, cond_channels=512, conv_pre_weight_norm=False,
                                  conv_post_weight_norm=False, conv_post_bias=False, device="cpu")
        disc = hg.VitsDiscriminator(device="cpu")
        for mod in (enc, dec, disc):
            mod.train()
        gen = torch.Generator().manual_seed(5)
        vt.HifiOnlyStep(enc, dec, disc).step(torch.randn(2, 513, 40, generator=gen).abs(), [40, 35],
                                              torch.randn(2, 1, 40 * 256, generator=gen), torch.randn(2, 512, generator=gen),
                                              eps=torch.randn(2, 192, 40, generator=gen), u=torch.tensor([0.3, 0.6]))

        # the xVAPitch text encoder: embedding, 2 relative-position transformer layers, prior projection, and back
        te = load("textenc", [('if dev.type != "cuda":', "if False:"),
                              ('dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")',
                               'dev = torch.device("cpu")')])
        enc_t = te.TextEncoder(50, 192, 192, 768, 2, 2, 3, 0.1, language_emb_dim=12, device="cpu")
        enc_t.train()
        li = torch.tensor([13, 8], dtype=torch.int32)
        xt, _ = enc_t.forward_cl(torch.randint(1, 50, (2, 13), generator=gen), li, torch.randn(2, 12, generator=gen))
        enc_t.stats_cl(xt, li)
        enc_t.backward_cl(torch.randn(2, 13, 204, generator=gen) + enc_t.stats_backward_cl(torch.randn(2, 13, 384, generator=gen)))
        enc_t.step_dropout()
    finally:
        capi.load, capi.call, ops._stream, ops._check3, ops.duration_scan = saved
        torch.cuda.Stream, torch.cuda.stream, torch.cuda.current_stream = saved_cuda[:3]
        torch.cuda.Event = saved_cuda[4]
        if saved_cuda[3] is not None:
            torch.Tensor.record_stream = saved_cuda[3]
        for k, v in saved_env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if streams:
        assert fp.FastPitch  # (modules were loaded with the experiments on)
        assert _FakeStream.waits > 100, "the stream experiments were not active in the dry run"
    return seq


def fingerprint(seq):
    """{section digest list}: one sha256 per 50 calls (so a change is localised) + the total."""
    lines = [json.dumps(s, sort_keys=True, default=str) for s in seq]
    chunks = [hashlib.sha256("\n".join(lines[i:i + 50]).encode()).hexdigest()[:16] for i in range(0, len(lines), 50)]
    return {"calls": len(lines), "sha256": hashlib.sha256("\n".join(lines).encode()).hexdigest(), "chunks": chunks,
            "names": [s[0] for s in seq]}


if __name__ == "__main__":
    fpr = fingerprint(record())
    if "--write" in sys.argv:
        json.dump(fpr, open(GOLDEN, "w"))
        print("written", GOLDEN)
    print(fpr["calls"], fpr["sha256"])
