"""xVAPitch text encoder on the B200: measured parity of the product path (the numbers the test bounds are set from) and
the time of one forward + backward at the full model's shape (10 layers, 256 + 12 channels, batch 32 x 160 tokens,
dropout 0.1), launched eagerly and replayed from a CUDA graph, next to the unmodified reference module under PyTorch
eager on the same GPU. Writes gpurun_out/r2ar_textenc.json."""
import json
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
out = {}
t_start = time.time()


def flush():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2ar_textenc.json"), "w"), indent=1)


def timed(fn, warm=3, steps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    from textenc_util import oracle_grads, rel, seeded_state
    from xva_trainer_b200 import capi, textenc

    dev = torch.device("cuda:0")
    # ---- 1. parity of the product path (same cases as tests/test_vits_text_encoder_gpu.py)
    import test_vits_text_encoder_gpu as T
    rows = []
    for (Tn, lens, layers, cfg) in T.CASES:
        sd, tokens, lang, seeds = T._case(Tn, lens, layers, cfg, 100 + Tn)
        want_out, want = oracle_grads(sd, tokens, lens, lang, layers, *seeds)
        m = T._build(sd, layers, **cfg)
        o, got, dlang = T._run(m, tokens, lens, lang, *seeds)
        num = sum(float((got[k].cpu() - want[k]).norm()) ** 2 for k in sd)
        den = sum(float(want[k].norm()) ** 2 for k in sd)
        floor = 1e-2 * max(float(want[k].norm()) for k in sd)
        per = sorted(((float((got[k].cpu() - want[k]).norm()) / max(float(want[k].norm()), floor), k) for k in sd), reverse=True)
        rows.append({"case": f"T={Tn} lens={lens} layers={layers} cfg={cfg}", "x": rel(o["x"], want_out["x"]),
                     "m_p": rel(o["m_p"], want_out["m_p"]), "logs_p": rel(o["logs_p"], want_out["logs_p"]),
                     "dlang": rel(dlang, want["lang"]), "grad_global": math.sqrt(num / den), "grad_worst": per[0][0],
                     "grad_worst_key": per[0][1], "grad_median": per[len(per) // 2][0]})
    out["parity_product_path"] = rows
    flush()
    # ---- 2. time of forward + backward at the full shape
    B, Tt, layers, hidden, lang_dim, vocab = 32, 160, 10, 256, 12, 200
    gen = torch.Generator().manual_seed(1)
    tokens = torch.randint(1, vocab, (B, Tt), generator=gen).to(dev)
    lens_l = [Tt - (7 * i) % 60 for i in range(B)]
    lens_l[0] = Tt
    lens = torch.tensor(lens_l, dtype=torch.int32, device=dev)
    lang = torch.randn(B, lang_dim, generator=gen).to(dev)
    dx = torch.randn(B, Tt, hidden + lang_dim, generator=gen).to(dev)
    dst = torch.randn(B, Tt, 2 * hidden, generator=gen).to(dev)
    m = textenc.TextEncoder(vocab, hidden, hidden, 768, 2, layers, 3, 0.1, language_emb_dim=lang_dim)
    m.train()

    def step():
        m.zero_grad()
        x, _ = m.forward_cl(tokens, lens, lang)
        m.stats_cl(x, lens)
        d = m.stats_backward_cl(dst)
        m.backward_cl(dx + d)
        m.step_dropout()

    capi.reset_launch_count()
    step()
    launches = capi.launch_count()
    ms = timed(step)
    out["engine"] = {"shape": f"B={B} T={Tt} layers={layers} C={hidden}+{lang_dim} ffn=768 heads=2 dropout=0.1 ragged ({sum(lens_l)} valid tokens)",
                     "ms_fwd_bwd_eager": ms, "own_kernel_launches": launches, "tokens_per_s_eager": B * Tt / ms * 1e3}
    flush()
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                step()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        msg = timed(g.replay)
        out["engine"]["ms_fwd_bwd_graph"] = msg
        out["engine"]["tokens_per_s_graph"] = B * Tt / msg * 1e3
    except Exception as e:  # noqa: BLE001
        out["engine"]["graph_error"] = repr(e)[:300]
    flush()
    # ---- 3. the unmodified reference module, PyTorch eager, same GPU and shape
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import ref_step
        ref_step.install_xvapitch()
        from python.xvapitch.model import TextEncoder as RefTE
        torch.manual_seed(0)
        r = RefTE(vocab, hidden, hidden, 768, 2, layers, 3, 0.1, language_emb_dim=lang_dim).to(dev).train()
        lang3 = lang.unsqueeze(-1)
        lens64 = lens.to(torch.int64)
        rx, rs = dx.transpose(1, 2).contiguous(), dst.transpose(1, 2).contiguous()

        def ref_step_fn():
            r.zero_grad(set_to_none=True)
            x, x_emb, mask = r(tokens, lens64, lang_emb=lang3)
            mp, lp = r(x, lens64, stats=True, x_mask=mask)
            ((x * rx).sum() + (torch.cat([mp, lp], 1) * rs).sum()).backward()

        out["reference_eager_b200"] = {"ms_fwd_bwd_fp32_tf32conv": timed(ref_step_fn, 2, 5)}
        flush()
        with torch.autocast("cuda", dtype=torch.float16):
            pass
        def ref_amp():
            r.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.float16):
                x, x_emb, mask = r(tokens, lens64, lang_emb=lang3)
                mp, lp = r(x, lens64, stats=True, x_mask=mask)
                loss = (x.float() * rx).sum() + (torch.cat([mp, lp], 1).float() * rs).sum()
            loss.backward()
        out["reference_eager_b200"]["ms_fwd_bwd_fp16_autocast"] = timed(ref_amp, 2, 5)
    except Exception as e:  # noqa: BLE001
        out["reference_eager_b200"] = {"error": repr(e)[:300]}
    out["wall_s"] = time.time() - t_start
    flush()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
