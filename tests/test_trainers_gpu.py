"""Facade test: a fake websocket records send() strings, a temp output_path collects the
files the Electron UI reads; asserts names / JSON schemas / protocol strings of the reference's trainers."""
import asyncio
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

# Device and synthetic FastPitch dataset of these tests. tests/test_trainers_cpu.py re-runs the host-logic ones on the CPU
# through the emulated C ABI (tests/cabi_emu.py) with DEV = "cpu" and a smaller dataset.
DEV = "cuda:0"
FP_SPEC = "synthetic:4x24x64x16"          # B x tokens x frames x utterances


def _fp_dir(spec=None):
    return (spec or FP_SPEC).replace(":", "_")


def _n_batches(spec=None):
    b, _, _, items = (spec or FP_SPEC).split(":")[1].split("x")[:4]
    return int(items) // int(b)


class FakeSocket:
    def __init__(self):
        self.sent = []

    async def send(self, msg):
        self.sent.append(msg)


class Log:
    def info(self, *a):
        pass


def test_fastpitch_handle_trainer_stages_and_files(lib, tmp_path, monkeypatch):
    from xva_trainer_b200 import trainers

    monkeypatch.setenv("XVA_B200_MAX_EPOCHS", "2")
    mm = trainers.ModelsManager(Log(), PROD=False)
    ws = FakeSocket()
    data = {"dataset_path": FP_SPEC, "output_path": str(tmp_path), "checkpoint": None, "num_workers": 0,
            "batch_size": 64, "epochs_per_checkpoint": 1, "force_stage": None}
    res = asyncio.run(trainers.handleTrainer(mm, data, ws, [0]))
    assert res == "move to hifi"
    assert ws.sent == ["Set stage to: 2 ", "Set stage to: 3 ", "Set stage to: 4 "]
    out = tmp_path / _fp_dir()
    names = sorted(os.listdir(out))
    assert "training.log" in names and "graphs.json" in names
    assert any(n.startswith("events.out.tfevents") for n in names)                 # TensorBoard scalars (xva_train.py:297,841-899)
    assert f"{_fp_dir()}.pt" in names and f"{_fp_dir()}.json" in names
    assert sum(n.startswith("FastPitch_checkpoint_") for n in names) <= 2          # only the last two are kept
    assert {n.split("_")[1] for n in names if n.startswith("Stage_")} == {"2", "3", "4"}
    graphs = json.load(open(out / "graphs.json"))
    assert set(graphs["stages"]) == {"1", "2", "3", "4", "5"}
    for s in ("2", "3", "4"):
        assert len(graphs["stages"][s]["loss"]) == 2 and graphs["stages"][s]["target_delta"] is not None
    ck = torch.load(out / [n for n in names if n.startswith("FastPitch_checkpoint_")][-1], map_location="cpu")
    assert set(ck) == {"epoch", "iteration", "avg_loss_per_epoch", "training_stage", "state_dict", "optimizer"}
    assert len(ck["state_dict"]) == 185 and ck["state_dict"]["decoder.layers.0.pos_ff.CoreNet.0.weight"].shape == (1536, 384, 3)
    half = torch.load(out / f"{_fp_dir()}.pt", map_location="cpu")
    assert half["proj.weight"].dtype == torch.float16
    log = open(out / "training.log").read()
    assert "Stage: 2" in log and "frames/s" in log
    # losses went down within a stage
    l3 = [v for _, v in graphs["stages"]["3"]["loss"]]
    assert l3[-1] < l3[0]


def test_fastpitch_handle_trainer_starts_at_the_aligner(lib, tmp_path, monkeypatch):
    """Batches that carry the alignment prior start at stage 1 like a new voice in the reference; the durations stage 2
    trains on are the aligner's (xva_train.py:1128-1160), not the ones the batches came with."""
    from xva_trainer_b200 import trainers

    monkeypatch.setenv("XVA_B200_MAX_EPOCHS", "2")
    mm = trainers.ModelsManager(Log(), PROD=False)
    ws = FakeSocket()
    data = {"dataset_path": FP_SPEC + ":prior", "output_path": str(tmp_path), "checkpoint": None,
            "num_workers": 0, "batch_size": 64, "epochs_per_checkpoint": 1, "force_stage": None}
    seen = {}
    orig = trainers.FastPitchTrainer.extract_durations

    def spy(self):
        before = [b[0][8].clone() for b in self.batches]
        orig(self)
        seen["changed"] = any(not torch.equal(a, b[0][8]) for a, b in zip(before, self.batches))
        seen["sums"] = [b[0][8].sum(1).cpu() for b in self.batches]
        seen["mel_lens"] = [b[0][3].cpu() for b in self.batches]

    monkeypatch.setattr(trainers.FastPitchTrainer, "extract_durations", spy)
    res = asyncio.run(trainers.handleTrainer(mm, data, ws, [0]))
    assert res == "move to hifi"
    assert ws.sent == ["Set stage to: 1 ", "Set stage to: 2 ", "Set stage to: 3 ", "Set stage to: 4 "]
    assert seen["changed"]
    for s_, l_ in zip(seen["sums"], seen["mel_lens"]):
        assert torch.equal(s_.long(), l_)                       # a hard alignment: durations add up to the mel length
    out = tmp_path / (_fp_dir() + "_prior")
    names = sorted(os.listdir(out))
    assert {n.split("_")[1] for n in names if n.startswith("Stage_")} == {"1", "2", "3", "4"}
    graphs = json.load(open(out / "graphs.json"))
    assert len(graphs["stages"]["1"]["loss"]) == 2 and graphs["stages"]["1"]["target_delta"] is not None
    log = open(out / "training.log").read()
    assert "Stage 1: Pre-training only the alignment." in log and "Extracting durations from alignments" in log


def test_hifigan_handle_trainer_files(lib, tmp_path, monkeypatch):
    from xva_trainer_b200 import trainers

    monkeypatch.setenv("XVA_B200_MAX_EPOCHS", "1")
    mm = trainers.ModelsManager(Log(), PROD=False)
    ws = FakeSocket()
    data = {"dataset_path": "synthetic:2x8x4", "output_path": str(tmp_path), "hifigan_checkpoint": None, "num_workers": 0,
            "batch_size": 2, "epochs_per_checkpoint": 1}
    res = asyncio.run(trainers.handleTrainerHiFi(mm, data, ws, [0]))
    assert res == "done"
    assert ws.sent == ["Set stage to: 5 ", "Finished training HiFi-GAN\n"]
    out = tmp_path / "synthetic_2x8x4"
    hifi = sorted(os.listdir(out / "hifi"))
    assert any(n.startswith("g_") and len(n) == 10 for n in hifi) and any(n.startswith("do_") and len(n) == 11 for n in hifi)
    assert os.path.exists(out / "synthetic_2x8x4.hg.pt")
    g = torch.load(out / "hifi" / [n for n in hifi if n.startswith("g_")][-1], map_location="cpu")
    assert list(g) == ["generator"] and len(g["generator"]) == 234
    do = torch.load(out / "hifi" / [n for n in hifi if n.startswith("do_")][-1], map_location="cpu")
    assert {"mpd", "msd", "optim_g", "optim_d", "steps", "epoch"} <= set(do)
    assert len(do["mpd"]) == 90 and len(do["msd"]) == 80
    graphs = json.load(open(out / "graphs.json"))
    assert len(graphs["stages"]["5"]["loss"]) == 1


def test_hifigan_trainer_on_a_voice_folder(lib, tmp_path, monkeypatch):
    """HiFiTrainer on metadata.csv + wavs/ (hifigan/xva_train.py:309-325): crops assembled on the host, both mels per batch
    on the device."""
    import numpy as np
    from scipy.io.wavfile import write as write_wav
    from xva_trainer_b200 import trainers

    voice = tmp_path / "voice"
    (voice / "wavs").mkdir(parents=True)
    rng = np.random.RandomState(1)
    lines = []
    for i in range(6):
        n = int(rng.randint(6000, 40000))
        t = np.arange(n) / 22050.0
        x = (8000 * np.sin(2 * np.pi * (110 + 40 * i) * t) + 500 * rng.randn(n)).astype(np.int16)
        write_wav(str(voice / "wavs" / f"u{i}.wav"), 22050, x)
        lines.append(f"u{i}|line {i}")
    (voice / "metadata.csv").write_text("\n".join(lines) + "\n")
    monkeypatch.setenv("XVA_B200_MAX_EPOCHS", "1")
    mm = trainers.ModelsManager(Log(), PROD=False)
    ws = FakeSocket()
    # a real voice is always fine-tuned from a generator checkpoint ("Don't ever train from scratch",
    # hifigan/xva_train.py:276-277): hand the trainer one in the reference's g_ file layout
    from xva_trainer_b200 import hifigan as hg
    base = tmp_path / "g_pretrained"
    torch.save({"generator": hg.Generator(trainers._Cfg(trainers.HIFI_CONFIG_V1), device=DEV).state_dict()}, base)
    data = {"dataset_path": str(voice), "output_path": str(tmp_path / "out"), "hifigan_checkpoint": str(base), "num_workers": 0,
            "batch_size": 3, "epochs_per_checkpoint": 1}
    os.makedirs(tmp_path / "out", exist_ok=True)
    res = asyncio.run(trainers.handleTrainerHiFi(mm, data, ws, [0]))
    assert res == "done" and ws.sent == ["Set stage to: 5 ", "Finished training HiFi-GAN\n"]
    log = open(tmp_path / "out" / "voice" / "training.log").read()
    assert "Training items: 6 | Data multiplier: 167 | Not found: 0 | Total: 1002" in log
    graphs = json.load(open(tmp_path / "out" / "voice" / "graphs.json"))
    assert len(graphs["stages"]["5"]["loss"]) == 1 and np.isfinite(graphs["stages"]["5"]["loss"][0][1])


# ------------------------------------------------------------------------------------------------ round-1 advisor findings
def _fp_data(tmp_path, spec=None, **kw):
    d = {"dataset_path": spec or FP_SPEC, "output_path": str(tmp_path), "checkpoint": None, "num_workers": 0, "batch_size": 64,
         "epochs_per_checkpoint": 1, "force_stage": None}
    d.update(kw)
    return d


def test_fastpitch_base_checkpoint_is_only_the_starting_point(lib, tmp_path, monkeypatch):
    """xva_train.py:284-288 + :380-385: a user-supplied base checkpoint seeds a run whose output folder is empty ("New
    voice": stage and iteration restart, weights kept); after that the run's own newest checkpoint wins, so the stages
    advance 2 -> 3 -> 4 instead of reloading the base (and its stage) at every stage change."""
    from xva_trainer_b200 import trainers

    monkeypatch.setenv("XVA_B200_MAX_EPOCHS", "1")
    mm = trainers.ModelsManager(Log(), PROD=False)
    res = asyncio.run(trainers.handleTrainer(mm, _fp_data(tmp_path / "a"), FakeSocket(), [0]))
    assert res == "move to hifi"
    a = tmp_path / "a" / _fp_dir()
    base = [n for n in os.listdir(a) if n.startswith("Stage_2_DONE_")]
    assert len(base) == 1
    base_ck = torch.load(a / base[0], map_location="cpu")
    assert base_ck["training_stage"] == 3
    ws = FakeSocket()
    other = FP_SPEC.rsplit("x", 1)[0] + "x" + str(2 * int(FP_SPEC.rsplit("x", 1)[1]))     # another voice: twice the utterances
    data = _fp_data(tmp_path / "b", spec=other, checkpoint=str(a / base[0]))
    res = asyncio.run(trainers.handleTrainer(mm, data, ws, [0]))
    assert res == "move to hifi"
    assert ws.sent == ["Set stage to: 2 ", "Set stage to: 3 ", "Set stage to: 4 "]
    log = open(tmp_path / "b" / _fp_dir(other) / "training.log").read()
    assert log.count("New voice") == 1 and log.count(f"Checkpoint: {a / base[0]}") == 1
    # a finished run on disk (training_stage 5) goes straight on to the vocoder (:356-359); so does force_stage 5
    ws2 = FakeSocket()
    assert asyncio.run(trainers.handleTrainer(mm, data, ws2, [0])) == "move to hifi"
    assert ws2.sent == []
    ws3 = FakeSocket()
    d3 = _fp_data(tmp_path / "c", force_stage=5)
    assert asyncio.run(trainers.handleTrainer(mm, d3, ws3, [0])) == "move to hifi" and ws3.sent == []


def test_fastpitch_nan_batch_is_skipped_before_the_update(lib, tmp_path, monkeypatch):
    """xva_train.py:825-832: a non-finite micro-batch is dropped BEFORE the optimizer step -- weights, LAMB moments and the
    tf32 weight copy never see it."""
    from xva_trainer_b200 import trainers

    monkeypatch.setenv("XVA_B200_MAX_EPOCHS", "1")
    batches = trainers._synthetic_fastpitch_batches(FP_SPEC, torch.device(DEV))
    assert len(batches) == _n_batches() and len(batches) >= 2
    batches[1][0][2][0, 3, 5] = float("nan")          # one NaN in the mel target of the second batch
    mm = trainers.ModelsManager(Log(), PROD=False)
    seen = {}
    orig = trainers.FastPitchTrainer.finish_epoch

    def spy(self):
        A = self.model.arena
        seen.setdefault("finite", []).append(bool(torch.isfinite(A.p).all() and torch.isfinite(A.m).all()
                                                  and torch.isfinite(A.v).all() and torch.isfinite(A.w).all()))
        seen["steps"] = self.optimizer.steps
        return orig(self)

    monkeypatch.setattr(trainers.FastPitchTrainer, "finish_epoch", spy)
    data = _fp_data(tmp_path, force_stage=3, batch_source=batches)
    with pytest.raises(RuntimeError):
        # force_stage stays 3 for every re-entry of this hand-driven call: stop after the first stage
        asyncio.run(_one_stage(trainers, mm, data))
    assert seen["finite"] and all(seen["finite"])
    assert seen["steps"] == _n_batches() - 1           # gam = 1, one batch skipped
    log = open(tmp_path / _fp_dir() / "training.log").read()
    assert log.count("loss is NaN") == 1


async def _one_stage(trainers, mm, data):
    trainer = mm.sync_init_model("fastpitch1_1", websocket=None, gpus=[0])
    await trainer.start(data, gpus=[0])


def test_hifigan_optimizer_state_is_torch_adamw_format_and_resumes(lib, tmp_path, monkeypatch):
    """do_ checkpoints carry optim_g / optim_d in torch.optim.AdamW's state_dict layout (hifigan/xva_train.py:583-584): a
    stock AdamW over parameters of the same shapes loads them, and a second run resumes from them."""
    from xva_trainer_b200 import trainers

    monkeypatch.setenv("XVA_B200_MAX_EPOCHS", "1")
    mm = trainers.ModelsManager(Log(), PROD=False)
    data = {"dataset_path": "synthetic:2x8x4", "output_path": str(tmp_path), "hifigan_checkpoint": None, "num_workers": 0,
            "batch_size": 2, "epochs_per_checkpoint": 1}
    assert asyncio.run(trainers.handleTrainerHiFi(mm, data, FakeSocket(), [0])) == "done"
    hifi = tmp_path / "synthetic_2x8x4" / "hifi"
    do_name = sorted(n for n in os.listdir(hifi) if n.startswith("do_"))[-1]
    do = torch.load(hifi / do_name, map_location="cpu")
    g = torch.load(hifi / do_name.replace("do_", "g_"), map_location="cpu")["generator"]
    for key in ("optim_g", "optim_d"):
        assert set(do[key]) == {"state", "param_groups"} and do[key]["param_groups"][0]["betas"] == (0.8, 0.99)
    # the generator's parameters in registration order = state_dict order without buffers (it has none)
    params = [torch.nn.Parameter(v.clone().float()) for v in g.values()]
    ref_opt = torch.optim.AdamW(params, 2e-4, betas=[0.8, 0.99])
    ref_opt.load_state_dict(do["optim_g"])
    st = ref_opt.state_dict()["state"]
    assert len(st) == len(params) and all(st[i]["exp_avg"].shape == params[i].shape for i in range(len(params)))
    assert float(st[0]["step"]) == 2.0                 # 4 items / batch 2 = 2 steps in the epoch
    # resume: the second run starts from these files (steps continue, moments are loaded, nothing raises)
    monkeypatch.setenv("XVA_B200_MAX_EPOCHS", "2")
    assert asyncio.run(trainers.handleTrainerHiFi(mm, data, FakeSocket(), [0])) == "done"
    do2 = torch.load(hifi / sorted(n for n in os.listdir(hifi) if n.startswith("do_"))[-1], map_location="cpu")
    assert do2["steps"] > do["steps"] and float(do2["optim_g"]["state"][0]["step"]) == 4.0
    log = open(tmp_path / "synthetic_2x8x4" / "training.log").read()
    assert "OPTIM NOT LOADED" not in log
