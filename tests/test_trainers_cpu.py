"""The trainer facades' host logic on the CPU: the very test functions of tests/test_trainers_gpu.py (stage sequence,
files and protocol strings; start at the aligner; base-checkpoint precedence and the stage-5 short cut; a NaN batch skipped
before the update) re-run with the product package's FastPitchTrainer driving the emulated C ABI (tests/cabi_emu.py) on a
smaller synthetic dataset. Nothing here measures kernels -- it is what `handleTrainer` does around them: checkpoints,
resume rules, logs, graphs.json, websocket strings, exports. CPU only; CUDA graphs off (XVA_TRAINER_GRAPH=0: the eager path
of the same micro-step)."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import cabi_emu  # noqa: E402
import test_trainers_gpu as T  # noqa: E402

FP_PATCHES = [('if self.device_.type != "cuda":', "if False:"),
              ('self.device_ = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")',
               'self.device_ = torch.device("cpu")')]
HG_PATCHES = [('if dev.type != "cuda":', "if False:")]
TR_PATCHES = [('self.device = torch.device(f"cuda:{gpu}")', 'self.device = torch.device("cpu")')]


@pytest.fixture(autouse=True)
def _drop_checkpoints(tmp_path):
    """A FastPitch checkpoint is 0.55 GB (weights + both LAMB moments) and every stage keeps several: delete them after
    each test instead of leaving ~3 GB per test in pytest's retained temp directories."""
    yield
    for root, _, files in os.walk(tmp_path):
        for f in files:
            if f.endswith(".pt"):
                try:
                    os.remove(os.path.join(root, f))
                except OSError:
                    pass


@pytest.fixture
def emulated(monkeypatch):
    """xva_trainer_b200.trainers (and the fastpitch / hifigan modules it drives) as private copies with the CUDA-only checks
    patched out, every C-ABI call executed by the emulator; the GPU test module pointed at the CPU and a small dataset."""
    import xva_trainer_b200

    monkeypatch.setenv("XVA_TRAINER_GRAPH", "0")
    monkeypatch.setattr(T, "DEV", "cpu")
    monkeypatch.setattr(T, "FP_SPEC", "synthetic:2x12x32x4")
    with cabi_emu.installed():
        fp = cabi_emu.load_module("fastpitch", FP_PATCHES)
        hg = cabi_emu.load_module("hifigan", HG_PATCHES)
        for name, mod in (("fastpitch", fp), ("hifigan", hg)):
            monkeypatch.setitem(sys.modules, f"xva_trainer_b200.{name}", mod)
            monkeypatch.setattr(xva_trainer_b200, name, mod, raising=False)
        tr = cabi_emu.load_module("trainers", TR_PATCHES)
        monkeypatch.setitem(sys.modules, "xva_trainer_b200.trainers", tr)
        monkeypatch.setattr(xva_trainer_b200, "trainers", tr, raising=False)
        yield tr


SLOW = pytest.mark.skipif(os.environ.get("XVA_TEST_SLOW", "0") != "1",
                          reason="~50 s on the CPU (every stage writes 0.5 GB checkpoints): XVA_TEST_SLOW=1 runs it; the same test "
                                 "runs on the GPU, and the base-checkpoint test below walks the same stage sequence")


@SLOW
def test_fastpitch_stages_files_and_protocol_strings(emulated, tmp_path, monkeypatch):
    T.test_fastpitch_handle_trainer_stages_and_files(None, tmp_path, monkeypatch)


@SLOW
def test_fastpitch_starts_at_the_aligner_and_extracts_durations(emulated, tmp_path, monkeypatch):
    T.test_fastpitch_handle_trainer_starts_at_the_aligner(None, tmp_path, monkeypatch)


def test_fastpitch_base_checkpoint_precedence_and_stage5_short_cut(emulated, tmp_path, monkeypatch):
    T.test_fastpitch_base_checkpoint_is_only_the_starting_point(None, tmp_path, monkeypatch)


def test_fastpitch_nan_batch_is_skipped_before_the_update(emulated, tmp_path, monkeypatch):
    T.test_fastpitch_nan_batch_is_skipped_before_the_update(None, tmp_path, monkeypatch)


@SLOW
def test_hifigan_trainer_files_and_optimizer_state_resume(emulated, tmp_path, monkeypatch):
    """HiFiTrainer through handleTrainerHiFi on the emulated ABI: g_ / do_ checkpoint files, the exported .hg.pt, websocket
    strings, torch.optim.AdamW-format optimizer state and the resume from it (the two HiFi-GAN tests of the GPU suite; several
    minutes on the CPU: every step is a whole G + MPD + MSD iteration)."""
    T.test_hifigan_handle_trainer_files(None, tmp_path / "files", monkeypatch)
    T.test_hifigan_optimizer_state_is_torch_adamw_format_and_resumes(None, tmp_path / "resume", monkeypatch)
