"""xVAPitch pitch predictor (textenc.RelativePositioningPitchEnergyEncoder) on the device vs the recording of the unmodified
reference module (tests/golden/vits_pitch_predictor.npz) and the CPU oracle + autograd.

STATUS: written after the round's GPU budget was spent -- these checks have NOT run on hardware yet. What has run:
  * the module's host code on the CPU through the emulated C ABI against the same golden and oracle
    (tests/test_vits_text_encoder_cpu.py::test_pitch_predictor_*), the method that predicted 12 / 12 for the text encoder
    (profiles/r02_textenc_gpu_tests.log);
  * on the B200: every kernel it launches -- the transformer layers are the text encoder's (tests/test_vits_text_encoder_gpu.py),
    LayerNorm at 708 / 780 channels (tests/test_rowops_gpu.py, profiles/r02_layernorm_1024_gpu_tests.log), the N = 1 / M = 1
    tap-GEMM shapes of the projection (HiFi-GAN conv_post, tests/test_hifigan_gpu.py).
Two precautions for a first run that happens unattended: the device work runs in a CHILD process
(tests/pitch_predictor_gpu_probe.py) that is killed after 240 s, so neither a CUDA fault nor a hang can reach the rest of the
suite; and the tests are marked xfail(strict=False): a failure is reported as xfailed, a pass as xpassed. Remove both after
the first green run."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="first hardware run pending: written after the round's GPU budget was spent; "
                                                     "host code verified on the CPU through the emulated C ABI")]

_cache = {}


def probe():
    if "out" not in _cache:
        _cache["out"] = None
        try:
            res = subprocess.run([sys.executable, os.path.join(HERE, "pitch_predictor_gpu_probe.py")], capture_output=True,
                                 text=True, timeout=240)
            lines = [ln for ln in res.stdout.splitlines() if ln.startswith("PITCH_PREDICTOR_PROBE ")]
            if res.returncode == 0 and lines:
                _cache["out"] = json.loads(lines[-1].split(" ", 1)[1])
            else:
                _cache["err"] = (res.stderr or res.stdout)[-1500:]
        except subprocess.TimeoutExpired:
            _cache["err"] = "the probe did not finish within 240 s (killed)"
    if _cache["out"] is None:
        pytest.fail("pitch-predictor probe failed: " + _cache.get("err", "?"))
    return _cache["out"]


@pytest.mark.parametrize("i", [0, 1])
def test_wiring_exact_with_fp32_checker_gemm(lib, i):
    e = probe()["exact"][i]
    assert e["same_keys"] and e["fwd"] < 2e-5 and e["grad_worst"] < 2e-4, e


@pytest.mark.parametrize("i", [0, 1])
def test_product_path_matches_the_oracle(lib, i):
    """Bounds = those of the text encoder's product path (measured there: forward 6e-4, gradient vector 9e-3, worst tensor
    2.9e-2, profiles/r02_textenc.txt); to be replaced by 2 x this module's own measurements after its first run."""
    e = probe()["product"][i]
    assert e["same_keys"] and e["pad_max"] == 0.0, e
    assert e["fwd"] < 3e-3 and e["grad_global"] < 2e-2 and e["grad_worst"] < 6e-2, e


def test_forward_matches_the_reference_golden_and_dead_state_survives_adamw(lib):
    o = probe()
    assert o["golden_fwd"] < 3e-3, o["golden_fwd"]
    assert o["keys_in_reference_order"] and o["dead_untouched"] and o["moved"] == o["trainable"], o
