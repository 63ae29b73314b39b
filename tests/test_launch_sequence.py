"""The Python host code still emits exactly the launches (entry points, scalar arguments, GEMM argument blocks) of the
tree whose GPU parity suite was last green -- see tests/launch_sequence.py. CPU only."""
import json

from launch_sequence import GOLDEN, fingerprint, record


def test_host_code_emits_the_gpu_validated_launch_sequence():
    want = json.load(open(GOLDEN))
    seq = record()
    got = fingerprint(seq)
    if got["sha256"] != want["sha256"]:
        first = next((i for i, (a, b) in enumerate(zip(got["names"], want["names"])) if a != b), None)
        bad_chunk = next((i for i, (a, b) in enumerate(zip(got["chunks"], want["chunks"])) if a != b), None)
        raise AssertionError(
            f"launch sequence changed: {got['calls']} calls (validated: {want['calls']}); first differing entry point at call "
            f"{first}; first differing 50-call chunk {bad_chunk} (calls {None if bad_chunk is None else bad_chunk * 50}..). "
            "If the change is intended, re-run the GPU parity suite and then `python tests/launch_sequence.py --write`.")


def test_stream_experiments_do_not_change_any_launch():
    """XVA_BWD_STREAMS / XVA_DISC_STREAMS / XVA_GEN_STREAMS (off by default) only choose the stream a
    launch goes to: with all three on, every entry point and every argument is the same as in the validated sequence."""
    want = json.load(open(GOLDEN))
    got = fingerprint(record(streams=True))
    assert got["calls"] == want["calls"] and got["sha256"] == want["sha256"]
