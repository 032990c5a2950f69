"""The example scripts (the reference's evaluation and training loops on this library) run end to end."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args):
    r = subprocess.run([sys.executable] + args, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_evaluate_example():
    out = _run(["examples/evaluate_synthetic.py", "--volumes", "2", "--slices", "9", "--batch", "8", "--mode", "fp32"])
    assert "'slices': 18" in out and "'dice'" in out


def test_train_example(tmp_path):
    ck = str(tmp_path / "best.pth")
    out = _run(["examples/train_head_synthetic.py", "--steps", "30", "--batch", "4", "--size", "64", "--save", ck])
    losses = [float(l.split("loss")[1].split()[0]) for l in out.splitlines() if l.startswith("step")]
    assert losses[-1] < losses[0]
    import torch
    sd = torch.load(ck)
    assert "decoder.0.0.cv1.conv.weight" in sd and "encoder.0.conv.weight" in sd and "output.bias" in sd


def test_overlapped_schedule_is_bit_identical_to_serial():
    """tools/soak.py: multi-stream + lanes + PDL pipeline vs the serial schedule, repeated with alternating inputs."""
    env = dict(os.environ, SOAK_B="48", SOAK_ITERS="12")
    r = subprocess.run([sys.executable, "tools/soak.py"], cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "bit-identical" in r.stdout
