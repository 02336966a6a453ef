"""CPU-side checks of bench.py's contract that need no GPU: the reference arm under torchrun lets rank 0 alone work, and
without a CUDA device the benchmark fails loudly instead of measuring anything on the CPU."""
import json
import os
import pathlib
import subprocess
import sys

import pytest
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent


def run(args, **env):
    e = dict(os.environ, **env)
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e,
                          timeout=300)


def test_reference_arm_other_ranks_exit_quietly():
    p = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    assert p.returncode == 0 and p.stdout == b""


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_gpu_is_an_error_not_a_cpu_measurement():
    p = run(["--steps", "1", "--warmup", "0"])
    assert p.returncode == 1
    line = json.loads(p.stdout.decode().strip().splitlines()[-1])
    assert "error" in line and "value" not in line
