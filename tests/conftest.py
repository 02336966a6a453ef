"""pytest configuration: the `gpu` marker, import paths and shared fixtures."""
import json
import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GOLDEN = ROOT / "tests" / "golden"
GOLDEN_CASES = sorted(p.name for p in GOLDEN.iterdir() if (p / "offtargets.txt").exists())


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class GoldenCase:
    """One committed fixture: inputs + what the unmodified reference printed for them."""

    def __init__(self, name: str):
        self.name = name
        d = GOLDEN / name
        self.offtargets = (d / "offtargets.txt").read_bytes()
        self.guides = (d / "guides.txt").read_bytes()
        self.expected = json.loads((d / "expected.json").read_text())
        self.seq_length = self.expected["seq_length"]
        self.slice_width = self.expected["slice_width"]
        self._img = None

    @property
    def issl(self) -> bytes:
        """The .issl image, rebuilt by the oracle's restatement of isslCreateIndex."""
        if self._img is None:
            from oracle import oracle
            self._img = oracle.create_index(self.offtargets, self.seq_length, self.slice_width)
        return self._img


_cases = {}


def golden_case(name: str) -> GoldenCase:
    if name not in _cases:
        _cases[name] = GoldenCase(name)
    return _cases[name]


@pytest.fixture(params=GOLDEN_CASES)
def golden(request) -> GoldenCase:
    return golden_case(request.param)
