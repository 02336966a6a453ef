"""Guide-side pre-filters of the pipeline: the oracle against the reference-made vectors
(CPU), the device pass against the oracle (GPU)."""
import json
import pathlib

import numpy as np
import pytest

import crackling_b200 as cb
from oracle import prefilters as pf

GOLDEN = json.loads((pathlib.Path(__file__).parent / "golden" / "prefilters.json").read_text())


def test_oracle_matches_reference_vectors():
    assert len(GOLDEN) >= 400
    seen = 0
    for v in GOLDEN:
        f, at = pf.filters(v["target23"])
        assert f == v["flags"] and at == v["at"]
        seen |= f
    assert seen == 15, "every filter must be exercised by the vectors"


def test_oracle_at_percentage_matches_reference_helper():
    ref = pathlib.Path("/root/reference/src/crackling/Helpers.py")
    if not ref.exists():
        pytest.skip("reference tree not present (GPU box)")
    import sys
    sys.path.insert(0, "/root/reference/src")
    from crackling.Helpers import AT_percentage
    rng = np.random.default_rng(5)
    for _ in range(500):
        s = "".join(rng.choice(list("ACGTN"), 20))
        assert pf.at_percentage(s) == AT_percentage(s)


@pytest.mark.gpu
def test_device_filters_match_oracle():
    from conftest import golden_case
    case = golden_case("w8_families")
    dev = cb.Device.from_index(cb.Index(case.issl), 0, "res32")
    rng = np.random.default_rng(6)
    targets = [v["target23"] for v in GOLDEN]
    targets += ["".join(rng.choice(list("ACGT"), 23)) for _ in range(5000)]
    text = ("\n".join(targets) + "\n").encode()
    flags, at, packed = dev.guide_filters(text)
    for i, t in enumerate(targets):
        f, a = pf.filters(t)
        assert flags[i] == f and at[i] == a, t
        assert int(packed[i]) == pf.encode20(t)
    assert np.array_equal(packed, cb.pack_guides(("\n".join(t[:20] for t in targets) + "\n").encode()))
    # the survivors go straight to the scorer, no temp file in between (ref Crackling.py:747-778)
    keep = packed[flags == 0][:64]
    mit, cfd = dev.score(keep, 4, 0, "and")
    assert mit.size == keep.size and np.all(mit > 0) and np.all(cfd > 0)
    with pytest.raises(cb.IsslError):
        dev.guide_filters(text[:-1])
    dev.close()
