#!/usr/bin/env python3
"""Generates tests/golden/prefilters.json in the build container: seeded 23-mers, the reference's own
AT_percentage (imported from /root/reference/src/crackling/Helpers.py) and the filter conditions of
Crackling.py:312-384 evaluated literally."""
import json, pathlib, random, sys

sys.path.insert(0, "/root/reference/src")
from crackling.Helpers import AT_percentage   # noqa: E402

rng = random.Random(20261018)
targets = []
for i in range(400):
    kind = i % 8
    body = "".join(rng.choice("ACGT") for _ in range(23))
    if kind == 0:
        body = body[:21] + "GG"
    elif kind == 1:
        body = "CC" + body[2:]
    elif kind == 2:
        body = "T" + body[1:21] + "GG"
    elif kind == 3:
        body = "CC" + body[2:22] + "A"
    elif kind == 4:
        p = rng.randrange(0, 20)
        body = body[:p] + "TTTT" + body[p + 4:]
        body = body[:23]
    elif kind == 5:
        body = "".join(rng.choice("AT") for _ in range(20)) + body[20:]
    elif kind == 6:
        body = "".join(rng.choice("GC") for _ in range(20)) + body[20:]
    else:
        body = body[:7] + "N" + body[8:]
    targets.append(body)
out = []
for t in targets:
    f = 0
    if t[19] != 'G':
        f |= 1
    if (t[-2:] == 'GG' and t[0] == 'T') or (t[:2] == 'CC' and t[-1] == 'A'):
        f |= 2
    at = AT_percentage(t[0:20])
    if at < 20 or at > 65:
        f |= 4
    if 'TTTT' in t:
        f |= 8
    out.append({"target23": t, "flags": f, "at": at})
pathlib.Path(__file__).with_name("prefilters.json").write_text(json.dumps(out, indent=0))
print(len(out), "vectors")
