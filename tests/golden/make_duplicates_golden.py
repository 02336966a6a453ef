#!/usr/bin/env python3
"""Generates tests/golden/duplicates.json in the build container by running the reference's OWN duplicate loop: the
source lines of "Process the last sequence" in /root/reference/src/crackling/Crackling.py (the `for guide in
processSequence(seq)` block, :231-241) are read from the reference tree at run time, dedented and exec'd with a stub
processSequence that yields seeded target23 strings and a stub Batchinator that records what the pipeline keeps.
Nothing of the reference is copied into this repository; only its outputs are committed."""
import json, pathlib, random, textwrap

src = pathlib.Path("/root/reference/src/crackling/Crackling.py").read_text().splitlines()
start = next(i for i, l in enumerate(src) if "# Process the last sequence" in l) + 1
end = start
while end < len(src) and (src[end].strip() == "" or src[end].startswith(" " * 12)):
    end += 1
block = textwrap.dedent("\n".join(src[start:end]))
assert block.startswith("for guide in processSequence(seq):") and "duplicateGuides.add" in block, block

rng = random.Random(20261019)
cases = []
for case in range(6):
    n = [0, 1, 50, 400, 2000, 3000][case]
    pool = ["".join(rng.choice("ACGT") for _ in range(21)) + "GG" for _ in range(max(1, n // (1 + case)))]
    targets = [rng.choice(pool) if rng.random() < 0.6 else "".join(rng.choice("ACGT") for _ in range(23)) for _ in range(n)]

    class Batchinator:
        def __init__(self):
            self.rows = []

        def recordEntry(self, guide):
            self.rows.append(guide[1])

    env = {"processSequence": lambda seq: ([t, i] for i, t in enumerate(seq)), "seq": targets, "candidateGuides": set(),
           "duplicateGuides": set(), "numIdentifiedGuides": 0, "numDuplicateGuides": 0, "guideBatchinator": Batchinator()}
    exec(block, env)
    cases.append({"targets": targets, "recorded": env["guideBatchinator"].rows, "duplicateGuides": sorted(env["duplicateGuides"]),
                  "numDuplicateGuides": env["numDuplicateGuides"], "numIdentifiedGuides": env["numIdentifiedGuides"]})
pathlib.Path(__file__).with_name("duplicates.json").write_text(json.dumps(cases))
print([len(c["targets"]) for c in cases], [c["numDuplicateGuides"] for c in cases])
