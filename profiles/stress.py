"""Command-line front end of the adversarial sweeps (tests/stress_cases.py) for sweeps larger than the bounded ones the GPU test
tier runs:  python profiles/stress.py [read|paste|geometry|dense_write|objects|fuse|linear ...] [CASES=n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import stress_cases as S  # noqa: E402

DEFAULT = {"read": 36, "paste": 60, "geometry": 12, "dense_write": 300, "objects": 30, "fuse": 0, "linear": 96}
names = [a for a in sys.argv[1:] if a in DEFAULT] or list(DEFAULT)
dev = torch.device("cuda:0")
rc = 0
for nm in names:
    n, bad, msgs = getattr(S, "stress_" + nm)(dev, int(os.environ.get("CASES", DEFAULT[nm])))
    for m in msgs:
        if "MISMATCH" in m or nm == "fuse":
            print(m)
    print(f"{nm} stress: {n} cases, {bad} mismatches")
    rc |= int(bad != 0)
sys.exit(rc)
