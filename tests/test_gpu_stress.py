"""The adversarial sweeps of tests/stress_cases.py with bounded case counts, under `-m gpu`: geometry knife edges, read patterns
(checkerboards, stripes, signed zeros), mask-pasting thresholds and non-finite probabilities, every dense-write kernel on run
structures random data never produces, the object regime with K objects per pixel / 140 objects, and the persistent tcgen05
projection against the tile-per-CTA kernel.  (Round 1 ran these only from profiles/ on the builder's lease.)"""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import stress_cases as S  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,cases", [("geometry", 8), ("read", 24), ("paste", 24), ("dense_write", 150), ("objects", 24), ("fuse", 0), ("linear", 36)])
def test_stress_sweep(cuda, name, cases):
    n, bad, msgs = getattr(S, "stress_" + name)(cuda, cases)
    assert n > 0
    assert bad == 0, "\n".join(m for m in msgs if "MISMATCH" in m or name == "fuse")[:4000]
