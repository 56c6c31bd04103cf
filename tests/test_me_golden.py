"""Full-resolution PU motion search (encoder/motion.cpp:571-1172): the oracle's restatement against the
committed results of the unmodified reference (tests/golden/me_golden.json, written by oracle/gen_me_golden.py) -- the pin that
travels to machines without /root/reference.  Every integer pattern (DIA, HEX, UMH, STAR, FULL) x every sub-pel level, the 24
inter PU shapes, windows that cut the patterns, far motion (UMH's grid, STAR's raster refinement)."""
import json
import os

import pytest

from harness import me_cases as mc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "me_golden.json")


@pytest.mark.parametrize("depth", [8, 10])
def test_oracle_motion_search_equals_reference_results(depth):
    gold = json.load(open(GOLDEN))["depths"][str(depth)]
    cases = mc.golden_cases()
    assert len(gold) == len(cases)
    O = mc.oracle_me(depth)
    lut = mc.make_lut(20)
    moved = far = 0
    for case, want in zip(cases, gold):
        got = mc.oracle_run(O, case, depth, lut)
        want = [tuple(r) for r in want]
        assert got == want, (case, [(i, a, b) for i, (a, b) in enumerate(zip(got, want)) if a != b][:3])
        moved += sum(1 for r in want if (r[0], r[1]) != (0, 0))
        far += sum(1 for r in want if abs(r[0]) > 80)
    assert moved > 1000 and far > 20
