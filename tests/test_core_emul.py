"""The product's warp-uniform search state machine (src/x265_b200/csrc/la_core.h), compiled for
the CPU and driven pass by pass exactly as the CUDA kernel drives it (tests/core_emul.cpp), against
the reference's golden traces.  Proves tie-breaking, hexagon walk, sub-pel refine, skip shortcut
and the list/bidir/intra decision without a GPU."""
import os

import pytest

from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_tables():
    assert po.oracle(8, emul=True).emul_check_tables() == 0


@pytest.mark.parametrize("name", ["tiny8", "tiny10", "odd8", "pool3_720p"])
def test_state_machine_replays_reference(name):
    t = po.Trace(os.path.join(GOLD, name + ".trace"))
    r = po.OracleReplay(t, emul=True)
    try:
        mm = r.run()
        assert not mm, mm[:5]
    finally:
        r.close()
