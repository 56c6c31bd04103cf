"""Encoder-level parity: the real x265 1.9 CLI with its lookahead running on the GPU
(oracle/_ref/x265_cu8: stock encoder objects + integration/x265_glue.cpp + libx265cu.so) must write
the SAME BITSTREAM as the stock CLI (oracle/_ref/x265_ref8) -- every slice-type decision, cuTree
QP offset and rate-control input downstream of the lookahead outputs is then identical.
Golden md5s of the stock CLI are committed (tests/golden/cli_md5.json, harness/x265_cli.py): 720p medium, 360p fast/b-adapt 1,
--weightb, VBV + ABR, and the 10-bit build; only option sets for which the stock binary is timing-stable (see harness/x265_cli.py)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "cli_md5.json")))


def _have(kind):
    from harness import x265_cli
    return os.path.exists(x265_cli.binary(kind, 8))


@pytest.mark.parametrize("name", sorted(GOLD))
def test_stock_cli_matches_golden(name):
    """the stock reference binary reproduces the committed md5 (pins the golden and the clip generator)"""
    if not _have("ref"):
        pytest.skip("oracle/_ref/x265_ref8 not built (reference tree absent)")
    from harness import x265_cli
    md5, size, _ = x265_cli.run_case("ref", name)
    assert (md5, size) == (GOLD[name]["md5"], GOLD[name]["bytes"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_gpu_lookahead_cli_bitstream_identical(name, built):
    if not _have("cu"):
        pytest.skip("oracle/_ref/x265_cu8 not built (reference tree absent at build time)")
    from harness import x265_cli
    md5, size, _ = x265_cli.run_case("cu", name)
    assert (md5, size) == (GOLD[name]["md5"], GOLD[name]["bytes"]), "bitstream differs from the stock x265 1.9 encoder"
