"""The cuTree kernel's per-CU source (csrc/x265cu_cutree_core.h: amount in double arithmetic, bipred weighting, bilinear
scatter with frame clipping, saturating adds, zero-row rule) and its phase scheduler, emulated on the CPU and compared
with the oracle's restatement of Lookahead::estimateCUPropagate on random adversarial frames, 8 and 10 bit."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("depth", [8, 10])
def test_cutree_core_emulation_matches_oracle(tmp_path, depth):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    exe = str(tmp_path / ("cutree_emul%d" % depth))
    subprocess.run(["g++", "-O2", "-Wall", "-DORACLE_DEPTH=%d" % depth, "-o", exe, os.path.join(ROOT, "tests", "cutree_emul.cpp"),
                    os.path.join(ROOT, "oracle", "_build", "oracle%d.o" % depth), "-lm"], check=True)
    r = subprocess.run([exe, "40"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout
    assert r.stdout.startswith("ok:")
