"""The cuTree kernel's op scheduler (csrc/x265cu_cutree_sched.h) on the CPU: random op lists executed in the given order
and as phases of mutually independent ops (any order inside a phase) must leave identical accumulators and outputs."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scheduler_preserves_sequential_semantics(tmp_path):
    exe = str(tmp_path / "cutree_sched_test")
    subprocess.run(["g++", "-O2", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cutree_sched_test.cpp")], check=True)
    r = subprocess.run([exe, "20000"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout
    assert r.stdout.startswith("ok:")
    # the point of the scheduler: far fewer phases than ops
    ops, phases = int(r.stdout.split()[4]), int(r.stdout.split()[7])
    assert phases < ops
