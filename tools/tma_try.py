import os, sys
sys.path.insert(0, os.getcwd())
os.environ["X265CU_PLAIN_TMA"] = "1"; os.environ["X265CU_SEARCH_MODE"] = "0"
from harness import replay
from oracle import pyoracle as po
t = po.Trace(replay.trace_path("tiny8"))
r = replay.CuReplay(t, device=0, check=True)
try:
    mm = r.run()
    print("mismatches", len(mm), mm[:2])
except Exception as e:
    print("EXC", e)
