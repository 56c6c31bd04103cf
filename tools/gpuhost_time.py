"""Wall time of x265's own lookahead with the GPU path bound in (libx265gpu) or of the unmodified reference (libx265ref).
usage: gpuhost_time.py <workload> [runs] [gpu|ref] [pool]"""
import sys, os
sys.path.insert(0, os.getcwd())
from harness import x265host as xh
from harness.workloads import WORKLOADS
name = sys.argv[1]
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 5
gpu = (sys.argv[3] if len(sys.argv) > 3 else "gpu") == "gpu"
depth, w, h, n, seed, pool, opts, _ = WORKLOADS[name]
if len(sys.argv) > 4:
    pool = int(sys.argv[4])
d = xh.LaDriver(depth, w, h, n, seed, opts, pool, gpu)
for i in range(runs):
    secs, types, stats = d.run()
    print(name, "gpu" if gpu else "ref", "pool", pool, "run", i, "ms %.2f fps %.1f" % (secs * 1e3, n / secs), flush=True)
d.close()
