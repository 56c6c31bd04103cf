import sys, os, time
sys.path.insert(0, os.getcwd())
from harness import x265host as xh
from harness.workloads import WORKLOADS
names = sys.argv[1:] or ["tiny8"]
for name in names:
    depth, w, h, n, seed, pool, opts, _ = WORKLOADS[name]
    d = xh.LaDriver(depth, w, h, n, seed, opts, pool, True)
    for level in (2, 1, 0, 0, 0):
        tr = "gpurun_out/%s_gpu_l%d.trace" % (name, level) if level else None
        secs, types, stats = d.run(trace=tr, level=level)
        mm = xh.compare_traces(tr, "tests/golden/%s.trace" % name, level) if tr else None
        print(name, "level", level, "secs %.4f fps %.1f" % (secs, n / secs), stats, "mismatches:", mm if mm is None else len(mm), flush=True)
        if mm:
            for x in mm[:10]: print("   ", x)
    d.close()
