#!/usr/bin/env python3
"""bench.py -- lookahead frames/s of the B200 lookahead cost-estimation path (BASELINE.json metric).

A "step" = one pass of the hot path over one synthetic clip: every pre-lookahead frame (lowres
init + AQ variance + intra estimate) and every frame-cost estimate (motion searches, bidir/intra
costs, weighted-prediction analysis) that x265 1.9's own lookahead issues for that clip, replayed in
the reference's order and batching from the golden trace (tests/golden/<workload>.trace), through
the host layer (x265cu::Lookahead / CostEstimateGroup) and the C ABI.

  value   device-resident: source pictures already in HBM, result arrays stay in HBM mirrors
          (only the per-estimate sums return); the host float decisions still run.
  e2e     the reference-facing call path with HOST buffers, with x265 1.9's OWN lookahead as the host
          (oracle/_ref/libx265gpu<depth>.so: the reference's objects with slicetype.cpp / lowres.cpp /
          picyuv.cpp bound to libx265cu.so, integration/): pictures are uploaded from x265's pinned
          PicYuv, every Lowres output array (planes, MVs, costs...) lands in x265's own arrays, and
          x265's slicetypeDecide / scenecut / slicetypePath / cuTree control flow run inside the clock.
  configs the same measurements for the other BASELINE configs (4K 8-bit, 4K 10-bit), N = 1.
  --impl reference   the UNMODIFIED x265 1.9 lookahead (oracle/_ref, C primitives) on the host
          cores, same clip and options, all threads.

One process per GPU (torchrun for N > 1): every rank runs its own independent stream (weak
scaling, no collective on the cost path); timing is barrier + synchronize on both sides, CUDA
events on the launching stream, MAX over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """`--impl reference`: the reference's own CPU lookahead on the host cores (rank 0 only; ONE CPU stream whatever --gpus)."""
    if rank != 0:
        return 0
    from harness.workloads import WORKLOADS, DESCRIPTIONS
    depth, w, h, nframes, seed, _pool, opts, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    pool = min(64, cores)          # one x265 thread pool holds at most 64 workers (threadpool.h:44)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "harness", "refrun.py"), args.workload, str(pool), str(args.warmup + args.steps), "ref"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0 or not r.stdout.strip():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref lookahead driver failed: " + r.stderr.strip()[-200:]}))
        return 0
    out = json.loads(r.stdout.strip().splitlines()[-1])
    kind = out["kind"]
    times = out["seconds"][args.warmup:] or out["seconds"]
    total = sum(times)
    value = nframes * len(times) / total
    line = {
        "impl": "reference", "metric": "lookahead_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8" if depth == 8 else "u16", "data": "synthetic",
        "config": {"workload": DESCRIPTIONS.get(args.workload, args.workload), "trace": args.workload, "frames_per_step": nframes,
                   "resolution": "%dx%d" % (w, h), "bit_depth": depth,
                   "streams": "1 CPU stream on all host cores whatever --gpus (a throughput baseline, not a per-stream ratio at N > 1)"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": pool, "kind": kind,
                         "sample": "whole workload (%d frames) per step, x265 1.9 Lookahead only (no frame encoders; harness/x265_la_driver.cpp, the "
                                   "driver the GPU arm's e2e uses), C primitives (no asm: no yasm/nasm in the image), pool of %d threads" % (nframes, pool)},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
class Runner:
    """pre-marshalled replay of a trace through the host layer (resident or host-buffer mode)."""

    def __init__(self, trace, clip, stream, device, resident, torch):
        from src.x265_b200 import abi
        self.abi, self.torch, self.resident = abi, torch, resident
        cfg = trace.cfg
        self.cfg = cfg
        slices = cfg["numCoopSlices"] if cfg["numCoopSlices"] > 1 else 0
        n = cfg["nframes"]
        la = None
        for s in ([slices] + list(range(2, 17)) if slices else [0]):
            la = abi.Lookahead(cfg["width"], cfg["height"], cfg["depth"], cfg["bframes"], cfg["lookahead"], s, cfg["pool"], cfg["weightp"],
                               cfg["aqmode"], cfg["aqStrength"], cfg["bFrameBias"], device, n + 2, stream=stream,
                               search_warps=env_int("X265CU_SEARCH_ROWS", 0), fps_num=cfg.get("fpsNum", 30), fps_denom=cfg.get("fpsDenom", 1),
                               qcompress=cfg.get("qCompress", 0.6), weighted_bipred=cfg.get("weightedBiPred", 0))
            if not slices or (la.numCoopSlices, la.numRowsPerSlice) == (cfg["numCoopSlices"], cfg["numRowsPerSlice"]):
                break
            la.close()
        self.la = la
        la.set_resident(resident)
        self.frames = {t: la.frame_alloc() for t in range(n)}
        self.inputs = {}
        self.keepalive = []
        dt = torch.uint8 if cfg["depth"] == 8 else torch.int16
        for t in range(n):
            y, u, v = clip.frames[t]
            if resident:
                # device copies with 8-byte aligned pitches (x265cu_frame_init device-pointer contract)
                def dev(a):
                    pitch = (a.shape[1] * a.itemsize + 63) // 64 * 64 // a.itemsize
                    d = torch.zeros((a.shape[0], pitch), dtype=dt, device="cuda")
                    d[:, :a.shape[1]] = torch.from_numpy(a.view("int16") if cfg["depth"] > 8 else a).to("cuda")
                    return d, pitch
                dy, py = dev(y)
                du, pu = dev(u)
                dv, _ = dev(v)
                self.keepalive += [dy, du, dv]
                self.inputs[t] = (dy.data_ptr(), py, du.data_ptr(), dv.data_ptr(), pu)
            else:
                # pinned host pictures (the encoder's PicYuv, registered once)
                for a in (y, u, v):
                    b = a.base if a.base is not None else a     # the picture is a view into its margin-padded buffer
                    abi.lib_cu().x265cu_host_register(b.ctypes.data, b.nbytes)
                self.keepalive += [y, u, v]
                self.inputs[t] = (y.ctypes.data, y.strides[0] // y.itemsize, u.ctypes.data, v.ctypes.data, u.strides[0] // u.itemsize)
        # pre-marshal the call sequence
        import ctypes as C
        cutree = env_int("X265CU_BENCH_CUTREE", 1) != 0
        self.prefetch = env_int("X265CU_BENCH_PREFETCH", 0) != 0
        self.npropagate = 0
        self.calls = []
        for e in trace.events:
            if e[0] == "P":
                # consecutive pre-lookahead frames are one PreLookaheadGroup list (slicetype.cpp:831-856)
                t = e[1]["poc"]
                if self.calls and self.calls[-1][0] == "P":
                    self.calls[-1][1].append(t)
                else:
                    self.calls.append(("P", [t]))
            elif e[0] in ("J", "B"):
                jobs = [e[1]] if e[0] == "J" else e[1]
                if not jobs:
                    continue
                lo = min(j["p0"] for j in jobs)
                hi = max(j["p1"] for j in jobs)
                fr = [self.frames.get(p) for p in range(lo, hi + 1)]
                tr = [(j["p0"] - lo, j["p1"] - lo, j["b"] - lo) for j in jobs]
                self.calls.append(("E", la.prepare_estimate(fr, tr), e[0] == "B"))
            elif e[0] in ("M", "X", "F") and cutree:
                # cuTree: the memsets, propagate steps and cuTreeFinish calls of Lookahead::cuTree, in place;
                # a run of them is marshalled once (like the estimates) so that Python's per-call cost stays out
                if not (self.calls and self.calls[-1][0] == "T"):
                    self.calls.append(("T", []))
                if e[0] == "M":
                    self.calls[-1][1].append(("M", self.frames[e[1]]))
                elif e[0] == "X":
                    x = e[1]
                    self.calls[-1][1].append(("X", self.frames[x["b"]], self.frames[x["p0"]], self.frames[x["p1"]], x["b"] - x["p0"],
                                              x["p1"] - x["b"], x["referenced"], x["avgDuration"]))
                    self.npropagate += 1
                else:
                    self.calls[-1][1].append(("F", self.frames[e[1]["poc"]], e[1]["avgDuration"], e[1]["ref0Distance"]))
        self.calls = [("T", la.prepare_cutree_sequence(c[1])) if c[0] == "T" else c for c in self.calls]
        self.calls = [("P", c[1], la.prepare_pre_lookahead_batch([(self.frames[t],) + tuple(self.inputs[t]) + (t,) for t in c[1]])) if c[0] == "P" else c
                      for c in self.calls]
        # each P call also carries the prepared list that follows it (None for the last)
        pidx = [i for i, c in enumerate(self.calls) if c[0] == "P"]
        for k, i in enumerate(pidx):
            nxt = self.calls[pidx[k + 1]][2] if k + 1 < len(pidx) else None
            self.calls[i] = self.calls[i] + (nxt,)
        self.units = sum(j["s0"] + j["s1"] for j in trace.jobs())
        self.njobs = sum(1 for _ in trace.jobs())

    def step(self):
        la = self.la
        for c in self.calls:
            if c[0] == "P":
                la.pre_lookahead_batch_prepared(c[2], True)
                # X265CU_BENCH_PREFETCH=1: the pictures of the NEXT list arrive in the input queue while this one's slicetypeDecide
                # runs (Lookahead::addPicture) and their uploads start now.  Off by default: measured 31.9 ms against 31.4 ms per
                # step -- the bulk uploads share the host-to-device copy engine with the small argument uploads of the estimate
                # batches, which then wait behind them on the critical path.
                if self.prefetch and c[3] is not None and not self.resident:
                    la.add_pictures_prepared(c[3])
            elif c[0] == "E":
                la.estimate_prepared(c[1], c[2])
            else:
                la.cutree_sequence_prepared(c[1])
        la.sync()          # every output, including the asynchronous plane copy-backs, is on the host

    def close(self):
        for f in self.frames.values():
            self.la.frame_free(f)
        if not self.resident:
            for a in self.keepalive:
                b = a.base if a.base is not None else a
                self.abi.lib_cu().x265cu_host_unregister(b.ctypes.data)
        self.la.close()


def multi_stream(torch, dist, world, rank, workload, device, nstreams, steps):
    """BASELINE configs[4]: nstreams independent encoder streams per GPU (64 over 8 GPUs), each an x265 Lookahead of its own
    (own thread pool, own x265cu context, own CUDA stream) bound to libx265cu.so -- the e2e arm, several at a time.  A single
    stream is bound by the dependent wavefront chain of its estimates and leaves most SMs idle, so streams overlap.  Streams
    are dealt to the ranks by harness/sharding.py; every stream's decisions (slice type of every picture) are digested and
    the digests gathered once: each must equal the digest of the reference's decisions (golden trace).  Wall clock between
    device-wide synchronizes around `steps` passes of every stream, max over ranks."""
    import threading
    from harness import sharding, replay
    from oracle import pyoracle as po
    total = nstreams * world
    mine = sharding.assign(total, world, rank)
    trace = po.Trace(replay.trace_path(workload))
    n = trace.cfg["nframes"]
    want = [0] * n
    for e in trace.events:
        if e[0] == "D":
            want[e[1]] = e[2]
    expected = sharding.digest(want)
    hosts = [X265Host(workload, device) for _ in mine]
    errs, last = [], {}

    def work(k, count):
        try:
            for _ in range(count):
                _, types, _ = hosts[k].d.run()
            last[mine[k]] = (n, sharding.digest(types))
        except Exception as ex:   # pragma: no cover
            errs.append(str(ex))

    def run_all(count):
        th = [threading.Thread(target=work, args=(k, count)) for k in range(len(hosts))]
        for t in th:
            t.start()
        for t in th:
            t.join()

    run_all(2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_all(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    for h in hosts:
        h.close()
    merged = sharding.gather_results(dist if world > 1 else None, last)
    bad = sorted(sid for sid, (_, dg) in merged.items() if dg != expected)
    return {"streams_per_gpu": nstreams, "streams": total, "value": total * steps * n / dt, "unit": "frames/s", "steps": steps, "seconds": dt,
            "what": "x265 1.9 Lookahead per stream bound to libx265cu.so (e2e: host pictures in, every Lowres array back), %d streams on %d GPU(s)" % (total, world),
            "timing": "wall clock between device-wide synchronizes around all passes of all streams, max over ranks",
            "parity": ("every stream's slice-type decisions == the reference's (digest %08x, %d streams gathered)" % (expected, len(merged)))
                      if not bad and len(merged) == total else "FAILED for streams %r (%d of %d gathered)" % (bad, len(merged), total),
            "errors": errs}


def pin_rank(local, nlocal):
    """Several ranks share the box's host cores: give each rank its own share (whole cores: hyper-thread siblings stay
    together), as a deployment with one encoder process per GPU would.  The library sizes its helper threads from
    the affinity mask it finds."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        groups, seen = [], set()
        for c in cpus:
            if c in seen:
                continue
            try:
                sib = open("/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list" % c).read().strip()
                g = set()
                for part in sib.split(","):
                    lo, _, hi = part.partition("-")
                    g |= set(range(int(lo), int(hi or lo) + 1))
                g &= set(cpus)
            except (OSError, ValueError):
                g = {c}
            g = g or {c}
            seen |= g
            groups.append(sorted(g))
        per = len(groups) // nlocal
        if per < 1:
            return
        mine = [c for g in groups[local * per:(local + 1) * per] for c in g]
        os.sched_setaffinity(0, mine)
    except (AttributeError, OSError):
        pass


def timed(torch, dist, world, fn, steps):
    """barrier + synchronize on both sides; CUDA events on the current (launching) stream; max over ranks"""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


class X265Host:
    """e2e arm: x265 1.9's OWN Lookahead (slicetypeDecide, slicetypeAnalyse, scenecut, slicetypePath, cuTree control flow, thread
    pool) as the host of the GPU path -- oracle/_ref/libx265gpu<depth>.so = the reference's objects with slicetype.cpp / lowres.cpp /
    picyuv.cpp bound to libx265cu.so (integration/).  Same lookahead-only driver, clip and pool size as `--impl reference`."""

    def __init__(self, workload, device):
        from harness import x265host as xh
        from harness.workloads import WORKLOADS
        self.xh = xh
        depth, w, h, n, seed, pool, opts, _ = WORKLOADS[workload]
        if not xh.available(depth, True):
            raise RuntimeError("oracle/_ref/libx265gpu%d.so is not built (integration/build_x265_cu.py needs the reference tree at build time)" % depth)
        os.environ["X265CU_DEVICE"] = str(device)
        self.workload, self.n, self.pool = workload, n, pool
        self.d = xh.LaDriver(depth, w, h, n, seed, opts, pool, True)

    def totals(self):
        import ctypes as C
        t = (C.c_longlong * 8)()
        self.d.L.x265glue_totals(t)
        return list(t)

    def parity(self, golden):
        import tempfile
        fd, path = tempfile.mkstemp(suffix=".trace")
        os.close(fd)
        try:
            _, _, stats = self.d.run(trace=path, level=1)
            mm = self.xh.compare_traces(path, golden, 1)
        finally:
            os.remove(path)
        return mm, stats

    def step(self):
        return self.d.run()[0]

    def close(self):
        self.d.close()


def hbm_peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    return (float(peaks.get("hbm_gbs", 6650.0)),
            "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)")


def measure(args, workload, torch, dist, world, local, rank, headline):
    """every number of one workload: parity gates, value (resident), e2e (x265 as the host), kernel times, rooflines"""
    from harness import replay
    from harness.workloads import DESCRIPTIONS
    from oracle import pyoracle as po
    from src.x265_b200 import abi

    trace = po.Trace(replay.trace_path(workload))
    cfg = trace.cfg
    nframes = cfg["nframes"]
    clip = replay.Clip(cfg)
    out = {"trace": workload, "workload": DESCRIPTIONS.get(workload, workload), "resolution": "%dx%d" % (cfg["width"], cfg["height"]),
           "bit_depth": cfg["depth"], "frames_per_step": nframes}

    # ---- parity gate on this very workload, through the replay call path (every array CRC of every estimate and cuTree step) ----
    parity = "skipped"
    if not args.no_parity:
        r = replay.CuReplay(trace, device=local, check=True, clip=clip)
        mm = r.run()
        nchk, nprop = r.njobs, r.npropagate
        r.close()
        if mm:
            raise SystemExit("bench.py: parity FAILED on %s: %d mismatches, first %r" % (workload, len(mm), mm[0]))
        parity = "bit-exact vs x265 1.9 reference trace: %d frames, %d estimates, %d cuTree propagate steps, every output array CRC" % (nframes, nchk, nprop)
    out["parity"] = parity

    stream = torch.cuda.current_stream().cuda_stream
    sampler = ClockSampler(local)

    # ---- value: device-resident replay ----
    res = Runner(trace, clip, stream, local, True, torch)
    for _ in range(args.warmup):
        res.step()
    res.la.stats_enable(True)
    res.la.stats(reset=True)
    sampler.start()
    ms_value = timed(torch, dist, world, res.step, args.steps)
    st = res.la.stats(reset=True)
    res.la.stats_enable(False)
    int_peak = (abi.C.c_double(), abi.C.c_double())
    abi.lib_cu().x265cu_int_peak(res.la.ctx, abi.C.byref(int_peak[0]), abi.C.byref(int_peak[1]))
    # SATD primitive throughput (the HBM-bound kernel).  One launch reads every one of the 4 planes (full-pel + the three
    # half-pel planes lowresMC reads) of every frame exactly ONCE: disjoint pairs (2t, 2t+1) per plane, so nothing is re-read
    # inside a launch; the L2 is flushed before every timed launch (a 512 MB buffer is read), so every byte comes from HBM.
    satd = None
    try:
        import numpy as np
        slots = np.array([abi.lib_host().x265cuh_frame_slot(res.frames[t]) for t in range(nframes)], np.int32)
        ev, od = slots[0:nframes - 1:2], slots[1:nframes:2]
        a = np.ascontiguousarray(np.concatenate([ev] * 4))
        b = np.ascontiguousarray(np.concatenate([od] * 4))
        pl = np.ascontiguousarray(np.repeat(np.arange(4, dtype=np.int32), len(ev)))
        flush = torch.zeros(128 << 20, dtype=torch.int32, device="cuda")     # 512 MB
        msf = abi.C.c_float()
        times = []
        for _ in range(7):
            flush.sum()                   # a READ of 512 MB: the L2 ends up full of clean lines of another buffer (a write would
            torch.cuda.synchronize()      # leave dirty lines whose write-back competes with the timed kernel's reads)
            abi.lib_cu().x265cu_pixelcmp_planes(res.la.ctx, 1, len(a), a.ctypes.data, pl.ctypes.data, b.ctypes.data, pl.ctypes.data, None, abi.C.byref(msf))
            times.append(msf.value)
        del flush
        med = sorted(times[2:])[len(times[2:]) // 2]
        pix = len(a) * res.la.nCU * 64
        satd = {"gpix_per_s": pix / (med * 1e-3) / 1e9, "pairs": len(a), "ms": med,
                "method": "one launch over the 4 planes of all frames, every plane read once, L2 flushed before each of 5 timed launches (median)",
                "bytes_per_launch": int(pix * 2 * (1 if cfg["depth"] == 8 else 2) + len(a) * res.la.nCU * 4)}
    except Exception as ex:  # pragma: no cover - reported, not fatal
        satd = {"error": str(ex)}
    ct_value = res.la.cutree_stats()
    npropagate = res.npropagate
    units, njobs = res.units, res.njobs
    res.close()

    # ---- e2e: x265's own lookahead as the host (host pictures in, every Lowres array + planes back, host decisions in the clock) ----
    e2e = None
    try:
        host = X265Host(workload, local)
    except RuntimeError as ex:
        host = None
        e2e = {"value": None, "unit": "frames/s", "error": str(ex)}
    if host:
        mmh, hstats = ([], None) if args.no_parity else host.parity(replay.trace_path(workload))
        if mmh:
            raise SystemExit("bench.py: x265-host parity FAILED on %s: %d differences, first %r" % (workload, len(mmh), mmh[0]))
        for _ in range(max(args.warmup - (0 if args.no_parity else 1), 1)):
            host.step()
        t0 = host.totals()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        secs = [host.step() for _ in range(args.steps)]
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        t1 = host.totals()
        tt = torch.tensor([sum(secs), wall], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        lookahead_s, wall_s = float(tt[0].item()), float(tt[1].item())
        host.close()
        e2e = {"value": world * nframes * args.steps / lookahead_s, "unit": "frames/s", "ms_per_step": 1e3 * lookahead_s / args.steps,
               "h2d_bytes_per_step": (t1[1] - t0[1]) // args.steps, "d2h_bytes_per_step": (t1[2] - t0[2]) // args.steps,
               "gpu_launches_per_step": (t1[3] - t0[3]) // args.steps,
               "host": "x265 1.9's own Lookahead (slicetypeDecide / slicetypeAnalyse / scenecut / slicetypePath / cuTree control flow / thread pool of "
                       "%d) bound to libx265cu.so; pictures in pinned PicYuv, every Lowres array and the padded planes back in x265's own arrays" % host.pool,
               "clock": "the lookahead-only driver's clock (first addPicture -> last decided picture; Lookahead::create / x265cu_open outside, as for "
                        "--impl reference), max over ranks; wall incl. open/close per step: %.2f ms" % (1e3 * wall_s / args.steps),
               "parity": "skipped" if args.no_parity else "trace of this host == reference trace: %d estimates, every array CRC, every cuTreeFinish, every "
                         "slice-type decision" % (hstats["jobs"] if hstats else 0),
               "lookahead_cache": "%d of %d non-batch requests served from estimates launched ahead" % (t1[5] - t0[5], t1[7] - t0[7])}

    # ---- auxiliary: the same call sequence replayed through the host layer with host buffers (no x265 in the loop) ----
    e2e_replay = None
    if headline and not args.quick:
        rp = Runner(trace, clip, stream, local, False, torch)
        for _ in range(args.warmup):
            rp.step()
        rp.la.stats(reset=True)
        ms_rp = timed(torch, dist, world, rp.step, args.steps)
        st_rp = rp.la.stats(reset=True)
        rp.close()
        e2e_replay = {"value": world * nframes * args.steps / (ms_rp * 1e-3), "unit": "frames/s", "ms_per_step": ms_rp / args.steps,
                      "h2d_bytes_per_step": st_rp["h2d"] // args.steps, "d2h_bytes_per_step": st_rp["d2h"] // args.steps,
                      "what": "the trace's calls replayed through x265cu::Lookahead with pinned host buffers, pre-marshalled (round 1's e2e)"}
    clocks = sampler.stop()

    # ---- rooflines ----
    hbm, peak_src = hbm_peak()
    P = 1 if cfg["depth"] == 8 else 2
    nCU = ((cfg["width"] // 2 + 7) // 8) * ((cfg["height"] // 2 + 7) // 8)
    Np = nCU * 64
    Nfull = cfg["width"] * cfg["height"]
    alg = {   # algorithmic bytes per step
        "lowres": 8 * Np * P * nframes,                        # A: per frame
        "var": (Nfull + Nfull // 2) * P * nframes,             # luma + Cb + Cr read once per frame
        "intra": (Np * P + 7 * nCU) * nframes,                 # B: per frame
        "search": (5 * Np * P + 8 * nCU) * units,              # C: per (frame, list, distance) searched
    }
    kms = {k: v / args.steps for k, v in st["ms"].items()}
    klaunch = {k: v / args.steps for k, v in st["launches"].items()}
    dom = max(("lowres", "intra", "search", "cost", "weight", "var", "cutree"), key=lambda k: kms[k])
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except (OSError, ValueError):
        pass
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    issue_peak = 148 * 4 * sm_mhz * 1e6 / 1e9           # G warp-instructions / s: 4 schedulers per SM, one instruction per cycle each
    roofline = None
    pk = prof.get("search", {}).get(workload) or prof.get("search", {}).get("c1_1080p" if cfg["width"] == 1920 else "c2_4k")
    if dom == "search" and pk and kms["search"] > 0:
        # the search family is bound by instruction issue (ncu: DRAM traffic below the algorithmic bytes, ALU pipe 60-65 %): warp
        # instructions per searched CU from the committed `ncu --set full` capture x the CUs searched per step / the measured kernel time
        ginst = pk["warp_inst_per_cu"] * nCU * units / 1e9
        achieved = ginst / (kms["search"] * 1e-3)
        roofline = {"bound": "issue", "kernel": "plain_search_kernel (+ refine_kernel / search_kernel for small batches with a close hint)",
                    "achieved": achieved, "peak": issue_peak, "unit": "Gwarp-inst/s", "frac": achieved / issue_peak,
                    "peak_source": "148 SMs x 4 schedulers x %.0f MHz (median SM clock sampled during the timed region)" % sm_mhz,
                    "traffic": pk["dram_bytes"] / pk["units"] * units / max(klaunch["search"], 1),
                    "inst_source": pk.get("source"),
                    "note": "per step, i.e. averaged over every search launch (small latency-bound launches included); the captured "
                            "launch alone issued %.2f of 4 warp-inst/cycle/SM" % pk.get("ipc", 0)}
    hbm_block = {}
    for k in ("search", "lowres", "var", "intra"):
        if kms.get(k, 0) > 0:
            ach = alg[k] / (kms[k] * 1e-3) / 1e9
            cap = prof.get(k, {}).get(workload) or {}
            per_launch_units = (units if k == "search" else nframes) / max(klaunch[k], 1)
            hbm_block[k] = {"achieved": ach, "frac": ach / hbm, "ms_per_step": kms[k], "launches_per_step": klaunch[k],
                            "algorithmic_bytes_per_step": alg[k],
                            # DRAM bytes of an average launch of this run, scaled from the committed ncu capture (per launch, like `achieved`)
                            "traffic": cap["dram_bytes"] / cap["units"] * per_launch_units if cap.get("dram_bytes") else None,
                            "captured_launch": {kk: cap[kk] for kk in ("kernel", "units", "dram_bytes", "us", "source") if kk in cap} or None}
    if roofline is None:
        d = hbm_block.get(dom, {"achieved": 0.0, "frac": 0.0})
        roofline = {"bound": "hbm", "kernel": dom + "_kernel", "achieved": d["achieved"], "peak": hbm, "unit": "GB/s", "frac": d["frac"],
                    "traffic": d.get("traffic"), "peak_source": peak_src}
    if satd and "gpix_per_s" in satd:
        roof_pix = hbm / (2 * P + 4.0 / 64)     # GB/s / (bytes per pixel pair) = Gpix/s
        satd["roofline_gpix_per_s"] = roof_pix
        satd["frac"] = satd["gpix_per_s"] / roof_pix
        satd["gb_per_s"] = satd["bytes_per_launch"] / (satd["ms"] * 1e-3) / 1e9
        cap = prof.get("pixel", {}).get(workload) or {}
        if cap.get("dram_bytes"):
            # DRAM bytes of this launch, scaled from the committed ncu capture of the same kernel (bytes per plane pair)
            satd["traffic"] = cap["dram_bytes"] / cap["units"] * satd.get("pairs", cap["units"])
            satd["traffic_source"] = cap.get("source")

    # ---- CPU baseline (reported, not the target): the reference's own lookahead on the host cores ----
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1 and rank == 0:
        cores = os.cpu_count() or 1
        pool = min(64, cores)
        rr = subprocess.run([sys.executable, os.path.join(ROOT, "harness", "refrun.py"), workload, str(pool), "1", "ref"],
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        try:
            o = json.loads(rr.stdout.strip().splitlines()[-1])
            sec = o["seconds"][-1]
            cpu_baseline = {"value": nframes / sec, "unit": "frames/s", "cores": o["threads"], "kind": o["kind"],
                            "sample": "the whole workload once (%d frames, %.1f s): x265 1.9 Lookahead only, same driver as the e2e arm, C primitives "
                                      "(asm build impossible: no yasm/nasm in the image)" % (nframes, sec)}
        except (ValueError, IndexError, KeyError):
            cpu_baseline = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "failed: " + rr.stderr.strip()[-160:]}

    out.update({
        "value": world * nframes * args.steps / (ms_value * 1e-3), "ms_per_step": ms_value / args.steps,
        "e2e": e2e, "e2e_replay": e2e_replay, "estimates_per_step": njobs, "searches_per_step": units,
        "gpu_launches": int(sum(st["launches"].values())),
        "kernel_ms_per_step": kms, "kernel_launches_per_step": klaunch,
        "roofline": roofline, "hbm_roofline": {"peak": hbm, "unit": "GB/s", "peak_source": peak_src, "kernels": hbm_block},
        "int_peak_gops": {"vabsdiff4": int_peak[0].value, "iadd": int_peak[1].value},
        "satd_8x8": satd, "cpu_baseline": cpu_baseline, "clocks": clocks,
        "cutree": "%d propagate steps per step on the GPU (estimateCUPropagate, SURVEY 8f-1), one launch per run of steps; "
                  "mirrors found stale: %d (re-uploaded)" % (npropagate, ct_value["reuploads"]),
    })
    return out, trace, clip


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1_1080p")
    ap.add_argument("--configs", default="c2_4k,c3_4k10", help="further BASELINE configs measured into the line's `configs` array (N = 1 only; '' = none)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline workload only, no auxiliary measurements")
    ap.add_argument("--streams", type=int, default=8, help="auxiliary multi-stream measurement: streams per GPU (0/1 = skip)")
    ap.add_argument("--profile-mode", action="store_true",
                    help="short run for ncu: device-resident runner only, 1 warm-up + 1 timed step, no parity/e2e/baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if (args.impl == "ours" and not args.profile_mode) else args.warmup
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        return reference_arm(args, rank, world)

    if world > 1 and env_int("X265CU_BENCH_PIN", 1):
        pin_rank(local, env_int("LOCAL_WORLD_SIZE", world))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    if rank == 0 or not os.path.exists(os.path.join(ROOT, "src", "x265_b200", "libx265cu.so")):
        ge.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL announces its version on stdout at communicator creation: keep stdout to the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    if args.profile_mode:
        from harness import replay
        from oracle import pyoracle as po
        trace = po.Trace(replay.trace_path(args.workload))
        clip = replay.Clip(trace.cfg)
        res = Runner(trace, clip, torch.cuda.current_stream().cuda_stream, local, True, torch)
        res.step()
        ms = timed(torch, dist, world, res.step, 1)
        st = res.la.stats()
        res.close()
        print(json.dumps({"profile_mode": True, "ms_per_step": ms, "launches": st["launches"]}))
        return 0

    head, trace, clip = measure(args, args.workload, torch, dist, world, local, rank, True)

    # ---- auxiliary: several independent streams per GPU (the production shape, BASELINE configs[4]) ----
    multi = None
    if args.streams > 1 and not args.quick:
        multi = multi_stream(torch, dist, world, rank, args.workload, local, args.streams, max(args.steps, 20))
    del trace, clip

    # ---- the other BASELINE configs, same measurements (N = 1: the target config of north_star is 4K rc-lookahead 40) ----
    configs = []
    if world == 1 and not args.quick:
        for wl in [w for w in args.configs.split(",") if w and w != args.workload]:
            c, _, _ = measure(args, wl, torch, dist, world, local, rank, False)
            configs.append(c)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    def compact(c):
        keep = ("trace", "workload", "resolution", "bit_depth", "frames_per_step", "value", "ms_per_step", "e2e", "estimates_per_step",
                "searches_per_step", "kernel_ms_per_step", "kernel_launches_per_step", "roofline", "hbm_roofline", "satd_8x8", "cpu_baseline", "parity")
        return {k: c[k] for k in keep if k in c}

    line = {
        "metric": "lookahead_frames_per_s", "value": head["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8" if head["bit_depth"] == 8 else "u16", "data": "synthetic",
        "config": {"workload": head["workload"], "trace": head["trace"], "frames_per_step": head["frames_per_step"],
                   "estimates_per_step": head["estimates_per_step"], "searches_per_step": head["searches_per_step"], "resolution": head["resolution"],
                   "streams": world, "parallelism": "%d independent stream(s), one per GPU, no collective on the cost path" % world,
                   "search_path": os.environ.get("X265CU_SEARCH_MODE", "2 (per search: plain wavefront kernel, or refine + commit for small batches with a close hint)"),
                   "lookahead_cache": os.environ.get("X265CU_LOOKAHEAD_CACHE", "1 (non-batch estimates predicted from the request history ride in one launch)"),
                   "l2": "working set per step (%d frames x 4 padded planes + sources, > 300 MB) exceeds the 126 MB L2; no explicit flush" % head["frames_per_step"],
                   "cutree": head["cutree"], "parity": head["parity"],
                   "value_is": "device-resident replay of the reference's call sequence (pictures in HBM, arrays stay in HBM)",
                   "e2e_is": "x265 1.9's own lookahead as the host, host pictures in, every Lowres array back (see e2e.host)"},
        "e2e": head["e2e"], "e2e_replay": head["e2e_replay"],
        "gpu_launches": head["gpu_launches"],
        "kernel_ms_per_step": head["kernel_ms_per_step"], "kernel_launches_per_step": head["kernel_launches_per_step"],
        "roofline": head["roofline"], "hbm_roofline": head["hbm_roofline"], "int_peak_gops": head["int_peak_gops"], "satd_8x8": head["satd_8x8"],
        "cpu_baseline": head["cpu_baseline"], "clocks": head["clocks"], "multi_stream": multi,
        "configs": [compact(head)] + [compact(c) for c in configs],
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
