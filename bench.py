#!/usr/bin/env python
"""bench.py -- scan-to-map registrations/s on synthetic HDL-64 sequences (BASELINE.json metric).

Workload "hdl64_batch_replay" (BASELINE config 4, SURVEY 8e "batch replay"): every GPU holds
S independent HDL-64 sequences (own map, own pose chain); one STEP advances every sequence by
one frame = S scan-to-map registrations (laserMapping.cpp:310-802 each), all inside the same
kernel launches.  No data-path collective; weak scaling (S per GPU fixed).

  value : registrations/s, inputs already resident in HBM (s2m_register_batch_dev)
  e2e   : the same through s2m_register_batch with pinned HOST buffers (H2D of the clouds and
          D2H of poses/counters inside the timed region)
  roofline : the fused association kernel (K4) -- algorithmic bytes of SURVEY 8d / CUDA-event time
  cpu_baseline : the oracle (restated reference path, 1 core) on a bounded sample of the workload

--impl reference times the reference path's CPU restatement (oracle/, one sequence per host
thread) on the same workload/metric; the real PCL+Ceres binary cannot be built here.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scan_to_map_registrations_per_s"
UNIT = "registrations/s"
N_WORLDS = 8          # distinct synthetic worlds/trajectories (BASELINE config 4: 8 sequences)
PREFILL = int(os.environ.get("S2M_BENCH_PREFILL", 120))  # untimed frames that build each sequence's map: 120 m driven, the
# local map (5 x 5 x 3 cubes) is then stationary in size (~110 k points) -- the steady state of BASELINE config 1
SENSOR = "HDL64"
LINE_RES, PLANE_RES = 0.4, 0.8  # aloam_velodyne_HDL_64.launch:11-12


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def make_worlds(n_frames, rank, threads):
    """N_WORLDS replayable HDL-64 sequences: truth poses + per-frame (corner, surf) clouds."""
    import harness
    from concurrent.futures import ThreadPoolExecutor

    def one(w):
        seed = 20261018 + w
        truth = harness.trajectory(seed, n_frames, 1.0)
        frames = [harness.features(SENSOR, harness.scan(seed, SENSOR, truth[f], f, 0.02)) for f in range(n_frames)]
        return (seed, truth, frames)

    harness.lib()
    with ThreadPoolExecutor(max_workers=min(N_WORLDS, max(1, threads))) as ex:  # ctypes releases the GIL
        return list(ex.map(one, range(N_WORLDS)))


def slot_odometry(worlds, n_slots, rank):
    """Slot s replays world s % N_WORLDS (frames PREFILL..) with its own odometry drift, starting on the truth at
    frame PREFILL (so every slot is a different registration problem: different guesses and poses).
    Row f of the result belongs to frame PREFILL + f."""
    import harness
    odo = []
    for s in range(n_slots):
        seed, truth = worlds[s % N_WORLDS][0], worlds[s % N_WORLDS][1]
        odo.append(harness.odometry(seed * 7919 + 104729 * (s // N_WORLDS) + 15485863 * rank, truth[PREFILL:], 0.02, 0.1))
    return odo


def mature_maps(pkg, worlds, device):
    """The map every sequence has after PREFILL frames: built once per world by the engine itself (one context, one
    slot per world, the frames registered at the true poses), downloaded, and later pushed into every slot that
    replays that world (s2m_map_upload) -- instead of prefilling 384 slots frame by frame."""
    if PREFILL == 0:
        return [(np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32)) for _ in worlds]
    nw = len(worlds)
    max_c = max(len(w[2][f][0]) for w in worlds for f in range(PREFILL)) + 64
    max_s = max(len(w[2][f][1]) for w in worlds for f in range(PREFILL)) + 64
    R = pkg.Registrar(LINE_RES, PLANE_RES, device=device, batch=nw, cap_corner_in=max_c, cap_surf_in=max_s,
                      cap_map_corner=1 << 19, cap_map_surf=1 << 19)
    for f in range(PREFILL):
        cs, ss = [w[2][f][0] for w in worlds], [w[2][f][1] for w in worlds]
        co = np.cumsum([0] + [len(c) for c in cs]).astype(np.int32)
        so = np.cumsum([0] + [len(c) for c in ss]).astype(np.int32)
        R.register_batch(np.concatenate(cs), co, np.concatenate(ss), so, np.array([w[1][f, :4] for w in worlds]),
                         np.array([w[1][f, 4:] for w in worlds]))
    maps = [(R.map_download(0, slot=i), R.map_download(1, slot=i)) for i in range(nw)]
    R.close()
    return maps


def pack_step(worlds, n_slots, f):
    cs = [worlds[s % N_WORLDS][2][f][0] for s in range(n_slots)]
    ss = [worlds[s % N_WORLDS][2][f][1] for s in range(n_slots)]
    co = np.cumsum([0] + [len(c) for c in cs]).astype(np.int32)
    so = np.cumsum([0] + [len(c) for c in ss]).astype(np.int32)
    return np.concatenate(cs), co, np.concatenate(ss), so


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons of one GPU, sampled during the timed region (NVML)."""

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.samples, self.reasons, self.stop_flag = device, [], set(), False
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"]}
        return {"sm_mhz": int(np.median(self.samples)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def k4_traffic():
    """DRAM bytes per K4 launch pair from the committed ncu --set full capture -- used only if that capture was
    taken from the kernels as they are now (sha256 of csrc/s2m_kernels.cu + s2m_math.cuh), else None."""
    import hashlib
    p = os.path.join(ROOT, "profiles", "k4_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        t = json.load(open(p))
        h = hashlib.sha256()
        for f in ("s2m_kernels.cu", "s2m_math.cuh"):
            h.update(open(os.path.join(ROOT, "sc-a-loam_b200", "csrc", f), "rb").read())
        return t if t.get("kernels_sha256") == h.hexdigest() else None
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
def cpu_baseline(worlds, maps, odo, n_run, max_seconds=25.0):
    """Oracle (restated reference path) on ONE core over a bounded sample: each world's replayed frames, starting
    from the same mature map, sequence after sequence."""
    import oracle
    regs, t_total = 0, 0.0
    per_phase = np.zeros(7)
    for w in range(N_WORLDS):
        O = oracle.Oracle(LINE_RES, PLANE_RES)
        frames = worlds[w][2]
        if len(maps[w][0]) + len(maps[w][1]):
            O.map_upload(maps[w][0], maps[w][1])
        od = odo[w]
        for f in range(n_run):
            c, s = frames[PREFILL + f]
            t0 = time.perf_counter()
            O.register(c, s, od[f, :4], od[f, 4:])
            dt = time.perf_counter() - t0
            if f >= 1:  # (the first frame re-filters the uploaded map)
                t_total += dt
                regs += 1
                per_phase += np.array(O.stats.t_ms[:7])
        if t_total > max_seconds:
            break
    return {"value": regs / t_total, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d registrations: frames %d..%d of %d HDL-64 worlds against their %d-frame maps, sequentially on "
                      "one core (oracle/ = CPU restatement of laserMapping.cpp:310-802; PCL/Ceres not installable here)"
                      % (regs, PREFILL + 1, PREFILL + n_run - 1, w + 1, PREFILL),
            "ms_per_registration": 1e3 * t_total / regs,
            "phase_ms": {k: round(float(v) / regs, 3) for k, v in
                         zip(["shift", "tree", "data", "solver", "add", "filter", "whole"], per_phase)}}


def run_reference(args, rank, world):
    """The reference path's CPU restatement on all host cores, same workload: every world's map is built by the
    restatement itself (PREFILL frames at the true poses), then one sequence per host thread is replayed."""
    if rank != 0:
        return 0
    import oracle
    cores = os.cpu_count() or 1
    n_seq = min(cores, 64)
    n_run = args.warmup + args.steps
    worlds = make_worlds(PREFILL + n_run, 0, cores)
    odo = slot_odometry(worlds, n_seq, 0)
    maps = [None] * N_WORLDS

    def build_map(w):
        O = oracle.Oracle(LINE_RES, PLANE_RES)
        truth, frames = worlds[w][1], worlds[w][2]
        for f in range(PREFILL):
            O.register(frames[f][0], frames[f][1], truth[f, :4], truth[f, 4:])
        maps[w] = (O.get_map(0), O.get_map(1))

    th = [threading.Thread(target=build_map, args=(w,)) for w in range(N_WORLDS)]
    [t.start() for t in th]
    [t.join() for t in th]
    oracles = [oracle.Oracle(LINE_RES, PLANE_RES) for _ in range(n_seq)]
    for s_, O in enumerate(oracles):
        cm, sm = maps[s_ % N_WORLDS]
        if len(cm) + len(sm):
            O.map_upload(cm, sm)

    def run_frames(s, f0, f1):
        frames = worlds[s % N_WORLDS][2]
        for f in range(f0, f1):
            c, su = frames[PREFILL + f]
            oracles[s].register(c, su, odo[s][f, :4], odo[s][f, 4:])

    def parallel(f0, f1):
        th = [threading.Thread(target=run_frames, args=(s, f0, f1)) for s in range(n_seq)]
        [t.start() for t in th]
        [t.join() for t in th]

    parallel(0, args.warmup)  # ctypes releases the GIL: one sequence per host thread
    t0 = time.perf_counter()
    parallel(args.warmup, n_run)
    dt = time.perf_counter() - t0
    value = n_seq * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "hdl64_batch_replay", "sensor": "HDL-64 synthetic 64x1900",
                       "line_res": LINE_RES, "plane_res": PLANE_RES, "sequences": n_seq,
                       "distinct_worlds": N_WORLDS, "prefill_frames": PREFILL,
                       "map_points_per_sequence": float(np.mean([len(m[0]) + len(m[1]) for m in maps]))},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_seq, "kind": "port",
                             "sample": "%d sequences x %d frames, one sequence per host thread (%d cores); oracle/ "
                                       "restatement of the reference path" % (n_seq, args.steps, cores)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
def prepare_worlds(pkg, n_frames, rank, world, device):
    """Each rank ray-casts its share of the N_WORLDS worlds and builds their mature maps; then all ranks exchange
    the replayed frames (PREFILL..) and the maps.  -> (worlds, maps); frames before PREFILL are dropped (None)."""
    if world == 1:
        worlds = make_worlds(n_frames, rank, os.cpu_count() or 1)
        maps = mature_maps(pkg, worlds, device)
        return worlds, maps
    import harness
    import torch.distributed as dist
    harness.set_threads(max(1, (os.cpu_count() or 1) // world))  # (torchrun exports OMP_NUM_THREADS=1)
    ids = list(range(rank, N_WORLDS, world))
    mine = []
    for w in ids:
        seed = 20261018 + w
        truth = harness.trajectory(seed, n_frames, 1.0)
        frames = [harness.features(SENSOR, harness.scan(seed, SENSOR, truth[f], f, 0.02)) for f in range(n_frames)]
        mine.append((seed, truth, frames))
    my_maps = mature_maps(pkg, mine, device) if mine else []
    share = {w: ((m[0], m[1], [None] * PREFILL + m[2][PREFILL:]), mp) for w, m, mp in zip(ids, mine, my_maps)}
    parts = [None] * world
    dist.all_gather_object(parts, share)
    merged = {}
    for p in parts:
        merged.update(p)
    return [merged[w][0] for w in range(N_WORLDS)], [merged[w][1] for w in range(N_WORLDS)]


class WcBuffer:
    """A host copy of `a` in write-combined pinned memory (cudaHostAlloc, cuda-python): what an application would use
    for buffers the CPU only fills and the GPU only reads; the DMA reads do not snoop the CPU caches."""

    def __init__(self, a):
        import ctypes
        from cuda.bindings import runtime as rt
        a = np.ascontiguousarray(a)
        err, ptr = rt.cudaHostAlloc(max(a.nbytes, 16), rt.cudaHostAllocWriteCombined | rt.cudaHostAllocPortable)
        if int(err) != 0:
            raise RuntimeError("cudaHostAlloc(write-combined) failed: %s" % err)
        self.ptr, self.nbytes, self._rt = int(ptr), a.nbytes, rt
        ctypes.memmove(self.ptr, a.ctypes.data, a.nbytes)

    def data_ptr(self):
        return self.ptr

    def __del__(self):
        if getattr(self, "ptr", 0):
            self._rt.cudaFreeHost(self.ptr)
            self.ptr = 0


class Lane:
    """One context holding S sequences split over `lanes` concurrent lanes (own stream + host thread
    each, inside the library): one C-ABI call per step advances all of them."""

    def __init__(self, pkg, torch, worlds, maps, S, lanes, variant, n_run, local_rank, args, host_buffers):
        """n_run frames are replayed: step f is frame PREFILL + f of every sequence; the maps start mature."""
        self.torch, self.S, self.n_frames = torch, S, n_run
        odo = slot_odometry(worlds, S, variant)
        steps = [pack_step(worlds, S, PREFILL + f) for f in range(n_run)]
        self.q_all = np.array([[odo[s][f, :4] for s in range(S)] for f in range(n_run)])
        self.t_all = np.array([[odo[s][f, 4:] for s in range(S)] for f in range(n_run)])
        self.host = host_buffers
        if host_buffers and args.staging == "wc":
            self.steps = [(WcBuffer(c), co, WcBuffer(su), so) for c, co, su, so in steps]
        elif host_buffers:
            self.steps = [(torch.from_numpy(c).pin_memory(), co, torch.from_numpy(su).pin_memory(), so) for c, co, su, so in steps]
        else:
            self.steps = [(torch.from_numpy(c).cuda(), co, torch.from_numpy(su).cuda(), so) for c, co, su, so in steps]
        self.h2d = float(np.mean([c.nbytes + su.nbytes for c, _, su, _ in steps[args.warmup:]]))
        max_c = max(int(np.diff(st[1]).max()) for st in steps)
        max_s = max(int(np.diff(st[3]).max()) for st in steps)
        self.stream = torch.cuda.Stream()  # a real stream: the library fences its lanes on it, the events time it
        self.R = pkg.Registrar(LINE_RES, PLANE_RES, device=local_rank, batch=S, lanes=lanes, cap_corner_in=max_c + 64,
                               cap_surf_in=max_s + 64, cap_map_corner=args.cap_map_corner, cap_map_surf=args.cap_map_surf)
        for sl in range(S):  # every sequence starts from the map its world has after PREFILL frames
            cm, sm = maps[sl % N_WORLDS]
            if len(cm) + len(sm):
                self.R.map_upload(cm, sm, slot=sl)
        self.R.set_stream(self.stream.cuda_stream)
        self.events = None
        self.pipelined = lanes >= 1 and not args.no_pipeline

    def step(self, f):
        c, co, su, so = self.steps[f]
        # host_buffers: the reference-facing call with HOST buffers (H2D + D2H inside); else device pointers
        return self.R.register_batch_ptr(c.data_ptr(), co, su.data_ptr(), so, self.q_all[f], self.t_all[f], not self.host)

    def run(self, f0, f1, timed):
        torch = self.torch
        with torch.cuda.stream(self.stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            if self.pipelined:
                # two frames in flight: the copy of frame f+1 overlaps the registration of frame f
                for f in range(f0, f1):
                    if f - f0 >= 2:
                        self.R.wait()
                    c, co, su, so = self.steps[f]
                    self.R.submit(c.data_ptr(), co, su.data_ptr(), so, self.q_all[f], self.t_all[f], device=not self.host)
                for _ in range(min(2, f1 - f0)):
                    self.R.wait()
            else:
                for f in range(f0, f1):
                    self.step(f)
            e1.record(self.stream)
        if timed:
            self.events = (e0, e1)

    def close(self):
        self.R.close()
        self.steps = None


def bind_near_gpu(device, world):
    """N > 1: run this rank (its lane threads inherit it) on the cores NVML reports as closest to its GPU, so that the
    pinned staging buffers are first touched -- allocated -- on that NUMA node and the copies do not cross sockets.
    Skipped for one rank (nothing to contend with) and when the environment disables it (S2M_NO_BIND=1)."""
    if world <= 1 or os.environ.get("S2M_NO_BIND"):
        return "none"
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return "none (empty NVML affinity)"
        os.sched_setaffinity(0, cpus)
        return "%d cores near GPU %d (NVML affinity)" % (len(cpus), device)
    except Exception as e:  # no NVML / not permitted: run unbound
        return "none (%s)" % type(e).__name__


def run_ours(args, rank, world, local_rank):
    binding = bind_near_gpu(local_rank, world)
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    pkg = load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    S, C = args.seqs, args.ctx
    n_frames = PREFILL + args.warmup + args.steps
    t_gen = time.perf_counter()
    worlds, maps = prepare_worlds(pkg, n_frames, rank, world, local_rank)
    t_gen = time.perf_counter() - t_gen

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    n_run = args.warmup + args.steps

    def timed_arm(host_buffers, sampler=None):
        lane = Lane(pkg, torch, worlds, maps, C * S, C, rank, n_run, local_rank, args, host_buffers)
        lane.run(0, args.warmup, False)  # untimed warm-up steps (the first one re-filters the uploaded map)
        barrier()
        if sampler:
            sampler.start()
        l0 = lane.R.launch_count()
        ncu_range = bool(os.environ.get("S2M_NCU_RANGE")) and not host_buffers  # ncu --profile-from-start off: only the timed steps
        if ncu_range:
            torch.cuda.profiler.start()
        wall0 = time.perf_counter()
        lane.run(args.warmup, n_run, True)
        barrier()
        wall = time.perf_counter() - wall0
        if ncu_range:
            torch.cuda.profiler.stop()
        if sampler:
            sampler.stop_flag = True
        ms = lane.events[0].elapsed_time(lane.events[1])
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        launches = lane.R.launch_count() - l0
        h2d = lane.h2d
        stats = [lane.R.batch_stats[s] for s in range(C * S)]
        shape = {k: float(np.mean([getattr(st, k) for st in stats])) for k in
                 ("n_corner_in", "n_surf_in", "n_corner_ds", "n_surf_ds", "n_map_corner", "n_map_surf")}
        shape["n_edge"] = float(np.mean([st.n_edge[1] for st in stats]))
        shape["n_plane"] = float(np.mean([st.n_plane[1] for st in stats]))
        shape["lm_iterations"] = float(np.mean([st.lm_iters[0] + st.lm_iters[1] for st in stats]))
        lane.close()
        torch.cuda.empty_cache()
        return ms, launches, wall, h2d, shape

    sampler = ClockSampler(local_rank)
    ms_dev, launches, wall_dev, _, shape = timed_arm(False, sampler)      # arm 1: inputs resident in HBM
    ms_e2e, _, wall_e2e, h2d, _ = timed_arm(True)                         # arm 2: pinned host buffers, end to end

    # arm 3 (untimed for the headline): one context with profiling on -> K4 roofline and phase split
    prof_steps = min(args.steps, 6)
    lane = Lane(pkg, torch, worlds, maps, S, 0, rank, args.warmup + prof_steps, local_rank, args, False)
    lane.run(0, args.warmup, False)
    torch.cuda.synchronize()
    lane.R.set_profiling(True, count_candidates=True)
    lane.R.k4_profile(reset=True)
    lane.run(args.warmup, args.warmup + prof_steps, False)
    k4_ms, k4_n, k4_bytes = lane.R.k4_profile(reset=False)
    phases = lane.R.phase_profile(reset=True)
    knn_fb = lane.R.knn_fallbacks() if hasattr(lane.R, "knn_fallbacks") else None
    lane.close()

    # arm 4: BASELINE config 5 -- the spatially sharded giant map (every rank takes part: NCCL allreduce inside)
    sharded = None
    if not args.no_sharded:
        from bench_sharded import run_sharded
        sharded = run_sharded(pkg, torch, dist, rank, world, local_rank, slots=args.sharded_slots,
                              fill_corner=args.sharded_fill_corner, fill_surf=args.sharded_fill_surf)

    # arm 5: BASELINE config 3 -- OS1-64 against a saturated window (a 1-GPU configuration: N = 1 only)
    os1 = None
    if world == 1 and not args.no_os1:
        from bench_os1 import run_os1_saturated
        os1 = run_os1_saturated(pkg, torch, local_rank, slots=args.os1_slots)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    regs = world * C * S * args.steps
    value = regs / (ms_dev * 1e-3)
    e2e_value = regs / (ms_e2e * 1e-3)
    peak, peak_src = measured_peak()
    ach = (k4_bytes / max(k4_n, 1)) / (1e-3 * k4_ms / max(k4_n, 1)) / 1e9 if k4_ms > 0 else 0.0
    traffic = k4_traffic()
    d2h = float(C * (S * 168 + 4))  # poses + per-slot counters + error flag read back every step
    # bytes_alg per registration = 2 (B_K4 + B_K5) + B_map with B_K4 per association launch as measured
    # above, B_K5 = 48 Nv E (E = LM evaluations executed after the first), B_map = 2*16 (Nc_in+Ns_in) + 3*16 (Nc+Ns)
    nv = shape["n_edge"] + shape["n_plane"]
    b_k4 = (k4_bytes / max(k4_n, 1)) / S
    b_k5 = 48.0 * nv * (shape["lm_iterations"] / 2.0)
    b_map = 32.0 * (shape["n_corner_in"] + shape["n_surf_in"]) + 48.0 * (shape["n_corner_ds"] + shape["n_surf_ds"])
    bytes_reg = 2.0 * (b_k4 + b_k5) + b_map
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "hdl64_batch_replay", "sensor": "HDL-64 synthetic 64x1900 (121600 rays/sweep)",
                   "line_res": LINE_RES, "plane_res": PLANE_RES, "sequences_per_gpu": C * S, "lanes_per_gpu": C,
                   "sequences_per_lane": S, "distinct_worlds": N_WORLDS, "prefill_frames": PREFILL,
                   "registrations_per_step": world * C * S, "parallelism": "independent sequences x%d GPUs, no collective" % world,
                   "l2": "no flush: each step touches >1 GB per lane (64 maps of ~110 k points, sort buffers, clouds), far larger than the 126 MB L2",
                   "maps": "mature: every sequence starts from the map its world has after %d frames (built by the engine, uploaded with s2m_map_upload); see per_registration_mean.n_map_*" % PREFILL,
                   "timing": "CUDA events around the K steps on the caller's stream (the library fences its lanes on it); max over ranks",
                   "cpu_binding": binding, "e2e_host_buffers": args.staging,
                   "per_registration_mean": shape, "datagen_s": round(t_gen, 1),
                   "phase_ms_per_step_single_lane": {k: round(v / prof_steps, 4) for k, v in phases.items()},
                   "host_wall_ms_per_step": round(1e3 * wall_dev / args.steps, 3)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d + C * 64 * 128, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "association = knn_group_kernel + fit_kernel, plus qgroup_kernel once per frame (transform + exact kNN5 + edge PCA / plane QR + residual/J + Huber + reduce)",
                     "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
                     "bytes_per_launch_algorithmic": k4_bytes / max(k4_n, 1), "launches": int(k4_n),
                     "avg_launch_us": 1e3 * k4_ms / max(k4_n, 1), "slots_per_launch": S,
                     "measured": "CUDA events on the launching stream, single context, profiling pass (the per-frame query sort of the grouped search is inside the bracket of the first launch)",
                     "knn_fallback_queries_profiling_pass": knn_fb,
                     "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                     "traffic_source": traffic.get("source") if traffic else None},
        "registration_roofline": {"bound": "hbm", "bytes_algorithmic_per_registration": bytes_reg,
                                  "achieved": bytes_reg * value / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": bytes_reg * value / 1e9 / peak,
                                  "what": "whole registration (rows A..W), SURVEY 8d bytes_alg x registrations/s of the device-resident arm"},
        "clocks": sampler.result(),
    }
    if sharded is not None:
        line["sharded"] = sharded
    if os1 is not None:
        line["os1_saturated"] = os1
    if world == 1 and not args.no_cpu_baseline:
        odo = slot_odometry(worlds, N_WORLDS, 0)
        line["cpu_baseline"] = cpu_baseline(worlds, maps, odo, n_run)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seqs", type=int, default=64, help="independent sequences per lane (<=64)")
    ap.add_argument("--no-pipeline", action="store_true", help="synchronous batch calls instead of submit/wait with two frames in flight")
    ap.add_argument("--ctx", "--lanes", dest="ctx", type=int, default=6, help="concurrent lanes per GPU inside the one context")
    ap.add_argument("--staging", choices=["pinned", "wc"], default="pinned", help="host buffers of the end-to-end arm: torch pinned memory, or write-combined pinned memory")
    ap.add_argument("--cap-map-corner", type=int, default=1 << 18)
    ap.add_argument("--cap-map-surf", type=int, default=1 << 18)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-os1", action="store_true", help="skip the OS1-64 / saturated-window arm (BASELINE config 3)")
    ap.add_argument("--os1-slots", type=int, default=16)
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded giant-map arm (BASELINE config 5)")
    ap.add_argument("--sharded-slots", type=int, default=16)
    ap.add_argument("--sharded-fill-corner", type=int, default=3_000_000)
    ap.add_argument("--sharded-fill-surf", type=int, default=1_000_000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    # the harness ray-caster uses OpenMP: share the host cores between the ranks of this node
    os.environ.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 1) // max(world, 1))))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    return run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
