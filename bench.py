#!/usr/bin/env python
"""bench.py -- GM-PHD updates/s of the RB-PHD-SLAM filter step on B200 (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one full filter step (predict + in-range split + GM-PHD update + prune/merge + weight update and
normalisation + state estimate + resampling when nEff triggers it) over one synthetic scene of exactly
P x C x M (particles x in-range components x measurements); one update = one (particle, component,
measurement) detection term, so a step is exactly P*C*M updates.  The device state is restored to the same
scene before every step (outside the timed region) so that all K steps do identical work.

  value     whole-job updates/s, device-timed (CUDA events on the filter's stream) with the scene resident in HBM
  e2e       the same metric through the C-ABI call phdslam_step() with HOST measurement / control buffers,
            wall clock, including the host<->device copies the call performs
  roofline  GM-PHD update kernel (update_dense_kernel): algorithmic bytes (28.99 B/update at
            C=256, M=64) / its CUDA-event duration, against the measured HBM copy bandwidth
  --impl reference  the CPU oracle (the reference's named CPU path src/scphd_cpu.cpp is an empty stub; its CUDA
            path cannot be built or run without a GPU toolchain of 2012), all host threads, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cuda-phdslam_b200"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[2]: the single-GPU configuration the updates/s metric is quoted on
    "synthetic_65536x256x64_phd": dict(P=65536, C=256, M=64, max_components=384),
    # smaller shapes for quick checks
    "synthetic_8192x256x64_phd": dict(P=8192, C=256, M=64, max_components=384),
    "synthetic_1024x64x32_phd": dict(P=1024, C=64, M=32, max_components=128),
    # BASELINE.json configs[3], one GPU's shard (1M particles over 8 GPUs): CPHD with the cardinality distribution
    "synthetic_16384x128x50_cphd": dict(P=16384, C=128, M=50, max_components=256, filter_type=1, max_cardinality=255),
    "synthetic_131072x128x50_cphd": dict(P=131072, C=128, M=50, max_components=256, filter_type=1, max_cardinality=255),
    # north_star target sentence: "the GM-PHD update at 1M particles on one GPU" (configs[3]'s particle count and C x M
    # shape on ONE B200; the 184 GB of dense update terms stream through the 32 GB update buffer in batches)
    "synthetic_1048576x128x50_phd": dict(P=1048576, C=128, M=50, max_components=256),
    "synthetic_16384x128x50_phd": dict(P=16384, C=128, M=50, max_components=256),      # the same shape, ncu-sized
    # BASELINE.json configs[4] per-GPU shape at a size one GPU's update buffer streams through: global resampling every step
    "synthetic_32768x128x100_phd": dict(P=32768, C=128, M=100, max_components=256, resample_threshold=1.0),
    "synthetic_262144x128x100_phd": dict(P=262144, C=128, M=100, max_components=256, resample_threshold=1.0),
    # BASELINE.json configs[4] at its literal size: 16 777 216 particles IN TOTAL, split over the GPUs (strong scaling), global
    # resampling every step.  A rank generates 262 144 distinct particles and tiles them on the device (import_tiled).
    # Maps reach 208 - 230 components after one step at this shape (max_components = 256).  Memory per GPU: the two map
    # buffers 2 x 6 KB per particle, 16 GB dense buffer, 16 GB candidate buffers, 1.6 GB snapshot of the distinct
    # particles: 139 GB at N = 2 (8.4 M particles per GPU).  N = 1 would need 206 GB for the map buffers alone: it does
    # not fit one B200 (180 GB), so the sweep starts at N = 2.
    "synthetic_16777216x128x100_phd": dict(P=16777216, C=128, M=100, max_components=256, resample_threshold=1.0, strong=1,
                                           scene_particles=262144, update_buffer_bytes=16 << 30),
    # SURVEY 8(f) rank 4, the mixed feature model: Cd constant-velocity features per particle next to the static map, a third
    # of the measurements observing them (no BASELINE config).  updates = particles x (C + Cd) x M.
    "synthetic_65536x128+16x50_mixed": dict(P=65536, C=128, M=50, max_components=256, Cd=16, max_components_dynamic=104,
                                            feature_model=2, std_ax_features=0.5, std_ay_features=0.4, cov_vx_birth=0.25,
                                            cov_vy_birth=0.36, tau=0.3, beta=4.0, ps=0.97),
    "synthetic_4096x64+24x32_mixed": dict(P=4096, C=64, M=32, max_components=128, Cd=24, max_components_dynamic=96,
                                          feature_model=2, std_ax_features=0.5, std_ay_features=0.4, cov_vx_birth=0.25,
                                          cov_vy_birth=0.36, tau=0.3, beta=4.0, ps=0.97),
    "synthetic_2097152x128x100_phd": dict(P=2097152, C=128, M=100, max_components=256, resample_threshold=1.0, strong=1,
                                          scene_particles=65536, update_buffer_bytes=16 << 30),   # the same path, quick-check size
}
NON_CFG_KEYS = ("P", "C", "M", "max_components", "strong", "scene_particles", "Cd")
DEFAULT_WORKLOAD = "synthetic_65536x256x64_phd"


def alg_bytes_per_update(C, M, n_card=0):
    """B_alg = [28*C + 28*(C*(M+1)+M) + 32 (+ 2*4*(N+1) cardinality in/out for CPHD)] / (C*M)"""
    return (28.0 * C + 28.0 * (C * (M + 1) + M) + 32.0 + 8.0 * n_card) / (C * M)


def load_mixed(filt, S, wl, n, sc, particle_seed=None):
    """Mixed feature model workloads: gives `filt` (PhdSlam or Oracle, n particles loaded from `sc`) its dynamic maps and
    returns the measurement set with every third measurement observing a dynamic feature.  Other workloads: sc["Z"]."""
    if not wl.get("Cd"):
        return sc["Z"]
    dsz, dm, base = S.make_dynamic_maps(n, wl["Cd"], seed=1, particle_seed=particle_seed)
    filt.set_maps_dynamic(dsz, dm)
    return S.mix_dynamic_measurements(sc["Z"], base)


def pairs_per_particle(wl):
    """(component, measurement) detection terms per particle and step: C x M, plus Cd x M in the mixed feature model"""
    return (wl["C"] + wl.get("Cd", 0)) * wl["M"]


def workload_config(name, wl, n_gpus):
    """the `config` object of a bench line: the same on both arms (what the workload is, not how a run went)"""
    strong = bool(wl.get("strong"))
    per = wl["P"] // n_gpus if strong else wl["P"]
    cfg = {"workload": name, "particles_per_gpu": per, "particles_total": wl["P"] if strong else wl["P"] * n_gpus,
           "components": wl["C"], "measurements": wl["M"], "filter": "CPHD" if wl.get("filter_type") == 1 else "PHD",
           "scaling": "strong" if strong else "weak",
           # timing rule: no L2 flush between the timed steps because every step's inputs and outputs exceed the 126 MB L2
           "cache": "inputs larger than L2 (map %.0f MB + dense update terms %.1f GB per step per GPU)"
                    % (per * wl["C"] * 24 / 1e6, per * (wl["C"] * (wl["M"] + 1) + wl["M"]) * 28 / 1e9)}
    if wl.get("Cd"):
        cfg["feature_model"] = "mixed (static + constant-velocity features)"
        cfg["dynamic_components"] = wl["Cd"]
    return cfg


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).
    A timed step lasts ~15 ms, far below nvidia-smi's 100 ms loop period, so the sampler polls NVML directly
    (pynvml, every 2 ms, only while `active` is set around a timed step); nvidia-smi is the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []          # (sm_mhz, reasons bitmask) per NVML sample
        self.smi_rows = []
        self.proc = None
        self.nvml = None
        self.handle = None
        self.sm_max = None
        self.active = False
        self.stop_flag = False
        self.t = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            h = pynvml.nvmlDeviceGetHandleByIndex(int(ids[self.index]) if self.index < len(ids) else self.index)
        return pynvml, h

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.sm_max = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        """Polls for the whole timed loop; every sample remembers whether a timed step was running when the query
        STARTED (an NVML query can take longer than a 13 ms step, so the tag is taken before, not after)."""
        n = self.nvml
        while not self.stop_flag:
            tag = bool(self.active)
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((sm, rs, tag))
            except Exception:
                pass
            time.sleep(0.002)

    def n_samples(self):
        return len(self.rows) if self.nvml is not None else len(self.smi_rows)

    def _read(self):
        for line in self.proc.stdout:
            self.smi_rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        self.stop_flag = True
        if self.nvml is not None:
            if self.t:
                self.t.join(timeout=2)
            n = self.nvml
            names = [("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown),
                     ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown),
                     ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap),
                     ("hw_power_brake_slowdown", n.nvmlClocksThrottleReasonHwPowerBrakeSlowdown)]
            in_step = [r for r in self.rows if r[2]]
            # normally the samples taken inside the timed steps; if the queries were too slow for that (fewer than 3
            # landed inside a step), every sample of the timed loop (steps + the state restores between them)
            rows = in_step if len(in_step) >= 3 else self.rows
            reasons = sorted(nm for nm, bit in names if any(r[1] & bit for r in rows))
            sm = [r[0] for r in rows]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                    "samples": len(sm), "samples_inside_timed_steps": len(in_step),
                    "how": "NVML polled every 2 ms during the timed loop" +
                           (", samples taken inside the timed steps" if rows is in_step else ", all samples of the loop")}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        rows = self.smi_rows
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "how": "nvidia-smi -lms 100 over the timed loop"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def load_oracle():
    """The CPU oracle for the timed baseline: rebuilt -O3 -march=native on THIS host; the portable
    -O3 -march=x86-64-v3 build that travelled with the repository is the fallback.  Same source, bit-identical results."""
    from oracle import oracle as O
    if O._lib is None:
        O.load(O.build_native())
    flags = "-O3 -march=native, built on this host" if (O.LIB_USED or "").endswith("_native.so") else "-O3 -march=x86-64-v3"
    return O, flags


def cpu_oracle_rate(wl, target_seconds=12.0, threads=None):
    """Times the CPU oracle (all host threads) on a bounded particle sample of the same workload.
    Returns (updates/s, cores, sample description, ms per sample step, particles in the sample).
    Nothing here touches libphdslam.so: the config is built in Python (scene_config_py), the scene in numpy."""
    from phdslam_b200 import scene as S
    O, flags = load_oracle()
    C, M = wl["C"], wl["M"]
    threads = threads or os.cpu_count() or 1
    Ps = max(threads, 8)
    rate = None
    for attempt in range(3):
        extra = {k: v for k, v in wl.items() if k not in NON_CFG_KEYS}
        cfg = S.scene_config_py(Ps, C, M, max_components=wl["max_components"], **extra)
        sc = S.make_scene(Ps, C, M, seed=0)
        o = O.Oracle(cfg, threads=threads)
        S.load_scene(o, sc)
        Zs = load_mixed(o, S, wl, Ps, sc)
        t0 = time.perf_counter()
        o.step(1, np.float32([1.0, 0.05]), Zs)
        dt = time.perf_counter() - t0
        rate = Ps * pairs_per_particle(wl) / dt
        o.close()
        if dt >= 0.4 * target_seconds or attempt == 2:
            break
        Ps = int(min(max(Ps * target_seconds / max(dt, 1e-3) * 0.8, Ps * 2), 65536))
    sample = "%d of the workload's particles (C=%d, M=%d), one full filter step, %d OpenMP threads, %.2f s; oracle %s" % (
        Ps, C, M, threads, dt, flags)
    return rate, threads, sample, dt * 1e3, Ps


def cpu_oracle_single_thread_rate(wl, seconds=4.0):
    """the same step on ONE thread (a scalar port's rate; BASELINE.md section 2 asks for both)"""
    r, _, _, _, _ = cpu_oracle_rate(wl, target_seconds=seconds, threads=1)
    return r


def run_reference(args, wl):
    rank, world, local = dist_env()
    if rank != 0:
        return
    C, M = wl["C"], wl["M"]
    from phdslam_b200 import scene as S      # numpy scene generator + the ctypes image of the config; no libphdslam.so
    O, flags = load_oracle()
    threads = os.cpu_count() or 1
    # size the per-step sample so that warmup+steps finish within a few minutes
    r0, _, _, _, _ = cpu_oracle_rate(wl, target_seconds=3.0, threads=threads)
    r1 = cpu_oracle_single_thread_rate(wl, seconds=3.0)
    budget = 140.0 / max(args.steps + args.warmup, 1)
    Ps = int(max(threads, min(wl["P"], r0 * min(budget, 20.0) / pairs_per_particle(wl))))
    extra = {k: v for k, v in wl.items() if k not in NON_CFG_KEYS}
    cfg = S.scene_config_py(Ps, C, M, max_components=wl["max_components"], **extra)
    sc = S.make_scene(Ps, C, M, seed=0)
    times = []
    for k in range(args.warmup + args.steps):
        o = O.Oracle(cfg, threads=threads)
        S.load_scene(o, sc)
        Zs = load_mixed(o, S, wl, Ps, sc)
        t0 = time.perf_counter()
        o.step(1, np.float32([1.0, 0.05]), Zs)
        dt = time.perf_counter() - t0
        o.close()
        if k >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = Ps * pairs_per_particle(wl) / (ms * 1e-3)
    sample = "%d of %d particles per step (C=%d, M=%d), full filter step, %d threads; oracle %s" % (Ps, wl["P"], C, M, threads, flags)
    assert not any("libphdslam" in l for l in open("/proc/self/maps")), "the reference arm must not load the product library"
    line = {
        "impl": "reference", "metric": "GM-PHD updates/s (particle x comp x meas)", "value": value, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, wl, args.gpus),
        "run": {"note": "CPU oracle port of the reference algorithm (its scphd_cpu.cpp is an empty stub); bounded sample"},
        "cpu_baseline": {"value": value, "unit": "updates/s", "cores": threads, "kind": "port", "sample": sample,
                         "single_thread_value": r1},
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if wl.get("Cd"):
        line["metric"] = "GM-PHD updates/s (particle x (static + dynamic comp) x meas)"
    print(json.dumps(line))


def time_production_mode(filt, cfg, u, Z, args, barrier, stream, torch, dist):
    """The same workload with update_mode = 1: the fused update emits only the prune survivors (nothing reads the dense
    update terms of the reference's layout after the prune), so this is the fastest CORRECT path for filter steps/s --
    same maps, same weights (tests/test_parity_gpu.py::test_update_modes_agree).  Device-timed, max over ranks."""
    cfg.set(update_mode=1)
    filt.setDeviceConfig(cfg)
    ms = []
    for k in range(2 + args.steps):
        filt.restore()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        filt.step(1, u, Z)
        e1.record(stream)
        barrier()
        if k >= 2:
            ms.append(e0.elapsed_time(e1))
    t = filt.timings()
    cfg.set(update_mode=0)
    filt.setDeviceConfig(cfg)
    tot = float(np.sum(ms))
    if dist is not None:
        tt = torch.tensor([tot], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tot = float(tt[0])
    per = tot / args.steps
    return {"update_mode": "fused (only prune survivors leave the SM; no dense update terms)", "ms_per_step": per,
            "filter_steps_per_s": 1e3 / per, "update_ms": t.update_ms, "merge_ms": t.merge_ms}


def check_exchange(filt, u, Z, P_total, torch, dist):
    """One extra, untimed step with injected uniforms that checks the sharded resampling exchange end to end:
    every rank checksums its particles (pose, map, cardinality: one uint64 each, computed on the device) before and after
    the global resampling; rank 0 gathers weights, ancestors and checksums, recomputes the ancestors with the CPU oracle's
    canonical resampler on the gathered weights (bit-exact), and requires checksum(offspring j) == checksum(ancestor of j)
    for EVERY offspring -- in particular for those whose ancestor lived on another GPU."""
    rank, world = dist.get_rank(), dist.get_world_size()
    filt.restore()
    est = filt.step_filter(1, u, Z)
    lw = filt.log_weights
    pre = filt.particle_checksums()
    uni = np.random.Generator(np.random.Philox(12345)).uniform(0.0, 1.0, P_total + 1)
    anc = filt.resampleParticles(uni)
    post = filt.particle_checksums()

    def gather(a, dtype):
        t = torch.from_numpy(np.ascontiguousarray(a).view(dtype)).cuda()
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return np.concatenate([o.cpu().numpy() for o in out])

    lw_all = gather(lw, np.float32)
    pre_all = gather(pre, np.int64)
    post_all = gather(post, np.int64)
    anc_all = gather(anc.astype(np.int32), np.int32)
    if rank != 0:
        return None
    from phdslam_b200 import scene as S
    from oracle import oracle as O
    ocfg = S.scene_config_py(P_total, 1, 1, max_components=32)
    o = O.Oracle(ocfg)
    o.log_weights = lw_all
    oanc = o.resampleParticles(uni)
    o.close()
    anc_ok = bool((oanc == anc_all).all())
    per = P_total // world                 # rank r owns [r*N/W, (r+1)*N/W); P_total = P * world
    own_j = np.arange(P_total) // per
    own_a = anc_all // per
    migrated = int((own_j != own_a).sum())
    copy_ok = bool((post_all == pre_all[anc_all]).all())
    return {"ancestors": "bit-exact against the CPU oracle's canonical resampler on the gathered weights" if anc_ok else "MISMATCH",
            "offspring_checked": int(P_total), "migrated_checked": migrated,
            "copies": "checksum(offspring) == checksum(ancestor) for every offspring" if copy_ok else "MISMATCH",
            "neff": float(est.neff), "ok": bool(anc_ok and copy_ok)}


def run_ours(args, wl):
    import torch
    import phdslam_b200 as PS
    from phdslam_b200 import scene as S
    rank, world, local = dist_env()
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch N>1 with torch.distributed.run" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the filter has no CPU fallback (use --impl reference for the CPU oracle)")
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P, C, M = wl["P"], wl["C"], wl["M"]
    strong = bool(wl.get("strong"))
    if strong:                               # strong scaling: wl["P"] particles in total
        if wl["P"] % world:
            raise SystemExit("the particle count of a strong-scaling workload must be a multiple of --gpus")
        P_total, P = wl["P"], wl["P"] // world
    else:
        P_total = P * world                  # weak scaling: P particles per GPU
    extra = {k: v for k, v in wl.items() if k not in NON_CFG_KEYS}
    cfg = S.scene_config(P_total, C, M, max_components=wl["max_components"], seed="0", **extra)
    filt = PS.PhdSlam(cfg, device=local)
    if world > 1:
        filt.dist_init(rank, world)
    assert filt.n_local == P
    # every rank holds P particles of the same landmark scene (same measurements Z), with its own pose / map jitter
    n_scene = min(P, int(wl.get("scene_particles", P)))
    sc = S.make_scene(n_scene, C, M, seed=0, particle_seed=rank)
    sc["log_weights"][:] = -np.log(np.float32(P_total))
    if n_scene < P:
        filt.import_tiled(sc)                # n_scene distinct particles, repeated on the device
    else:
        S.load_scene(filt, sc)
    Z = load_mixed(filt, S, wl, n_scene, sc, particle_seed=rank)
    filt.snapshot()
    u = np.float32([1.0, 0.05])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        filt.synchronize()

    stream = torch.cuda.ExternalStream(filt.stream, device=torch.device("cuda", local))
    for _ in range(args.warmup):
        filt.restore()
        filt.step(1, u, Z)
    barrier()
    t_before = filt.timings()
    l0, mig0 = t_before.launches, t_before.migrated_in
    h2d0, d2h0 = t_before.h2d_bytes, t_before.d2h_bytes
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms, wall_ms, upd_ms, mrg_ms, other, dyn_ms = [], [], [], [], [], []
    n_resampled = 0
    for k in range(args.steps):
        filt.restore()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.active = True
        t0 = time.perf_counter()
        e0.record(stream)
        est, res = filt.step(1, u, Z)
        e1.record(stream)
        barrier()
        t1 = time.perf_counter()
        sampler.active = False
        n_resampled += int(res)
        dev_ms.append(e0.elapsed_time(e1))
        wall_ms.append((t1 - t0) * 1e3)
        t = filt.timings()
        upd_ms.append(t.update_ms)
        mrg_ms.append(t.merge_ms)
        other.append((t.predict_ms, t.weights_ms, t.estimate_ms, t.resample_ms if res else 0.0))
        dyn_ms.append(t.dynamic_ms)
    t_after = filt.timings()
    l_timed, mig_timed = t_after.launches, t_after.migrated_in
    # bytes the timed phdslam_step() calls moved between host and device, counted inside the library at every copy call
    h2d_per_step = (t_after.h2d_bytes - h2d0) / float(args.steps)
    d2h_per_step = (t_after.d2h_bytes - d2h0) / float(args.steps)
    # clocks under load are part of the contract: if (almost) no sample landed during the timed loop -- NVML can be slow
    # right after a profiler run -- all ranks run extra UNTIMED steps under the sampler until there are enough
    need_more = torch.tensor([1 if (rank == 0 and sampler.n_samples() < 3) else 0], device="cuda", dtype=torch.int32)
    if world > 1:
        dist.broadcast(need_more, 0)
    if int(need_more.item()):
        for _ in range(40):
            filt.restore()
            sampler.active = True
            filt.step(1, u, Z)
            filt.synchronize()
            sampler.active = False
            if world == 1 and sampler.n_samples() >= 10:
                break
    clocks = sampler.stop() if rank == 0 else None
    try:
        production = time_production_mode(filt, cfg, u, Z, args, barrier, stream, torch, dist if world > 1 else None)
    except Exception as e:       # the extra key must never cost the headline line (every rank takes the same path)
        production = {"error": str(e)[:200]}
    exchange_check = check_exchange(filt, u, Z, P_total, torch, dist) if world > 1 else None
    launches = l_timed - l0
    # restore() launches no kernels (cudaMemcpyAsync only), so `launches` counts the timed steps' kernels
    dev_total, wall_total = float(np.sum(dev_ms)), float(np.sum(wall_ms))
    exchange = None
    if world > 1:
        tt = torch.tensor([dev_total, wall_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_total, wall_total = float(tt[0]), float(tt[1])
        # global resampling: particles this rank received over NVLink (pose 24 B + size 4 B + ancestor 4 B + the map block
        # of 6 planes x Cmax floats [+ cardinality]) against the time of the whole resampling phase (CDF scan, search,
        # local gather and the send/recv ring), max over ranks
        cmax = (wl["max_components"] + 31) // 32 * 32
        rec = 32 + 24 * cmax + (4 * (wl.get("max_cardinality", -1) + 1) if wl.get("filter_type") == 1 else 0)
        mig = float(mig_timed - mig0) / max(args.steps, 1)
        rs = float(np.mean([o[3] for o in other]))
        mm = torch.tensor([mig, rs], device="cuda", dtype=torch.float64)
        dist.all_reduce(mm, op=dist.ReduceOp.MAX)
        mig, rs = float(mm[0]), float(mm[1])
        exchange = {"mode": ("nvlink peer-window push: the gather kernel writes offspring into the owner's buffers (CUDA IPC)"
                             if filt.dist_p2p else "nccl send/recv ring (pack, send, unpack)"),
                    "migrated_particles_per_step_max_rank": mig, "bytes_per_step_max_rank": mig * rec, "resample_phase_ms": rs,
                    "achieved_GBps_lower_bound": (mig * rec / (rs * 1e-3) / 1e9) if rs > 0 else None,
                    "nvlink5_peak_GBps_per_direction": 900.0,
                    "note": "bytes received by the busiest rank / the whole resampling phase (scan + search + gather + ring)"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    updates_per_step = float(P_total) * pairs_per_particle(wl)
    ms_per_step = dev_total / args.steps
    value = updates_per_step / (ms_per_step * 1e-3)
    e2e_value = updates_per_step / (wall_total / args.steps * 1e-3)
    peak, peak_src = measured_peaks()
    upd = float(np.mean(upd_ms))
    balg = alg_bytes_per_update(C, M, wl.get("max_cardinality", -1) + 1 if wl.get("filter_type") == 1 else 0)
    achieved = (float(P) * C * M * balg) / (upd * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get("update_dense_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    # the prune + merge kernel is not HBM-bound (175 MB of DRAM reads per 8192 particles): its bound is the SM issue rate.
    # warp-instructions per particle come from the committed ncu capture, the duration is measured live.
    merge_roof = None
    try:
        wi = json.load(open(tp)).get(args.workload, {}).get("merge_fast_warp_instructions_per_particle")
        if wi and clocks and clocks.get("sm_max_mhz"):
            mrg = float(np.mean(mrg_ms))
            peak_issue = 148 * 4 * clocks["sm_max_mhz"] * 1e6 / 1e9          # G warp-instructions/s: 148 SMs x 4 schedulers
            ach = float(P) * wi / (mrg * 1e-3) / 1e9
            merge_roof = {"bound": "issue", "kernel": "merge_fast_kernel", "achieved": ach, "peak": peak_issue,
                          "unit": "G warp-instr/s", "frac": ach / peak_issue, "kernel_ms": mrg,
                          "warp_instructions_per_particle": wi}
    except Exception:
        merge_roof = None
    cpu_rate, cores, sample, _, _ = cpu_oracle_rate(wl) if not args.no_cpu_baseline else (None, 0, "skipped", 0, 0)
    oth = np.mean(np.array(other), axis=0)
    line = {
        "metric": "GM-PHD updates/s (particle x comp x meas)", "value": value, "unit": "updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, wl, world),
        "run": {"update_mode": "dense (reference-equivalent update terms materialised in HBM)",
                "distinct_particles_per_gpu": n_scene, "steps_that_resampled": n_resampled},
        "filter_steps_per_s": 1e3 / ms_per_step,
        "phase_ms": {"update": upd, "merge": float(np.mean(mrg_ms)), "predict": float(oth[0]), "weights": float(oth[1]),
                     "estimate": float(oth[2]), "resample": float(oth[3])},
        "roofline": {"bound": "hbm", "kernel": "update_dense_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_update": balg, "kernel_ms": upd,
                     "timing": "CUDA events on the filter's stream around the kernel, averaged over the timed steps"},
        "cpu_baseline": {"value": cpu_rate, "unit": "updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": e2e_value, "unit": "updates/s", "h2d_bytes_per_step": h2d_per_step, "d2h_bytes_per_step": d2h_per_step,
                "ms_per_step": wall_total / args.steps,
                "bytes": "counted inside libphdslam.so at every host<->device copy of the timed calls (measurement upload; "
                         "term counts, update status, estimate and CDF total coming back)"},
        "production": production,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if wl.get("Cd"):
        line["phase_ms"]["dynamic"] = float(np.mean(dyn_ms))      # dyn_pre_kernel + dyn_update_kernel (not in update / merge)
        line["metric"] = "GM-PHD updates/s (particle x (static + dynamic comp) x meas)"
    if exchange is not None:
        line["exchange"] = exchange
        line["exchange_check"] = exchange_check
    if merge_roof is not None:
        line["roofline_merge"] = merge_roof
    print(json.dumps(line))
    sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
