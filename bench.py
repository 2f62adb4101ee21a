#!/usr/bin/env python
"""bench.py -- headline benchmark of the stair-step geometry hot path (BASELINE.json metric:
Mpoints/s and frames/s at 1024x768 on 1/2/4/8 B200, % of HBM roofline).

    python bench.py --gpus N --steps K --warmup W          # our CUDA path
    python bench.py --impl reference ...                   # the reference's own CPU code on the host cores

A "step" is one pass of the whole chain (transform -> height-band labels -> per-step reductions -> Stairs
records on the host) over one batch of synthetic 1024x768 frames that is already resident in HBM
(BASELINE.json configs[2]: 4096 frames of 3-8-step staircases with L515-class noise). One process per GPU;
under torchrun every rank owns its own batch (no collective on the data path: frames are independent), the
timed region is bracketed by barrier + synchronize and the MAX over ranks is reported.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H = 1024, 768
N = W * H
ALGO_BYTES_PER_POINT = 13  # 12 B packed f32 xyz read + 1 B u8 label written
NOISY = dict(noise_sigma=0.0025, dropout=0.03, n_holes=3)
BASE_SEED = 20261017


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region (B200_PROFILING.md recipe). NVML in a sampling thread
    (every 5 ms: the timed region is a few hundred ms); `nvidia-smi -lms` as the fallback when pynvml is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, device):
        self.device, self.proc, self.path, self.thread = device, None, None, None
        self.sm, self.power, self.bits, self.smax = [], [], 0, None

    def _nvml_loop(self):
        import pynvml as nv
        h = self.handle
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1e3)
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                self.bits |= int(get(h))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            self.handle = nv.nvmlDeviceGetHandleByIndex(self.device)
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            self.stop_flag = False
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if not self.sm:
                return {"sm_mhz": None, "sm_max_mhz": self.smax, "reasons": ["no samples"]}
            return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.smax, "power_w_max": max(self.power), "samples": len(self.sm),
                    "source": "nvml", "reasons": sorted(name for bit, name in self.REASONS if self.bits & bit)}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, smax, reasons, power = [], [], set(), []
        with open(self.path) as f:
            for line in f:
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1]))
                    smax.append(float(c[2]))
                    power.append(float(c[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power), "samples": len(sm),
                "source": "nvidia-smi", "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def base_scene(S):
    return S.default_scene(W, H, **NOISY)


def cpu_reference_leg(S, xyz_sample, xf, threads, target_cpu_seconds):
    """The reference's own Pointcloud::process (oracle/_ref, compiled from the reference sources by path) on a
    bounded sample of the workload, frame-parallel over `threads` host threads. Falls back to the C port."""
    import helpers
    cfg = S.default_config(W, H)
    nf = xyz_sample.shape[0]
    ref = helpers.load_ref(cfg)
    secs = C.c_double()
    failed = C.c_int()
    if ref is not None:
        kind = "reference"
        ref.ssd_ref_process_timed(C.byref(xf), helpers.ptr(xyz_sample), min(nf, threads), threads, 1, C.byref(secs), C.byref(failed))  # warm
        per_frame = max(secs.value * threads / max(1, min(nf, threads)), 1e-3)
        repeat = max(1, int(round(target_cpu_seconds / (per_frame * nf))))
        rc = ref.ssd_ref_process_timed(C.byref(xf), helpers.ptr(xyz_sample), nf, threads, repeat, C.byref(secs), C.byref(failed))
        assert rc == 0
        frames = nf * repeat
        t = secs.value
    else:
        kind = "port"
        orc = helpers.load_oracle()
        threads = 1
        t0 = time.perf_counter()
        rc = orc.ssd_oracle_process_batch(C.byref(cfg), C.byref(xf), helpers.ptr(xyz_sample), nf, None)
        t = time.perf_counter() - t0
        assert rc == 0
        frames, repeat = nf, 1
    return {"value": frames * N / t / 1e6, "unit": "Mpoints/s", "frames_per_s": frames / t, "cores": threads, "kind": kind,
            "host_cores": os.cpu_count(),
            "sample": f"{nf} frames of the same synthetic batch x{repeat} passes, Pointcloud::process per frame, "
                      f"{threads} host thread(s) (threading in the harness; shim 3x3 close instead of OpenCV's)",
            "seconds": t, "failed_frames": failed.value if ref is not None else 0}


def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return
    import stair_step_detector_b200 as S
    base = base_scene(S)
    xf = S.scene_transform(base)
    nf = 64
    xyz = np.stack([S.deproject_host(sc, S.synth_depth_host(sc))
                    for sc in (S.randomize_scene(base, BASE_SEED, i, 3, 8) for i in range(nf))]).reshape(nf, N, 3)
    threads = os.cpu_count() or 1
    # each "step" is one bounded sample; keep the whole run within a few minutes
    res = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_leg(S, xyz, xf, threads, target_cpu_seconds=0.0)
        if i >= args.warmup:
            res.append(r)
    t = sum(r["seconds"] for r in res)
    frames = nf * len(res)
    val = frames * N / t / 1e6
    cb = dict(res[-1])
    cb["value"] = val
    line = {"impl": "reference", "metric": "Mpoints/s", "value": val, "unit": "Mpoints/s", "frames_per_s": frames / t,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / len(res),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[2] sample: 64 synthetic 1024x768 frames of 3-8-step staircases with L515-class noise per step",
                       "width": W, "height": H, "frames_per_step": nf},
            "cpu_baseline": cb,
            "e2e": {"value": val, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def run_ours(args):
    rank, world, local = dist_env()
    import stair_step_detector_b200 as S
    from stair_step_detector_b200 import _abi as A

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier_sync():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    if world > 1:
        # one process per GPU: run on the CPUs (and allocate the pinned host buffers on the memory) next to that GPU
        try:
            import pynvml as nv
            nv.nvmlInit()
            nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(local))
        except Exception:
            pass

    frames = args.frames
    cfg = S.default_config(W, H)
    base = base_scene(S)
    xf = S.scene_transform(base)
    det = S.Detector(cfg, xf, device=local, max_frames=frames)
    d_xyz = det.malloc(frames * N * 12)
    # every rank owns its own frames: global frame id = rank*frames + i (camera pose fixed by the calibration,
    # staircase geometry varies per frame)
    det.synth_frames(base, BASE_SEED, rank * frames, frames, 3, 8, d_xyz)

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        det.process_device(d_xyz, frames)

    # ---- timed: device-resident input -> Stairs records on the host ----
    sampler = ClockSampler(local)
    barrier_sync()
    sampler.start()
    t_wall0 = time.perf_counter()
    dev_ms, launches = [], 0
    for _ in range(args.steps):
        det.process_device(d_xyz, frames)
        t = det.timing()
        dev_ms.append(t.total_ms)
        launches += t.n_launches
    barrier_sync()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop()
    my_ms = sum(dev_ms)
    n_steps_found = int(det.n_steps_all(frames).sum())

    # ---- roofline of the dominant kernel: per-launch CUDA events on the launching streams ----
    # (one extra pass right after the timed region, chunks serialised on one stream so that an event bracket is the
    #  kernel's own duration and not its duration while sharing the GPU with the other streams' kernels)
    det.process_device(d_xyz, frames, flags=A.FLAG_STAGE_TIMING | A.FLAG_SINGLE_STREAM)
    stages = det.stage_times()
    stage_total = det.timing().total_ms
    det.process_device(d_xyz, frames, flags=A.FLAG_STAGE_TIMING)
    stages_overlapped = det.stage_times()

    # ---- e2e: the reference-facing call with HOST buffers (pinned), host->device copies and the device->host read of the
    # results inside the timed region. The reference's process() receives the z16 DEPTH FRAME (Camera::DepthFrame,
    # pointcloud.cpp:608,138), so that is the host buffer of the headline e2e: 2 bytes per pixel cross PCIe and the
    # deprojection runs on the GPU (ssd_gpu_process_depth_host). The same through the packed-vertex entry point
    # (12 bytes per pixel, ssd_gpu_process_host) is reported next to it.
    intr = S.scene_intrinsics(base)
    e2e_frames = min(args.e2e_frames, frames)
    d_depth = det.malloc(e2e_frames * N * 2)
    d_tmp = det.malloc(e2e_frames * N * 12)
    det.synth_frames(base, BASE_SEED, rank * frames, e2e_frames, 3, 8, d_tmp, d_depth)  # same frames as the device-resident batch
    det.free(d_tmp)
    h_depth, h_depth_handle = S.pinned_empty((e2e_frames, N), np.uint16)
    det.d2h(h_depth, d_depth)
    det.free(d_depth)
    for _ in range(2):
        det.process_depth_host_ptr(h_depth_handle, intr, e2e_frames)
    e2e_steps_found = int(det.n_steps_all(e2e_frames).sum())
    barrier_sync()
    e2e_ms = []
    for _ in range(max(3, args.steps // 2)):
        det.process_depth_host_ptr(h_depth_handle, intr, e2e_frames)
        e2e_ms.append(det.timing().total_ms)
    barrier_sync()
    my_e2e = sum(e2e_ms) / len(e2e_ms)
    # the same depth frames already resident in HBM (no PCIe): what the fused deprojection buys on the device
    d_depth2 = det.malloc(e2e_frames * N * 2)
    det.h2d(d_depth2, h_depth)
    for _ in range(2):
        det.process_depth_device(d_depth2, intr, e2e_frames)
    dd_ms = []
    for _ in range(3):
        det.process_depth_device(d_depth2, intr, e2e_frames)
        dd_ms.append(det.timing().total_ms)
    det.free(d_depth2)
    my_dd = sum(dd_ms) / len(dd_ms)
    # packed vertices through PCIe
    v_frames = min(256, e2e_frames)
    h_xyz, h_handle = S.pinned_empty((v_frames, N, 3), np.float32)
    det.d2h(h_xyz, d_xyz)
    det.process_host_ptr(h_handle, v_frames)
    v_ms = []
    for _ in range(3):
        det.process_host_ptr(h_handle, v_frames)
        v_ms.append(det.timing().total_ms)
    my_e2e_v = sum(v_ms) / len(v_ms)

    # ---- max over ranks ----
    if dist is not None:
        from stair_step_detector_b200 import sharding
        (my_ms, my_e2e, wall_ms, my_e2e_v, my_dd), (launches, n_steps_found) = sharding.reduce_timing(
            dist, [my_ms, my_e2e, wall_ms, my_e2e_v, my_dd], [launches, n_steps_found], device=f"cuda:{local}")

    if rank == 0:
        ms_per_step = my_ms / args.steps
        total_frames = frames * world
        fps = total_frames / (ms_per_step * 1e-3)
        mpts = fps * N / 1e6
        peak, peak_src = peaks()
        k_ms, k_n = stages["transform_bin"]
        pts_per_launch = det.chunk_frames * N
        achieved = ALGO_BYTES_PER_POINT * pts_per_launch / (k_ms / k_n * 1e-3) / 1e9 if k_n and k_ms > 0 else None
        chain_achieved = ALGO_BYTES_PER_POINT * (fps / world) * N / 1e9
        roofline = {"bound": "hbm", "kernel": "k_transform_bin", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak if achieved else None, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": ALGO_BYTES_PER_POINT * pts_per_launch, "launches_timed": k_n,
                    "avg_launch_ms": k_ms / k_n if k_n else None,
                    "chain": {"achieved": chain_achieved, "frac": chain_achieved / peak, "frac_of_nominal_8000": chain_achieved / 8000.0,
                              "note": "whole chain per GPU: 13 B/point x points/s"},
                    "stage_ms_sum": {k: v[0] for k, v in stages.items()}, "stage_total_ms": stage_total,
                    "stage_ms_sum_overlapped": {k: v[0] for k, v in stages_overlapped.items()},
                    "note": "per-launch CUDA events on the launching stream; chunks serialised on one stream for this pass "
                            "(the timed region overlaps chunks on several streams)"}
        traffic_file = os.path.join(ROOT, "profiles", "traffic_latest.json")
        if os.path.exists(traffic_file):
            try:
                with open(traffic_file) as f:
                    tj = json.load(f)
                per = tj["per_kernel"]
                # measured DRAM bytes per frame (one ncu --set full launch) scaled to this run's launch size
                roofline["traffic"] = per["k_transform_bin"]["bytes_per_frame"] * det.chunk_frames
                roofline["traffic_source"] = tj.get("source")
                roofline["chain"]["traffic_bytes_per_frame"] = {k: v["bytes_per_frame"] for k, v in per.items()}
            except Exception:
                pass
        if roofline["frac"] and roofline["frac"] > 1.0:
            roofline["note"] += ("; frac > 1: the peak is the driver's COPY bandwidth (equal read and write streams), "
                                 "this kernel reads 12 bytes for every byte it writes and a read-dominated stream runs faster than a copy")
        e2e_fps = e2e_frames * world / (my_e2e * 1e-3)
        line = {"metric": "Mpoints/s", "value": mpts, "unit": "Mpoints/s", "frames_per_s": fps, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "wall_ms_per_step": wall_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "configs[2]: batch of synthetic 1024x768 L515-resolution frames of 3-8-step staircases "
                                       "(depth noise sigma 2.5 mm, 3 % dropouts, 3 holes), device-resident, per GPU",
                           "width": W, "height": H, "frames_per_gpu": frames, "global_frames": total_frames,
                           "chunk_frames": det.chunk_frames, "parallelism": f"frames sharded over {world} GPU(s), no collective",
                           "cache": f"inputs larger than L2 ({frames * N * 12 / 1e9:.1f} GB of vertices per GPU per step)",
                           "stairs_found": n_steps_found},
                "clocks": clocks,
                "e2e": {"value": e2e_fps * N / 1e6, "unit": "Mpoints/s", "frames_per_s": e2e_fps, "frames_per_step": e2e_frames * world,
                        "ms_per_step": my_e2e, "h2d_bytes_per_step": e2e_frames * N * 2,
                        "d2h_bytes_per_step": e2e_frames * (32 + 72 * A.MAX_STEPS + 16), "stairs_found_per_gpu": e2e_steps_found,
                        "input": "z16 depth frames in pinned host memory (what the reference's Pointcloud::process receives), "
                                 "deprojected on the GPU",
                        "call": "ssd_gpu_process_depth_host",
                        "device_resident_depth": {"value": e2e_frames * world / (my_dd * 1e-3) * N / 1e6, "unit": "Mpoints/s",
                                                  "frames_per_s": e2e_frames * world / (my_dd * 1e-3), "ms_per_step": my_dd,
                                                  "call": "ssd_gpu_process_depth_device (z16 frames in HBM, 2 B/point, deprojected "
                                                          "inside the point kernels: no PCIe in this figure)"},
                        "vertices": {"value": v_frames * world / (my_e2e_v * 1e-3) * N / 1e6, "unit": "Mpoints/s",
                                     "frames_per_s": v_frames * world / (my_e2e_v * 1e-3), "frames_per_step": v_frames * world,
                                     "ms_per_step": my_e2e_v, "h2d_bytes_per_step": v_frames * N * 12,
                                     "call": "ssd_gpu_process_host (packed f32 vertices in pinned host memory; PCIe-bound)"}},
                "gpu_launches": launches,
                "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            sample = np.ascontiguousarray(h_xyz[:min(32, v_frames)])
            line["cpu_baseline"] = cpu_reference_leg(S, sample, xf, os.cpu_count() or 1, target_cpu_seconds=20.0)
        emit(line)

    S.free_pinned(h_handle)
    S.free_pinned(h_depth_handle)
    det.free(d_xyz)
    det.close()
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version banner on stdout at
    communicator creation): point fd 1 at stderr for the duration of the run and keep the real stdout for emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=4096, help="frames per GPU per step (BASELINE config: 4096)")
    ap.add_argument("--e2e-frames", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
