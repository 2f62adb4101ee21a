"""GPU box: per-kernel times of ONE frame through the chain (the reference's real use: one frame per call).
    python tools/latency.py [--w 1024 --h 768]"""
import argparse, json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import stair_step_detector_b200 as S
from stair_step_detector_b200 import _abi as A
ap = argparse.ArgumentParser()
ap.add_argument("--w", type=int, default=1024)
ap.add_argument("--h", type=int, default=768)
ap.add_argument("--frames", type=int, default=1)
a = ap.parse_args()
N = a.w * a.h
cfg = S.default_config(a.w, a.h)
base = S.default_scene(a.w, a.h, noise_sigma=0.0025, dropout=0.03, n_holes=3)
xf = S.scene_transform(base)
with S.Detector(cfg, xf, max_frames=a.frames) as det:
    d = det.malloc(a.frames * N * 12)
    det.synth_frames(base, 1, 0, a.frames, 0 if a.frames == 1 else 3, 0 if a.frames == 1 else 8, d)
    for _ in range(10):
        det.process_device(d, a.frames)
    ev, wall = [], []
    for _ in range(100):
        t0 = time.perf_counter()
        det.process_device(d, a.frames)
        wall.append((time.perf_counter() - t0) * 1e6)
        ev.append(det.timing().total_ms * 1e3)
    st = {}
    for _ in range(20):
        det.process_device(d, a.frames, flags=A.FLAG_STAGE_TIMING | A.FLAG_SINGLE_STREAM)
        for k, v in det.stage_times().items():
            st.setdefault(k, []).append(v[0] * 1e3)
    print(json.dumps({"size": [a.w, a.h], "frames": a.frames, "events_us_median": round(statistics.median(ev), 1), "events_us_min": round(min(ev), 1),
                      "wall_us_median": round(statistics.median(wall), 1), "launches": det.timing().n_launches,
                      "stage_us_median": {k: round(statistics.median(v), 1) for k, v in st.items()}, "steps": int(det.n_steps_all(a.frames)[0])}))
    det.free(d)
