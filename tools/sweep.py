"""Scratch tuning harness (GPU box): chunk-size sweep + per-stage CUDA-event times.
    python tools/sweep.py --frames 1024 --chunks 4,8,16,32,64"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import stair_step_detector_b200 as S
from stair_step_detector_b200 import _abi as A

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=1024)
ap.add_argument("--chunks", default="4,8,16,32,64,128")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--streams", default="2")
ap.add_argument("--warm", type=int, default=3)
ap.add_argument("--w", type=int, default=1024)
ap.add_argument("--h", type=int, default=768)
args = ap.parse_args()
W, H = args.w, args.h
N = W * H
cfg = S.default_config(W, H)
base = S.default_scene(W, H, noise_sigma=0.0025, dropout=0.03, n_holes=3)
xf = S.scene_transform(base)
for ns, cf in [(int(n), int(c)) for n in args.streams.split(",") for c in args.chunks.split(",")]:
    os.environ["SSD_GPU_CHUNK_FRAMES"] = str(cf)
    os.environ["SSD_GPU_STREAMS"] = str(ns)
    det = S.Detector(cfg, xf, max_frames=args.frames)
    d = det.malloc(args.frames * N * 12)
    det.synth_frames(base, 1, 0, args.frames, 3, 8, d)
    for _ in range(args.warm):
        det.process_device(d, args.frames)
    ts = []
    for _ in range(args.reps):
        det.process_device(d, args.frames)
        ts.append(det.timing().total_ms)
    det.process_device(d, args.frames, flags=A.FLAG_STAGE_TIMING | A.FLAG_SINGLE_STREAM)
    st_serial = det.stage_times()
    det.process_device(d, args.frames, flags=A.FLAG_STAGE_TIMING)
    st = det.stage_times()
    tot = det.timing().total_ms
    ms = min(ts)
    fps = args.frames / (ms * 1e-3)
    print(json.dumps({"streams": ns, "chunk": det.chunk_frames, "ms_best": round(ms, 3), "ms_med": round(float(np.median(ts)), 3), "kfps": round(fps / 1e3, 1),
                      "Gpts": round(fps * N / 1e9, 1), "chain_GBs": round(13 * fps * N / 1e9, 0), "staged_total": round(tot, 3),
                      "stages": {k: round(v[0], 3) for k, v in st.items()},
                      "serial": {k: round(v[0], 3) for k, v in st_serial.items()}, "steps": int(det.n_steps_all(args.frames).sum()), "exact_frac": det.stats().n_exact_fallback / max(1, det.stats().n_points), "quad_fast": det.stats().n_quad_fast, "quad_exact": det.stats().n_quad_exact}), flush=True)
    det.free(d)
    det.close()
