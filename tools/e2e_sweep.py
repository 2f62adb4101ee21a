"""Scratch (GPU box): e2e throughput of the host-input entry points against the host chunk size.
    python tools/e2e_sweep.py --frames 1024 --chunks 32,64,128,256,512"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import stair_step_detector_b200 as S

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=1024)
ap.add_argument("--chunks", default="32,64,128,256,512")
ap.add_argument("--streams", default="3")
ap.add_argument("--reps", type=int, default=4)
args = ap.parse_args()
W, H = 1024, 768
N = W * H
cfg = S.default_config(W, H)
base = S.default_scene(W, H, noise_sigma=0.0025, dropout=0.03, n_holes=3)
xf = S.scene_transform(base)
intr = S.scene_intrinsics(base)
h_depth = hd = None
for ns in [int(x) for x in args.streams.split(",")]:
    for hc in [int(c) for c in args.chunks.split(",")]:
        os.environ["SSD_GPU_HOST_CHUNK_FRAMES"] = str(hc)
        os.environ["SSD_GPU_STREAMS"] = str(ns)
        det = S.Detector(cfg, xf, max_frames=args.frames)
        if h_depth is None:
            d_depth = det.malloc(args.frames * N * 2)
            d_tmp = det.malloc(args.frames * N * 12)
            det.synth_frames(base, 1, 0, args.frames, 3, 8, d_tmp, d_depth)
            h_depth, hd = S.pinned_empty((args.frames, N), np.uint16)
            det.d2h(h_depth, d_depth)
            det.free(d_tmp)
            det.free(d_depth)
        for _ in range(2):
            det.process_depth_host_ptr(hd, intr, args.frames)
        ts = []
        for _ in range(args.reps):
            det.process_depth_host_ptr(hd, intr, args.frames)
            ts.append(det.timing().total_ms)
        ms = min(ts)
        print(json.dumps({"streams": ns, "host_chunk": hc, "ms_best": round(ms, 3), "ms_med": round(float(np.median(ts)), 3),
                          "kfps": round(args.frames / ms, 2), "h2d_GBs": round(args.frames * N * 2 / ms / 1e6, 1),
                          "steps": int(det.n_steps_all(args.frames).sum())}), flush=True)
        det.close()
