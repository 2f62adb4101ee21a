"""GPU box: device-resident throughput of the BASELINE.json configs that are not the bench line -- configs[3]
(descending view, rotated 180 degrees, occluders, 1024 frames) and configs[4] (4096x3072, 12 steps, extended range) --
next to configs[2]. Parity of the same configs is what tests/test_gpu_parity.py checks; this only times them.
    python tools/configs_bench.py > gpurun_out/<tag>/configs.jsonl"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import stair_step_detector_b200 as S
from stair_step_detector_b200 import _abi as A

NOISY = dict(noise_sigma=0.0025, dropout=0.03, n_holes=3)
CASES = [
    ("configs[2] batch 1024x768, 3-8 steps", (1024, 768), {}, dict(**NOISY), 2048, (3, 8)),
    ("configs[3] descending (rotate180, 2 occluders) 1024x768", (1024, 768), {}, dict(rotate180=1, n_occluders=2, **NOISY), 1024, (3, 8)),
    ("configs[4] hi-res 4096x3072, 12 steps, range y<3.7 z<2.3", (4096, 3072), dict(y_max=3.7, z_max=2.3),
     dict(n_steps=12, riser=0.17, tread=0.26, cam_height=3.2, cam_pitch_deg=48.0, first_riser_y=0.5, **NOISY), 64, (12, 12)),
]
if os.environ.get("CONFIGS_ONLY"):  # e.g. CONFIGS_ONLY=2: only the hi-res case (profiling)
    CASES = [CASES[int(i)] for i in os.environ["CONFIGS_ONLY"].split(",")]
for name, (w, h), ck, sk, frames, (smin, smax) in CASES:
    N = w * h
    cfg = S.default_config(w, h, **ck)
    base = S.default_scene(w, h, **sk)
    xf = S.scene_transform(base)
    det = S.Detector(cfg, xf, max_frames=frames)
    d = det.malloc(frames * N * 12)
    det.synth_frames(base, 1, 0, frames, smin, smax, d)
    for _ in range(3):
        det.process_device(d, frames)
    ts = []
    for _ in range(5):
        det.process_device(d, frames)
        ts.append(det.timing().total_ms)
    det.process_device(d, frames, flags=A.FLAG_STAGE_TIMING | A.FLAG_SINGLE_STREAM)
    st = det.stage_times()
    ms = min(ts)
    fps = frames / (ms * 1e-3)
    nst = det.n_steps_all(frames)
    print(json.dumps({"config": name, "frames": frames, "chunk": det.chunk_frames, "ms_best": round(ms, 3), "kfps": round(fps / 1e3, 2),
                      "Gpoints_s": round(fps * N / 1e9, 1), "chain_GBs_13B": round(13 * fps * N / 1e9, 0),
                      "steps_per_frame_mean": round(float(nst.mean()), 2), "frames_with_steps": int((nst > 0).sum()),
                      "serial_stage_ms": {k: round(v[0], 3) for k, v in st.items()}}), flush=True)
    det.free(d)
    det.close()
