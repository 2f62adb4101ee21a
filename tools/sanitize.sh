#!/bin/bash
# GPU box: compute-sanitizer over a small end-to-end run of the chain (all seven kernels + the depth path).
#   bash tools/sanitize.sh <tag>
TAG=${1:-san}; OUT=gpurun_out/$TAG; mkdir -p $OUT
cat > /tmp/san_run.py <<'P'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import helpers, stair_step_detector_b200 as S
W, H = 640, 480
cfg = S.default_config(W, H)
base = S.default_scene(W, H, noise_sigma=0.0025, dropout=0.03, n_holes=3)
xf = S.scene_transform(base)
scenes = [S.randomize_scene(base, 7, i, 3, 8) for i in range(6)]
depth = np.stack([S.synth_depth_host(sc) for sc in scenes])
xyz = np.stack([S.deproject_host(sc, d) for sc, d in zip(scenes, depth)])
orc = helpers.load_oracle()
with S.Detector(cfg, xf, max_frames=len(scenes)) as det:
    det.process_host(xyz)
    a = [det.steps(f)[0] for f in range(len(scenes))]
    labels = [det.labels(f) for f in range(len(scenes))]
    det.process_depth_host(depth, S.scene_intrinsics(base))
    b = [det.steps(f)[0] for f in range(len(scenes))]
for f in range(len(scenes)):
    o = helpers.oracle_process(orc, cfg, xf, xyz[f])
    assert np.array_equal(labels[f], o.labels) and len(a[f]) == len(o.steps) == len(b[f])
print("sanitizer run ok:", [len(x) for x in a])
P
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_run.py > $OUT/$tool.log 2>&1
  tail -4 $OUT/$tool.log
done
