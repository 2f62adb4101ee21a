"""GPU box, debug build (-DSSD_OL_PROF, SSD_GPU_LIB=build/libssd_olprof.so): cycle stamps of k_outline's phases for one plateau of one frame."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stair_step_detector_b200 as S
W, H = 1024, 768; N = W * H
cfg = S.default_config(W, H); base = S.default_scene(W, H, noise_sigma=0.0025, dropout=0.03, n_holes=3); xf = S.scene_transform(base)
with S.Detector(cfg, xf, max_frames=1) as det:
    d = det.malloc(N * 12); det.synth_frames(base, 1, 0, 1, 0, 0, d)
    for _ in range(3): det.process_device(d, 1)
