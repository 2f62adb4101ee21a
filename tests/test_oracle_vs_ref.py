"""CPU: the oracle against the reference's own code (oracle/_ref, compiled by path from /root/reference in the
build container; the prebuilt library travels to the GPU box). Skipped when neither is available."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from stair_step_detector_b200 import _abi as A

NOISY = dict(noise_sigma=0.0025, dropout=0.03, n_holes=3)
# status bits the harness around the compiled reference does not raise (it reports only NO_STEPS / DEGENERATE_QUAD; the quirks
# themselves -- all-zero ground step, NaN mean, wrap -- are in the compared RESULTS)
QUIRKS = A.STATUS_BEV_OOB | A.STATUS_INVALID_FRONT_EDGE | A.STATUS_EMPTY_MEAN | A.STATUS_HMIN_WRAP | A.STATUS_TOO_MANY_PLATEAUS


def get_ref(S, w, h):
    ref = H.load_ref(S.default_config(w, h))
    if ref is None:
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return ref


@pytest.mark.parametrize("w,h,n", [(320, 240, 10), (640, 480, 6), (1024, 768, 3)])
def test_frames_random_scenes(S, oracle, w, h, n):
    ref = get_ref(S, w, h)
    cfg = S.default_config(w, h)
    base = S.default_scene(w, h, randomize_camera=1, **NOISY)
    for i in range(n):
        sc = S.randomize_scene(base, 99, i, 3, 8)
        if i % 3 == 2:
            sc.rotate180, sc.n_occluders = 1, 2
        xf = S.scene_transform(sc)
        xyz = S.deproject_host(sc, S.synth_depth_host(sc))
        o = H.oracle_process(oracle, cfg, xf, xyz)
        r = H.ref_process(ref, cfg, xf, xyz)
        assert not H.compare_results(r, o, tol=1e-12), i
        assert (o.info["status"] & ~QUIRKS) == (r.info["status"] & ~QUIRKS)
    assert oracle.ssd_oracle_sort_ties() == 0  # no rank tie that the reference's unstable sort could resolve differently


HIRES = dict(n_steps=12, riser=0.17, tread=0.26, cam_height=3.2, cam_pitch_deg=48.0, first_riser_y=0.5)


@pytest.mark.parametrize("w,h,ck,sk,n", [(2560, 1920, {}, {}, 2), (4096, 3072, dict(y_max=3.7, z_max=2.3), HIRES, 2)])
def test_frames_hires(S, oracle, w, h, ck, sk, n):
    """BASELINE.json configs[4] (4096x3072, extended range: 241 height bins, ~83 points per BestLine list) and 2560x1920
    (more than 12 'smallest distances' per line fit): the restatement against the compiled reference, as pinned as the
    small sizes."""
    cfg = S.default_config(w, h, **ck)
    ref = H.load_ref(cfg)
    if ref is None:
        pytest.skip("oracle/_ref not built and /root/reference absent")
    base = S.default_scene(w, h, **sk, **NOISY)
    for i in range(n):
        sc = S.randomize_scene(base, 4141, i, sk.get("n_steps", 3), sk.get("n_steps", 8))
        xf = S.scene_transform(base)
        xyz = S.deproject_host(sc, S.synth_depth_host(sc))
        o = H.oracle_process(oracle, cfg, xf, xyz)
        r = H.ref_process(ref, cfg, xf, xyz)
        assert not H.compare_results(r, o, tol=1e-12), i
        assert (o.info["status"] & ~QUIRKS) == (r.info["status"] & ~QUIRKS)
        assert o.info["n_plateaus"] >= 3
    assert oracle.ssd_oracle_sort_ties() == 0


@pytest.mark.parametrize("w,h,n", [(320, 240, 6), (640, 480, 4), (1024, 768, 2)])
def test_overlay_projection(S, oracle, w, h, n):
    """drawStairStep (pointcloud.cpp:583-597): the quadrilaterals the reference hands to drawQuadrilateral, recorded by
    the harness' stub, against the oracle's restatement -- bit-exact f32 pixels; every step is drawn twice (depth and
    infrared viewport) with the same corners; labelling arguments as detectStairs passes them."""
    ref = get_ref(S, w, h)
    cfg = S.default_config(w, h)
    base = S.default_scene(w, h, randomize_camera=1, **NOISY)
    try:
        for i in range(n):
            sc = S.randomize_scene(base, 777, i, 3, 8)
            if i % 3 == 2:
                sc.rotate180, sc.n_occluders = 1, 2
            xf = S.scene_transform(sc)
            intr = S.scene_intrinsics(sc) if i % 2 else None  # odd frames: the scene's intrinsics, even: the harness default
            ref.ssd_ref_set_intrinsics(C.byref(intr) if intr is not None else None)
            if intr is None:
                intr = A.Intrinsics(fx=w, fy=w, ppx=w * 0.5, ppy=h * 0.5)
            xyz = S.deproject_host(sc, S.synth_depth_host(sc))
            o = H.oracle_process(oracle, cfg, xf, xyz)
            a_inv = np.empty(9)
            assert oracle.ssd_oracle_inverse3(xf.a, a_inv.ctypes.data_as(C.POINTER(C.c_double))) == 0
            assert np.array_equal(a_inv, H.ref_a_inv(ref, xf))
            assert np.array_equal(a_inv, S.inverse3(xf.a))  # host library: same inverse
            ov = H.oracle_overlay(oracle, xf, a_inv, intr)
            r = H.ref_process(ref, cfg, xf, xyz)
            px, label, z = H.ref_overlay(ref)
            ns = r.info["n_steps"]
            assert ns > 0 and len(px) == 2 * ns and len(ov) == ns
            assert np.array_equal(px[:ns].view(np.uint32), ov.view(np.uint32)), i
            assert np.array_equal(px[ns:].view(np.uint32), ov.view(np.uint32)), i
            assert np.isfinite(ov).all()
            for k in range(ns):  # second pass labels: external-world corners and height (pointcloud.cpp:388-392)
                assert np.array_equal(label[ns + k], r.steps[k]["quad"][:2]) and z[ns + k] == r.steps[k]["height"]
    finally:
        ref.ssd_ref_set_intrinsics(None)


def test_derived_constants(S, oracle):
    for (w, h) in ((320, 240), (640, 480), (1024, 768)):
        ref = get_ref(S, w, h)
        cfg = S.default_config(w, h)
        d = H.Derived()
        assert oracle.ssd_oracle_derive(C.byref(cfg), C.byref(d)) == 0
        mh, my, nb = C.c_int(), C.c_int(), C.c_int()
        xr, hir = C.c_double(), C.c_double()
        ref.ssd_ref_derived(C.byref(mh), C.byref(my), C.byref(xr), C.byref(hir), C.byref(nb))
        assert (d.min_height, d.min_img_y_extent, d.n_bins) == (mh.value, my.value, nb.value)
        assert (d.xy_ratio, d.height_interval_reciprocal) == (xr.value, hir.value)
        assert d.min_height == 15 and d.n_bins == 121
        # Projection2D round trips through the reference
        rng = np.random.default_rng(0)
        px = rng.uniform(0, w, (1000, 2))
        out = np.empty_like(px)
        ref.ssd_ref_image_to_world(H.ptr(px), len(px), H.ptr(out))
        exp = np.stack([cfg.x_min + px[:, 0] * d.x_to_world, cfg.y_max - px[:, 1] * d.y_to_world], 1)
        assert np.array_equal(out, exp)


def blob(rng, w, h, kind):
    yy, xx = np.mgrid[0:h, 0:w]
    cx, cy = w * rng.uniform(0.4, 0.6), h * rng.uniform(0.3, 0.8)
    hw, hh = w * rng.uniform(0.15, 0.45), h * rng.uniform(0.08, 0.3)
    th = rng.uniform(-0.15, 0.15)
    u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
    v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
    img = ((np.abs(u) < hw) & (np.abs(v) < hh)).astype(np.uint8) * 255
    if kind >= 1:
        img[rng.random((h, w)) < 0.25] = 0
    if kind >= 2:
        img[rng.random((h, w)) < 0.002] = 255
    if kind == 3:
        img[:, : w // 2 - 30] = 0
    return img


def test_outline_front_edge_fuzz(S, oracle):
    w, h = 320, 240
    ref = get_ref(S, w, h)
    rng = np.random.default_rng(5)
    nvalid = 0
    for i in range(60):
        img = blob(rng, w, h, i % 4)
        q1, q2 = (C.c_double * 8)(), (C.c_double * 8)()
        v1, v2 = C.c_int(), C.c_int()
        assert oracle.ssd_oracle_detect_outline(H.ptr(img), w, h, 20, 4 / 3, q1, C.byref(v1)) == 0
        assert ref.ssd_ref_detect_outline(H.ptr(img), w, h, 20, 4 / 3, q2, C.byref(v2)) == 0, ref.ssd_ref_last_error()
        assert v1.value == v2.value and q1[:] == q2[:], i
        nvalid += v1.value
        l1, r1, l2, r2 = ((C.c_double * 2)() for _ in range(4))
        assert oracle.ssd_oracle_detect_front_edge(H.ptr(img), w, h, l1, r1, C.byref(v1)) == 0
        assert ref.ssd_ref_detect_front_edge(H.ptr(img), w, h, l2, r2, C.byref(v2)) == 0
        assert v1.value == v2.value and l1[:] == l2[:] and r1[:] == r2[:], i
    assert nvalid > 10


def test_points_in_quad_fuzz(S, oracle):
    ref = get_ref(S, 320, 240)
    rng = np.random.default_rng(9)
    quads = []
    for _ in range(200):
        c = rng.uniform(-0.3, 0.3, 2)
        wq, d = rng.uniform(0.2, 0.6), rng.uniform(0.1, 0.5)
        q = np.array([[-wq, -d], [wq, -d], [-wq, d], [wq, d]]) * 0.5
        th = rng.uniform(-0.8, 0.8)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        quads.append(q @ R.T + c + rng.normal(0, 0.03, (4, 2)))
    quads += [np.array([[0, 0], [1, 0], [0, 1], [1, 1.0]]), np.array([[0, 0], [1, 0], [1, 1], [0, 1.0]]),
              np.array([[0, 0], [1, 0], [0, 0], [1, 0.0]]), np.array([[0, 0], [0, 0], [0, 1], [0, 1.0]])]
    thrown = 0
    for q in quads:
        xy = rng.uniform(-0.8, 0.8, (4000, 2))
        xy[:8] = np.repeat(q, 2, axis=0)
        a, b = np.empty(len(xy), np.uint8), np.empty(len(xy), np.uint8)
        sa, sb = C.c_int(), C.c_int()
        qq = (C.c_double * 8)(*q.ravel())
        oracle.ssd_oracle_points_in_quad(qq, H.ptr(xy), len(xy), H.ptr(a), C.byref(sa))
        ref.ssd_ref_points_in_quad(qq, H.ptr(xy), len(xy), H.ptr(b), C.byref(sb))
        assert sa.value == sb.value
        assert np.array_equal(a, b)
        thrown += sa.value
    assert thrown >= 3


def test_close_matches_opencv(S, oracle):
    """the 3x3 MORPH_CLOSE restatements (oracle and the reference-build shim) against opencv-python"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    ref = H.load_ref(S.default_config(320, 240))
    for (w, h) in ((320, 240), (37, 19), (64, 5), (3, 3)):
        for dens in (0.02, 0.3, 0.7):
            img = (rng.random((h, w)) < dens).astype(np.uint8) * 255
            exp = cv2.morphologyEx(img, cv2.MORPH_CLOSE, None)
            a = img.copy()
            assert oracle.ssd_oracle_close(H.ptr(a), w, h) == 0
            assert np.array_equal(a, exp)
            if ref is not None:
                b = img.copy()
                assert ref.ssd_ref_close(H.ptr(b), w, h) == 0
                assert np.array_equal(b, exp)
