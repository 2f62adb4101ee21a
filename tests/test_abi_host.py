"""CPU: the C-ABI library loads, exports every symbol include/ssd_gpu.h declares, fails loudly without a GPU,
and its host-only entry points (configuration, transformation builder, serializer, scene source) behave."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from stair_step_detector_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header="ssd_gpu.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ssd_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(built_lib):
    names = declared_functions()
    assert len(names) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", A.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (ssd_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    unbound = [n for n in names if n not in A.PROTOTYPES]
    assert not unbound, unbound


def test_scene_library_exports_its_header_and_nothing_of_the_product(built_lib):
    """the synthetic input source lives in its own library (include/ssd_scene.h): a process that only generates input or
    runs the CPU reference (bench.py --impl reference) never maps libssd_gpu.so"""
    names = declared_functions("ssd_scene.h")
    assert len(names) == 7
    out = subprocess.run(["nm", "-D", "--defined-only", A.SCENE_LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (ssd_[a-z0-9_]+)", out))
    assert sorted(exported) == names
    assert not [n for n in names if n not in A.SCENE_PROTOTYPES]
    prod = subprocess.run(["nm", "-D", "--defined-only", A.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert not re.findall(r" T (ssd_scene_[a-z0-9_]+|ssd_synth[a-z0-9_]*|k_synth[a-z0-9_]*)", prod)


def test_struct_layouts_match_header():
    assert C.sizeof(A.Config) == 8 + 9 * 8 + 8
    assert C.sizeof(A.Transform) == 19 * 8
    assert C.sizeof(A.Step) == 72
    assert C.sizeof(A.FrameInfo) == 32
    assert C.sizeof(A.Plateau) == 32 + 64 + 8
    assert C.sizeof(A.Intrinsics) == 32
    assert C.sizeof(A.Overlay) == 32
    assert A.Scene.seed.offset % 8 == 0


def test_inverse3_and_transform_ex(S):
    """host side of the overlay projection: boost::qvm::inverse of a 3x3 (adjugate / determinant) and the camera
    transformation's _aInv as the triangle ctor holds it (_a = transposed(_aInv), transformation.cpp:128-132)"""
    rng = np.random.default_rng(3)
    for _ in range(20):
        a = rng.normal(size=(3, 3))
        inv = S.inverse3(a).reshape(3, 3)
        assert np.abs(inv @ a - np.eye(3)).max() < 1e-9
    with pytest.raises(S.SsdError):
        S.inverse3(np.zeros((3, 3)))
    sc = S.default_scene(640, 480, cam_yaw_deg=5.0, cam_roll_deg=-2.0)
    xf, a_inv = S.scene_transform_ex(sc)
    xf0 = S.scene_transform(sc)
    assert list(xf.a) == list(xf0.a) and list(xf.b) == list(xf0.b)
    assert np.array_equal(a_inv.reshape(3, 3), np.array(xf.a[:]).reshape(3, 3).T)
    assert np.abs(S.inverse3(xf.a) - a_inv).max() < 1e-12  # a rotation: inverse == transpose up to rounding


def test_no_gpu_means_loud_failure(S):
    if S.lib().ssd_gpu_device_count() > 0:
        pytest.skip("a GPU is present")
    cfg = S.default_config(320, 240)
    xf = S.scene_transform(S.default_scene(320, 240))
    with pytest.raises(S.SsdError, match="no CUDA device"):
        S.Detector(cfg, xf)


def test_config_defaults_are_the_reference_values(S):
    c = S.default_config(640, 480)  # configuration.h:27-52
    assert (c.x_min, c.x_max, c.y_min, c.y_max, c.z_min, c.z_max) == (-0.6, 0.6, 0.1, 1.3, -0.1, 1.1)
    assert (c.height_interval, c.min_height_above_ground, c.min_step_depth, c.min_peak_points) == (0.01, 0.05, 0.1, 2000)


def test_transform_is_a_rigid_map_onto_the_calibration_plane(S):
    sc = S.default_scene(640, 480, cam_roll_deg=4.0, cam_yaw_deg=-9.0, cam_pitch_deg=42.0)
    w = (C.c_double * 9)()
    c = (C.c_double * 9)()
    S.scene_lib().ssd_scene_calibration_points(C.byref(sc), w, c)
    xf = S.make_transform(np.array(w[:]).reshape(3, 3), np.array(c[:]).reshape(3, 3))
    a = np.array(xf.a[:]).reshape(3, 3)
    assert np.allclose(a @ a.T, np.eye(3), atol=1e-14) and abs(np.linalg.det(a) - 1) < 1e-14
    cam = np.array(c[:]).reshape(3, 3)
    world = cam @ a.T + np.array(xf.b[:])
    assert np.abs(world[:, 2]).max() < 1e-6          # the marks lie in the plane z = 0
    ext = world[:, :2] @ np.array(xf.ext_a[:]).reshape(2, 2).T + np.array(xf.ext_b[:])
    assert np.abs(ext - np.array(w[:]).reshape(3, 3)[:, :2]).max() < 1e-6  # external world = scene coordinates
    # camera below the plane / degenerate triangle are rejected, not silently accepted
    bad = A.Transform()
    assert S.lib().ssd_make_transform(w, (C.c_double * 9)(*([0.0] * 9)), C.byref(bad)) != 0


def test_serialize_wire_format(S):
    assert S.serialize([]) == '["stairs",["stairSteps",0]]'
    q = np.array([[-0.45, 0.2785], [0.4494, 0.2785], [-0.45, 0.45], [0.4494, 0.45]])
    line = S.serialize([(0.0040004, q), (float("nan"), np.zeros((4, 2)))])
    assert line == ('["stairs",["stairSteps",2],[[["height",0.004],["quadrilateral",[-0.450,0.279],[0.449,0.279],[-0.450,0.450],'
                    '[0.449,0.450]]],[["height",nan],["quadrilateral",[0.000,0.000],[0.000,0.000],[0.000,0.000],[0.000,0.000]]]]]'
                    ) or "-nan" in line
    import json
    j = json.loads(S.serialize([(0.25, q)]))  # what print-stairs.py / the ROS wrapper index into
    assert j[1][1] == 1 and j[2][0][0][1] == 0.25 and j[2][0][1][1] == [-0.45, 0.279]


def test_wire_format_through_the_consumers_index_paths(S):
    """The line protocol as its two consumers read it: print-stairs.py:54-72 (jdata[0], jdata[1][1], jdata[2][i],
    step[1][1..4]) and the ROS publisher stair_step_detector.py:33-57 (step[0][1], step[1][1..4] -> Point2 x, y)."""
    import json
    rng = np.random.default_rng(3)
    steps = [(float(h), rng.uniform(-0.6, 1.3, (4, 2))) for h in (0.004, 0.177, 0.35)]
    jdata = json.loads(S.serialize(steps))
    assert jdata[0] == "stairs" and jdata[1][0] == "stairSteps" and jdata[1][1] == len(steps) == len(jdata[2])
    for (h, q), step in zip(steps, jdata[2]):
        assert step[0][0] == "height" and step[1][0] == "quadrilateral"
        msg = dict(height=step[0][1], quadrilateral=[dict(x=step[1][k][0], y=step[1][k][1]) for k in (1, 2, 3, 4)])
        assert abs(msg["height"] - h) <= 0.0005 + 1e-12  # three decimals (stairs.cpp:58)
        for k in range(4):
            assert abs(msg["quadrilateral"][k]["x"] - q[k, 0]) <= 0.0005 + 1e-12
            assert abs(msg["quadrilateral"][k]["y"] - q[k, 1]) <= 0.0005 + 1e-12
    empty = json.loads(S.serialize([]))
    assert empty[1][1] == 0 and len(empty) == 2  # print-stairs.py:60 / stair_step_detector.py:38 never index jdata[2] then


def test_scene_source_geometry(S):
    sc = S.default_scene(320, 240)
    d = S.synth_depth_host(sc)
    assert d.shape == (240, 320) and d.min() > 0
    xyz = S.deproject_host(sc, d)
    xf = S.scene_transform(sc)
    a, b = np.array(xf.a[:]).reshape(3, 3), np.array(xf.b[:])
    world = xyz.reshape(-1, 3).astype(np.float64) @ a.T + b
    z = world[:, 2]
    # ground + three treads at multiples of the riser, to depth-quantisation accuracy
    for k in range(4):
        assert (np.abs(z - (0.004 + k * 0.173)) < 1e-3).sum() > 1500
    noisy = S.default_scene(320, 240, noise_sigma=0.0025, dropout=0.03, n_holes=3)
    dn = S.synth_depth_host(noisy)
    assert 0.02 < (dn == 0).mean() < 0.08
    r1 = S.randomize_scene(noisy, 1, 7, 3, 8)
    r2 = S.randomize_scene(noisy, 1, 7, 3, 8)
    assert bytes(r1) == bytes(r2) and 3 <= r1.n_steps <= 8
    assert r1.cam_pitch_deg == noisy.cam_pitch_deg  # camera fixed unless randomize_camera is set
