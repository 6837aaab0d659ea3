"""Pair types that no reference golden reaches (round 2): sphere-box (dedicated routine), cylinder / capsule / box against
a box or a cylinder through the general convex routine (mjc_Convex = MPR; capsule-box and box-box are stand-ins for
MuJoCo's dedicated multi-contact routines).

A synthetic single-tree model (the engines support one kinematic tree): a heavy box `base` on the floor (slide z), a
massless carriage on it (slide x) and a `probe` body (slide z) whose geom is dropped onto the base - probe and base are
grandchild / grandparent, i.e. a collidable pair. (No rotational dof on purpose: a capsule standing on its end or MPR's
choice of ONE point of a face-face contact would make the comparison chaotic; depth, normal and friction are what is compared.)
Every shape is run upright (geometric anchor) and tilted (euler 10 20 0).  Checks:
  * the fp64 oracle settles where the geometry says (the probe rests ON the box: gap = static penetration of a soft contact),
  * the fp32 engine core (serial emulation build, CPU tier; CUDA engine, GPU tier) follows the oracle.
The oracle is unpinned for these routines (no MuJoCo here); the geometric check is what anchors it."""
import ctypes
import os
import tempfile

import numpy as np
import pytest

from helpers import ROOT

XML = """
<mujoco model="drop">
  <option timestep="0.002" integrator="Euler" cone="pyramidal"/>
  <worldbody>
    <geom name="floor" type="plane" size="0 0 1" condim="3"/>
    <body name="base" pos="0 0 0.05">
      <joint name="bz" type="slide" axis="0 0 1"/>
      <geom name="base_geom" type="{btype}" size="{bsize}" mass="20" condim="3"/>
      <body name="carriage" pos="0 0 0">
        <joint name="cx" type="slide" axis="1 0 0" damping="5"/>
        <geom name="c_mass" type="sphere" size="0.01" mass="0.1" contype="0" conaffinity="0"/>
        <body name="probe" pos="0.05 0.02 {z0}">
          <joint name="pz" type="slide" axis="0 0 1"/>
          <geom name="probe_geom" type="{gtype}" size="{gsize}" mass="1" condim="3" {extra}/>
        </body>
      </body>
    </body>
  </worldbody>
  <actuator>
    <motor name="m_cx" joint="cx" gear="1"/>
  </actuator>
</mujoco>
"""
# geom type -> (size attribute, height of the geom's lowest point below its centre at zero rotation)
SHAPES = {"sphere": ("0.06", 0.06), "capsule": ("0.04 0.08", 0.12), "cylinder": ("0.05 0.07", 0.07), "box": ("0.06 0.05 0.04", 0.04)}


BASES = {"box": "0.4 0.3 0.05", "cylinder": "0.4 0.05"}       # both: top face at z = +0.05 of the base body


def _compile(gtype, tilt=0, btype="box"):
    from loco_mujoco_b200 import mjcf, modelpack
    size, low = SHAPES[gtype]
    xml = XML.format(btype=btype, bsize=BASES[btype], gtype=gtype, gsize=size, z0=0.05 + low + (0.06 if tilt else 0.03), extra='euler="10 20 0"' if tilt else "")
    with tempfile.NamedTemporaryFile("w", suffix=".xml", delete=False) as f:
        f.write(xml)
        path = f.name
    try:
        m = mjcf.compile_model(mjcf.XmlHandle(path), fuse_static=False)
    finally:
        os.unlink(path)
    return m, modelpack.pack(m), low


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _oracle_rollout(oracle, blobs, nv, q0, ctrl, n, nsub=5):
    lib = oracle.lib
    lib.ref_create.restype = ctypes.c_void_p
    mi, mr = [np.ascontiguousarray(x) for x in blobs]
    sim = ctypes.c_void_p(lib.ref_create(_p(mi), len(mi), _p(mr), len(mr)))
    assert sim.value
    v0 = np.zeros(nv)
    lib.ref_reset(sim, _p(q0), _p(v0))
    out = []
    q, v = np.zeros(nv), np.zeros(nv)
    lib.ref_ncon.argtypes = [ctypes.c_void_p]
    ncon = 0
    for k in range(n):
        c = np.ascontiguousarray(ctrl[k], dtype=np.float64)
        lib.ref_step(sim, _p(c), nsub)
        lib.ref_get_state(sim, _p(q), _p(v))
        ncon = max(ncon, lib.ref_ncon(sim))
        out.append(np.concatenate([q, v]))
    lib.ref_destroy(sim)
    return np.array(out), ncon


@pytest.mark.parametrize("btype", list(BASES))
@pytest.mark.parametrize("tilt", [0, 1])
@pytest.mark.parametrize("gtype", list(SHAPES))
def test_probe_settles_on_the_box_and_fp32_core_follows(oracle, emu, gtype, tilt, btype):
    m, blobs, low = _compile(gtype, tilt, btype)
    assert m.n_dropped_pairs == 0 and m.npair >= 2
    nv, n = m.nv, 120
    q0 = np.zeros(nv)
    ctrl = np.zeros((n, m.nu))
    ctrl[40:, 0] = 1.5                       # then push the carriage: the probe slides over the box (friction, rolling)
    ref, ncon = _oracle_rollout(oracle, blobs, nv, q0, ctrl, n)
    assert ncon >= 2                         # floor-box and probe-box contacts
    # geometric anchor at rest (step 39, before the push): probe bottom sits on the box top within the soft-contact penetration
    pz = ref[39, 2] - ref[39, 0] * 0         # qpos: bz, cx, pz (pz is relative to the base)
    gap = pz + 0.03                          # initial clearance was 0.03 -> gap = penetration (negative) at rest
    if not tilt:
        assert -3e-3 < gap < 1e-4, gap
    assert np.abs(ref[39, nv:]).max() < 2e-2  # at rest
    # fp32 core (serial emulation build of the CUDA source)
    sim = emu.emu_create(_p(blobs[0]), len(blobs[0]), _p(blobs[1]), len(blobs[1]))
    assert sim
    v0 = np.zeros(nv)
    emu.emu_reset(sim, _p(q0), _p(v0))
    q, v = np.zeros(nv), np.zeros(nv)
    worst = 0.0
    for k in range(n):
        c = np.ascontiguousarray(ctrl[k], dtype=np.float64)
        emu.emu_step(sim, _p(c), 5)
        emu.emu_get_state(sim, _p(q), _p(v))
        worst = max(worst, float(np.abs(q - ref[k, :nv]).max()))
    emu.emu_destroy(sim)
    assert worst < 3e-3, worst


@pytest.mark.gpu
@pytest.mark.parametrize("btype", list(BASES))
@pytest.mark.parametrize("tilt", [0, 1])
@pytest.mark.parametrize("gtype", list(SHAPES))
def test_cuda_engine_follows_oracle_on_the_synthetic_pairs(oracle, gtype, tilt, btype):
    import torch
    from loco_mujoco_b200.engine import CudaEngine
    from loco_mujoco_b200.task import TaskSpec, OBS_QPOS, OBS_QVEL, REWARD_NONE
    m, blobs, low = _compile(gtype, tilt, btype)
    nv, n, N = m.nv, 100, 32
    table = np.zeros((1, 1, 2 * nv))
    spec = TaskSpec([OBS_QPOS] * nv + [OBS_QVEL] * nv, list(range(nv)) * 2, [], REWARD_NONE, [], [], np.zeros(m.nu), np.ones(m.nu),
                    5, table, 0, [nv, nv], use_absorbing=False)          # (recenter indices out of range: nothing recentred)
    eng = CudaEngine(blobs, spec.pack(), N, device=0, seed=0, warps_per_block=8)
    eng.reset()
    rng = np.random.RandomState(3)
    ctrl = np.zeros((n, N, m.nu), dtype=np.float32)
    ctrl[30:] = rng.uniform(-2, 2, (1, N, m.nu))           # every env pushes its carriage differently
    refs = [_oracle_rollout(oracle, blobs, nv, np.zeros(nv), ctrl[:, i].astype(np.float64), n)[0] for i in range(0, N, 4)]
    worst = 0.0
    for k in range(n):
        obs, _, _, _ = eng.step(torch.tensor(ctrl[k], device=eng.device), auto_reset=False)
        o = obs.cpu().numpy()
        for j, i in enumerate(range(0, N, 4)):
            worst = max(worst, float(np.abs(o[i, :nv] - refs[j][k, :nv]).max()))
    assert worst < 3e-3, worst


# ----------------------------------------------------------------------------------------------------------------------
# the restated pair routines against brute-force geometry (oracle entry points ref_debug_*; no simulation involved)
# ----------------------------------------------------------------------------------------------------------------------
def _rand_rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _box_edges(p, R, s):
    import itertools
    V = {sg: p + R @ (np.array(sg) * s) for sg in itertools.product([-1, 1], repeat=3)}
    E = []
    for sg in V:
        for ax in range(3):
            if sg[ax] == -1:
                o = list(sg)
                o[ax] = 1
                E.append((V[sg], V[tuple(o)]))
    return list(V.values()), E


def _seg_seg(p1, q1, p2, q2):
    d1, d2, r = q1 - p1, q2 - p2, p1 - p2
    a, e, f, c, b = d1 @ d1, d2 @ d2, d2 @ r, d1 @ r, d1 @ d2
    den = a * e - b * b
    s = float(np.clip((b * f - c * e) / den, 0, 1)) if den > 1e-15 else 0.0
    t = (b * s + f) / e
    if t < 0:
        t, s = 0.0, float(np.clip(-c / a, 0, 1))
    elif t > 1:
        t, s = 1.0, float(np.clip((b - c) / a, 0, 1))
    c1, c2 = p1 + d1 * s, p2 + d2 * t
    return np.linalg.norm(c1 - c2), c1, c2


def _pt_box(x, p, R, s):
    loc = R.T @ (x - p)
    c = np.clip(loc, -s, s)
    return np.linalg.norm(loc - c), p + R @ c


def test_box_box_edge_branch_is_the_exact_closest_feature(oracle):
    """Whenever `box_box_edge` claims a pair (separated boxes whose closest features are two edges), its dist / position /
    normal are those of the brute-force closest points over all 144 edge pairs and 16 vertex-box pairs."""
    lib = oracle.lib
    lib.ref_debug_box_box_edge.restype = ctypes.c_int
    lib.ref_debug_box_box_edge.argtypes = [ctypes.c_double] + [ctypes.c_void_p] * 7
    rng = np.random.default_rng(0)
    claimed = 0
    for trial in range(4000):
        s1, s2 = rng.uniform(0.02, 0.2, 3), rng.uniform(0.02, 0.2, 3)
        R1, R2 = _rand_rot(rng), _rand_rot(rng)
        p1 = np.zeros(3)
        p2 = rng.normal(size=3)
        p2 *= (np.linalg.norm(s1) + np.linalg.norm(s2)) * rng.uniform(0.5, 1.1) / np.linalg.norm(p2)
        out = np.zeros(7)
        n = lib.ref_debug_box_box_edge(0.02, _p(p1), _p(np.ascontiguousarray(R1)), _p(s1), _p(p2), _p(np.ascontiguousarray(R2)), _p(s2), _p(out))
        VA, EA = _box_edges(p1, R1, s1)
        VB, EB = _box_edges(p2, R2, s2)
        best = (1e9, None, None, "")
        for a0, a1 in EA:
            for b0, b1 in EB:
                d, c1, c2 = _seg_seg(a0, a1, b0, b1)
                if d < best[0] - 1e-12:
                    best = (d, c1, c2, "edge")
        for v in VA:
            d, c = _pt_box(v, p2, R2, s2)
            if d < best[0] - 1e-9:
                best = (d, v, c, "vertex")
        for v in VB:
            d, c = _pt_box(v, p1, R1, s1)
            if d < best[0] - 1e-9:
                best = (d, c, v, "vertex")
        if n == 1 and out[0] > 1e-6:                      # claimed, separated: must be THE closest feature pair
            claimed += 1
            assert best[3] == "edge"
            assert abs(out[0] - best[0]) < 1e-9, (trial, out[0], best[0])
            assert np.abs(out[1:4] - 0.5 * (best[1] + best[2])).max() < 1e-8
            assert np.abs(out[4:7] - (best[2] - best[1]) / best[0]).max() < 1e-6
        if n == 0:                                        # "farther apart than the margin"
            assert best[0] > 0.02 - 1e-9
    assert claimed > 40


def test_sphere_box_matches_point_to_box_geometry(oracle):
    lib = oracle.lib
    lib.ref_debug_sphere_box.restype = ctypes.c_int
    lib.ref_debug_sphere_box.argtypes = [ctypes.c_double, ctypes.c_void_p, ctypes.c_double] + [ctypes.c_void_p] * 4
    rng = np.random.default_rng(1)
    hits = 0
    for trial in range(2000):
        s2, R2, p2, r = rng.uniform(0.02, 0.2, 3), _rand_rot(rng), rng.normal(size=3) * 0.05, rng.uniform(0.01, 0.1)
        p1 = p2 + R2 @ (rng.uniform(-1.6, 1.6, 3) * s2)
        out = np.zeros(7)
        n = lib.ref_debug_sphere_box(0.01, _p(p1), r, _p(p2), _p(np.ascontiguousarray(R2)), _p(s2), _p(out))
        d, c = _pt_box(p1, p2, R2, s2)
        if d < 1e-9:
            continue                                      # centre inside the box: convention case
        assert (n == 1) == (d - r <= 0.01), (trial, d - r)
        if n == 1:
            hits += 1
            nrm = (c - p1) / d                            # from the sphere (geom 1) to the box (geom 2)
            assert abs(out[0] - (d - r)) < 1e-12 and np.abs(out[4:7] - nrm).max() < 1e-9
            assert np.abs(out[1:4] - (c - nrm * 0.5 * (d - r))).max() < 1e-9      # midway between the two surfaces
    assert hits > 300
