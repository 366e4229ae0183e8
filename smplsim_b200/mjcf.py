"""MJCF -> ModelDesc extractor (host side, stdlib xml only).

Replaces, for the hot path, what the reference obtains from
``mujoco.MjModel.from_xml_string`` (smpl_sim/envs/base_env.py:139-142) plus the
property scraping in ``HumanoidEnv.setup_humanoid_properties``
(smpl_sim/envs/humanoid_env.py:262-308): the constant kinematic tree, the
per-body inertials derived from geom densities, the collision geoms and the
sensor/actuator ordering.  The model class handled is exactly the one the
reference ships (smpl_sim/data/assets/mjcf/smpl_humanoid.xml, /smpl_humanoid.xml):

* one floor plane on the world body,
* one tree rooted at a free-joint body,
* every other body carries up to three hinge joints anchored at its origin,
* box / capsule / sphere geoms with ``density`` (``inertiafromgeom`` = auto).

Compile-time semantics restated from MuJoCo's documented behaviour
(SURVEY.md Appendix A.1): angles in degrees, local coordinates, capsule
``fromto`` frames, density -> mass/inertia formulas.
"""
from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE, GEOM_BOX = 0, 2, 3, 6
_GEOM_TYPES = {"plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE, "box": GEOM_BOX}


def _floats(s: Optional[str], n: Optional[int] = None, default=None) -> np.ndarray:
    if s is None:
        if default is None:
            raise ValueError("missing attribute")
        return np.asarray(default, dtype=np.float64)
    v = np.asarray([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.size != n:
        raise ValueError(f"expected {n} numbers, got {v.size}: {s!r}")
    return v


def quat_to_mat(q: np.ndarray) -> np.ndarray:
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def quat_z2vec(vec: np.ndarray) -> np.ndarray:
    """Quaternion (wxyz) rotating +z onto ``vec`` (shortest arc)."""
    v = vec / np.linalg.norm(vec)
    axis = np.cross([0.0, 0.0, 1.0], v)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        # parallel or anti-parallel
        return np.array([1.0, 0, 0, 0]) if v[2] > 0 else np.array([0.0, 1.0, 0, 0])
    axis /= s
    ang = math.atan2(s, v[2])
    return np.concatenate([[math.cos(ang / 2)], axis * math.sin(ang / 2)])


@dataclass
class Geom:
    name: str
    type: int
    body: int            # robot body index (-1 = world)
    pos: np.ndarray      # centre in body frame
    mat: np.ndarray      # 3x3 geom frame in body frame
    size: np.ndarray     # box: half extents; capsule: (r, half_len, 0); sphere: (r,0,0)
    density: float
    margin: float
    friction: np.ndarray
    contype: int
    conaffinity: int
    condim: int


@dataclass
class Body:
    name: str
    parent: int
    pos: np.ndarray
    quat: np.ndarray
    joint_names: List[str] = field(default_factory=list)
    joint_axes: List[np.ndarray] = field(default_factory=list)
    joint_range: List[np.ndarray] = field(default_factory=list)     # radians
    joint_limited: List[bool] = field(default_factory=list)
    joint_armature: List[float] = field(default_factory=list)
    free: bool = False
    geoms: List[int] = field(default_factory=list)


@dataclass
class ParsedMJCF:
    bodies: List[Body]
    geoms: List[Geom]                # robot geoms only, MuJoCo id = index + 1
    floor: Geom
    actuator_joints: List[str]
    excludes: List[tuple]
    sensors: List[tuple]             # (type, body name)
    solref: np.ndarray
    solimp: np.ndarray


def _geom_inertia(g: Geom):
    """mass, inertia (3x3, about geom centre, in geom frame).  SURVEY.md A.1."""
    if g.type == GEOM_BOX:
        sx, sy, sz = g.size
        m = g.density * 8.0 * sx * sy * sz
        I = np.diag([m / 3 * (sy * sy + sz * sz), m / 3 * (sx * sx + sz * sz), m / 3 * (sx * sx + sy * sy)])
    elif g.type == GEOM_CAPSULE:
        r, hl = g.size[0], g.size[1]
        H = 2.0 * hl
        mc = g.density * math.pi * r * r * H
        ms = g.density * 4.0 / 3.0 * math.pi * r ** 3
        m = mc + ms
        izz = mc * r * r / 2 + 0.4 * ms * r * r
        ixx = mc * (3 * r * r + H * H) / 12 + ms * (0.4 * r * r + 0.375 * r * H + 0.25 * H * H)
        I = np.diag([ixx, ixx, izz])
    elif g.type == GEOM_SPHERE:
        r = g.size[0]
        m = g.density * 4.0 / 3.0 * math.pi * r ** 3
        I = np.eye(3) * 0.4 * m * r * r
    else:
        raise NotImplementedError(f"geom type {g.type}")
    return m, I


def parse_mjcf(xml_text: str) -> ParsedMJCF:
    root = ET.fromstring(xml_text)
    comp = root.find("compiler")
    angle_deg = True
    if comp is not None:
        if comp.get("angle", "degree") == "radian":
            angle_deg = False
        if comp.get("coordinate", "local") != "local":
            raise NotImplementedError("global coordinates")
    if root.find("option") is not None and len(root.find("option").attrib):
        # the reference models have none; engine options are fixed in ModelDesc
        raise NotImplementedError("<option> overrides are not supported")

    jdef: Dict[str, str] = {}
    gdef: Dict[str, str] = {}
    d = root.find("default")
    if d is not None:
        if d.find("default") is not None:
            raise NotImplementedError("nested default classes")
        if d.find("joint") is not None:
            jdef = dict(d.find("joint").attrib)
        if d.find("geom") is not None:
            gdef = dict(d.find("geom").attrib)

    def gattr(e, k, default=None):
        return e.get(k, gdef.get(k, default))

    def jattr(e, k, default=None):
        return e.get(k, jdef.get(k, default))

    bodies: List[Body] = []
    geoms: List[Geom] = []
    floor: Optional[Geom] = None

    def make_geom(e, body_idx) -> Geom:
        tname = e.get("type", gdef.get("type", "sphere"))
        if tname not in _GEOM_TYPES:
            raise NotImplementedError(f"geom type {tname!r}")
        t = _GEOM_TYPES[tname]
        size = _floats(e.get("size"), default=[0, 0, 0])
        pos = _floats(e.get("pos"), 3, default=[0, 0, 0])
        quat = _floats(e.get("quat"), 4, default=[1, 0, 0, 0])
        if t == GEOM_CAPSULE:
            if e.get("fromto") is not None:
                ft = _floats(e.get("fromto"), 6)
                a, b = ft[:3], ft[3:]
                pos = 0.5 * (a + b)
                quat = quat_z2vec(b - a)
                size3 = np.array([size[0], 0.5 * np.linalg.norm(b - a), 0.0])
            else:
                size3 = np.array([size[0], size[1], 0.0])
        elif t == GEOM_BOX:
            size3 = size[:3].copy()
        elif t == GEOM_SPHERE:
            size3 = np.array([size[0], 0.0, 0.0])
        else:
            size3 = np.resize(size, 3).astype(np.float64)
        fr = _floats(gattr(e, "friction"), default=[1, 0.005, 0.0001])
        fr = np.concatenate([fr, [1, 0.005, 0.0001][len(fr):]])[:3]
        return Geom(
            name=e.get("name", f"geom{len(geoms)}"), type=t, body=body_idx, pos=pos, mat=quat_to_mat(quat),
            size=size3, density=float(gattr(e, "density", "1000")), margin=float(gattr(e, "margin", "0")),
            friction=fr, contype=int(gattr(e, "contype", "1")), conaffinity=int(gattr(e, "conaffinity", "1")),
            condim=int(gattr(e, "condim", "3")))

    def walk(e, parent_idx):
        nonlocal floor
        idx = len(bodies)
        b = Body(name=e.get("name", f"body{idx}"), parent=parent_idx,
                 pos=_floats(e.get("pos"), 3, default=[0, 0, 0]), quat=_floats(e.get("quat"), 4, default=[1, 0, 0, 0]))
        bodies.append(b)
        for c in e:
            if c.tag == "freejoint" or (c.tag == "joint" and c.get("type") == "free"):
                b.free = True
                b.joint_names.append(c.get("name", b.name))
            elif c.tag == "joint":
                if c.get("type", jdef.get("type", "hinge")) != "hinge":
                    raise NotImplementedError("only hinge / free joints")
                jp = _floats(c.get("pos"), 3, default=[0, 0, 0])
                if np.abs(jp).max() > 0:
                    raise NotImplementedError("hinge anchors must sit at the body origin")
                ax = _floats(jattr(c, "axis"), 3, default=[0, 0, 1])
                rng = _floats(jattr(c, "range"), 2, default=[0, 0])
                if angle_deg:
                    rng = np.deg2rad(rng)
                lim = jattr(c, "limited", "auto")
                limited = (lim == "true") or (lim == "auto" and rng[0] < rng[1])
                if abs(float(jattr(c, "damping", "0"))) > 0 or abs(float(jattr(c, "stiffness", "0"))) > 0:
                    raise NotImplementedError("passive joint damping/stiffness")
                b.joint_names.append(c.get("name"))
                b.joint_axes.append(ax / np.linalg.norm(ax))
                b.joint_range.append(rng)
                b.joint_limited.append(bool(limited))
                b.joint_armature.append(float(jattr(c, "armature", "0")))
            elif c.tag == "geom":
                g = make_geom(c, idx)
                b.geoms.append(len(geoms))
                geoms.append(g)
            elif c.tag == "body":
                walk(c, idx)
            elif c.tag in ("inertial",):
                raise NotImplementedError("explicit <inertial>")

    wb = root.find("worldbody")
    top = [c for c in wb if c.tag == "body"]
    if len(top) != 1:
        raise NotImplementedError("exactly one kinematic tree expected")
    for c in wb:
        if c.tag == "geom":
            g = make_geom(c, -1)
            if g.type != GEOM_PLANE or floor is not None:
                raise NotImplementedError("world geoms: exactly one plane supported")
            floor = g
    if floor is None:
        raise ValueError("no floor plane")
    walk(top[0], -1)
    if not bodies[0].free:
        raise NotImplementedError("root body must carry the free joint")
    for b in bodies[1:]:
        if b.free:
            raise NotImplementedError("free joint below the root")
        if len(b.joint_axes) > 3:
            raise NotImplementedError("more than three hinges per body")

    acts = []
    a = root.find("actuator")
    if a is not None:
        for m in a:
            if m.tag != "motor" or float(m.get("gear", "1").split()[0]) != 1.0:
                raise NotImplementedError("only gear-1 motors")
            if m.get("ctrlrange") or m.get("forcerange"):
                raise NotImplementedError("actuator ranges")
            acts.append(m.get("joint"))
    excl = []
    c = root.find("contact")
    if c is not None:
        for e in c:
            if e.tag == "exclude":
                excl.append((e.get("body1"), e.get("body2")))
    sens = []
    s = root.find("sensor")
    if s is not None:
        for e in s:
            sens.append((e.tag, e.get("objname")))
    solimp = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
    if gdef.get("solimp") is not None:
        v = _floats(gdef.get("solimp"))
        solimp[:len(v)] = v
    return ParsedMJCF(bodies=bodies, geoms=geoms, floor=floor, actuator_joints=acts, excludes=excl, sensors=sens,
                      solref=_floats(gdef.get("solref"), 2, default=[0.02, 1.0]), solimp=solimp)


def body_inertials(p: ParsedMJCF):
    """Per-body mass, COM (body frame) and 3x3 inertia about the COM (body frame)."""
    nb = len(p.bodies)
    mass = np.zeros(nb)
    ipos = np.zeros((nb, 3))
    inertia = np.zeros((nb, 3, 3))
    for bi, b in enumerate(p.bodies):
        ms, cs, Is = [], [], []
        for gi in b.geoms:
            g = p.geoms[gi]
            m, I = _geom_inertia(g)
            ms.append(m)
            cs.append(g.pos)
            Is.append(g.mat @ I @ g.mat.T)
        if not ms:
            raise NotImplementedError(f"body {b.name} has no geom (massless bodies unsupported)")
        M = float(np.sum(ms))
        com = np.sum([m * c for m, c in zip(ms, cs)], axis=0) / M
        Ib = np.zeros((3, 3))
        for m, c, I in zip(ms, cs, Is):
            r = c - com
            Ib += I + m * (np.dot(r, r) * np.eye(3) - np.outer(r, r))
        mass[bi], ipos[bi], inertia[bi] = M, com, Ib
    return mass, ipos, inertia
