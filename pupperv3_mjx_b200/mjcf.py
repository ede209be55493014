"""MJCF -> flat constant tables for the Pupper robot family (host side, NumPy, float64).

Replaces, for the hot path only, what the reference gets from ``brax.io.mjcf.load`` ->
``mujoco.MjModel.from_xml_path`` -> ``mjx.put_model`` (reference ``environment.py:165``) plus the
id lookups at ``environment.py:17-29,183-203``.  MuJoCo's compiler is not available in this
image, so the subset of its behaviour the model file needs is restated here:

* default classes, body/geom/site/joint ordering (depth-first bodies; geoms grouped by body);
* quaternion normalisation, ``<inertial>`` handling, ``autolimits``;
* the ``mj_setConst`` constants at ``qpos0`` (``dof_invweight0``, ``body_invweight0``,
  ``stat.meaninertia``) -- SURVEY.md Appendix A.12;
* contact parameter mixing (friction = max, solimp/solref = solmix-weighted mean) -- A.5.

Only the topology world -> base (free joint) -> 4 legs x 3 hinge links with sphere colliders is
supported; anything else raises ``UnsupportedModelError``.
"""

from __future__ import annotations

import dataclasses
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional

import numpy as np

MJ_MINVAL = 1e-15

_GEOM_DEFAULTS = dict(
    type="sphere", contype="1", conaffinity="1", condim="3", friction="1 0.005 0.0001",
    solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2", solmix="1", margin="0", gap="0",
    priority="0", pos="0 0 0", quat="1 0 0 0", size="0 0 0",
)
_JOINT_DEFAULTS = dict(
    type="hinge", armature="0", damping="0", frictionloss="0", pos="0 0 0", axis="0 0 1",
    solreflimit="0.02 1", solimplimit="0.9 0.95 0.001 0.5 2",
    solreffriction="0.02 1", solimpfriction="0.9 0.95 0.001 0.5 2", stiffness="0", margin="0",
)
_GENERAL_DEFAULTS = dict(
    gainprm="1 0 0", biasprm="0 0 0", biastype="none", gaintype="fixed", dyntype="none",
    forcelimited="auto", ctrllimited="auto", gear="1 0 0 0 0 0",
)


class UnsupportedModelError(ValueError):
    pass


def _floats(s: str, n: Optional[int] = None, fill: Optional[List[float]] = None) -> np.ndarray:
    v = [float(x) for x in s.split()]
    if fill is not None and len(v) < len(fill):
        v = v + list(fill[len(v):])
    a = np.array(v, dtype=np.float64)
    if n is not None and a.shape[0] != n:
        raise UnsupportedModelError(f"expected {n} numbers, got {s!r}")
    return a


def _qnorm(q: np.ndarray) -> np.ndarray:
    return q / np.linalg.norm(q)


def quat_mul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
    ])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


@dataclasses.dataclass
class _Geom:
    name: str
    body: int
    type: str
    pos: np.ndarray
    quat: np.ndarray
    size: np.ndarray
    contype: int
    conaffinity: int
    condim: int
    friction: np.ndarray
    solref: np.ndarray
    solimp: np.ndarray
    solmix: float
    priority: int


@dataclasses.dataclass
class CompiledModel:
    """Flat tables in MuJoCo id order. All float64 here; cast to f32 at the ABI."""

    nbody: int
    body_names: List[str]
    body_parent: np.ndarray
    body_pos: np.ndarray
    body_quat: np.ndarray
    body_ipos: np.ndarray
    body_iquat: np.ndarray
    body_mass: np.ndarray
    body_inertia: np.ndarray
    body_invweight0: np.ndarray  # [nbody, 2]
    nq: int
    nv: int
    nu: int
    dof_bodyid: np.ndarray
    dof_parentid: np.ndarray
    dof_armature: np.ndarray
    dof_damping: np.ndarray
    dof_frictionloss: np.ndarray
    dof_invweight0: np.ndarray
    dof_solref: np.ndarray
    dof_solimp: np.ndarray
    jnt_names: List[str]
    jnt_range: np.ndarray  # [12, 2] hinges only
    jnt_solref: np.ndarray
    jnt_solimp: np.ndarray
    actuator_names: List[str]
    actuator_gainprm: np.ndarray  # [nu, 10]
    actuator_biasprm: np.ndarray  # [nu, 10]
    actuator_forcerange: np.ndarray  # [nu, 2]
    ngeom: int
    geom_names: List[str]
    geom_bodyid: np.ndarray
    geom_type: List[str]
    geom_friction: np.ndarray  # [ngeom, 3]
    geom_collides: np.ndarray  # bool
    body_geomadr: np.ndarray
    body_geomnum: np.ndarray
    floor_geomid: int
    sphere_geomid: np.ndarray
    sphere_body: np.ndarray
    sphere_pos: np.ndarray
    sphere_radius: np.ndarray
    box_geomid: np.ndarray
    box_pos: np.ndarray
    box_mat: np.ndarray
    box_size: np.ndarray
    plane_sphere_solref: np.ndarray
    plane_sphere_solimp: np.ndarray
    sphere_box_solref: np.ndarray
    sphere_box_solimp: np.ndarray
    sphere_sphere_solref: np.ndarray
    sphere_sphere_solimp: np.ndarray
    site_names: List[str]
    site_body: np.ndarray
    site_pos: np.ndarray
    timestep: float
    gravity: np.ndarray
    impratio: float
    tolerance: float
    ls_tolerance: float
    iterations: int
    ls_iterations: int
    meaninertia: float
    max_geom_pairs: int
    max_contact_points: int
    frictionloss_rows: bool
    keyframes: Dict[str, np.ndarray]
    qpos0: np.ndarray
    M0: np.ndarray
    geoms: List["_Geom"] = dataclasses.field(default_factory=list)

    # -- name lookups the reference does through mujoco (environment.py:17-29) ------------------
    def body_id(self, name: str) -> int:
        if name not in self.body_names:
            raise AssertionError("Body not found.")
        return self.body_names.index(name)

    def site_id(self, name: str) -> int:
        if name not in self.site_names:
            raise AssertionError("Site not found.")
        return self.site_names.index(name)

    def body_geom_ids(self, name: str) -> np.ndarray:
        b = self.body_id(name)
        return self.body_geomadr[b] + np.arange(self.body_geomnum[b])


def _resolve_defaults(root: ET.Element):
    """Returns {class_name: {tag: attrib-dict}}; '' is the top-level class."""
    classes: Dict[str, Dict[str, Dict[str, str]]] = {"": {"geom": {}, "joint": {}, "general": {}, "site": {}}}

    def walk(node: ET.Element, parent_cls: str):
        name = node.get("class", "") if node is not top else ""
        if node is not top:
            classes[name] = {k: dict(v) for k, v in classes[parent_cls].items()}
        for child in node:
            if child.tag == "default":
                continue
            tag = "general" if child.tag in ("general", "motor", "position") else child.tag
            classes[name].setdefault(tag, {}).update(child.attrib)
        for child in node:
            if child.tag == "default":
                walk(child, name)

    tops = root.findall("default")
    for top in tops:
        walk(top, "")
    return classes


def compile_model(xml: "str | ET.ElementTree | ET.Element", frictionloss_rows: bool = True) -> CompiledModel:
    """Compile an MJCF file path / XML string / ElementTree into ``CompiledModel``."""
    if isinstance(xml, ET.ElementTree):
        root = xml.getroot()
    elif isinstance(xml, ET.Element):
        root = xml
    else:
        s = str(xml)
        root = ET.fromstring(s) if s.lstrip().startswith("<") else ET.parse(s).getroot()

    compiler = root.find("compiler")
    if compiler is not None and compiler.get("angle", "degree") != "radian":
        raise UnsupportedModelError("only <compiler angle='radian'> is supported")
    classes = _resolve_defaults(root)

    def attrs(elem: ET.Element, tag: str, base: Dict[str, str], childclass: str) -> Dict[str, str]:
        cls = elem.get("class", childclass)
        out = dict(base)
        out.update(classes.get(cls, classes[""]).get(tag, {}))
        out.update(elem.attrib)
        return out

    # ---- options ------------------------------------------------------------------------------
    opt = root.find("option")
    oa = opt.attrib if opt is not None else {}
    if oa.get("cone", "pyramidal") != "pyramidal":
        raise UnsupportedModelError("only cone='pyramidal' is supported")
    if oa.get("solver", "Newton") != "Newton":
        raise UnsupportedModelError("only the Newton solver is supported")
    if oa.get("integrator", "Euler") != "Euler":
        raise UnsupportedModelError("only the Euler integrator is supported")
    iterations = int(oa.get("iterations", "100"))
    if iterations != 1:
        raise UnsupportedModelError("only iterations=1 (single unrolled Newton body) is supported")
    eulerdamp_disabled = False
    if opt is not None:
        flag = opt.find("flag")
        if flag is not None and flag.get("eulerdamp", "enable") == "disable":
            eulerdamp_disabled = True
    if not eulerdamp_disabled:
        raise UnsupportedModelError("only <flag eulerdamp='disable'/> is supported")

    customs = {n.get("name"): float(n.get("data")) for n in root.findall("custom/numeric")}
    max_geom_pairs = int(customs.get("max_geom_pairs", -1))
    max_contact_points = int(customs.get("max_contact_points", -1))

    # ---- bodies (depth-first document order) ----------------------------------------------------
    world = root.find("worldbody")
    body_names, body_parent = ["world"], [0]
    body_pos, body_quat = [np.zeros(3)], [np.array([1.0, 0, 0, 0])]
    body_ipos, body_iquat = [np.zeros(3)], [np.array([1.0, 0, 0, 0])]
    body_mass, body_inertia = [0.0], [np.zeros(3)]
    body_joints: List[List[Dict[str, str]]] = [[]]
    geoms_by_body: List[List[_Geom]] = [[]]
    sites_by_body: List[List] = [[]]

    def read_geoms_sites(elem: ET.Element, bid: int, childclass: str):
        for g in elem.findall("geom"):
            a = attrs(g, "geom", _GEOM_DEFAULTS, childclass)
            geoms_by_body[bid].append(_Geom(
                name=a.get("name", ""), body=bid, type=a["type"], pos=_floats(a["pos"], 3),
                quat=_qnorm(_floats(a["quat"], 4)), size=_floats(a["size"], None, [0, 0, 0]),
                contype=int(a["contype"]), conaffinity=int(a["conaffinity"]), condim=int(a["condim"]),
                friction=_floats(a["friction"], None, [1, 0.005, 0.0001]),
                solref=_floats(a["solref"], 2), solimp=_floats(a["solimp"], None, [0.9, 0.95, 0.001, 0.5, 2]),
                solmix=float(a["solmix"]), priority=int(a["priority"])))
            if float(a["margin"]) != 0 or float(a["gap"]) != 0:
                raise UnsupportedModelError("geom margin/gap must be 0")
        for s in elem.findall("site"):
            sites_by_body[bid].append((s.get("name", ""), _floats(s.get("pos", "0 0 0"), 3)))

    read_geoms_sites(world, 0, "")

    def walk_body(elem: ET.Element, parent: int, childclass: str):
        bid = len(body_names)
        cc = elem.get("childclass", childclass)
        body_names.append(elem.get("name", f"body{bid}"))
        body_parent.append(parent)
        body_pos.append(_floats(elem.get("pos", "0 0 0"), 3))
        body_quat.append(_qnorm(_floats(elem.get("quat", "1 0 0 0"), 4)))
        inertial = elem.find("inertial")
        if inertial is None or inertial.get("diaginertia") is None:
            raise UnsupportedModelError(f"body {body_names[-1]}: explicit <inertial diaginertia=...> required")
        body_ipos.append(_floats(inertial.get("pos", "0 0 0"), 3))
        body_iquat.append(_qnorm(_floats(inertial.get("quat", "1 0 0 0"), 4)))
        body_mass.append(float(inertial.get("mass")))
        body_inertia.append(_floats(inertial.get("diaginertia"), 3))
        joints = []
        for j in elem:
            if j.tag == "freejoint":
                joints.append({"type": "free", "name": j.get("name", "")})
            elif j.tag == "joint":
                joints.append(attrs(j, "joint", _JOINT_DEFAULTS, cc))
        body_joints.append(joints)
        geoms_by_body.append([])
        sites_by_body.append([])
        read_geoms_sites(elem, bid, cc)
        for child in elem.findall("body"):
            walk_body(child, bid, cc)

    for b in world.findall("body"):
        walk_body(b, 0, "")

    nbody = len(body_names)
    if nbody != 14:
        raise UnsupportedModelError(f"expected 14 bodies (world, base, 4x3 leg links), got {nbody}")
    body_parent = np.array(body_parent)
    expect_parent = [0, 0] + sum([[1, 2 + 3 * k, 3 + 3 * k] for k in range(4)], [])
    if list(body_parent) != expect_parent:
        raise UnsupportedModelError(f"unsupported kinematic tree: parents {list(body_parent)}")
    if len(body_joints[1]) != 1 or body_joints[1][0]["type"] != "free":
        raise UnsupportedModelError("base body must carry exactly one free joint")

    # ---- joints / dofs ----------------------------------------------------------------------------
    nv, nq = 18, 19
    dof_bodyid = np.array([1] * 6 + list(range(2, 14)))
    dof_parentid = np.array([-1, 0, 1, 2, 3, 4] + sum([[5, 6 + 3 * k, 7 + 3 * k] for k in range(4)], []))
    dof_armature, dof_damping, dof_frictionloss = np.zeros(nv), np.zeros(nv), np.zeros(nv)
    jnt_names = [body_joints[1][0]["name"]]
    jnt_range = np.zeros((12, 2))
    jnt_solref = jnt_solimp = dof_solref = dof_solimp = None
    for b in range(2, 14):
        js = body_joints[b]
        if len(js) != 1 or js[0]["type"] != "hinge":
            raise UnsupportedModelError(f"body {body_names[b]} must carry exactly one hinge joint")
        a = js[0]
        if not np.allclose(_floats(a["axis"], 3), [0, 0, 1]) or not np.allclose(_floats(a["pos"], 3), 0):
            raise UnsupportedModelError("hinge joints must have axis='0 0 1' and pos='0 0 0'")
        if float(a["stiffness"]) != 0 or float(a["margin"]) != 0:
            raise UnsupportedModelError("joint stiffness/margin must be 0")
        if a.get("limited", "auto") == "false" or "range" not in a:
            raise UnsupportedModelError("hinge joints must be limited with a range")
        d = b - 2 + 6
        jnt_names.append(a.get("name", ""))
        dof_armature[d] = float(a["armature"])
        dof_damping[d] = float(a["damping"])
        dof_frictionloss[d] = float(a["frictionloss"])
        jnt_range[b - 2] = _floats(a["range"], 2)
        sl, il = _floats(a["solreflimit"], 2), _floats(a["solimplimit"], None, [0.9, 0.95, 0.001, 0.5, 2])
        sf, if_ = _floats(a["solreffriction"], 2), _floats(a["solimpfriction"], None, [0.9, 0.95, 0.001, 0.5, 2])
        for cur, new in ((jnt_solref, sl), (jnt_solimp, il), (dof_solref, sf), (dof_solimp, if_)):
            if cur is not None and not np.allclose(cur, new):
                raise UnsupportedModelError("per-joint solref/solimp must be uniform")
        jnt_solref, jnt_solimp, dof_solref, dof_solimp = sl, il, sf, if_

    # ---- geoms (grouped by body, document order inside a body) --------------------------------------
    geoms: List[_Geom] = [g for bl in geoms_by_body for g in bl]
    ngeom = len(geoms)
    body_geomnum = np.array([len(bl) for bl in geoms_by_body])
    body_geomadr = np.concatenate([[0], np.cumsum(body_geomnum)[:-1]])
    collides = np.array([(g.contype | g.conaffinity) != 0 for g in geoms])
    floor_ids = [i for i, g in enumerate(geoms) if collides[i] and g.type == "plane"]
    if len(floor_ids) != 1 or geoms[floor_ids[0]].body != 0:
        raise UnsupportedModelError("exactly one colliding world plane (the floor) is required")
    floor = geoms[floor_ids[0]]
    if not np.allclose(floor.pos, 0) or not np.allclose(floor.quat, [1, 0, 0, 0]):
        raise UnsupportedModelError("the floor plane must be z=0 with identity orientation")
    sphere_ids = [i for i, g in enumerate(geoms) if collides[i] and g.type == "sphere"]
    box_ids = [i for i, g in enumerate(geoms) if collides[i] and g.type == "box"]
    other = [i for i in range(ngeom) if collides[i] and i not in floor_ids + sphere_ids + box_ids]
    if other:
        raise UnsupportedModelError(f"unsupported colliding geom types: {[geoms[i].type for i in other]}")
    exp_sphere_bodies = sum([[3 + 3 * k, 4 + 3 * k] for k in range(4)], [])
    if [geoms[i].body for i in sphere_ids] != exp_sphere_bodies:
        raise UnsupportedModelError("expected one colliding sphere on link2 and link3 of each leg")
    if any(geoms[i].body != 0 for i in box_ids):
        raise UnsupportedModelError("colliding boxes must be static world geoms")
    for i in floor_ids + sphere_ids + box_ids:
        g = geoms[i]
        if g.condim != 3 or g.contype != 1 or g.conaffinity != 1 or g.priority != 0:
            raise UnsupportedModelError("colliding geoms must have condim=3, contype=conaffinity=1, priority=0")

    def mix(g1: _Geom, g2: _Geom):
        m = g1.solmix / (g1.solmix + g2.solmix)
        solref = m * g1.solref + (1 - m) * g2.solref if (g1.solref[0] > 0 and g2.solref[0] > 0) \
            else np.minimum(g1.solref, g2.solref)
        return solref, m * g1.solimp + (1 - m) * g2.solimp

    def uniform_mix(pairs):
        out = None
        for g1, g2 in pairs:
            r = mix(g1, g2)
            if out is not None and not (np.allclose(out[0], r[0]) and np.allclose(out[1], r[1])):
                raise UnsupportedModelError("contact solref/solimp must be uniform per pair type")
            out = r
        return out if out is not None else (np.array([0.02, 1.0]), np.array([0.9, 0.95, 0.001, 0.5, 2]))

    sph = [geoms[i] for i in sphere_ids]
    ps_ref, ps_imp = uniform_mix([(floor, s) for s in sph])
    sb_ref, sb_imp = uniform_mix([(s, geoms[b]) for s in sph for b in box_ids])
    ss_ref, ss_imp = uniform_mix([(sph[i], sph[j]) for i in range(8) for j in range(i + 1, 8)])

    # ---- sites ----------------------------------------------------------------------------------------
    site_names, site_body, site_pos = [], [], []
    for b, sl in enumerate(sites_by_body):
        for name, pos in sl:
            site_names.append(name)
            site_body.append(b)
            site_pos.append(pos)

    # ---- actuators -------------------------------------------------------------------------------------
    act = root.find("actuator")
    actuator_names, gainprm, biasprm, forcerange = [], [], [], []
    for i, a_el in enumerate(list(act) if act is not None else []):
        a = attrs(a_el, "general", _GENERAL_DEFAULTS, "")
        if a_el.tag != "general" or a["biastype"] != "affine" or a["gaintype"] != "fixed" or a["dyntype"] != "none":
            raise UnsupportedModelError("actuators must be <general biastype='affine'> position-style servos")
        if a.get("joint") != jnt_names[1 + i]:
            raise UnsupportedModelError("actuator i must drive hinge joint i")
        if "ctrlrange" in a and a.get("ctrllimited", "auto") != "false":
            raise UnsupportedModelError("ctrlrange is not supported")
        if not np.allclose(_floats(a["gear"], None, [1, 0, 0, 0, 0, 0]), [1, 0, 0, 0, 0, 0]):
            raise UnsupportedModelError("actuator gear must be 1")
        actuator_names.append(a.get("name", ""))
        gainprm.append(_floats(a["gainprm"], None, [0] * 10))
        biasprm.append(_floats(a["biasprm"], None, [0] * 10))
        limited = a["forcelimited"] == "true" or (a["forcelimited"] == "auto" and "forcerange" in a)
        forcerange.append(_floats(a["forcerange"], 2) if limited else np.array([-np.inf, np.inf]))
    if len(actuator_names) != 12:
        raise UnsupportedModelError("expected 12 actuators")

    # ---- keyframes ---------------------------------------------------------------------------------------
    keyframes = {k.get("name", f"key{i}"): _floats(k.get("qpos"), nq) for i, k in enumerate(root.findall("keyframe/key"))}

    m = CompiledModel(
        nbody=nbody, body_names=body_names, body_parent=body_parent,
        body_pos=np.array(body_pos), body_quat=np.array(body_quat), body_ipos=np.array(body_ipos),
        body_iquat=np.array(body_iquat), body_mass=np.array(body_mass), body_inertia=np.array(body_inertia),
        body_invweight0=np.zeros((nbody, 2)), nq=nq, nv=nv, nu=12, dof_bodyid=dof_bodyid,
        dof_parentid=dof_parentid, dof_armature=dof_armature, dof_damping=dof_damping,
        dof_frictionloss=dof_frictionloss, dof_invweight0=np.zeros(nv), dof_solref=dof_solref,
        dof_solimp=dof_solimp, jnt_names=jnt_names, jnt_range=jnt_range, jnt_solref=jnt_solref,
        jnt_solimp=jnt_solimp, actuator_names=actuator_names, actuator_gainprm=np.array(gainprm),
        actuator_biasprm=np.array(biasprm), actuator_forcerange=np.array(forcerange), ngeom=ngeom,
        geom_names=[g.name for g in geoms], geom_bodyid=np.array([g.body for g in geoms]),
        geom_type=[g.type for g in geoms], geom_friction=np.array([g.friction for g in geoms]),
        geom_collides=collides, body_geomadr=body_geomadr, body_geomnum=body_geomnum,
        floor_geomid=floor_ids[0], sphere_geomid=np.array(sphere_ids),
        sphere_body=np.array([geoms[i].body for i in sphere_ids]),
        sphere_pos=np.array([geoms[i].pos for i in sphere_ids]),
        sphere_radius=np.array([geoms[i].size[0] for i in sphere_ids]),
        box_geomid=np.array(box_ids, dtype=int),
        box_pos=np.array([geoms[i].pos for i in box_ids]).reshape(-1, 3),
        box_mat=np.array([quat_to_mat(geoms[i].quat) for i in box_ids]).reshape(-1, 3, 3),
        box_size=np.array([geoms[i].size for i in box_ids]).reshape(-1, 3),
        plane_sphere_solref=ps_ref, plane_sphere_solimp=ps_imp, sphere_box_solref=sb_ref,
        sphere_box_solimp=sb_imp, sphere_sphere_solref=ss_ref, sphere_sphere_solimp=ss_imp,
        site_names=site_names, site_body=np.array(site_body), site_pos=np.array(site_pos),
        timestep=float(oa.get("timestep", "0.002")), gravity=_floats(oa.get("gravity", "0 0 -9.81"), 3),
        impratio=float(oa.get("impratio", "1")), tolerance=float(oa.get("tolerance", "1e-8")),
        ls_tolerance=float(oa.get("ls_tolerance", "0.01")), iterations=iterations,
        ls_iterations=int(oa.get("ls_iterations", "50")), meaninertia=0.0,
        max_geom_pairs=max_geom_pairs, max_contact_points=max_contact_points,
        frictionloss_rows=bool(frictionloss_rows) and bool(np.any(dof_frictionloss > 0)),
        keyframes=keyframes, qpos0=np.zeros(nq), M0=np.zeros((nv, nv)), geoms=geoms,
    )
    if len(m.site_names) != 5 or m.site_body[0] != 1 or list(m.site_body[1:]) != [4, 7, 10, 13]:
        raise UnsupportedModelError("expected an IMU site on the base and one foot site per lower leg")
    if m.box_pos.shape[0] > 32:
        raise UnsupportedModelError("at most 32 obstacle boxes are supported")
    m.qpos0[:3] = m.body_pos[1]
    m.qpos0[3:7] = m.body_quat[1]
    set_const(m)
    return m


# ---- mj_setConst restated (Appendix A.12) --------------------------------------------------------------
def forward_kinematics(m: CompiledModel, qpos: np.ndarray):
    """World pose of every body for ``qpos`` (float64). Returns xpos, xquat, xmat, xipos, ximat."""
    xpos, xquat = np.zeros((m.nbody, 3)), np.zeros((m.nbody, 4))
    xquat[0] = [1, 0, 0, 0]
    for b in range(1, m.nbody):
        p = m.body_parent[b]
        pos = xpos[p] + quat_to_mat(xquat[p]) @ m.body_pos[b]
        quat = quat_mul(xquat[p], m.body_quat[b])
        if b == 1:
            pos, quat = qpos[:3].copy(), _qnorm(qpos[3:7])
        else:
            ang = qpos[7 + b - 2]
            quat = quat_mul(quat, np.array([np.cos(ang / 2), 0, 0, np.sin(ang / 2)]))
        xpos[b], xquat[b] = pos, quat
    xmat = np.array([quat_to_mat(q) for q in xquat])
    xipos = np.array([xpos[b] + xmat[b] @ m.body_ipos[b] for b in range(m.nbody)])
    ximat = np.array([quat_to_mat(quat_mul(xquat[b], m.body_iquat[b])) for b in range(m.nbody)])
    return xpos, xquat, xmat, xipos, ximat


def body_jacobian(m: CompiledModel, xpos, xmat, point: np.ndarray, body: int) -> np.ndarray:
    """6 x nv Jacobian [jacp; jacr] of ``point`` attached to ``body`` (MuJoCo dof conventions)."""
    J = np.zeros((6, m.nv))
    if body == 0:
        return J
    J[0:3, 0:3] = np.eye(3)
    for i in range(3):
        a = xmat[1][:, i]
        J[0:3, 3 + i] = np.cross(a, point - xpos[1])
        J[3:6, 3 + i] = a
    b = body
    while b >= 2:
        axis = xmat[b][:, 2]
        d = 6 + b - 2
        J[0:3, d] = np.cross(axis, point - xpos[b])
        J[3:6, d] = axis
        b = m.body_parent[b]
    return J


def mass_matrix(m: CompiledModel, qpos: np.ndarray, body_mass=None, body_inertia=None, body_ipos=None):
    """Joint-space inertia by summing J^T [m, I] J over bodies (independent of the CRBA in the oracle)."""
    mm = m if body_ipos is None else dataclasses.replace(m, body_ipos=body_ipos)
    xpos, xquat, xmat, xipos, ximat = forward_kinematics(mm, qpos)
    mass = m.body_mass if body_mass is None else body_mass
    inertia = m.body_inertia if body_inertia is None else body_inertia
    M = np.diag(m.dof_armature).astype(np.float64)
    for b in range(1, m.nbody):
        J = body_jacobian(m, xpos, xmat, xipos[b], b)
        Iw = ximat[b] @ np.diag(inertia[b]) @ ximat[b].T
        M += mass[b] * J[:3].T @ J[:3] + J[3:].T @ Iw @ J[3:]
    return M


def set_const(m: CompiledModel) -> None:
    xpos, xquat, xmat, xipos, ximat = forward_kinematics(m, m.qpos0)
    M0 = mass_matrix(m, m.qpos0)
    A = np.linalg.inv(M0)
    m.M0 = M0
    m.meaninertia = float(np.mean(np.diag(M0)))
    inv = np.diag(A).copy()
    inv[0:3] = inv[0:3].mean()
    inv[3:6] = inv[3:6].mean()
    m.dof_invweight0 = inv
    bw = np.zeros((m.nbody, 2))
    for b in range(1, m.nbody):
        J = body_jacobian(m, xpos, xmat, xipos[b], b)
        S = J @ A @ J.T
        bw[b, 0] = np.trace(S[:3, :3]) / 3.0
        bw[b, 1] = np.trace(S[3:, 3:]) / 3.0
    m.body_invweight0 = bw


# ---- canonical MJCF emitter ----------------------------------------------------------------------------
def _fmt(a) -> str:
    return " ".join(repr(float(x)) for x in np.atleast_1d(a))


def to_xml(m: CompiledModel) -> str:
    """Emit a canonical, default-free MJCF of the compiled model (``compile_model(to_xml(m))`` reproduces
    ``m``).  Geoms that never collide (the reference's visual meshes / visual plane) are emitted as
    non-colliding placeholder spheres so geom ids and counts -- which the reference's reward code and DR
    shapes depend on (``rewards.py:131-138``, ``test_domain_randomization.py:74``) -- are preserved without
    needing mesh assets."""
    root = ET.Element("mujoco", model="pupper_v3_canonical")
    ET.SubElement(root, "compiler", angle="radian", autolimits="true")
    opt = ET.SubElement(root, "option", cone="pyramidal", impratio=_fmt(m.impratio), iterations=str(m.iterations),
                        ls_iterations=str(m.ls_iterations), timestep=_fmt(m.timestep), gravity=_fmt(m.gravity),
                        tolerance=_fmt(m.tolerance), ls_tolerance=_fmt(m.ls_tolerance))
    ET.SubElement(opt, "flag", eulerdamp="disable")
    custom = ET.SubElement(root, "custom")
    ET.SubElement(custom, "numeric", name="max_contact_points", data=str(m.max_contact_points))
    ET.SubElement(custom, "numeric", name="max_geom_pairs", data=str(m.max_geom_pairs))
    world = ET.SubElement(root, "worldbody")

    def emit_geoms(parent: ET.Element, bid: int):
        for gi in range(m.body_geomadr[bid], m.body_geomadr[bid] + m.body_geomnum[bid]):
            g = m.geoms[gi]
            a = dict(pos=_fmt(g.pos), quat=_fmt(g.quat), condim=str(g.condim), friction=_fmt(g.friction),
                     solref=_fmt(g.solref), solimp=_fmt(g.solimp), solmix=_fmt(g.solmix))
            if g.name:
                a["name"] = g.name
            if m.geom_collides[gi]:
                a.update(type=g.type, size=_fmt(g.size), contype=str(g.contype), conaffinity=str(g.conaffinity))
            else:
                a.update(type="sphere", size="0.001", contype="0", conaffinity="0", group="1")
            ET.SubElement(parent, "geom", a)

    def emit_sites(parent: ET.Element, bid: int):
        for si, sb in enumerate(m.site_body):
            if sb == bid:
                ET.SubElement(parent, "site", name=m.site_names[si], pos=_fmt(m.site_pos[si]))

    elems = {0: world}
    emit_geoms(world, 0)
    for b in range(1, m.nbody):
        e = ET.SubElement(elems[int(m.body_parent[b])], "body", name=m.body_names[b], pos=_fmt(m.body_pos[b]),
                          quat=_fmt(m.body_quat[b]))
        elems[b] = e
        ET.SubElement(e, "inertial", pos=_fmt(m.body_ipos[b]), quat=_fmt(m.body_iquat[b]), mass=_fmt(m.body_mass[b]),
                      diaginertia=_fmt(m.body_inertia[b]))
        if b == 1:
            ET.SubElement(e, "freejoint", name=m.jnt_names[0])
        else:
            d = b + 4
            ET.SubElement(e, "joint", name=m.jnt_names[b - 1], type="hinge", axis="0 0 1", pos="0 0 0",
                          range=_fmt(m.jnt_range[b - 2]), limited="true", armature=_fmt(m.dof_armature[d]),
                          damping=_fmt(m.dof_damping[d]), frictionloss=_fmt(m.dof_frictionloss[d]),
                          solreflimit=_fmt(m.jnt_solref), solimplimit=_fmt(m.jnt_solimp),
                          solreffriction=_fmt(m.dof_solref), solimpfriction=_fmt(m.dof_solimp))
        emit_geoms(e, b)
        emit_sites(e, b)
    act = ET.SubElement(root, "actuator")
    for i, name in enumerate(m.actuator_names):
        a = dict(name=name, joint=m.jnt_names[1 + i], biastype="affine", gainprm=_fmt(m.actuator_gainprm[i][:3]),
                 biasprm=_fmt(m.actuator_biasprm[i][:3]))
        if np.all(np.isfinite(m.actuator_forcerange[i])):
            a.update(forcelimited="true", forcerange=_fmt(m.actuator_forcerange[i]))
        ET.SubElement(act, "general", a)
    kf = ET.SubElement(root, "keyframe")
    for name, q in m.keyframes.items():
        ET.SubElement(kf, "key", name=name, qpos=_fmt(q))
    ET.indent(root)
    return ET.tostring(root, encoding="unicode")
