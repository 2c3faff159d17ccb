#!/usr/bin/env python3
"""One-off extractor: reads the DualUR5 MJCF of the reference and emits the
kinematic / inertial parameter table `irl_control_b200/dual_ur5_model.py`.

The GPU box has no /root/reference, and this image has no MuJoCo, so the
topology the reference gets from `load_model_from_path` (mujoco_app.py:17)
must travel with the repo as plain data. Only what the OSC path needs is
kept: the body tree, hinge joints, <inertial> elements, sites, actuators and
sensors. Geoms, meshes, materials and equality constraints are dropped (no
collision / rendering / constraint solving on this path) - except that a body
WITHOUT an <inertial> element gets its mass, centre of mass and inertia from
its mesh geoms at the default density of 1000 kg/m^3, as MuJoCo's compiler
does (inertiafromgeom="auto"): base_link_ur5right / base_link_ur5left
(dual_ur5.xml:63-64, 164-165) carry only the `link0` mesh and are welded to the
rotating stand, so they contribute to M[stand][stand] and to qfrc_bias.

    python tools/extract_dual_ur5.py /root/reference/irl_control/scenes/dual_ur5.xml
"""
import os
import struct
import sys
import math
import xml.etree.ElementTree as ET

DEFAULT_DENSITY = 1000.0      # mjCGeom default density [kg / m^3]


def read_stl(path):
    """Triangles [(v0, v1, v2)] of a binary or ASCII STL."""
    data = open(path, "rb").read()
    n = struct.unpack_from("<I", data, 80)[0] if len(data) >= 84 else -1
    if n >= 0 and len(data) == 84 + 50 * n:
        tris = []
        for i in range(n):
            f = struct.unpack_from("<12f", data, 84 + 50 * i)
            tris.append((f[3:6], f[6:9], f[9:12]))
        return tris
    verts = [tuple(float(x) for x in ln.split()[1:4]) for ln in data.decode("ascii", "replace").splitlines()
             if ln.strip().startswith("vertex")]
    return [tuple(verts[i:i + 3]) for i in range(0, len(verts) - 2, 3)]


def mesh_inertial(tris, density=DEFAULT_DENSITY):
    """(pos, quat_wxyz, mass, diaginertia) of a uniform-density mesh the way MuJoCo 2.0 / 2.1 (the engine behind
    mujoco_py; `mjCMesh::Process`, later called inertia="legacy") computes it: the surface is cut into pyramids with
    their apex at the area-weighted centroid of the faces, every pyramid counts with its ABSOLUTE volume, the centre of
    mass and the second moments are the exact integrals over those pyramids, and the geom is given the principal
    frame of that tensor.  MuJoCo itself is absent from this image: restated from its published algorithm, unpinned."""
    import numpy as np
    T = np.asarray(tris, dtype=np.float64)                       # [F][3][3]
    cen = T.mean(axis=1)
    area = 0.5 * np.linalg.norm(np.cross(T[:, 1] - T[:, 0], T[:, 2] - T[:, 0]), axis=1)
    apex = (cen * area[:, None]).sum(0) / area.sum()
    A = T - apex                                                  # vertices relative to the apex
    vol = np.abs(np.einsum("fi,fi->f", A[:, 0], np.cross(A[:, 1], A[:, 2]))) / 6.0
    com = apex + (vol[:, None] * (A.sum(axis=1) / 4.0)).sum(0) / vol.sum()
    # second moments of a tetrahedron (apex at the origin o = 0, vertices a, b, c) about the origin:
    #   int r r^T dV = V / 20 * (a a^T + b b^T + c c^T + s s^T),  s = a + b + c
    C = np.zeros((3, 3))
    for f in range(T.shape[0]):
        P = T[f] - com
        o = apex - com
        pts = np.vstack([P, o[None]])
        ssum = pts.sum(0)
        C += vol[f] / 20.0 * (pts.T @ pts + np.outer(ssum, ssum))
    I = np.trace(C) * np.eye(3) - C
    w, V = np.linalg.eigh(I)
    order = np.argsort(-w)                                        # MuJoCo orders the principal moments descending
    w, V = w[order], V[:, order]
    if np.linalg.det(V) < 0:
        V[:, 2] = -V[:, 2]
    # rotation matrix -> quaternion (w x y z)
    tr = np.trace(V)
    if tr > 0:
        S = math.sqrt(tr + 1.0) * 2
        q = [0.25 * S, (V[2, 1] - V[1, 2]) / S, (V[0, 2] - V[2, 0]) / S, (V[1, 0] - V[0, 1]) / S]
    else:
        i = int(np.argmax(np.diag(V)))
        j, k = (i + 1) % 3, (i + 2) % 3
        S = math.sqrt(1.0 + V[i, i] - V[j, j] - V[k, k]) * 2
        q = [0.0] * 4
        q[0] = (V[k, j] - V[j, k]) / S
        q[1 + i] = 0.25 * S
        q[1 + j] = (V[j, i] + V[i, j]) / S
        q[1 + k] = (V[k, i] + V[i, k]) / S
    if q[0] < 0:
        q = [-x for x in q]
    mass = density * vol.sum()
    return ([float(x) for x in com], [float(x) for x in q], float(mass), [float(density * x) for x in w])


def floats(s, n=None, default=None):
    if s is None:
        return default
    v = [float(x) for x in s.split()]
    assert n is None or len(v) == n, (s, n)
    return v


def euler_xyz_to_quat(e):
    # MuJoCo default eulerseq="xyz" (intrinsic x-y'-z''), angle="radian" in the
    # scene files that include this model (gain_test_scene.xml:2).
    def axq(axis, a):
        q = [math.cos(a / 2), 0.0, 0.0, 0.0]
        q[1 + axis] = math.sin(a / 2)
        return q

    def mul(a, b):
        w1, x1, y1, z1 = a
        w2, x2, y2, z2 = b
        return [w1*w2 - x1*x2 - y1*y2 - z1*z2,
                w1*x2 + x1*w2 + y1*z2 - z1*y2,
                w1*y2 - x1*z2 + y1*w2 + z1*x2,
                w1*z2 + x1*y2 - y1*x2 + z1*w2]
    q = axq(0, e[0])
    q = mul(q, axq(1, e[1]))
    q = mul(q, axq(2, e[2]))
    return q


def frame_quat(el):
    if el.get("quat") is not None:
        return floats(el.get("quat"), 4)
    if el.get("euler") is not None:
        return euler_xyz_to_quat(floats(el.get("euler"), 3))
    return [1.0, 0.0, 0.0, 0.0]


def main(path):
    root = ET.parse(path).getroot()
    bodies = []
    # mesh name -> file, resolved like the including scene does (<compiler meshdir="../meshes/"/>, gain_test_scene.xml:2)
    meshdir = os.path.join(os.path.dirname(os.path.abspath(path)), "..", "meshes")
    mesh_file = {m.get("name"): os.path.normpath(os.path.join(meshdir, m.get("file"))) for m in root.iter("mesh")}

    def walk(el, parent):
        name = el.get("name")
        rec = {
            "name": name, "parent": parent,
            "pos": floats(el.get("pos"), 3, [0.0, 0.0, 0.0]),
            "quat": frame_quat(el),
            "inertial": None, "joints": [], "sites": [],
        }
        for ch in el:
            if ch.tag == "inertial":
                rec["inertial"] = {
                    "pos": floats(ch.get("pos"), 3, [0.0, 0.0, 0.0]),
                    "quat": frame_quat(ch),
                    "mass": float(ch.get("mass")),
                    "diaginertia": floats(ch.get("diaginertia"), 3),
                }
            elif ch.tag == "joint":
                assert ch.get("type", "hinge") == "hinge"
                rec["joints"].append({
                    "name": ch.get("name"),
                    "axis": floats(ch.get("axis"), 3),
                    "pos": floats(ch.get("pos"), 3, [0.0, 0.0, 0.0]),
                    "range": floats(ch.get("range"), 2, None),
                })
            elif ch.tag == "site":
                rec["sites"].append({
                    "name": ch.get("name"),
                    "pos": floats(ch.get("pos"), 3, [0.0, 0.0, 0.0]),
                    "quat": frame_quat(ch),
                })
        geoms = [g for g in el if g.tag == "geom"]
        if rec["inertial"] is None and geoms:
            # inertiafromgeom="auto": no <inertial> -> from the geoms (here always ONE mesh geom at the body origin)
            assert len(geoms) == 1 and geoms[0].get("type") == "mesh" and geoms[0].get("mass") is None, name
            assert floats(geoms[0].get("pos"), 3, [0.0, 0.0, 0.0]) == [0.0, 0.0, 0.0] and geoms[0].get("quat") is None
            ipos, iquat, mass, diag = mesh_inertial(read_stl(mesh_file[geoms[0].get("mesh")]),
                                                    float(geoms[0].get("density", DEFAULT_DENSITY)))
            rec["inertial"] = {"pos": ipos, "quat": iquat, "mass": mass, "diaginertia": diag}
        bodies.append(rec)
        for ch in el:
            if ch.tag == "body":
                walk(ch, name)

    for b in root.find("worldbody"):
        if b.tag == "body":
            walk(b, "world")
    actuators = [(a.tag, a.get("name"), a.get("joint")) for a in root.find("actuator")]
    sensors = [(s.tag, s.get("name"), s.get("site")) for s in root.find("sensor")]

    out = []
    w = out.append
    w('"""DualUR5 kinematic / inertial parameters (generated data, do not edit).')
    w("")
    w("Generated by tools/extract_dual_ur5.py from the reference scene")
    w("`irl_control/scenes/dual_ur5.xml:51-297` (bodies 51-263, actuators 267-287,")
    w("sensors 289-297). Frames are (pos, quat wxyz) relative to the parent body;")
    w("`euler=` attributes were converted with MuJoCo's default intrinsic xyz")
    w("sequence in radians. Geoms/meshes/equalities are intentionally absent: the")
    w("OSC path never touches them - except that bodies without an <inertial> element")
    w("(base_link_ur5right / base_link_ur5left) carry the mass, centre of mass and")
    w("principal inertia of their `link0` mesh at MuJoCo's default density of 1000 kg/m^3.")
    w('"""')
    w("")
    w("# (name, parent, pos, quat_wxyz, inertial|None, joints, sites)")
    w("#   inertial = (pos, quat_wxyz, mass, diaginertia)")
    w("#   joint    = (name, axis, pos, range|None)     -- all hinge")
    w("#   site     = (name, pos, quat_wxyz)")
    w("BODIES = [")
    for b in bodies:
        inert = None
        if b["inertial"]:
            i = b["inertial"]
            inert = (tuple(i["pos"]), tuple(i["quat"]), i["mass"], tuple(i["diaginertia"]))
        joints = [(j["name"], tuple(j["axis"]), tuple(j["pos"]),
                   tuple(j["range"]) if j["range"] else None) for j in b["joints"]]
        sites = [(s["name"], tuple(s["pos"]), tuple(s["quat"])) for s in b["sites"]]
        w("    (%r, %r, %r, %r,\n     %r,\n     %r,\n     %r)," % (
            b["name"], b["parent"], tuple(b["pos"]), tuple(b["quat"]), inert, joints, sites))
    w("]")
    w("")
    w("# (kind, name, joint)  -- ctrl index == position in this list")
    w("ACTUATORS = [")
    for a in actuators:
        w("    %r," % (a,))
    w("]")
    w("")
    w("# (kind, name, site)   -- each sensor contributes 3 scalars to sensordata")
    w("SENSORS = [")
    for s in sensors:
        w("    %r," % (s,))
    w("]")
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
