"""Quaternion / Euler helpers used by the host layer (w, x, y, z order).

The reference takes these from the third-party package `transforms3d`
(`irl_control/utils.py:3`, `irl_control/osc.py:4-7`), which is neither vendored
in the reference tree nor installed in this image.  Only the static-frame
'sxyz' convention is used on the OSC path, so that is all that is provided:

    euler2quat(ai, aj, ak)   <- utils.py:14-15,53,57,67
    quat2euler(q)            <- utils.py:30,33 ; osc.py:117
    qmult / qconjugate       <- osc.py:116-117
    normalized_vector        <- osc.py:115

Scalar (single pose) versions live here; the batched CUDA kernel carries its
own device copies (csrc/irlosc_device.cuh) and is checked against oracle/.
"""
import math

import numpy as np

_EPS = float(np.finfo(np.float64).eps)
_EPS4 = 4.0 * _EPS


def normalized_vector(v):
    v = np.asarray(v, dtype=np.float64).reshape(-1)
    return v / math.sqrt(float((v ** 2).sum()))


def qconjugate(q):
    q = np.asarray(q, dtype=np.float64)
    return np.array([q[0], -q[1], -q[2], -q[3]])


def qmult(a, b):
    aw, ax, ay, az = (float(t) for t in a)
    bw, bx, by, bz = (float(t) for t in b)
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
    ])


def quat2mat(q):
    w, x, y, z = (float(t) for t in q)
    nq = w * w + x * x + y * y + z * z
    if nq < _EPS:
        return np.eye(3)
    s = 2.0 / nq
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return np.array([
        [1.0 - (yy + zz), xy - wz, xz + wy],
        [xy + wz, 1.0 - (xx + zz), yz - wx],
        [xz - wy, yz + wx, 1.0 - (xx + yy)],
    ])


def mat2euler(m):
    """Static-frame x-y-z ('sxyz') angles of a rotation matrix."""
    m = np.asarray(m, dtype=np.float64)
    cy = math.sqrt(m[0, 0] * m[0, 0] + m[1, 0] * m[1, 0])
    if cy > _EPS4:
        ax = math.atan2(m[2, 1], m[2, 2])
        ay = math.atan2(-m[2, 0], cy)
        az = math.atan2(m[1, 0], m[0, 0])
    else:
        ax = math.atan2(-m[1, 2], m[1, 1])
        ay = math.atan2(-m[2, 0], cy)
        az = 0.0
    return ax, ay, az


def quat2euler(q):
    return mat2euler(quat2mat(q))


def euler2quat(ai, aj, ak):
    hi, hj, hk = 0.5 * float(ai), 0.5 * float(aj), 0.5 * float(ak)
    ci, si = math.cos(hi), math.sin(hi)
    cj, sj = math.cos(hj), math.sin(hj)
    ck, sk = math.cos(hk), math.sin(hk)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return np.array([
        cj * cc + sj * ss,
        cj * sc - sj * cs,
        cj * ss + sj * cc,
        cj * cs - sj * sc,
    ])
