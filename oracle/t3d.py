"""TEST INFRASTRUCTURE ONLY - restatement of the `transforms3d` calls on the OSC path.

transforms3d is a third-party dependency of the reference (unpinned,
`requirements.in:3`), not vendored under /root/reference and not installed in
this image; its call sites on the path are

    osc.py:115  normalized_vector        (transforms3d.utils)
    osc.py:116  qmult                    (transforms3d.derivations.quaternions)
    osc.py:116  qconjugate               (transforms3d.quaternions)
    osc.py:117  quat2euler               (transforms3d.euler, default axes 'sxyz')
    utils.py:14 euler2quat               (transforms3d.euler, default axes 'sxyz')

Published algorithm (transforms3d 0.4.x): quaternions are (w, x, y, z);
`quat2euler(q) = mat2euler(quat2mat(q))`; `mat2euler` for 'sxyz' uses
cy = hypot(M00, M10) with the 4*eps gimbal switch; `euler2quat` is the
half-angle product in static x, y, z order.  PARITY UNPINNED against
transforms3d itself: the package is unavailable here, so these functions are
anchored on mathematical identities (tests/test_oracle.py: round trips,
composition with rotation matrices) and on scipy's independent `Rotation`
implementation with extrinsic "xyz" axes (same convention) - not on
transforms3d outputs.
"""
import math

import numpy as np

_FLOAT_EPS = np.finfo(np.float64).eps
_EPS4 = _FLOAT_EPS * 4.0


def normalized_vector(vec):
    vec = np.asarray(vec).squeeze()
    return vec / math.sqrt((vec ** 2).sum())


def qconjugate(q):
    return np.array(q) * np.array([1.0, -1, -1, -1])


def qmult(q1, q2):
    w1, x1, y1, z1 = q1
    w2, x2, y2, z2 = q2
    w = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2
    x = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2
    y = w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2
    z = w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2
    return w, x, y, z


def quat2mat(q):
    w, x, y, z = q
    Nq = w * w + x * x + y * y + z * z
    if Nq < _FLOAT_EPS:
        return np.eye(3)
    s = 2.0 / Nq
    X, Y, Z = x * s, y * s, z * s
    wX, wY, wZ = w * X, w * Y, w * Z
    xX, xY, xZ = x * X, x * Y, x * Z
    yY, yZ, zZ = y * Y, y * Z, z * Z
    return np.array([[1.0 - (yY + zZ), xY - wZ, xZ + wY],
                     [xY + wZ, 1.0 - (xX + zZ), yZ - wX],
                     [xZ - wY, yZ + wX, 1.0 - (xX + yY)]])


def mat2euler(mat, axes='sxyz'):
    assert axes == 'sxyz'
    M = np.asarray(mat, dtype=np.float64)[:3, :3]
    cy = math.sqrt(M[0, 0] * M[0, 0] + M[1, 0] * M[1, 0])
    if cy > _EPS4:
        ax = math.atan2(M[2, 1], M[2, 2])
        ay = math.atan2(-M[2, 0], cy)
        az = math.atan2(M[1, 0], M[0, 0])
    else:
        ax = math.atan2(-M[1, 2], M[1, 1])
        ay = math.atan2(-M[2, 0], cy)
        az = 0.0
    return ax, ay, az


def quat2euler(quaternion, axes='sxyz'):
    return mat2euler(quat2mat(quaternion), axes)


def euler2quat(ai, aj, ak, axes='sxyz'):
    assert axes == 'sxyz'
    ai, aj, ak = ai / 2.0, aj / 2.0, ak / 2.0
    ci, si = math.cos(ai), math.sin(ai)
    cj, sj = math.cos(aj), math.sin(aj)
    ck, sk = math.cos(ak), math.sin(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    q = np.empty((4,))
    q[0] = cj * cc + sj * ss
    q[1] = cj * sc - sj * cs
    q[2] = cj * ss + sj * cc
    q[3] = cj * cs - sj * sc
    return q


def euler2mat(ai, aj, ak, axes='sxyz'):
    assert axes == 'sxyz'
    return quat2mat(euler2quat(ai, aj, ak))
