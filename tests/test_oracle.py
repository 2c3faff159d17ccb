"""CPU: the numpy oracle against the golden vectors produced by the reference itself,
plus identities that anchor the transforms3d restatement (parity unpinned there)."""
import math

import numpy as np
import pytest

from conftest import GOLDEN_CASES, GOLDEN_CASES_F4, GOLDEN_CASES_FALLBACK, GOLDEN_CASES_GENERIC, golden_oracle_batch, load_golden
from oracle import osc_numpy, t3d, ref_harness


@pytest.mark.parametrize("case", GOLDEN_CASES + GOLDEN_CASES_F4 + GOLDEN_CASES_GENERIC + GOLDEN_CASES_FALLBACK)
def test_oracle_matches_reference_golden(case):
    g, layout = load_golden(case)
    batch = golden_oracle_batch(g)
    B = g["dq"].shape[0]
    for i in range(B):
        inst = {k: v[i] for k, v in batch.items()}
        if g["index_error"][i]:
            with pytest.raises(IndexError):
                osc_numpy.osc_step(layout, inst)
            continue
        o = osc_numpy.osc_step(layout, inst)
        assert o["pinv"] == bool(g["pinv"][i])
        scale = np.abs(g["u_all"][i]).max()
        assert np.abs(o["u_all"] - g["u_all"][i]).max() <= 1e-10 * scale
        assert np.abs(o["ctrl"] - g["ctrl"][i]).max() <= 1e-10 * scale


def test_golden_covers_both_branches_and_quirks():
    pinv = np.concatenate([load_golden(c)[0]["pinv"] for c in GOLDEN_CASES])
    assert pinv.any() and (~pinv).any()
    assert load_golden("insertion_vel_s6")[0]["index_error"].all()
    tv = load_golden("gain_test_vel_s4")[0]["target_vel"]
    nz = (tv != 0).all(axis=-1)
    assert nz.any() and (~nz).any()                                      # both branches


@pytest.mark.reference
@pytest.mark.skipif(not ref_harness.reference_available(), reason="needs /root/reference")
def test_oracle_matches_live_reference():
    """Fresh seeds through the unmodified reference (not just the committed fixtures)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden
    from irl_control_b200.synthetic import build_scenario, synth_batch, oracle_inputs
    for scenario, seed in (("gain_test", 101), ("admit_test", 102), ("worst_case", 103)):
        _, _, _, layout = build_scenario(scenario)
        st = synth_batch(layout, 6, seed=seed)
        runner = make_golden.reference_runner(scenario)
        ob = oracle_inputs(st, layout)
        sn = {k: v.numpy() for k, v in st.items()}
        for i in range(6):
            r = runner.run({k: v[i] for k, v in sn.items()}, sn["target_xyz"][i], sn["target_quat"][i],
                           max_vel=sn["max_vel"][i])
            o = osc_numpy.osc_step(layout.as_dict(), {k: v[i] for k, v in ob.items()})
            scale = np.abs(r["u_all"]).max()
            assert np.abs(o["u_all"] - r["u_all"]).max() <= 1e-10 * scale
            assert o["pinv"] == r["pinv"]


# ---- transforms3d restatement: mathematical anchors ---------------------------------
def _rot(axis, a):
    c, s = math.cos(a), math.sin(a)
    return {0: np.array([[1, 0, 0], [0, c, -s], [0, s, c]]),
            1: np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]),
            2: np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])}[axis]


def test_t3d_static_xyz_convention():
    rng = np.random.default_rng(0)
    for _ in range(50):
        a, b, c = rng.uniform(-1.5, 1.5, 3)
        R = _rot(2, c) @ _rot(1, b) @ _rot(0, a)      # static x, then y, then z
        q = t3d.euler2quat(a, b, c)
        assert abs(np.linalg.norm(q) - 1) < 1e-14
        assert np.abs(t3d.quat2mat(q) - R).max() < 1e-14
        assert np.allclose(t3d.quat2euler(q), (a, b, c), atol=1e-13)
        assert np.allclose(t3d.mat2euler(R), (a, b, c), atol=1e-13)


def test_t3d_quaternion_algebra():
    rng = np.random.default_rng(1)
    for _ in range(50):
        p, q = rng.normal(size=4), rng.normal(size=4)
        p, q = t3d.normalized_vector(p), t3d.normalized_vector(q)
        pq = np.array(t3d.qmult(p, q))
        assert np.abs(t3d.quat2mat(pq) - t3d.quat2mat(p) @ t3d.quat2mat(q)).max() < 1e-14
        assert np.allclose(np.array(t3d.qmult(q, t3d.qconjugate(q))), [1, 0, 0, 0], atol=1e-15)
    assert np.array_equal(t3d.quat2mat([0, 0, 0, 0]), np.eye(3))
    # gimbal branch of mat2euler (cy <= 4 eps)
    ax, ay, az = t3d.mat2euler(_rot(1, math.pi / 2) @ _rot(0, 0.3))
    assert az == 0.0 and abs(ay - math.pi / 2) < 1e-12


def test_host_rotations_equal_oracle_rotations():
    from irl_control_b200 import rotations as R
    rng = np.random.default_rng(2)
    for _ in range(100):
        e = rng.uniform(-3, 3, 3)
        q = rng.normal(size=4)
        assert np.array_equal(R.euler2quat(*e), t3d.euler2quat(*e))
        assert R.quat2euler(q) == t3d.quat2euler(q)
        assert np.array_equal(R.qmult(q, e.tolist() + [1.0]), np.array(t3d.qmult(q, e.tolist() + [1.0])))
        assert np.array_equal(R.normalized_vector(q), t3d.normalized_vector(q))


def test_t3d_restatement_agrees_with_scipy_rotations():
    """Independent third-party anchor for the transforms3d restatement (transforms3d itself is absent):
    scipy's `Rotation` with extrinsic axes 'xyz' is the same static-frame x-y-z convention transforms3d calls
    'sxyz' (its default); scipy quaternions are scalar-last."""
    Rotation = pytest.importorskip("scipy.spatial.transform").Rotation
    rng = np.random.default_rng(7)
    for _ in range(200):
        e = rng.uniform([-np.pi, -np.pi / 2 + 1e-3, -np.pi], [np.pi, np.pi / 2 - 1e-3, np.pi])
        r = Rotation.from_euler("xyz", e)
        x, y, z, w = r.as_quat()
        q = t3d.euler2quat(*e)
        sgn = 1.0 if q[0] * w >= 0 else -1.0                         # q and -q are the same rotation
        assert np.abs(q - sgn * np.array([w, x, y, z])).max() < 1e-14
        assert np.abs(t3d.quat2mat(q) - r.as_matrix()).max() < 1e-14
        assert np.abs(t3d.euler2mat(*e) - r.as_matrix()).max() < 1e-14
        assert np.allclose(t3d.quat2euler(q), r.as_euler("xyz"), atol=1e-12)
        # Hamilton product order: qmult(a, b) is "b then a" as rotations
        e2 = rng.uniform(-1.0, 1.0, 3)
        r2 = Rotation.from_euler("xyz", e2)
        qa = t3d.qmult(q, t3d.euler2quat(*e2))
        assert np.abs(t3d.quat2mat(qa) - (r * r2).as_matrix()).max() < 1e-14
        assert np.abs(np.asarray(t3d.qconjugate(q)) - np.array([q[0], -q[1], -q[2], -q[3]])).max() == 0.0
    # the pose error of osc.py:115-117 as a whole: euler(conj(q_d * conj(q_ee)))
    for _ in range(50):
        qd, qe = rng.normal(size=4), rng.normal(size=4)
        qe /= np.linalg.norm(qe)
        rd = Rotation.from_quat(np.roll(qd / np.linalg.norm(qd), -1))
        re_ = Rotation.from_quat(np.roll(qe, -1))
        want = (rd * re_.inv()).inv().as_euler("xyz")
        q_r = t3d.qmult(t3d.normalized_vector(qd), t3d.qconjugate(qe))
        got = np.array(t3d.quat2euler(t3d.qconjugate(q_r)))
        assert np.allclose(got, want, atol=1e-11)
