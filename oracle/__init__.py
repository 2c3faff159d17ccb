"""TEST INFRASTRUCTURE ONLY - CPU oracle for the DualUR5 OSC hot path.

Nothing under `oracle/` is part of the shipped product.  It may be imported
only by `tests/`, by `__graft_entry__.smoke()` (as the checker) and by
`bench.py`'s `cpu_baseline` / `--impl reference` legs (as the timed CPU
baseline).  The product package `irl_control_b200` never imports it and fails
loudly when its CUDA library is missing.

Contents
--------
osc_numpy.py    clean-room float64 numpy restatement of the reference control
                law (`irl_control/osc.py:35-39,41-68,70-99,101-118,132-210`,
                `robot.py:44-72`, `device.py:36,66-74,115-170`,
                `utils.py:10-67`), batched over instances with a plain loop.
t3d.py          restatement of the five `transforms3d` functions the path
                calls (third-party, unpinned in `requirements.in:3`, absent
                from /root/reference and from this image).
ref_harness.py  drives the UNMODIFIED reference sources from /root/reference
                through stub `mujoco_py` / `transforms3d` modules and a fake
                `sim`; only usable where /root/reference exists (this
                container).  Used to pin osc_numpy.py and to generate
                `tests/golden/*.npz`.

Parity pinning
--------------
The reference ships no golden vectors, KATs or fixtures for this path
(`irl_control/tests/run_tests.py:1` is `assert True`).  The oracle is
therefore pinned against outputs of the reference itself run here:
`tests/golden/make_golden.py` calls the real `OSC.generate` via
ref_harness.py and stores inputs + outputs; `tests/test_oracle.py` checks
osc_numpy.py against those vectors.  The one boundary that cannot be pinned
is `transforms3d` (not installed anywhere reachable): t3d.py is written from
the published algorithm of transforms3d 0.4.x ('sxyz' static-frame Euler
convention, w-x-y-z quaternions) - "parity unpinned" for those five
functions only; DESIGN.md repeats this.
"""
