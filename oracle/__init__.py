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
sequence_numpy.py  restatement of the caller loops around `generate`:
                `examples/insertion_task.py` (run_sequence, go_to_waypoint,
                grip, send_forces, set_waypoint_targets, object placement)
                and the waypoint cycling of `examples/gain_test.py:134-162`.
ref_harness.py  drives the UNMODIFIED reference sources from /root/reference
                through stub `mujoco_py` / `transforms3d` modules and a fake
                `sim`: `OSC.generate` (ReferenceRunner), the insertion demo's
                own methods and `GainTest.run` on pose streams
                (drive_reference_sequence, drive_reference_gain_test,
                reference_object_placement).  Its reference-driving parts only
                work where /root/reference exists (this container); `FakeSim`
                alone is plain numpy and is also used by
                tests/test_dropin_golden.py.  Used to pin the restatements
                and to generate `tests/golden/*.npz`.

Parity pinning
--------------
The reference ships no golden vectors, KATs or fixtures for this path
(`irl_control/tests/run_tests.py:1` is `assert True`).  The oracle is
therefore pinned against outputs of the reference itself run here:
`tests/golden/make_golden.py` calls the real `OSC.generate` via
ref_harness.py and stores inputs + outputs; `tests/test_oracle.py` checks
osc_numpy.py against those vectors (21 files: 18 of `OSC.generate` - every
shipped configuration plus the branches no shipped YAML takes - and 3 of the
caller loops).  The one boundary that cannot be pinned
is `transforms3d` (not installed anywhere reachable): t3d.py is written from
the published algorithm of transforms3d 0.4.x ('sxyz' static-frame Euler
convention, w-x-y-z quaternions) and cross-checked against scipy's
independent `Rotation` implementation - "parity unpinned" against
transforms3d itself for those functions only; DESIGN.md repeats this.
"""
