"""TEST INFRASTRUCTURE ONLY - stages the UNMODIFIED reference for boxes without /root/reference.

The reference is pure Python; its hot path is four files (`irl_control/{osc,robot,device,utils}.py`) plus the YAML /
MJCF data they are configured from.  `/root/reference` exists in the build container only, so `build()` copies those
files verbatim into the git-ignored `oracle/_ref/` (it travels to the GPU box with the snapshot, like the built
`libirlosc.so`; nothing of it enters the history).  `oracle/ref_harness.py` falls back to this copy, so that on the
GPU box `bench.py --impl reference` times the real `OSC.generate` (`kind: "reference"`) and the reference-marked
tests run instead of skipping.

    python oracle/stage_ref.py            # copy (no-op when /root/reference is absent)
    python oracle/stage_ref.py --check    # verify that the staged files are byte-identical to the source
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("IRL_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
# python sources, robot / action-sequence configs, the scene descriptions and one mesh (no other meshes, no images)
PATTERNS = (("irl_control", (".py",)), ("irl_control/examples", (".py",)), ("irl_control/input_devices", (".py",)),
            ("irl_control/robot_configs", (".yaml",)), ("irl_control/action_sequence_configs", (".yaml",)),
            ("irl_control/scenes", (".xml",)),
            ("irl_control/meshes/ur5", ("link0.stl",)))        # the one mesh whose volume enters the dynamics model


def _files():
    for rel, exts in PATTERNS:
        d = os.path.join(SRC, rel)
        if not os.path.isdir(d):
            continue
        for name in sorted(os.listdir(d)):
            if name.endswith(exts) and os.path.isfile(os.path.join(d, name)):
                yield os.path.join(rel, name)


def stage(verbose: bool = False) -> int:
    """Copies the files; returns how many are staged (0 when the reference is not mounted)."""
    if not os.path.isfile(os.path.join(SRC, "irl_control", "osc.py")):
        return 0
    n = 0
    for rel in _files():
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            if verbose:
                print("[stage_ref]", rel)
        n += 1
    with open(os.path.join(DST, "STAGED_FROM"), "w") as fh:
        fh.write("verbatim copies of %d files of %s (ir-lab/irl_control); made by oracle/stage_ref.py, git-ignored\n" % (n, SRC))
    return n


def check() -> bool:
    ok = True
    for rel in _files():
        dst = os.path.join(DST, rel)
        same = os.path.isfile(dst) and filecmp.cmp(os.path.join(SRC, rel), dst, shallow=False)
        ok = ok and same
        if not same:
            print("[stage_ref] differs or missing:", rel)
    return ok


if __name__ == "__main__":
    if "--check" in sys.argv:
        sys.exit(0 if check() else 1)
    print("[stage_ref] %d files staged in %s" % (stage(verbose=True), DST))
