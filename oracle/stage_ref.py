"""Stage the UNMODIFIED reference where the GPU box can see it: /root/reference -> oracle/_ref/ (git-ignored).

TEST / BASELINE INFRASTRUCTURE ONLY.  `/root/reference` exists in the build container but not on the GPU box;
`gpurun` ships the working tree (including git-ignored paths), so a byte-for-byte copy of the files the hot path's
callers import travels with it.  Nothing is edited, and nothing under oracle/_ref enters the git history.  Only
`tests/`, `__graft_entry__` and the reference legs of `bench.py` (`--impl reference`, `cpu_baseline`,
`reference_cuda`, `producer`) load it, through `oracle/ref_driver.py`.

    python oracle/stage_ref.py            # copy (idempotent), print what was staged
    python oracle/stage_ref.py --check    # exit 1 unless the staged files equal the source byte for byte

Staged: PredictAndGenerate.py (SbsProcessor, inference_worker, nibba_woka, main_func), SupportFunction.py,
Combine_Clips.py, Check_Clips.py and depth_anything_v2/ (the depth producer's model code; no checkpoints exist).
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("VRSBS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["PredictAndGenerate.py", "SupportFunction.py", "Combine_Clips.py", "Check_Clips.py", "LICENSE.MD"]
TREES = ["depth_anything_v2"]


def _pairs():
    for f in FILES:
        yield os.path.join(SRC, f), os.path.join(DST, f)
    for t in TREES:
        for root, _dirs, names in os.walk(os.path.join(SRC, t)):
            for n in names:
                if n.endswith((".pyc", ".pth")):
                    continue
                s = os.path.join(root, n)
                yield s, os.path.join(DST, os.path.relpath(s, SRC))


def source_available():
    return os.path.isfile(os.path.join(SRC, "PredictAndGenerate.py"))


def staged():
    return os.path.isfile(os.path.join(DST, "PredictAndGenerate.py"))


def stage(verbose=False):
    """Copy the reference files (only those that differ).  Returns the number of files staged."""
    if not source_available():
        raise RuntimeError(f"reference not present at {SRC}")
    n = 0
    for s, d in _pairs():
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
        n += 1
        if verbose:
            print(f"  {os.path.relpath(d, HERE)}")
    return n


def check():
    return source_available() and all(os.path.exists(d) and filecmp.cmp(s, d, shallow=False) for s, d in _pairs())


if __name__ == "__main__":
    if "--check" in sys.argv:
        ok = check()
        print("oracle/_ref matches the reference" if ok else "oracle/_ref is missing or differs from the reference")
        sys.exit(0 if ok else 1)
    print(f"staged {stage(verbose=True)} files from {SRC} into {DST}")
