"""Stage the UNMODIFIED reference sources of the hot path into the git-ignored `baseline/_ref/` so that the reference itself
(not a restatement) can be timed on the GPU box, where /root/reference does not exist (`bench.py --impl reference`,
`cpu_baseline.kind: "reference"`, tests/test_gpu_eager_baseline.py).  TEST / BENCH INFRASTRUCTURE, never imported by
lft_b200.

    python oracle/stage_reference.py        (run by __graft_entry__.build() whenever /root/reference is present)

The three files are copied byte for byte (their sha256 is recorded next to them); nothing under baseline/_ref/ is
tracked by git (.gitignore), it only travels with the gpurun snapshot like the built .so files."""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("LFT_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["model/LFT.py", "utils/utils.py", "option.py"]   # LFT.py:1-283, utils.py:91-157 (LFdivide / LFintegrate), option.py (args)


def stage() -> bool:
    if not os.path.isdir(REF):
        return False
    rec = {}
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        rec[f] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump(rec, open(os.path.join(DST, "STAGED.json"), "w"), indent=1)
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged" if ok else f"{REF} not present: nothing staged", DST)
    sys.exit(0)
