"""TEST / BENCH INFRASTRUCTURE — recipe that stages the UNMODIFIED reference decoder under oracle/_ref/.

The reference (wojlin/WEFAX) is four Python files; its file-decoding path needs `wefax.py`, `config.py`,
`progress_bar.py` and `config/config.json`.  This script copies exactly those, byte for byte, from
/root/reference (read-only mount of the build container) into oracle/_ref/, which is git-ignored (it never
enters the history) but travels to the GPU box with the snapshot, like a compiled reference binary would.
There `bench.py --impl reference` and the `cpu_baseline` leg time the reference's own
`Demodulator.process()` on the host cores (oracle/ref_runner.py: matplotlib stubbed, `time.sleep` patched
out and reported separately).  Nothing on the product path imports anything from here.

    python oracle/make_ref.py          # idempotent; prints what it staged
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DST = os.path.join(HERE, "_ref")
FILES = ("wefax.py", "config.py", "progress_bar.py", os.path.join("config", "config.json"))


def stage(src_root: str = "/root/reference") -> dict | None:
    """Copies the reference's decode-path files; returns {file: sha256} or None when the mount is absent."""
    if not os.path.isfile(os.path.join(src_root, "wefax.py")):
        return None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(src_root, rel), os.path.join(REF_DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src_root, "sha256": manifest,
                   "note": "unmodified copies made by oracle/make_ref.py; git-ignored"}, fh, indent=1)
    return manifest


if __name__ == "__main__":
    m = stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    if m is None:
        print("reference mount not found: nothing staged")
        sys.exit(1)
    for k, v in m.items():
        print(f"{v[:16]}  oracle/_ref/{k}")
