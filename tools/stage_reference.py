#!/usr/bin/env python
"""Stage the UNMODIFIED reference checkout under the git-ignored ``baseline/_ref/`` so that it travels to the
GPU box with the gpurun snapshot (``/root/reference`` does not exist there).

What is staged: the reference's Python package tree (``src/``) plus ``requirements.txt`` and ``LICENSE``, byte
for byte, and a manifest with the sha256 of every file (``baseline/_ref/MANIFEST.json``).  Nothing under
``baseline/_ref`` is product source: it is the reference arm — the drop-in tests (tests/test_gpu_dropin.py) run the
untouched model from it next to the patched one, and ``bench.py`` times its CPU functions as
``cpu_baseline_reference_python``.  The package itself never imports it.

    python tools/stage_reference.py            # no-op when /root/reference is absent (GPU box)
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("LCR_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
KEEP = ("src", "requirements.txt", "LICENSE")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "src")):
        if verbose:
            print(f"stage_reference: {SRC} not present — keeping {DST} as shipped")
        return os.path.isdir(os.path.join(DST, "src"))
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for item in KEEP:
        s, d = os.path.join(SRC, item), os.path.join(DST, item)
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
            for base, _, files in os.walk(d):
                for fn in sorted(files):
                    p = os.path.join(base, fn)
                    manifest[os.path.relpath(p, DST)] = _sha(p)
        elif os.path.isfile(s):
            shutil.copyfile(s, d)
            manifest[item] = _sha(d)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"stage_reference: {len(manifest)} files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
