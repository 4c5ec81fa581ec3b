#!/usr/bin/env python
"""ORACLE build recipe (test infrastructure): stage the UNMODIFIED reference modules of the tile
path under oracle/_ref/ so that the CPU arm of bench.py (`--impl reference`, `cpu_baseline`) can
run the reference itself on the GPU box, where /root/reference does not exist.

    python oracle/make_ref.py            (also run by __graft_entry__.build())

oracle/_ref/ is git-ignored (no reference source enters the history) and NOT gpurun-ignored (it
travels with the snapshot like the built .so files). Files are copied byte for byte; a manifest
with their sha256 is written next to them. Nothing under oracle/_ref is ever imported by the
product path (cerberus_b200/)."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
# the packages infer/tile.py -> models/run_desc.py -> loader/postproc.py import
DIRS = ("infer", "loader", "misc", "models", "run_utils")
EXT = (".py", ".yml")


def main():
    if not os.path.isdir(SRC):
        print("make_ref: %s not present (GPU box): keeping the prebuilt %s" % (SRC, DST))
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for d in DIRS:
        for cur, _, files in os.walk(os.path.join(SRC, d)):
            for f in sorted(files):
                if not f.endswith(EXT):
                    continue
                src = os.path.join(cur, f)
                rel = os.path.relpath(src, SRC)
                dst = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    for f in ("dataset.yml",):
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    print("make_ref: staged %d reference files under %s" % (len(manifest), DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
