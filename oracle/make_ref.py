#!/usr/bin/env python
"""Stage the UNMODIFIED reference under oracle/_ref/ so that it travels to the GPU box.  TEST INFRASTRUCTURE.

The reference is pure Python (no C/C++/CUDA sources): there is nothing to compile.  What the GPU box
lacks is the checkout itself -- `gpurun` ships only /root/repo.  This recipe copies the reference's
Python sources (and four small .bin sweeps used as realistic loader inputs) byte for byte from
/root/reference into oracle/_ref/, which is git-ignored (never committed) but not gpurun-ignored.
`oracle.ref_loader` then imports the real modules from there, `tests/test_gpu_reference_dropin.py`
runs the reference's own models on the CUDA ops through `b200pc.dropin.install()`, and
`bench.py --impl reference` / `cpu_baseline` time the reference's own torch-CPU code
(kind "reference") instead of the port.

  python oracle/make_ref.py          # idempotent; called by __graft_entry__.build() when /root/reference exists
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("B200PC_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")

# directories whose *.py files are staged (relative to the reference root)
PY_DIRS = ["Utils", "Models", "Dataset", "PointINet20230424/models", "PointINet20230424/data",
           "PolyPCI/Utils", "PolyPCI/Models", "PolyPCI/Dataset"]
# realistic loader inputs: two KITTI sweeps (f32 x4) and two nuScenes sweeps (f32 x5)
BIN_GLOBS = [("PointINet20230424/data/demo_data/original", 2), ("Demos/20230508test/demo_data/Inputs", 2)]


def stage(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("make_ref: %s not present; nothing staged" % SRC)
        return False
    manifest = {}
    for d in PY_DIRS:
        sd = os.path.join(SRC, d)
        if not os.path.isdir(sd):
            continue
        os.makedirs(os.path.join(DST, d), exist_ok=True)
        for f in sorted(os.listdir(sd)):
            if f.endswith(".py") or f.endswith(".txt"):
                shutil.copyfile(os.path.join(sd, f), os.path.join(DST, d, f))
                manifest[os.path.join(d, f)] = hashlib.sha256(open(os.path.join(sd, f), "rb").read()).hexdigest()
    for d, count in BIN_GLOBS:
        sd = os.path.join(SRC, d)
        if not os.path.isdir(sd):
            continue
        bins = sorted(f for f in os.listdir(sd) if f.endswith(".bin"))[:count]
        os.makedirs(os.path.join(DST, d), exist_ok=True)
        for f in bins:
            shutil.copyfile(os.path.join(sd, f), os.path.join(DST, d, f))
            manifest[os.path.join(d, f)] = hashlib.sha256(open(os.path.join(sd, f), "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print("make_ref: staged %d files under %s" % (len(manifest), DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
