#!/usr/bin/env python
"""Recipe for oracle/_ref/ — the reference's OWN Python modules of the hot path, staged so that they can run
where /root/reference does not exist (the GPU box).  TEST / BENCH INFRASTRUCTURE, NOT PRODUCT.

    python oracle/make_ref.py            # build container only (needs /root/reference)

What it does: imports the five modules of the path from /root/reference

    federatedml/secureprotol/jzf_flashe.py  jzf_aes_prp.py  jzf_aes.py  jzf_quantize.py  jzf_aciq.py

under oracle/ref_shim (a ~10-line stand-in for pycryptodome's AES-256-ECB over `cryptography`, because
pycryptodome 3.9.9 is a third-party dependency that is not in this image), asks the interpreter which files
under /root/reference that import actually loaded (their package __init__ files, jzf_twocomplement.py,
arch/api/utils/log_utils.py, federatedml/util/consts.py, ...), and copies exactly those files, unmodified and
with their relative paths, into oracle/_ref/.  A MANIFEST.json records each file's sha256.

oracle/_ref/ is listed in .gitignore (reference sources never enter the history) but not in .gpurunignore,
so the copy travels to the GPU box like a built .so.  Consumers: bench.py (`--impl reference` and the
`cpu_baseline` leg) through oracle/ref_driver.py, and tests that compare the device with the reference itself
when the copy is present.  When /root/reference is absent this script leaves an existing copy alone.
"""
import hashlib
import importlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
OUT = os.path.join(HERE, "_ref")
MODULES = ["federatedml.secureprotol.jzf_flashe", "federatedml.secureprotol.jzf_aes_prp",
           "federatedml.secureprotol.jzf_aes", "federatedml.secureprotol.jzf_quantize",
           "federatedml.secureprotol.jzf_aciq"]


def loaded_reference_files():
    """Import the path's modules from the reference tree in a clean interpreter state and return the
    reference files that were loaded (relative paths)."""
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, os.path.join(HERE, "ref_shim"))
    before = set(sys.modules)
    try:
        for m in MODULES:
            importlib.import_module(m)
        files = set()
        for name in set(sys.modules) - before:
            f = getattr(sys.modules[name], "__file__", None)
            if f and os.path.abspath(f).startswith(REF_ROOT + os.sep):
                files.add(os.path.relpath(os.path.abspath(f), REF_ROOT))
        return sorted(files)
    finally:
        sys.path.remove(REF_ROOT)
        sys.path.remove(os.path.join(HERE, "ref_shim"))
        for name in set(sys.modules) - before:       # do not leave reference modules imported in the caller
            del sys.modules[name]


def build(force=False):
    if not os.path.isdir(REF_ROOT):
        return OUT if os.path.exists(os.path.join(OUT, "MANIFEST.json")) else None
    files = loaded_reference_files()
    manifest = {"source": REF_ROOT, "modules": MODULES, "files": {}}
    for rel in files:
        with open(os.path.join(REF_ROOT, rel), "rb") as fh:
            manifest["files"][rel] = hashlib.sha256(fh.read()).hexdigest()
    mpath = os.path.join(OUT, "MANIFEST.json")
    if not force and os.path.exists(mpath):
        try:
            if json.load(open(mpath))["files"] == manifest["files"] and all(os.path.exists(os.path.join(OUT, r)) for r in files):
                return OUT
        except Exception:
            pass
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    for rel in files:
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_ROOT, rel), dst)
    with open(mpath, "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    return OUT


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    if out is None:
        print("no /root/reference here and no staged copy: nothing to do")
    else:
        m = json.load(open(os.path.join(out, "MANIFEST.json")))
        print("%s: %d reference files" % (out, len(m["files"])))
        for rel in sorted(m["files"]):
            print("  ", rel)
