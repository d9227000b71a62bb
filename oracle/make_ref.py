"""Build the reference's own hot-path modules for the CPU arm on the GPU box:   python -m oracle.make_ref

/root/reference is a Python code base and exists only in the build container.  Like a C reference that is compiled from its
sources where they lie into `oracle/_ref/*.so`, this recipe BYTE-COMPILES the four files `oracle/ref_shim.py` executes —
model/utils.py, model/modules.py, model/backbones/dit.py, model/cfm.py — with `py_compile`, straight from /root/reference into
the git-ignored (NOT gpurun-ignored) `oracle/_ref/f5_tts/...*.code` (CPython code objects, the .pyc format under a
neutral extension: snapshot tools commonly drop *.pyc).  No reference source text is copied anywhere; the compiled
files never enter the repository history and the product package never reads them.  With them present, `bench.py --impl
reference` and `cpu_baseline` on the GPU box (same image, same CPython) time the REFERENCE ITSELF (`kind: "reference"`)
instead of the oracle port.  Run by `__graft_entry__.build()` whenever /root/reference is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("F5_REFERENCE_ROOT", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")
FILES = ("model/utils.py", "model/modules.py", "model/backbones/dit.py", "model/cfm.py")


def staged() -> bool:
    return all(os.path.isfile(os.path.join(DST_ROOT, "f5_tts", rel[:-3] + ".code")) for rel in FILES)


def stage(quiet: bool = False) -> bool:
    src_pkg = os.path.join(SRC_ROOT, "src", "f5_tts")
    if not os.path.isfile(os.path.join(src_pkg, FILES[-1])):
        if not quiet:
            print(f"make_ref: {SRC_ROOT} not present; keeping whatever is built in {DST_ROOT}")
        return staged()
    digests = {}
    for rel in FILES:
        src, dst = os.path.join(src_pkg, rel), os.path.join(DST_ROOT, "f5_tts", rel[:-3] + ".code")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile=f"<reference>/src/f5_tts/{rel}", doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        digests[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(DST_ROOT, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC_ROOT, "python": sys.version.split()[0], "source_sha256": digests}, f, indent=1)
    if not quiet:
        print(f"make_ref: byte-compiled {len(FILES)} reference modules into {DST_ROOT}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
