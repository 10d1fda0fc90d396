#!/usr/bin/env python
"""
TEST / BENCH INFRASTRUCTURE — builds ``oracle/_ref/``: the UNMODIFIED reference's two hot-path
modules, byte-compiled from the sources where they lie under /root/reference.

    /root/reference/qnmfits/qnm.py      ->  oracle/_ref/qnmfits/qnm.code
    /root/reference/qnmfits/qnmfits.py  ->  oracle/_ref/qnmfits/qnmfits.code

(``.code`` = the marshalled code object of the module, what a ``.pyc`` holds behind its
16-byte header; the snapshot that carries the repository to the GPU box drops ``*.pyc``.)

Nothing of the reference is copied into the repository: ``oracle/_ref/`` is a build output
(git-ignored; it travels to the GPU box like the built ``.so`` files), and it holds compiled
code objects only.  ``oracle/ref_loader.py`` executes them as the modules ``qnmfits.qnm`` / ``qnmfits.qnmfits`` when
/root/reference itself is absent, which is how ``bench.py --impl reference`` times the
reference's own functions on the GPU box's host cores, and how the GPU-box tests can compare
against the real thing.  Run by ``__graft_entry__.build()``; a no-op without /root/reference.
"""
import marshal
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("QNMFITS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref", "qnmfits")


def main():
    src_dir = os.path.join(REFERENCE_ROOT, "qnmfits")
    if not os.path.isfile(os.path.join(src_dir, "qnmfits.py")):
        return 0                                     # GPU box: use what was built in the container
    os.makedirs(OUT, exist_ok=True)
    for name in ("qnm", "qnmfits"):
        src = os.path.join(src_dir, name + ".py")
        dst = os.path.join(OUT, name + ".code")
        if not os.path.isfile(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            with open(src, "rb") as fh:
                text = fh.read()
            with warnings.catch_warnings():      # the plotting code has invalid escape sequences
                warnings.simplefilter("ignore")
                code = compile(text, src, "exec", dont_inherit=True)
            with open(dst, "wb") as fh:
                marshal.dump(code, fh)
    with open(os.path.join(OUT, "BUILT_FROM"), "w") as fh:
        fh.write(f"{src_dir} (python {sys.version.split()[0]}, magic {sys.implementation.cache_tag})\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
