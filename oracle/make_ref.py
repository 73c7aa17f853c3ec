"""Recipe for ``oracle/_ref/`` (git-ignored, shipped to the GPU box with the snapshot).

The reference is a pure-Python ComfyUI node pack: nothing compiles.  What the GPU box needs is the reference's own code
for the path (``src/nodes``: nodes_vadv.py / nodes_adv.py / models/float/{FMT,FLOAT,generator,styledecoder,encoder}.py /
options/) so that ``bench.py --impl reference`` times the REAL ``FloatSampleMotionSequenceRD_VA.sample_rd_sequence_va`` on the
box's host cores and the PSNR gate decodes with the REAL ``Generator``.  This script copies that tree, unmodified, from
``/root/reference`` into ``oracle/_ref/reference/src/nodes`` - an artefact like a built ``.so``, never committed - and is
called by ``__graft_entry__.build()`` whenever ``/root/reference`` exists.  Missing third-party packages (seconohe, comfy,
timm, torchdiffeq, ...) are stood in for by ``oracle/refshim.py``.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("FLOAT_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref", "reference")


def make_ref(verbose: bool = True) -> bool:
    src_nodes = os.path.join(SRC, "src", "nodes")
    if not os.path.isfile(os.path.join(src_nodes, "models", "float", "FMT.py")):
        if verbose:
            print(f"make_ref: {SRC} not present - keeping whatever oracle/_ref already holds")
        return os.path.isdir(os.path.join(DST, "src", "nodes"))
    dst_nodes = os.path.join(DST, "src", "nodes")
    if os.path.isdir(dst_nodes):
        shutil.rmtree(dst_nodes)
    shutil.copytree(src_nodes, dst_nodes, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in ("LICENSE.md", "pyproject.toml"):
        if os.path.isfile(os.path.join(SRC, f)):
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(dst_nodes))
        print(f"make_ref: copied {n} files of the unmodified reference into {os.path.relpath(dst_nodes, os.path.dirname(HERE))}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
