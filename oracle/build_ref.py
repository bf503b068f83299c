"""Recipe: vendor the UNMODIFIED reference's pure-Python path into oracle/_ref/ (git-ignored, travels with gpurun).

TEST / BENCH INFRASTRUCTURE ONLY.  The reference's weight-only path is pure Python + torch, nothing is compiled: the
"build" is a verbatim copy of the packages the path imports

    neural_compressor/{__init__,version}.py, common/, torch/, transformers/

(about 4 MB; evaluation/, tensorflow/, jax/ are not on the path and are left out) from `/root/reference` into
`oracle/_ref/neural_compressor/`.  `oracle/_ref/` is listed in .gitignore, so no reference source ever enters the
history; it is NOT in .gpurunignore, so the copy travels to the GPU box, where `bench.py --impl reference` and the
`cpu_baseline` leg run the real reference on the host cores (`cpu_baseline.kind = "reference"`).  `oracle/ref_loader.py`
imports it from there when `/root/reference` itself is absent.

Run by `__graft_entry__.build()` whenever `/root/reference` is present (i.e. in the build container).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC_ROOT = os.environ.get("B200WOQ_REFERENCE_ROOT", "/root/reference")
PARTS = ["__init__.py", "version.py", "common", "torch", "transformers"]


def vendored() -> bool:
    return os.path.isfile(os.path.join(DEST, "neural_compressor", "torch", "quantization", "quantize.py"))


def build(force: bool = False) -> str:
    src = os.path.join(SRC_ROOT, "neural_compressor")
    if not os.path.isdir(src):
        if vendored():
            return DEST
        raise RuntimeError(f"{src} not found and oracle/_ref is empty: the reference arm is unavailable")
    stamp = os.path.join(DEST, ".stamp")
    if vendored() and os.path.exists(stamp) and not force:
        return DEST
    dst = os.path.join(DEST, "neural_compressor")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst, exist_ok=True)
    for part in PARTS:
        s, d = os.path.join(src, part), os.path.join(dst, part)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy2(s, d)
    with open(stamp, "w") as f:
        f.write("verbatim copy of /root/reference/neural_compressor/{%s}\n" % ",".join(PARTS))
    return DEST


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
