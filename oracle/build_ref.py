"""Recipe for oracle/_ref: the reference's own implementation of the hot path, as a build product.

    python oracle/build_ref.py          (authoring container only: needs /root/reference)

The reference is pure Python (train_test_code/unet.py and the modules it and the training step import), so
"building" it means byte-compiling those files, from where they lie under /root/reference, into ONE archive of
code objects, oracle/_ref/reference_modules.bin (marshal).  No reference source is copied into the repository:
oracle/_ref/ is git-ignored (it still travels to the GPU box with the tree, like the built libfluorounet.so).  Test / bench infrastructure
only: it is the CPU baseline `bench.py --impl reference` times (kind "reference") and a cross-check of the
oracle port; the product never imports it.
"""
import os
import py_compile
import sys

REF = "/root/reference/train_test_code"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
# unet.py:38 imports util; util.py:15 imports dice; dice.py:12 imports ncc; the LR schedule of train.py:337
MODULES = ["unet", "util", "dice", "ncc", "warm_restarts_lr"]


ARCHIVE = os.path.join(OUT, "reference_modules.bin")


def build(force=False):
    """Returns ARCHIVE when the compiled reference is present (building it if the sources are here), else None."""
    if not os.path.isdir(REF):
        return ARCHIVE if os.path.exists(ARCHIVE) else None
    srcs = [os.path.join(REF, m + ".py") for m in MODULES]
    if not force and os.path.exists(ARCHIVE) and all(os.path.getmtime(ARCHIVE) >= os.path.getmtime(s) for s in srcs):
        return ARCHIVE
    import marshal
    os.makedirs(OUT, exist_ok=True)
    blob = {"python": list(sys.version_info[:2]), "modules": {}}
    for m, src in zip(MODULES, srcs):
        with open(src, "rb") as f:
            blob["modules"][m] = compile(f.read(), f"reference:train_test_code/{m}.py", "exec", dont_inherit=True)
    with open(ARCHIVE + ".tmp", "wb") as f:
        marshal.dump(blob, f)
    os.replace(ARCHIVE + ".tmp", ARCHIVE)
    with open(os.path.join(OUT, "BUILD_INFO"), "w") as f:
        f.write(f"byte-compiled from {REF} with python {sys.version.split()[0]}: {' '.join(MODULES)}\n")
    return ARCHIVE


_loaded = None


def load():
    """The compiled reference modules {"unet", "dice", "util", ...} (executed once, registered in sys.modules under
    their reference names so that `import util` inside unet resolves), or None when oracle/_ref was not built or was
    built by another interpreter version: the caller then falls back to the oracle port."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.exists(ARCHIVE):
        return None
    import marshal
    import types
    try:
        with open(ARCHIVE, "rb") as f:
            blob = marshal.load(f)
        if tuple(blob["python"]) != tuple(sys.version_info[:2]):
            return None
        mods = {}
        for m in ("ncc", "dice", "util", "unet", "warm_restarts_lr"):      # dependency order (dice -> ncc, util -> dice, unet -> util)
            if m in sys.modules and getattr(sys.modules[m], "__reference_module__", False) is False:
                return None                       # an unrelated module of that name is already imported: do not shadow it
            mod = types.ModuleType(m)
            mod.__file__ = f"reference:train_test_code/{m}.py"
            mod.__reference_module__ = True
            sys.modules[m] = mod
            exec(blob["modules"][m], mod.__dict__)
            mods[m] = mod
        _loaded = mods
        return mods
    except Exception:
        for m in MODULES:
            if getattr(sys.modules.get(m), "__reference_module__", False):
                del sys.modules[m]
        return None


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
