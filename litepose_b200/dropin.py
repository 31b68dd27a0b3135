"""Run reference scripts (valid.py ...) unchanged on the sm_100a path.

The reference resolves ``models.pose_mobilenet``, ``core.group`` and ``utils.transforms`` through ``sys.path``, and
``valid.py`` itself pushes ``<reference>/lib`` to the FRONT of ``sys.path`` (``import _init_paths``, _init_paths.py:14-23)
- after PYTHONPATH and sitecustomize have been processed.  ``<reference>/lib/models`` is a regular package, so plain path
shadowing loses against it (the model would silently run stock PyTorch eager).  ``install()`` therefore binds the three
top-level names with a meta-path finder, which is consulted before any ``sys.path`` entry whatever its order:

    python -m litepose_b200.dropin /path/to/reference/valid.py --cfg ... --superconfig ...      # runner
    PYTHONPATH=<repo>/litepose_b200/dropin_site:<repo> python valid.py --cfg ...                # sitecustomize

Everything the mirrors do not restate (core.inference, utils.utils, the rest of the model zoo ...) still resolves from
the unmodified reference tree through the mirrors' dynamic ``__path__`` (litepose_b200/lib/_dynpath.py)."""
import importlib.abc
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "litepose_b200", "lib")
MIRRORS = ("models", "core", "utils")


class _MirrorFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        if fullname in MIRRORS:
            pkg = os.path.join(LIB, fullname)
            return importlib.util.spec_from_file_location(fullname, os.path.join(pkg, "__init__.py"),
                                                          submodule_search_locations=[pkg])
        return None


def install():
    """Idempotent.  Raises if one of the names is already bound to something else (import order problem)."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    for name in MIRRORS:
        mod = sys.modules.get(name)
        f = getattr(mod, "__file__", None) or ""
        if mod is not None and not os.path.abspath(f).startswith(LIB):
            raise ImportError("litepose_b200.dropin.install(): %r is already imported from %s - install the drop-in "
                              "before the reference's packages are imported" % (name, f or "a namespace package"))
    if not any(isinstance(f, _MirrorFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _MirrorFinder())


def main(argv=None):
    import runpy
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        sys.stderr.write("usage: python -m litepose_b200.dropin <reference script.py> [its arguments]\n")
        return 2
    install()
    script = os.path.abspath(argv[0])
    sys.argv = [script] + argv[1:]
    sys.path.insert(0, os.path.dirname(script))          # what ``python script.py`` does
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
