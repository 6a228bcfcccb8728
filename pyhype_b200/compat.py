"""Run code written against the reference's import paths unchanged.

The package mirrors pyHype's module layout for everything on the hot path's boundary
(``pyhype.solvers``, ``pyhype.solver_config``, ``pyhype.states[.primitive|.conservative]``,
``pyhype.fluids``, ``pyhype.mesh.base|rectangular``, ``pyhype.initial_conditions.base|supersonic_flood``,
``pyhype.boundary_conditions.base|bc``).  ``install()`` makes ``import pyhype...`` resolve to
``pyhype_b200...`` so that e.g. the reference's ``examples/dmr/dmr.py`` runs on the B200 engine as it is:

    python -m pyhype_b200.compat examples/dmr/dmr.py          # from the root of a pyHype checkout

or, inside a script, ``import pyhype_b200.compat as c; c.install()`` before the first ``import pyhype``.
Modules that are outside the hot path (``pyhype.utils.visualizer`` ...) are not provided and raise
``ModuleNotFoundError`` as usual.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.util
import sys

_TARGET = "pyhype_b200"


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, real_name):
        self.real_name = real_name

    def create_module(self, spec):
        return importlib.import_module(self.real_name)   # the very same module object, under a second name

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def __init__(self, alias):
        self.alias = alias

    def find_spec(self, fullname, path=None, target=None):
        if fullname != self.alias and not fullname.startswith(self.alias + "."):
            return None
        real = _TARGET + fullname[len(self.alias):]
        try:
            real_spec = importlib.util.find_spec(real)
        except ModuleNotFoundError:
            return None
        if real_spec is None:
            return None
        return importlib.util.spec_from_loader(fullname, _AliasLoader(real), is_package=real_spec.submodule_search_locations is not None)


def install(alias: str = "pyhype") -> None:
    """Route ``import <alias>[.x.y]`` to ``pyhype_b200[.x.y]``.  Refuses to shadow a real package of that name that is
    already imported (so it can never silently swap engines under a running program)."""
    if any(isinstance(f, _AliasFinder) and f.alias == alias for f in sys.meta_path):
        return
    mod = sys.modules.get(alias)
    if mod is not None and not getattr(mod, "__name__", "").startswith(_TARGET):
        raise RuntimeError(f"'{alias}' is already imported from {getattr(mod, '__file__', '?')}; install the alias first")
    sys.meta_path.insert(0, _AliasFinder(alias))


def uninstall(alias: str = "pyhype") -> None:
    sys.meta_path[:] = [f for f in sys.meta_path if not (isinstance(f, _AliasFinder) and f.alias == alias)]
    for name in [n for n in sys.modules if n == alias or n.startswith(alias + ".")]:
        del sys.modules[name]


def main(argv=None) -> int:
    import os
    import runpy

    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print("usage: python -m pyhype_b200.compat SCRIPT.py [args...]", file=sys.stderr)
        return 2
    install()
    script = argv[0]
    sys.argv = argv
    # like `python script.py`, plus the working directory (the reference's examples import `examples.<name>.config`
    # relative to the checkout root, makefile:1-23 runs them with `python -m examples...` from there)
    for p in (os.getcwd(), os.path.dirname(os.path.abspath(script))):
        if p not in sys.path:
            sys.path.insert(0, p)
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
