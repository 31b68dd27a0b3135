"""``__path__`` of the mirror packages (``models``, ``core``, ``utils``): this repo's directory first, then every
same-named package directory found on ``sys.path`` AT LOOKUP TIME.  The reference adds ``<reference>/lib`` to
``sys.path`` from inside ``valid.py`` (``import _init_paths``, _init_paths.py:14-23), i.e. after a mirror that was
bound early (sitecustomize / ``python -m litepose_b200.dropin``) has been imported, so a list computed once at import
time would miss it and ``core.inference`` / ``utils.utils`` / the rest of the model zoo could not be resolved."""
import os
import sys


class DynPath(list):
    def __init__(self, here, name, marker=None):
        super().__init__([here])
        self._here, self._name, self._marker = os.path.abspath(here), name, marker

    def _scan(self):
        for p in list(sys.path):
            if not p:
                continue
            cand = os.path.join(p, self._name)
            if cand in self or os.path.abspath(cand) == self._here or not os.path.isdir(cand):
                continue
            if self._marker and not os.path.exists(os.path.join(cand, self._marker)):
                continue
            self.append(cand)

    def __iter__(self):
        self._scan()
        return super().__iter__()

    def __len__(self):
        self._scan()
        return super().__len__()
