"""Import the UNMODIFIED reference modules from /root/reference/objnerf (CPU only).

TEST INFRASTRUCTURE ONLY.  Nothing under ``openobj_b200/`` may import this file.
It is used (a) by ``oracle/make_golden.py`` to freeze golden vectors into
``tests/golden/`` and (b) by ``tests/test_oracle_vs_reference.py`` to pin the
restatement in ``oracle/openobj_oracle.py`` against the real code whenever
``/root/reference`` is present, and (c) by ``oracle/ref_loop.py`` for bench.py's reference arm; on the GPU box, where
/root/reference does not exist, the modules come from ``oracle/_ref`` (byte-identical files, see oracle/build_ref.py).

The reference imports visualisation / geometry packages that are not installed in
this image and are not on the hot path:
``imgviz, open3d, trimesh, skimage.measure, matplotlib.pyplot, natsort, bidict``.
They are replaced by permissive stubs in ``sys.modules``; the reference files
themselves are never modified or copied.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
# /root/reference in the build container; on the GPU box the byte-identical files oracle/build_ref.py placed in oracle/_ref/
REF_ROOT = os.environ.get("OPENOBJ_REFERENCE") or (
    "/root/reference" if os.path.isfile("/root/reference/objnerf/vmap.py") else os.path.join(_HERE, "_ref"))
REF_OBJNERF = os.path.join(REF_ROOT, "objnerf")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_OBJNERF, "vmap.py"))


class _Anything(types.ModuleType):
    """Module stub that answers any attribute access with another stub."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        sub = _Anything(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return _Anything(self.__name__ + "()")


class _Inverse:
    def __init__(self, owner):
        self._o = owner

    def __setitem__(self, value, key):
        # bidict.inv[value] = key  : re-key the entry that maps to `value`
        for k in [k for k, v in self._o.items() if v == value]:
            dict.__delitem__(self._o, k)
        dict.__setitem__(self._o, key, value)

    def __getitem__(self, value):
        for k, v in self._o.items():
            if v == value:
                return k
        raise KeyError(value)


class _Bidict(dict):
    """10-line stand-in for bidict.bidict (only what vmap.py:65,198,226 uses)."""

    @property
    def inv(self):
        return _Inverse(self)


_modules = None


def load():
    """Return a dict of the reference's hot-path modules (imported once)."""
    global _modules
    if _modules is not None:
        return _modules
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_OBJNERF)
    for name in ("imgviz", "open3d", "trimesh", "skimage", "skimage.measure",
                 "matplotlib", "matplotlib.pyplot", "natsort"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    # sample_points_bbox (trainer.py:148) builds an open3d OrientedBoundingBox only to
    # read .R/.center/.extent back: give the stub a plain record type for it.
    class _OBB:
        def __init__(self, center, R, extent):
            self.center, self.R, self.extent = center, R, extent
    sys.modules["open3d"].geometry.OrientedBoundingBox = _OBB
    if "bidict" not in sys.modules:
        b = types.ModuleType("bidict")
        b.bidict = _Bidict
        sys.modules["bidict"] = b
    # The reference uses bare module names (``import model``).  Import them under
    # a private sys.path entry and remember/restore any name clashes.
    clash = {}
    names = ["model", "embedding", "render_rays", "loss", "utils", "cfg",
             "vis", "trainer", "vmap"]
    for n in names:
        if n in sys.modules:
            clash[n] = sys.modules.pop(n)
    sys.path.insert(0, REF_OBJNERF)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import importlib
            mods = {n: importlib.import_module(n) for n in names}
    finally:
        sys.path.remove(REF_OBJNERF)
        for n in names:
            sys.modules.pop(n, None)
        sys.modules.update(clash)
    # the reference modules reference each other through the names bound at
    # import time, so removing them from sys.modules is safe.
    _modules = mods
    return mods


def make_cfg(config_json=None, device="cpu", **overrides):
    """reference cfg.Config with both devices forced to `device`."""
    m = load()
    if config_json is None:
        config_json = os.path.join(REF_OBJNERF, "configs", "Replica", "room_0.json")
    c = m["cfg"].Config(config_json)
    c.training_device = device
    c.data_device = device
    for k, v in overrides.items():
        setattr(c, k, v)
    return c
