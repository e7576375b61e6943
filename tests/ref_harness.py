"""Import the UNMODIFIED reference from the staged copy ``baseline/_ref`` (tools/stage_reference.py) for the drop-in
tests and bench.py's reference-python leg.  Test infrastructure: nothing in the package imports this.

The reference's scripts import packages this image does not have (matplotlib, gradio, pycocotools, seaborn —
SURVEY.md §8c); ``stub_missing_modules()`` puts inert stand-ins into ``sys.modules`` so that ``train_custom.py``,
``app_gradio.py`` and ``visualize.py`` import unchanged and their function bodies can be driven with synthetic loaders.
"""
import importlib
import os
import sys
import types
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

_STUBS = ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors", "matplotlib.cm", "gradio", "pycocotools",
          "pycocotools.mask", "pycocotools.coco", "seaborn", "wandb", "dotenv")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "src", "custom_maskrcnn.py"))


class _Stub(types.ModuleType):
    """Module whose every attribute is a MagicMock (context managers, calls and chained attributes all work)."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []          # lets `import a.b` resolve through sys.modules
        self._mock = mock.MagicMock(name=name)

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return getattr(self._mock, item)


def stub_missing_modules():
    """Stand-ins for the third-party packages the reference scripts import but this image lacks."""
    made = []
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            if name.split(".")[0] not in ("wandb", "dotenv"):      # present here, but they phone home / read .env: always stub
                importlib.import_module(name)
                continue
        except Exception:
            pass
        sys.modules[name] = _Stub(name)
        made.append(name)
    plt = sys.modules.get("matplotlib.pyplot")
    if isinstance(plt, _Stub):                                      # the few pyplot calls whose RESULTS the scripts use
        import numpy as np

        def subplots(nrows=1, ncols=1, **kw):
            fig = mock.MagicMock(name="figure")
            fig.canvas.renderer.buffer_rgba.return_value = np.zeros((8, 8, 4), np.uint8)
            n = nrows * ncols
            axes = mock.MagicMock(name="axes") if n == 1 else [mock.MagicMock(name=f"axes{i}") for i in range(n)]
            return fig, axes

        plt._mock.subplots.side_effect = subplots
        plt._mock.cm.tab20.side_effect = lambda i: (0.1, 0.2, 0.3, 1.0)
    for name in made:                                               # parent.child attribute access
        if "." in name:
            parent, child = name.rsplit(".", 1)
            if isinstance(sys.modules.get(parent), _Stub):
                setattr(sys.modules[parent], child, sys.modules[name])
    return made


def purge():
    """Forget every imported reference module (and our drop-in patches of them)."""
    for m in [k for k in sys.modules if k == "src" or k.startswith("src.") or k in
              ("custom_maskrcnn", "dataset", "visualize", "train_custom", "app_gradio", "train_transfer", "explain_predictions")]:
        del sys.modules[m]
    for p in (REF, os.path.join(REF, "src")):
        while p in sys.path:
            sys.path.remove(p)


def import_reference(scripts: bool = False):
    """Make ``src.*`` (and, like the reference's scripts do with sys.path.append('src'), ``custom_maskrcnn``) resolve to
    the staged reference.  Returns the ``src.custom_maskrcnn`` module."""
    if not available():
        raise RuntimeError("baseline/_ref is not staged: run tools/stage_reference.py where /root/reference exists")
    purge()
    sys.path.insert(0, REF)
    if scripts:
        stub_missing_modules()
        sys.path.insert(1, os.path.join(REF, "src"))
    return importlib.import_module("src.custom_maskrcnn")
