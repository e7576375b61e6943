"""livecell-instance-segmentation_b200 — B200-native region pipeline (RPN proposals -> RoIAlign ->
mask paste) of the custom Mask R-CNN in jakubradziejewski/livecell-instance-segmentation.

The directory name carries a hyphen (it mirrors the reference repo's name), so import it through the
alias package ``livecell_instance_segmentation_b200`` at the repo root:

    import livecell_instance_segmentation_b200 as lcr

Sub-modules are imported lazily: ``synth`` needs numpy only; everything else needs torch, and every
compute entry point needs the in-tree CUDA library ``csrc/liblcr.so`` (built by ``build.py`` /
``__graft_entry__.build()``) and a B200 — there is no CPU fallback.
"""
import importlib as _importlib

__version__ = "0.1.0"

_LAZY = {
    "synth", "build", "_lib", "_ext", "ops", "roi_align", "pipeline", "dist", "install", "tv_rpn",
}


def __getattr__(name):
    if name in _LAZY:
        mod = _importlib.import_module(f"{__name__}.{name}")
        globals()[name] = mod
        return mod
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
