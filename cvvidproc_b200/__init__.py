"""cvvidproc_b200 -- B200-native drop-in for the hot path of UkoeHB/CvVidProc.

Same public names as the reference's package (PySources/cvvidproc/__init__.py:3), re-exported from this package's own
pybind11 extension `_core` (C++ host layer over the CUDA C ABI, include/cvvp.h):

    import cvvidproc_b200 as cvvidproc
    bg = cvvidproc.GetVideoBackground(cvvidproc.VidBgPack(vid_path, frame_limit=1000, vid_is_grayscale=True))

The extension is built in-tree by `__graft_entry__.build()`.  There is no CPU fallback: without the built CUDA
library the names below raise ImportError on first use.
"""
_NAMES = ("VidBgPack", "GetVideoBackground", "HighlightObjectsPack", "AssignObjectsPack", "VidObjectTrackPack", "TrackObjects")

try:
    from ._core import (  # noqa: F401
        AssignObjectsPack,
        GetVideoBackground,
        HighlightObjectsPack,
        TrackObjects,
        VidBgPack,
        VidObjectTrackPack,
        __doc__ as _core_doc,
    )
except ImportError as _exc:  # not built yet (fresh checkout): fail loudly on use, keep submodules importable
    _import_error = _exc

    def __getattr__(name):
        if name in _NAMES:
            raise ImportError(
                "cvvidproc_b200._core is not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                f"(there is no CPU fallback): {_import_error}"
            )
        raise AttributeError(name)

__all__ = list(_NAMES)
