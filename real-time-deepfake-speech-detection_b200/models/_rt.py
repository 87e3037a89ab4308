"""Locate the runtime package whether ``models`` is imported as a top-level package (drop-in:
this directory on sys.path) or as ``real-time-deepfake-speech-detection_b200.models``."""
try:  # package-relative
    from .. import rtdf_runtime as runtime  # type: ignore
except (ImportError, ValueError):  # top-level `models`
    import rtdf_runtime as runtime  # type: ignore

engine_for = runtime.engine_for
Engine = runtime.Engine
native = runtime.native
