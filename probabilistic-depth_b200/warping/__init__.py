from . import homography  # noqa: F401
