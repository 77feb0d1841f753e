from . import img_utils  # noqa: F401
