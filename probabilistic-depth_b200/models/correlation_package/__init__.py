from . import correlation  # noqa: F401
