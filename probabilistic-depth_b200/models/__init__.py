from . import correlation_native  # noqa: F401
from . import correlation_package  # noqa: F401
