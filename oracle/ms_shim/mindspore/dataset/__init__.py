from . import audio  # noqa: F401
