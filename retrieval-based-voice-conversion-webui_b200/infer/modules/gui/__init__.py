from .realtime_tail import RealtimeTail  # noqa: F401
