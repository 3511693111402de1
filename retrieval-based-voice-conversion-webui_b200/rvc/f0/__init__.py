from .gen import Generator, post_process  # noqa: F401
