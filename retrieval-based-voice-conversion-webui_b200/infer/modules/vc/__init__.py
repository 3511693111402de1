from .pipeline import Pipeline  # noqa: F401
from .modules import VC  # noqa: F401
