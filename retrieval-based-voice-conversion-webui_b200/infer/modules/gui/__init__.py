from .realtime_tail import RealtimeTail  # noqa: F401
from .torchgate import TorchGate  # noqa: F401
from .resample import Resample  # noqa: F401
from .realtime_block import RealtimeBlock  # noqa: F401
