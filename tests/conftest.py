import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    # the fp32 CPU oracle runs tiny convolutions; on many-core hosts the default thread count only adds fork/join overhead
    import torch
    torch.set_num_threads(min(16, os.cpu_count() or 1))
