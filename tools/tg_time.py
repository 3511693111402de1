import os, sys, time, torch
sys.path.insert(0, "retrieval-based-voice-conversion-webui_b200")
from infer.modules.gui import TorchGate
tg = TorchGate(sr=48000, n_fft=1920, prop_decrease=0.9).to("cuda:0")
xn = torch.randn(1, 130560, device="cuda") * 0.1
x = xn[:, -9600:].contiguous()
for _ in range(3): tg(x, xn)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): tg(x, xn)
b.record(); torch.cuda.synchronize()
print(f"TorchGate apply (9600 samples vs 130560-sample noise reference): {a.elapsed_time(b)/20*1e3:.1f} us eager (RVCB_TG_FP32={os.environ.get('RVCB_TG_FP32','')})")
