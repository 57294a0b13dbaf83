"""Run the library's fp32 Conv1d (forward, backward) at the JCU discriminator's layer shapes; meant to run under
`ncu --cache-control none --metrics gpu__time_duration.sum` (warm-L2 per-kernel durations; scripts/conv1d_ncu_summary.py
folds the CSV).   python scripts/bench_conv1d.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mixgan_tts_b200.discriminator import conv1d_frames

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
# (name, rows T, Cin, Cout, k, stride, act, rowbias)
LAYERS = [("input_proj", 800, 160, 160, 1, 1, 0, False), ("conv0", 800, 160, 64, 3, 1, 1, False), ("conv1", 800, 64, 128, 5, 2, 1, False),
          ("conv2", 400, 128, 512, 5, 2, 1, False), ("cond0", 200, 512, 128, 5, 1, 1, True), ("cond1", 200, 128, 1, 3, 1, 1, False),
          ("mlp0", 1, 256, 1024, 1, 1, 2, False), ("mlp2", 1, 1024, 512, 1, 1, 0, False)]
for name, T, Cin, Cout, k, stride, act, rb in LAYERS:
    x = torch.randn(B, T, Cin, device="cuda", requires_grad=True)
    w = (torch.randn(Cout, Cin, k, device="cuda") / (Cin * k) ** 0.5).requires_grad_(True)
    b = torch.randn(Cout, device="cuda", requires_grad=True)
    r = torch.randn(B, Cin, device="cuda", requires_grad=True) if rb else None
    for _ in range(3):
        y = conv1d_frames(x, w, b, r, stride, act)
        y.backward(torch.ones_like(y))
    torch.cuda.synchronize()
    print(name, "done")
