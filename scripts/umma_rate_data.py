"""tcgen05.mma throughput against operand CONTENT and FORMAT (zeros / random, bf16 / fp16) with every SM pair busy for a
long time: the question is whether the tensor pipe of a power-managed B200 is data dependent."""
import sys, torch
sys.path.insert(0, '.')
from mixgan_tts_b200 import _lib
lib = _lib.load_debug()
torch.cuda.set_device(0)

def run(name, fill, fp16, reps, n=128, ksteps=16, grid=74):
    cyc = torch.zeros(2 * grid, dtype=torch.int64, device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b_lbo, b_kadv = (1024, 2048) if n == 128 else (2048, 4096)
    for it in range(3):
        if it == 2: e0.record()
        _lib.check(lib.mgb_probe_umma_rate_data(1, grid, n, ksteps, reps, 2, 2048, 128, 4096, b_lbo, 128, b_kadv, 96 * 1024, fill, fp16,
                                                _lib.ptr(cyc), _lib.ptr(st), None), "rate", lib)
        if it == 2: e1.record()
        torch.cuda.synchronize()
    c = cyc[cyc > 0].double()
    per = c.mean().item() / (reps * ksteps)
    ms = e0.elapsed_time(e1)
    tf = 2.0 * 256 * n * 16 * reps * ksteps * grid / (ms * 1e-3) / 1e12
    print(f"{name:34s} N={n}: {per:6.1f} cyc/MMA (floor {n // 2}), {ms:8.2f} ms wall, {tf:7.1f} TFLOP/s, implied SM clock {c.mean().item() / (ms * 1e-3) / 1e6:6.0f} MHz")

for reps in (2048, 65536):
    print(f"-- {reps * 16} MMAs per CTA pair")
    for n in (128, 256):
        run("bf16 zeros", 0, 0, reps, n)
        run("bf16 random", 1, 0, reps, n)
        run("fp16 zeros", 0, 1, reps, n)
        run("fp16 random", 1, 1, reps, n)
