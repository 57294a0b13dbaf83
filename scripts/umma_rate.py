"""tcgen05.mma issue-rate microbenchmark (no-swizzle K-major operands, as the fused kernel uses them)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
from mixgan_tts_b200 import _lib
lib = _lib.load_debug()
torch.cuda.set_device(0)

def run(name, cta2, grid, n, ksteps, reps, nacc, a_lbo, a_kadv, b_lbo, b_kadv, b_off=96 * 1024):
    cyc = torch.zeros(2 * grid, dtype=torch.int64, device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    for _ in range(2):
        _lib.check(lib.mgb_probe_umma_rate(cta2, grid, n, ksteps, reps, nacc, a_lbo, 128, a_kadv, b_lbo, 128, b_kadv, b_off,
                                           _lib.ptr(cyc), _lib.ptr(st), None), "rate")
        torch.cuda.synchronize()
    c = cyc[cyc > 0].double()
    per = c.mean().item() / (reps * ksteps)
    M = 256 if cta2 else 128
    ideal = (128 * n) / 256.0 if not cta2 else (256 * n) / 512.0
    print(f"{name:52s} grid={grid:3d} M={M} N={n:3d}: {per:7.1f} cyc/MMA (floor {ideal:.0f}) status={int(st)}")

for grid in (1, 74):
    run("1cta N=128 A_LBO=2080 B_LBO=2048", 0, grid, 128, 4, 512, 2, 2080, 4160, 2048, 4096)
    run("1cta N=128 A_LBO=2048 B_LBO=2048", 0, grid, 128, 4, 512, 2, 2048, 4096, 2048, 4096)
    run("1cta N=256 A_LBO=2048 B_LBO=4096", 0, grid, 256, 4, 512, 2, 2048, 4096, 4096, 8192)
    run("2cta N=128 A_LBO=2080 B_LBO=1024 (conv)", 1, grid, 128, 4, 512, 2, 2080, 4160, 1024, 2048)
    run("2cta N=128 A_LBO=2048 B_LBO=1024", 1, grid, 128, 4, 512, 2, 2048, 4096, 1024, 2048)
    run("2cta N=256 A_LBO=2048 B_LBO=2048 (res/skip)", 1, grid, 256, 2, 1024, 2, 2048, 4096, 2048, 4096)
    run("2cta N=256 A_LBO=2080 B_LBO=2048", 1, grid, 256, 2, 1024, 2, 2080, 4160, 2048, 4096)
    run("2cta N=256 same-acc chain", 1, grid, 256, 16, 128, 1, 2048, 4096, 2048, 4096)
    run("2cta N=128 same-acc chain", 1, grid, 128, 16, 128, 1, 2048, 4096, 1024, 2048)
print("-- round 2")
for grid in (74,):
    run("1cta N=128 chain16 nacc=1", 0, grid, 128, 16, 128, 1, 2048, 4096, 2048, 4096)
    run("1cta N=128 k4 nacc=1", 0, grid, 128, 4, 512, 1, 2048, 4096, 2048, 4096)
    run("1cta N=128 k48 nacc=2", 0, grid, 128, 16, 128, 2, 2048, 4096, 2048, 4096)
    run("2cta N=128 k4 nacc=1", 1, grid, 128, 4, 512, 1, 2048, 4096, 1024, 2048)
    run("2cta N=128 k16 nacc=2", 1, grid, 128, 16, 128, 2, 2048, 4096, 1024, 2048)
    run("2cta N=256 k2 nacc=1", 1, grid, 256, 2, 1024, 1, 2048, 4096, 2048, 4096)
    run("2cta N=256 k16 nacc=2", 1, grid, 256, 16, 128, 2, 2048, 4096, 2048, 4096)
    run("1cta N=256 k16 nacc=2", 0, grid, 256, 16, 128, 2, 2048, 4096, 4096, 8192)
