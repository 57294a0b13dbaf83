"""Per-SM shared-memory fill rate of cp.async.bulk as a function of the copy size (B200)."""
import ctypes as C, sys
sys.path.insert(0, ".")
import torch
from mixgan_tts_b200 import _lib
lib = _lib.load_debug()
src = torch.zeros(32 << 20, dtype=torch.uint8, device="cuda")      # 32 MB: L2-resident after the first pass
st = torch.zeros(1, dtype=torch.int32, device="cuda")
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
print("grid copy_bytes copies/slot slots -> B/clk per SM (mean), chip B/clk")
for grid in (1, 51, 102, 148):
    for copy, per in ((2048, 8), (2048, 24), (4096, 12), (8192, 6), (16384, 3), (49152, 1), (2048, 48), (32768, 3)):
        slot = copy * per
        slots = max(2, min(4, (192 * 1024) // slot))
        iters = max(64, (8 << 20) // slot)
        cyc = torch.zeros(grid, dtype=torch.int64, device="cuda")
        for _ in range(2):
            rc = lib.mgb_probe_bulk_rate(_lib.ptr(src), src.numel(), grid, copy, per, slots, iters, _lib.ptr(cyc), _lib.ptr(st), stream)
            assert rc == 0, lib.mgb_last_error()
            torch.cuda.synchronize()
        c = cyc.double().mean().item()
        print(f"{grid:4d} {copy:6d} {per:3d} {slots:2d} -> {slot * iters / c:7.1f} B/clk/SM  {grid * slot * iters / c:8.0f} chip  status {int(st.item())}")
