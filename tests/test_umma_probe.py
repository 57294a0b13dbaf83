"""Pin the tcgen05 shared-memory descriptor conventions the bf16 path relies on (csrc/tc05.cuh):
K-major / no-swizzle core-matrix layout, LBO = K-direction stride, SBO = 8-row-group stride,
a +16 B start address = one-row shift (the k=3 convolution taps), K advance by descriptor."""
import ctypes as C
import os

import pytest
import torch

from mixgan_tts_b200 import _lib

pytestmark = pytest.mark.gpu


def _image(mat: torch.Tensor) -> torch.Tensor:
    """[rows, K] bf16 -> bytes of the [K/8][rows][8] core-matrix image."""
    rows, K = mat.shape
    return mat.reshape(rows, K // 8, 8).permute(1, 0, 2).contiguous()


def run_probe(A, Bm, *, n, ksteps, shift=0, swap=False, bulk=True, a_rows=None):
    """A [rows,K] bf16 (rows >= 128+shift), Bm [n,K] bf16.  Returns D [128,n] fp32 and status."""
    lib = _lib.load_debug()
    dev = A.device
    rows = A.shape[0]
    a_img, b_img = _image(A), _image(Bm)
    a_lbo, a_sbo = rows * 16, 128
    b_lbo, b_sbo = n * 16, 128
    if swap:
        a_lbo, a_sbo, b_lbo, b_sbo = a_sbo, a_lbo, b_sbo, b_lbo
    D = torch.full((128, n), float("nan"), device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = lib.mgb_probe_umma(
        _lib.ptr(a_img), a_img.numel() * 2, _lib.ptr(b_img), b_img.numel() * 2,
        shift * 16, a_lbo, a_sbo, 2 * rows * 16,
        0, b_lbo, b_sbo, 2 * n * 16,
        n, ksteps, int(bulk), _lib.ptr(D), _lib.ptr(st),
        C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "mgb_probe_umma")
    torch.cuda.synchronize()
    return D, int(st.item())


def _mats(rows, n, K, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = (torch.randn(rows, K, generator=g) * 0.5).bfloat16().cuda()
    Bm = (torch.randn(n, K, generator=g) * 0.5).bfloat16().cuda()
    return A, Bm


def _expect(A, Bm, shift):
    return A[shift:shift + 128].float() @ Bm.float().t()


def _report(line):
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/umma_probe.txt", "a") as f:
        f.write(line + "\n")
    print(line)


@pytest.mark.parametrize("bulk", [True, False])
@pytest.mark.parametrize("n,ksteps", [(128, 1), (128, 4), (256, 4), (64, 2)])
def test_descriptor_convention(n, ksteps, bulk):
    K = 16 * ksteps
    A, Bm = _mats(128, n, K)
    D, st = run_probe(A, Bm, n=n, ksteps=ksteps, bulk=bulk)
    err = (D - _expect(A, Bm, 0)).abs().max().item()
    _report(f"convention n={n} ksteps={ksteps} bulk={bulk}: status={st} err={err:.3e}")
    assert st == 0, "probe timed out"
    assert err < 1e-3, "canonical LBO/SBO convention wrong"


@pytest.mark.parametrize("shift", [1, 2, 7, 8, 9])
def test_row_shift_by_start_address(shift):
    n, ksteps = 128, 4
    A, Bm = _mats(144, n, 16 * ksteps, seed=shift)
    D, st = run_probe(A, Bm, n=n, ksteps=ksteps, shift=shift)
    err = (D - _expect(A, Bm, shift)).abs().max().item()
    err0 = (D - _expect(A, Bm, 0)).abs().max().item()
    _report(f"row shift {shift}: status={st} err={err:.3e} (err vs unshifted {err0:.3e})")
    assert st == 0 and err < 1e-3


@pytest.mark.parametrize("n,ksteps", [(128, 1), (128, 8), (256, 4), (64, 4)])
def test_two_cta_pair_conventions(n, ksteps):
    """cta_group::2: M = 256 split by rows across the CTA pair, B split by rows in halves (n/2 each),
    each CTA reads its accumulator rows from its own TMEM; commit multicast reaches both CTAs."""
    lib = _lib.load_debug()
    K = 16 * ksteps
    g = torch.Generator(device="cpu").manual_seed(n + ksteps)
    A = (torch.randn(256, K, generator=g) * 0.5).bfloat16().cuda()
    Bm = (torch.randn(n, K, generator=g) * 0.5).bfloat16().cuda()
    a_img = torch.stack([_image(A[:128]), _image(A[128:])]).contiguous()
    b_img = torch.stack([_image(Bm[:n // 2]), _image(Bm[n // 2:])]).contiguous()
    D = torch.full((256, n), float("nan"), device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = lib.mgb_probe_umma_2cta(_lib.ptr(a_img), 128 * K * 2, _lib.ptr(b_img), (n // 2) * K * 2,
                                 128 * 16, 128, 2 * 128 * 16, (n // 2) * 16, 128, 2 * (n // 2) * 16,
                                 n, ksteps, _lib.ptr(D), _lib.ptr(st), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "mgb_probe_umma_2cta")
    torch.cuda.synchronize()
    ref = A.float() @ Bm.float().t()
    err = (D - ref).abs().max().item()
    err_top = (D[:128] - ref[:128]).abs().max().item()
    err_bot = (D[128:] - ref[128:]).abs().max().item()
    _report(f"2cta n={n} ksteps={ksteps}: status={int(st.item())} err={err:.3e} (rows 0-127 {err_top:.3e}, rows 128-255 {err_bot:.3e})")
    assert int(st.item()) == 0 and err < 1e-3


@pytest.mark.parametrize("n", [128, 256])
def test_aliased_operand_lbo0_sbo0(n):
    """LBO = SBO = 0 makes every 8-row group and both k-chunks of the A operand read the SAME 128-byte core
    matrix: A[r][k] = block[r % 8][k % 8].  The fused kernel uses this as a 128-byte "ones" operand that adds the
    convolution bias inside the MMA."""
    lib = _lib.load_debug()
    g = torch.Generator(device="cpu").manual_seed(n)
    block = (torch.randn(8, 8, generator=g) * 0.5).bfloat16().cuda()
    Bm = (torch.randn(n, 16, generator=g) * 0.5).bfloat16().cuda()
    a_img = torch.zeros(128 * 16, dtype=torch.bfloat16, device="cuda")     # 4 KB image, only the first 128 B matter
    a_img[:64] = block.reshape(-1)
    b_img = _image(Bm)
    D = torch.full((128, n), float("nan"), device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = lib.mgb_probe_umma(_lib.ptr(a_img), a_img.numel() * 2, _lib.ptr(b_img), b_img.numel() * 2,
                            0, 0, 0, 0, 0, n * 16, 128, 2 * n * 16, n, 1, 1, _lib.ptr(D), _lib.ptr(st),
                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "mgb_probe_umma")
    torch.cuda.synchronize()
    A = block.float().repeat(16, 2)                       # [128, 16]
    err = (D - A @ Bm.float().t()).abs().max().item()
    _report(f"aliased A operand (LBO=SBO=0) n={n}: status={int(st.item())} err={err:.3e}")
    assert int(st.item()) == 0 and err < 1e-3


def _image_mn(mat: torch.Tensor) -> torch.Tensor:
    """[rows(M or N), K] bf16 -> the [rows/8][K][8] image: 8 consecutive ROWS contiguous (16 B), k running at 16 B —
    i.e. the frames-major activation image [C/8][frames][8] seen as an operand whose K axis is the frame axis."""
    rows, K = mat.shape
    return mat.t().reshape(K, rows // 8, 8).permute(1, 0, 2).contiguous()


@pytest.mark.parametrize("which", ["a", "b", "ab"])
@pytest.mark.parametrize("n,ksteps,shift", [(128, 4, 0), (256, 8, 0), (128, 4, 1), (128, 4, 2)])
def test_mn_major_operands(which, n, ksteps, shift):
    """MN-major (no swizzle) operands as the weight-gradient GEMMs of the training path read them: the activation image
    [C/8][frames][8] with the FRAME axis as K.  Convention pinned here: LBO = 128 B (next 8-frame group), SBO = byte
    stride between 8-channel chunks, K advance of one MMA (16 frames) = 256 B, a +16 B start = one frame later."""
    lib = _lib.load_debug()
    K = 16 * ksteps
    Kp = K + 8                                  # room for the shifted read
    g = torch.Generator(device="cpu").manual_seed(n + ksteps + shift)
    A = (torch.randn(128, Kp, generator=g) * 0.5).bfloat16().cuda()
    Bm = (torch.randn(n, Kp, generator=g) * 0.5).bfloat16().cuda()
    a_mn, b_mn = "a" in which, "b" in which
    a_img = _image_mn(A) if a_mn else _image(A[:, :K].contiguous())
    b_img = _image_mn(Bm) if b_mn else _image(Bm[:, :K].contiguous())
    a_desc = (shift * 16, 128, Kp * 16, 256) if a_mn else (0, 128 * 16, 128, 2 * 128 * 16)
    b_desc = (shift * 16, 128, Kp * 16, 256) if b_mn else (0, n * 16, 128, 2 * n * 16)
    D = torch.full((128, n), float("nan"), device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    flags = 1 | (2 if a_mn else 0) | (4 if b_mn else 0)
    rc = lib.mgb_probe_umma(_lib.ptr(a_img), a_img.numel() * 2, _lib.ptr(b_img), b_img.numel() * 2,
                            *a_desc, *b_desc, n, ksteps, flags, _lib.ptr(D), _lib.ptr(st),
                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "mgb_probe_umma")
    torch.cuda.synchronize()
    Ae = A[:, shift:shift + K] if a_mn else A[:, :K]
    Be = Bm[:, shift:shift + K] if b_mn else Bm[:, :K]
    err = (D - Ae.float() @ Be.float().t()).abs().max().item()
    _report(f"MN-major {which} n={n} ksteps={ksteps} shift={shift}: status={int(st.item())} err={err:.3e}")
    assert int(st.item()) == 0 and err < 1e-3


def test_bulk_copy_size_sets_the_shared_memory_fill_rate():
    """The finding behind the tensor-map TMA boxes: one cp.async.bulk costs ~60-110 cycles of TMA-engine time
    whatever its size, so a ring refilled with 2 KB copies fills at a fraction of the rate of one refilled with 32 KB copies."""
    lib = _lib.load_debug()
    src = torch.zeros(32 << 20, dtype=torch.uint8, device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rate = {}
    for copy, per, slots in ((2048, 8, 4), (32768, 3, 2)):
        cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
        iters = (8 << 20) // (copy * per)
        for _ in range(2):
            _lib.check(lib.mgb_probe_bulk_rate(_lib.ptr(src), src.numel(), 1, copy, per, slots, iters, _lib.ptr(cyc), _lib.ptr(st),
                                               stream), "mgb_probe_bulk_rate")
            torch.cuda.synchronize()
        assert int(st.item()) == 0
        rate[copy] = copy * per * iters / float(cyc.item())
    _report(f"bulk copy fill rate, one SM: 2 KB copies {rate[2048]:.1f} B/clk, 32 KB copies {rate[32768]:.1f} B/clk")
    assert rate[32768] > 2.5 * rate[2048]
