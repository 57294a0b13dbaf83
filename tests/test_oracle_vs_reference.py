"""Live check of the oracle against the reference itself (build container only;
skipped where ``/root/reference`` does not exist, e.g. on the GPU box)."""
import pytest
import torch

from oracle import ref_loader

from helpers import Case, rel_l2

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


@pytest.mark.parametrize("model,multi,B,T", [("naive", False, 2, 33), ("shallow", True, 1, 150)])
def test_oracle_matches_live_reference(model, multi, B, T):
    c = Case("LJSpeech", model, multi, B, T, wseed=3, iseed=11, layers=4)
    gd = ref_loader.build_reference_diffusion(c.args, c.pc, c.mc, c.tc, c.W)
    K = gd.num_timesteps
    noises = c.t("noises")
    seq = [noises[i] for i in reversed(range(K))]
    if model == "shallow":
        with ref_loader.injected_noise(seq, [c.t("start_noise")]):
            ref = gd(None, c.t("cond"), c.t("spk"), c.t("pad_mask"), coarse_mel=c.t("coarse_mel"))[0]
    else:
        gd.cond, gd.spk_emb = c.t("cond").transpose(1, 2), c.t("spk")
        with ref_loader.injected_noise(seq):
            ref = gd.sampling(noise=c.t("x_T"))[-1] * (~c.t("pad_mask").unsqueeze(-1))
    final, _, _, _ = c.oracle_forward()
    assert rel_l2(final, ref) < 1e-6


def test_reference_import_restores_cuda_visible_devices(monkeypatch):
    import os
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "3,5")
    ref_loader.load()
    assert os.environ["CUDA_VISIBLE_DEVICES"] == "3,5"


def test_multi_speaker_without_embedding_raises_like_reference():
    # reference behaviour: F.linear(None, ...) raises TypeError
    c = Case("LJSpeech", "naive", True, 1, 16, wseed=1, iseed=2, layers=1, timesteps=1)
    from oracle.denoiser import denoiser_forward
    with pytest.raises(TypeError):
        denoiser_forward(c.oracle.W, c.t("x_T"), torch.zeros(1, dtype=torch.long),
                         c.t("cond").transpose(1, 2), None)


@pytest.mark.parametrize("model,multi", [("naive", False), ("shallow", True)])
def test_training_branch_oracle_matches_live_reference(model, multi):
    """model/diffusion.py:201-225 with randint / randn_like / noise_like injected in call order."""
    from mixgan_tts_b200 import synth
    c = Case("LJSpeech", model, multi, 3, 40, wseed=3, iseed=12, layers=3)
    ex = {k: torch.from_numpy(v) for k, v in synth.make_train_extras(77, c.B, c.T, c.K).items()}
    gd = ref_loader.build_reference_diffusion(c.args, c.pc, c.mc, c.tc, c.W)
    with ref_loader.injected_noise(noise_like_seq=[ex["post_noise"]], randn_like_seq=[ex["noise_t"], ex["noise_prev"]],
                                   randint_seq=[ex["t"].clone()]):
        with torch.no_grad():
            ref = gd(ex["mel"], c.t("cond"), c.t("spk"), c.t("pad_mask"), coarse_mel=c.t("coarse_mel"))
    out = c.oracle.forward_training(ex["mel"], c.t("cond"), c.t("spk"), c.t("pad_mask"), t=ex["t"], noise_t=ex["noise_t"],
                                    noise_prev=ex["noise_prev"], post_noise=ex["post_noise"], coarse_mel=c.t("coarse_mel"))
    for a, b in zip(out[:4], ref[:4]):
        assert rel_l2(a, b) < 1e-6
    assert torch.equal(out[4], ref[4])


def test_length_regulator_oracle_matches_live_reference():
    """The real LengthRegulator + pad + get_mask_from_lengths + duration rounding, run here, against the numpy restatement."""
    import numpy as np
    from oracle.length_regulator import durations_from_log, length_regulate, mask_from_lengths
    le = ref_loader.load_module("model.linguistic_encoder")
    tools = ref_loader.load_module("utils.tools")
    lr = le.LengthRegulator()
    g = np.random.default_rng(9)
    for B, S, D, max_len in [(2, 5, 3, None), (3, 17, 8, 40), (2, 9, 4, 7)]:
        x = g.standard_normal((B, S, D)).astype(np.float32)
        dur = g.integers(-2, 7, (B, S)).astype(np.int64)
        ref, ref_len = lr(torch.from_numpy(x), torch.from_numpy(dur), max_len)
        out, ml = length_regulate(x, dur, max_len)
        assert np.array_equal(ml, ref_len.numpy())
        assert np.array_equal(out.view(np.uint32), ref.numpy().view(np.uint32))
        assert np.array_equal(mask_from_lengths(ml), tools.get_mask_from_lengths(ref_len).numpy())
    log_d = (g.standard_normal(2000) * 1.5 + 0.5).astype(np.float32)
    for c in (1.0, 0.8, 1.7):
        ref = torch.clamp((torch.round(torch.exp(torch.from_numpy(log_d)) - 1) * c), min=0).long().numpy()
        assert np.array_equal(durations_from_log(log_d, c), ref)
    # autograd of the reference expansion = segment sum
    from oracle.length_regulator import length_regulate_backward
    x = torch.randn(2, 6, 4, requires_grad=True)
    dur = torch.tensor([[2, 0, 3, 1, -1, 2], [1, 1, 1, 0, 4, 0]])
    o, _ = lr(x, dur, 6)
    w = torch.randn_like(o)
    (o * w).sum().backward()
    assert np.allclose(length_regulate_backward(w.numpy(), dur.numpy()), x.grad.numpy(), rtol=1e-6, atol=1e-6)
