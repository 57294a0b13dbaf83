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
