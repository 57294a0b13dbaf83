"""Pin the oracle: it must reproduce the golden vectors the REAL reference produced
(tests/golden/make_golden.py), and the reference's schedule buffers bit for bit."""
import numpy as np
import pytest
import torch

from mixgan_tts_b200 import synth
from oracle import schedule
from oracle.denoiser import denoiser_forward

from helpers import GOLDEN_CASES, TRAIN_CASES, TRAIN_KEYS, golden_case, load_golden, rel_l2, train_case

# fp32 on a different host CPU may pick different oneDNN kernels: allow rounding-level drift.
TOL = 2e-5


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_weights_generator_is_stable(name):
    g = load_golden(name)
    c = golden_case(name)
    assert synth.weights_digest(c.W) == str(g["weights_sha256"]), \
        "synthetic weight generator drifted: regenerate tests/golden with make_golden.py"


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_schedule_buffers_bit_exact(name):
    g = load_golden(name)
    c = golden_case(name)
    assert c.oracle.K == int(g["K"])
    for k, v in c.oracle.buf.items():
        ref = g[f"sched_{k}"]
        assert v.numpy().dtype == np.float32
        assert np.array_equal(v.numpy().view(np.uint32), ref.view(np.uint32)), k


def test_schedule_known_answers_survey_table():
    # SURVEY.md §8a table, printed from the reference (LJSpeech naive, K=4, vpsde, beta in [0.1, 40])
    b = schedule.diffusion_buffers(schedule.noise_schedule("vpsde", 4, 0.1, 40, 0.008))
    np.testing.assert_allclose(b["betas"], [0.7196944356, 0.9768468738, 0.9980875850, 0.9998420477], rtol=2e-7)
    np.testing.assert_allclose(b["alphas_cumprod"], [0.2803055644, 6.489953026e-3, 1.241165046e-5, 1.960629881e-9], rtol=2e-7)
    np.testing.assert_allclose(b["posterior_mean_coef1"], [1.0, 0.5205591321, 0.08040717244, 3.522460582e-3], rtol=2e-7)
    np.testing.assert_allclose(b["posterior_mean_coef2"], [0.0, 0.1102251783, 0.04344818369, 0.01256833225], rtol=2e-7)
    np.testing.assert_allclose(b["posterior_log_variance_clipped"],
                               [-46.05170059, -0.3458428085, -8.412964642e-3, -1.703891467e-4], rtol=2e-7)
    k1 = schedule.diffusion_buffers(schedule.noise_schedule("vpsde", 1, 0.1, 40, 0.008))
    assert k1["posterior_mean_coef1"][0] == 1.0 and k1["posterior_mean_coef2"][0] == 0.0
    np.testing.assert_allclose(k1["sqrt_alphas_cumprod"], [4.4279e-5], rtol=1e-4)


@pytest.mark.parametrize("mode", ["linear", "cosine", "vpsde"])
def test_schedule_modes_shapes(mode):
    # linear: np.linspace(1e-4, max_beta, K) (utils/tools.py:432) — max_beta is the last beta itself
    b = schedule.noise_schedule(mode, 8, 0.1, 0.06 if mode == "linear" else 40, 0.008)
    assert b.shape == (8,) and b.dtype == np.float64 and np.all(b > 0) and np.all(b <= 1.0)
    buf = schedule.diffusion_buffers(b)
    assert all(v.dtype == np.float32 and v.shape == (8,) for v in buf.values())
    assert buf["posterior_mean_coef1"][0] == 1.0 and buf["posterior_mean_coef2"][0] == 0.0


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_denoiser_forward_matches_reference_golden(name):
    g = load_golden(name)
    c = golden_case(name)
    out = denoiser_forward(c.oracle.W, c.t("x_T"), torch.from_numpy(g["denoiser_t"]),
                           c.t("cond").transpose(1, 2), c.t("spk"))
    assert out.shape == g["denoiser_out"].shape
    assert rel_l2(out, g["denoiser_out"]) < TOL


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_p_sample_matches_reference_golden(name):
    g = load_golden(name)
    c = golden_case(name)
    K = c.oracle.K
    t = torch.full((c.B,), K - 1, dtype=torch.long)
    out, _ = c.oracle.p_sample(c.t("x_T"), t, c.t("cond").transpose(1, 2), c.t("spk"), c.t("noises")[K - 1])
    assert rel_l2(out, g["p_sample_out"]) < TOL


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_full_inference_matches_reference_golden(name):
    g = load_golden(name)
    c = golden_case(name)
    final, states, x0s, start = c.oracle_forward()
    assert final.shape == g["final_mel"].shape
    assert rel_l2(final, g["final_mel"]) < TOL
    if "state_after_first_step" in g.files:
        assert rel_l2(states[1], g["state_after_first_step"]) < TOL
    # padded frames are zeroed by the final mask (model/diffusion.py:200)
    pad = c.t("pad_mask")
    assert float(final[pad].abs().max()) == 0.0 if pad.any() else True


@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_training_branch_matches_reference_golden(name):
    """model/diffusion.py:201-225 (forward values) with every random draw injected."""
    g = load_golden(name)
    c, ex = train_case(name)
    assert synth.weights_digest(c.W) == str(g["weights_sha256"])
    out = c.oracle.forward_training(ex["mel"], c.t("cond"), c.t("spk"), c.t("pad_mask"), t=ex["t"], noise_t=ex["noise_t"],
                                    noise_prev=ex["noise_prev"], post_noise=ex["post_noise"], coarse_mel=c.t("coarse_mel"))
    assert np.array_equal(out[4].numpy(), g["t"])
    for k, v in zip(TRAIN_KEYS, out):
        assert v.shape == g[k].shape, k
        assert rel_l2(v, g[k]) < TOL, k
    # t = 0 rows: x_{t-1} is the clean (normalised) mel and the posterior sample equals x_start exactly (coef1 = 1, sigma = 0)
    b0 = int(np.nonzero(g["t"] == 0)[0][0])
    valid = ~c.t("pad_mask")[b0]
    assert torch.equal(out[2][b0][valid], c.oracle.norm_spec(ex["mel"])[b0][valid])


@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_oracle_gradients_vs_reference_golden(name):
    """torch autograd through the oracle reproduces the REAL reference's gradients (tests/golden/grad_*.npz)."""
    from helpers import check_grads_against_golden, grad_golden_name, oracle_training_grads
    g = load_golden(grad_golden_name(name))
    c, ex = train_case(name)
    assert synth.weights_digest(c.W) == str(g["weights_sha256"])
    probe = {k: torch.from_numpy(v) for k, v in synth.grad_probe(TRAIN_CASES[name][6] + 2000, c.B, c.T).items()}
    loss, _, grads, gcond, gspk = oracle_training_grads(c, ex, probe)
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * max(1.0, abs(float(g["loss"])))
    check_grads_against_golden(g, grads, gcond, gspk, tol=5e-5)


# ---- LengthRegulator / duration rounding / mask: pinned by outputs of the REAL reference -------------------------------
def test_length_regulator_oracle_matches_reference_golden():
    """tests/golden/length_regulator.npz was made by tests/golden/make_golden_lr.py from model.linguistic_encoder.LengthRegulator,
    utils.tools.pad / get_mask_from_lengths and the rounding expression of linguistic_encoder.py:310-314."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_lr import LR_CASES, logd_case, lr_case
    from oracle.length_regulator import durations_from_log, length_regulate, mask_from_lengths
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "length_regulator.npz"))
    for name in LR_CASES:
        x, dur, max_len = lr_case(name)
        out, ml = length_regulate(x, dur, max_len)
        assert np.array_equal(ml, g[f"{name}/mel_len"]), name
        assert out.shape == g[f"{name}/out"].shape, name
        assert np.array_equal(out.view(np.uint32), g[f"{name}/out"].view(np.uint32)), name
        if f"{name}/mask" in g.files:
            assert np.array_equal(mask_from_lengths(ml), g[f"{name}/mask"]), name
            assert np.array_equal(mask_from_lengths(ml, out.shape[1]), g[f"{name}/mask_w"]), name
    log_d, controls = logd_case()
    for c in controls:
        assert np.array_equal(durations_from_log(log_d, c), g[f"dur_from_log/{c}"]), c
    # the cropped case really crops: some utterance is longer than max_len and keeps its true length
    assert g["cropped/mel_len"].max() > g["cropped/out"].shape[1]


# ---- JCU discriminator: oracle pinned by outputs and gradients of the REAL reference -------------------------------------
def _jcu_oracle_run(name):
    import os
    import sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_jcu import JCU_CASES, WEIGHT_STD
    from mixgan_tts_b200 import configs, synth
    from oracle.discriminator import fm_loss, jcu_forward, lsgan_jcu_loss
    multi, B, T, wseed, iseed = JCU_CASES[name]
    mc = configs.make_configs("LJSpeech", "naive", multi)[2]
    Wn = synth.make_discriminator_weights(wseed, multi_speaker=multi, weight_std=WEIGHT_STD)
    W = {k: torch.from_numpy(v).requires_grad_(True) for k, v in Wn.items()}
    inp = synth.make_discriminator_inputs(iseed, B, T, 4, multi_speaker=multi)
    tt = lambda k: None if inp[k] is None else torch.from_numpy(inp[k])
    preds = tt("x_t_prev_preds").requires_grad_(True)
    run = lambda prev: jcu_forward(W, mc["discriminator"], tt("x_ts"), prev, tt("spk"), tt("t"), multi_speaker=multi)
    fc, fu = run(preds)
    rc, ru = run(tt("x_t_prevs"))
    fcd, fud = run(preds.detach())
    d_loss = lsgan_jcu_loss(rc[-1], ru[-1], 1.0) + lsgan_jcu_loss(fcd[-1], fud[-1], 0.0)
    d_grads = dict(zip(W, torch.autograd.grad(d_loss, list(W.values()), retain_graph=True, allow_unused=True)))
    n_layers = mc["discriminator"]["n_layer"] + mc["discriminator"]["n_cond_layer"]
    g_loss = lsgan_jcu_loss(fc[-1], fu[-1], 1.0) + fm_loss(rc, ru, fc, fu, n_layers)
    g_preds = torch.autograd.grad(g_loss, preds)[0]
    return Wn, (fc, fu, rc, ru), d_loss, d_grads, g_loss, g_preds


def test_jcu_discriminator_oracle_matches_reference_golden():
    import os
    import torch
    from mixgan_tts_b200 import synth
    from helpers import rel_l2
    for name in ("jcu_lj_B3_T50", "jcu_spk_B2_T37"):
        g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
        Wn, feats, d_loss, d_grads, g_loss, g_preds = _jcu_oracle_run(name)
        assert synth.weights_digest(Wn) == str(g["weights_sha256"])
        for lst, key in zip(feats, ("fake_cond", "fake_uncond", "real_cond", "real_uncond")):
            for i, f in enumerate(lst):
                assert tuple(f.shape) == g[f"{key}/{i}"].shape
                assert rel_l2(f.detach(), g[f"{key}/{i}"]) < 2e-5, (name, key, i)
        assert abs(float(d_loss) - float(g["d_loss"])) < 2e-5 * abs(float(g["d_loss"]))
        assert abs(float(g_loss) - float(g["g_loss"])) < 2e-5 * abs(float(g["g_loss"]))
        assert rel_l2(g_preds, g["g_grad_preds"]) < 2e-5
        for k, gr in d_grads.items():
            flat = gr.detach().reshape(-1).double()
            assert abs(float(flat.norm()) - float(g[f"gnorm/{k}"])) < 2e-5 * max(float(g[f"gnorm/{k}"]), 1e-6), (name, k)
            idx = torch.from_numpy(synth.grad_sample_index(flat.numel()))
            assert rel_l2(flat[idx], g[f"gsample/{k}"]) < 2e-5, (name, k)
