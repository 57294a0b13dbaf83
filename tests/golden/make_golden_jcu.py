"""Golden vectors of the REAL reference's JCUDiscriminator (model/mixgantts.py:186-288): forward features and the gradients of
the D-step and G-step losses of train.py:126-184 -> tests/golden/jcu_*.npz

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden_jcu.py

Weights and inputs are regenerated from seeds by ``mixgan_tts_b200.synth`` (only the reference's OUTPUTS are committed).
For each case: every feature map of ``D(x_ts, x_t_prevs, s, t)`` and ``D(x_ts, x_t_prev_preds, s, t)``; the D loss
(``d_loss_fn`` real + fake, model/loss.py:21-24) and its gradient with respect to every discriminator parameter (norm and a
strided sample of each, ``synth.grad_sample_index``); the
generator-side loss (``g_loss_fn`` + ``get_fm_loss``, model/loss.py:26-28,221-227) and its gradient with respect to
``x_t_prev_preds`` (what flows back into the Denoiser)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mixgan_tts_b200 import configs, synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name: (multi_speaker, B, T, weight seed, input seed)
JCU_CASES = {
    "jcu_lj_B3_T50": (False, 3, 50, 11, 12),
    "jcu_spk_B2_T37": (True, 2, 37, 13, 14),       # odd length: the stride-2 layers see 37 -> 19 -> 10 frames
}
WEIGHT_STD = 0.1


def build_reference(multi, wseed):
    mm = ref_loader.load_module("model.mixgantts")
    args, pc, mc, tc = configs.make_configs("LJSpeech", "naive", multi)
    D = mm.JCUDiscriminator(pc, mc, tc)
    W = synth.make_discriminator_weights(wseed, multi_speaker=multi, weight_std=WEIGHT_STD)
    D.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}, strict=True)
    return D, W, mc


def main():
    torch.set_num_threads(1)
    loss_mod = None
    for name, (multi, B, T, wseed, iseed) in JCU_CASES.items():
        D, W, mc = build_reference(multi, wseed)
        inp = synth.make_discriminator_inputs(iseed, B, T, 4, multi_speaker=multi)
        tt = lambda k: None if inp[k] is None else torch.from_numpy(inp[k])
        x_ts, prevs, spk, t = tt("x_ts"), tt("x_t_prevs"), tt("spk"), tt("t")
        preds = tt("x_t_prev_preds").requires_grad_(True)
        out = {"weights_sha256": synth.weights_digest(W), "torch_version": torch.__version__}
        fc, fu = D(x_ts, preds, spk, t)
        rc, ru = D(x_ts, prevs, spk, t)
        for i, f in enumerate(fc):
            out[f"fake_cond/{i}"] = f.detach().numpy()
        for i, f in enumerate(fu):
            out[f"fake_uncond/{i}"] = f.detach().numpy()
        for i, f in enumerate(rc):
            out[f"real_cond/{i}"] = f.detach().numpy()
        for i, f in enumerate(ru):
            out[f"real_uncond/{i}"] = f.detach().numpy()
        import torch.nn.functional as F
        # D step (train.py:139-146; inputs detached there): LSGAN real -> 1, fake -> 0
        jcu = lambda c, u, lab: 0.5 * (F.mse_loss(c, torch.full_like(c, lab)) + F.mse_loss(u, torch.full_like(u, lab)))
        fc_d, fu_d = D(x_ts, preds.detach(), spk, t)
        d_loss = jcu(rc[-1], ru[-1], 1.0) + jcu(fc_d[-1], fu_d[-1], 0.0)
        D.zero_grad()
        d_loss.backward(retain_graph=True)
        out["d_loss"] = np.float64(d_loss.item())
        for k, p in D.named_parameters():       # norm + strided sample of every parameter gradient (the full set is 7.5 MB)
            gflat = p.grad.detach().reshape(-1).double()
            out[f"gnorm/{k}"] = np.float64(gflat.norm().item())
            out[f"gsample/{k}"] = gflat[torch.from_numpy(synth.grad_sample_index(gflat.numel()))].numpy()
        # G step (train.py:159-182): adversarial + feature matching, gradient into x_t_prev_preds
        n_layers = mc["discriminator"]["n_layer"] + mc["discriminator"]["n_cond_layer"]
        fm, wgt = 0, 4.0 / (n_layers + 1)
        for j in range(len(fc) - 1):
            fm = fm + wgt * 0.5 * (F.l1_loss(rc[j].detach(), fc[j]) + F.l1_loss(ru[j].detach(), fu[j]))
        g_loss = jcu(fc[-1], fu[-1], 1.0) + fm
        D.zero_grad()
        g_loss.backward()
        out["g_loss"] = np.float64(g_loss.item())
        out["g_grad_preds"] = preds.grad.detach().numpy().copy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "d_loss", float(d_loss), "g_loss", float(g_loss), "logit rms", float(fc[-1].pow(2).mean().sqrt()))


if __name__ == "__main__":
    main()
