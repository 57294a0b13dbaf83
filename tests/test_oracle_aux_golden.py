"""The aux-decoder and HiFi-GAN oracles against the goldens the REAL reference modules produced
(tests/golden/make_golden_aux.py: transformer.Models.Decoder + nn.Linear + transformer.Layers.PostNet; hifigan.models.Generator),
and — when /root/reference is present — against those modules live.  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

from mixgan_tts_b200 import synth
from oracle import aux_decoder as oa, hifigan as oh, ref_loader
from helpers import load_golden, rel_l2

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_aux import AUX_CASES, VOC_CASES  # noqa: E402


@pytest.mark.parametrize("name", list(AUX_CASES))
def test_aux_decoder_oracle_matches_reference_golden(name):
    dataset, B, T, wseed, iseed = AUX_CASES[name]
    g = load_golden(name)
    Wn = synth.make_auxdec_weights(wseed)
    assert synth.weights_digest(Wn) == str(g["weights_sha256"])
    W = {k: torch.from_numpy(np.asarray(v)) for k, v in Wn.items()}
    inp = synth.make_auxdec_inputs(iseed, B, T)
    coarse, dec, mel = oa.aux_decoder_forward(W, torch.from_numpy(inp["x"]), torch.from_numpy(inp["pad_mask"]))
    assert rel_l2(coarse, g["coarse"]) < 2e-5
    assert rel_l2(mel, g["mel_before"]) < 2e-5
    assert rel_l2(dec[:, ::7, ::5], g["dec_sample"]) < 2e-5
    pad = torch.from_numpy(inp["pad_mask"])
    assert float(dec[pad].abs().max()) == 0.0          # masked_fill after every sub-layer (Layers.py:27,30)


@pytest.mark.parametrize("name", list(VOC_CASES))
def test_hifigan_oracle_matches_reference_golden(name):
    B, T, wseed, iseed = VOC_CASES[name]
    g = load_golden(name)
    Wn = synth.make_hifigan_weights(wseed)
    assert synth.weights_digest(Wn) == str(g["weights_sha256"])
    W = {k: torch.from_numpy(v) for k, v in Wn.items()}
    mel = torch.from_numpy(synth.make_mel(iseed, B, T))
    wav = oh.generator_forward(W, mel.transpose(1, 2), synth.HIFIGAN_CFG).squeeze(1)
    assert tuple(wav.shape) == (B, T * 256)
    assert rel_l2(wav, g["wav"]) < 2e-5
    assert float(wav.abs().max()) < 0.95               # the final tanh is not saturated, so it cannot hide errors


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")
def test_aux_and_vocoder_oracles_match_live_reference():
    from make_golden_aux import build_aux_reference, build_voc_reference
    ref, Wn, mc = build_aux_reference("AISHELL3", 41)      # max_seq_len 1500
    W = {k: torch.from_numpy(np.asarray(v)) for k, v in Wn.items()}
    inp = synth.make_auxdec_inputs(42, 2, 70)
    with torch.no_grad():
        want = ref(torch.from_numpy(inp["x"]), torch.from_numpy(inp["pad_mask"]))
    got = oa.aux_decoder_forward(W, torch.from_numpy(inp["x"]), torch.from_numpy(inp["pad_mask"]), max_seq_len=mc["max_seq_len"])
    for a, b in zip(got, want):
        assert rel_l2(a, b) < 2e-6
    gen, Wv = build_voc_reference(43)
    mel = torch.from_numpy(synth.make_mel(44, 1, 21))
    with torch.no_grad():
        want = gen(mel.transpose(1, 2))
    got = oh.generator_forward({k: torch.from_numpy(v) for k, v in Wv.items()}, mel.transpose(1, 2), synth.HIFIGAN_CFG)
    assert rel_l2(got, want) < 2e-6


def test_weight_norm_checkpoint_keys_fold_like_remove_weight_norm():
    """A checkpoint saved with weight norm applied (weight_g / weight_v, what generator_LJSpeech.pth.tar holds) loads into the
    drop-in Generator as g * v / ||v|| — what the reference's remove_weight_norm() leaves (hifigan/models.py:168-175)."""
    from mixgan_tts_b200 import Generator
    W = synth.make_hifigan_weights(5)
    sd = {}
    g0 = np.random.default_rng(0)
    for k, v in W.items():
        if k.endswith(".weight"):
            t = torch.from_numpy(v)
            norm = t.reshape(t.shape[0], -1).norm(dim=1).reshape([-1] + [1] * (t.dim() - 1))
            scale = torch.from_numpy(g0.uniform(0.5, 2.0, size=(t.shape[0],)).astype(np.float32)).reshape(norm.shape)
            sd[k[:-len("weight")] + "weight_v"] = t * scale      # any rescaling of v is undone by the normalisation
            sd[k[:-len("weight")] + "weight_g"] = norm
        else:
            sd[k] = torch.from_numpy(v)
    gen = Generator(synth.HIFIGAN_CFG)
    gen.load_state_dict(sd, strict=True)
    for k, v in W.items():
        got = dict(gen.state_dict())[k]
        assert rel_l2(got, v) < 1e-6, k
    assert gen.hop == 256
