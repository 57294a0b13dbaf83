"""GPU parity of the stages either side of the diffusion decoder, through the C ABI:
the aux decoder (FastSpeech2 decoder + mel_linear + PostNet) and the HiFi-GAN generator on the tcgen05 engine (fp16 operands,
fp32 accumulation / residual sums / LayerNorm / softmax statistics) against the goldens the REAL reference modules produced
and against the CPU oracle.  Tolerance: the fp32 bar of the north star, 1e-3 relative L2."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

from mixgan_tts_b200 import AuxDecoder, Generator, _lib, configs, synth
from oracle import aux_decoder as oa, hifigan as oh
from helpers import load_golden, rel_l2

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_aux import AUX_CASES, VOC_CASES  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-3          # aux decoder (measured 3e-4)
# HiFi-GAN, ~75 convolutions deep, by operand precision of the tcgen05 GEMMs (measured over the cases of this file):
TOL_VOC = {"mixed": 1e-3,     # default: fp16 operands, hi + lo pairs in conv_pre / transposed convs / conv_post: 2.5e-4 .. 6.5e-4
           "fp16": 2e-3,      # fp16 operands everywhere: 4.6e-4 .. 1.2e-3 (an fp16-rounding emulation of the oracle gives the same)
           "fp16x3": 1e-4}    # hi + lo pairs everywhere, the parity mode: 1.0e-5 .. 1.7e-5


def build_aux(dataset, wseed):
    _, pc, mc, _ = configs.make_configs(dataset, "shallow")
    m = AuxDecoder(pc, mc)
    W = synth.make_auxdec_weights(wseed, {"max_seq_len": mc["max_seq_len"]})
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in W.items()}, strict=True)
    return m.cuda().eval(), W, mc


def aux_status(m, B, T):
    st = C.c_int(-1)
    ws = m._ws[(torch.device("cuda", torch.cuda.current_device()), B, T)]
    torch.cuda.synchronize()
    _lib.check(_lib.load().mgb_auxdec_debug_status(C.byref(m.dims), B, T, _lib.ptr(ws), C.byref(st)), "status")
    return st.value


@pytest.mark.parametrize("name", list(AUX_CASES))
def test_aux_decoder_vs_reference_golden(name):
    dataset, B, T, wseed, iseed = AUX_CASES[name]
    g = load_golden(name)
    m, W, mc = build_aux(dataset, wseed)
    inp = synth.make_auxdec_inputs(iseed, B, T)
    x, pad = torch.from_numpy(inp["x"]).cuda(), torch.from_numpy(inp["pad_mask"]).cuda()
    coarse, dec, mel = m(x, pad, return_intermediate=True)
    assert aux_status(m, B, T) == 0
    assert rel_l2(dec[:, ::7, ::5], g["dec_sample"]) < TOL
    assert rel_l2(mel, g["mel_before"]) < TOL
    assert rel_l2(coarse, g["coarse"]) < TOL
    assert float(dec[pad].abs().max()) == 0.0          # masked_fill rows are exactly zero
    again = m(x, pad)
    assert torch.equal(again, coarse)                  # deterministic


@pytest.mark.parametrize("B,T", [(1, 1), (2, 5), (1, 127), (2, 128), (3, 129), (2, 257), (1, 1100)])
def test_aux_decoder_edge_shapes_vs_oracle(B, T):
    m, W, mc = build_aux("LJSpeech", 50 + T % 7)
    Wt = {k: torch.from_numpy(np.asarray(v)) for k, v in W.items()}
    inp = synth.make_auxdec_inputs(60 + T, B, T, min_len_frac=0.3)
    x, pad = torch.from_numpy(inp["x"]), torch.from_numpy(inp["pad_mask"])
    want = oa.aux_decoder_forward(Wt, x, pad, max_seq_len=mc["max_seq_len"])      # T = 1100 > max_seq_len: recomputed table
    got = m(x.cuda(), pad.cuda(), return_intermediate=True)
    assert aux_status(m, B, T) == 0
    for a, b in zip(got, want):
        assert rel_l2(a, b) < TOL, (B, T)
    # lens= is the same thing as the prefix mask; no mask = every frame valid
    got2 = m(x.cuda(), lens=torch.from_numpy(inp["lens"]).cuda())
    assert torch.equal(got2, got[0])
    full = oa.aux_decoder_forward(Wt, x, torch.zeros_like(pad), max_seq_len=mc["max_seq_len"])[0]
    assert rel_l2(m(x.cuda()), full) < TOL


def test_aux_decoder_shard_equivalence_and_errors():
    m, W, mc = build_aux("LJSpeech", 3)
    inp = synth.make_auxdec_inputs(4, 5, 210)
    x, pad = torch.from_numpy(inp["x"]).cuda(), torch.from_numpy(inp["pad_mask"]).cuda()
    full = m(x, pad)
    parts = torch.cat([m(x[:2], pad[:2]), m(x[2:], pad[2:])])
    assert torch.equal(parts, full)                    # utterances never interact: shards reproduce the batch bit for bit
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x.cpu(), pad.cpu())
    with pytest.raises(ValueError):
        m(x[:, :, :128], pad)
    m.train()
    with pytest.raises(RuntimeError, match="inference-only"):
        m(x, pad)


def build_voc(wseed, precision="mixed"):
    gen = Generator(synth.HIFIGAN_CFG, precision=precision)
    W = synth.make_hifigan_weights(wseed)
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}, strict=True)
    return gen.cuda().eval(), W


def voc_status(gen, B, T):
    st = C.c_int(-1)
    ws = gen._ws[(torch.device("cuda", torch.cuda.current_device()), B, T)]
    torch.cuda.synchronize()
    _lib.check(_lib.load().mgb_hifigan_debug_status(C.byref(gen.dims), B, T, _lib.ptr(ws), C.byref(st)), "status")
    return st.value


@pytest.mark.parametrize("precision", ["mixed", "fp16", "fp16x3"])
@pytest.mark.parametrize("name", list(VOC_CASES))
def test_hifigan_vs_reference_golden(name, precision):
    B, T, wseed, iseed = VOC_CASES[name]
    g = load_golden(name)
    gen, W = build_voc(wseed, precision)
    mel = torch.from_numpy(synth.make_mel(iseed, B, T)).cuda()
    wav = gen(mel.transpose(1, 2))                     # the reference's [B, n_mel, T] signature
    assert voc_status(gen, B, T) == 0
    assert tuple(wav.shape) == (B, 1, T * 256)
    assert rel_l2(wav.squeeze(1), g["wav"]) < TOL_VOC[precision]
    assert torch.equal(gen.forward_frames(mel), wav.squeeze(1))


@pytest.mark.parametrize("precision", ["mixed", "fp16x3"])
@pytest.mark.parametrize("B,T", [(1, 1), (3, 7), (2, 64), (1, 130)])
def test_hifigan_edge_shapes_vs_oracle(B, T, precision):
    gen, W = build_voc(70 + T % 5, precision)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    mel = torch.from_numpy(synth.make_mel(80 + T, B, T))
    want = oh.generator_forward(Wt, mel.transpose(1, 2), synth.HIFIGAN_CFG).squeeze(1)
    got = gen.forward_frames(mel.cuda())
    assert voc_status(gen, B, T) == 0
    assert rel_l2(got, want) < TOL_VOC[precision], (B, T)


def test_hifigan_shard_equivalence():
    gen, W = build_voc(9)
    mel = torch.from_numpy(synth.make_mel(10, 4, 50)).cuda()
    full = gen.forward_frames(mel)
    parts = torch.cat([gen.forward_frames(mel[:1]), gen.forward_frames(mel[1:])])
    assert torch.equal(parts, full)
