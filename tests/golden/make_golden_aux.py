"""Golden vectors of the REAL reference's aux decoder (transformer.Models.Decoder + nn.Linear + transformer.Layers.PostNet as
model/mixgantts.py:26-31,139-143 wires them) and HiFi-GAN generator (hifigan.models.Generator, weight norm removed as
utils/model.py:99 does) -> tests/golden/auxdec_*.npz, tests/golden/hifigan_*.npz

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden_aux.py

Weights and inputs are regenerated from seeds by ``mixgan_tts_b200.synth`` (only the reference's OUTPUTS are committed)."""
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

# name: (dataset, B, T, weight seed, input seed)
AUX_CASES = {
    "auxdec_lj_B3_T150": ("LJSpeech", 3, 150, 21, 22),
    "auxdec_lj_B2_T301": ("LJSpeech", 2, 301, 23, 24),      # three query tiles / key blocks, ragged lengths
}
# name: (B, T, weight seed, input seed)
VOC_CASES = {
    "hifigan_B2_T40": (2, 40, 31, 32),
    "hifigan_B1_T33": (1, 33, 33, 34),
}


class _AuxRef(torch.nn.Module):
    """The three sub-modules MixGANTTS owns for the aux decoder, under their reference names."""

    def __init__(self, preprocess_config, model_config):
        super().__init__()
        models = ref_loader.load_module("transformer.Models")
        layers = ref_loader.load_module("transformer.Layers")
        self.decoder = models.Decoder(model_config)
        self.mel_linear = torch.nn.Linear(model_config["transformer"]["decoder_hidden"],
                                          preprocess_config["preprocessing"]["mel"]["n_mel_channels"])
        self.postnet = layers.PostNet()

    def forward(self, output, mel_masks):          # model/mixgantts.py:139-143
        dec = self.decoder(output, mel_masks)
        mel = self.mel_linear(dec)
        return self.postnet(mel) + mel, dec, mel


def build_aux_reference(dataset, wseed):
    _, pc, mc, _ = configs.make_configs(dataset, "shallow")
    ref = _AuxRef(pc, mc)
    W = synth.make_auxdec_weights(wseed, {"max_seq_len": mc["max_seq_len"]})
    ref.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in W.items()}, strict=True)
    return ref.eval(), W, mc


def build_voc_reference(wseed):
    hm = ref_loader.load_module("hifigan.models")
    hp = ref_loader.load_module("hifigan")
    gen = hm.Generator(hp.AttrDict(synth.HIFIGAN_CFG))
    gen.eval()
    gen.remove_weight_norm()
    W = synth.make_hifigan_weights(wseed)
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}, strict=True)
    return gen, W


def main():
    torch.set_num_threads(4)
    for name, (dataset, B, T, wseed, iseed) in AUX_CASES.items():
        ref, W, mc = build_aux_reference(dataset, wseed)
        inp = synth.make_auxdec_inputs(iseed, B, T)
        with torch.no_grad():
            coarse, dec, mel = ref(torch.from_numpy(inp["x"]), torch.from_numpy(inp["pad_mask"]))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), coarse=coarse.numpy(), dec_sample=dec.numpy()[:, ::7, ::5],
                            mel_before=mel.numpy(), weights_sha256=synth.weights_digest(W), torch_version=torch.__version__)
        print(name, "coarse", tuple(coarse.shape), float(coarse.abs().mean()), "dec rms", float(dec.pow(2).mean().sqrt()))
    for name, (B, T, wseed, iseed) in VOC_CASES.items():
        gen, W = build_voc_reference(wseed)
        mel = synth.make_mel(iseed, B, T)
        with torch.no_grad():
            wav = gen(torch.from_numpy(mel).transpose(1, 2)).squeeze(1)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), wav=wav.numpy(), weights_sha256=synth.weights_digest(W),
                            torch_version=torch.__version__)
        print(name, "wav", tuple(wav.shape), "rms", float(wav.pow(2).mean().sqrt()), "max", float(wav.abs().max()))


if __name__ == "__main__":
    main()
