"""Golden vectors of the REAL reference's LengthRegulator / duration rounding / mask helper -> tests/golden/length_regulator.npz

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden_lr.py

Calls ``model.linguistic_encoder.LengthRegulator`` (linguistic_encoder.py:383-416, through ``utils.tools.pad`` :374-392),
``utils.tools.get_mask_from_lengths`` (:144-153) and the duration expression of ``LinguisticEncoder.forward``
(linguistic_encoder.py:310-314, restated verbatim below because it is inline in a method that needs the whole encoder).
Inputs are regenerated from seeds by ``lr_case`` so that only the reference's OUTPUTS are committed.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name: (B, S, D, max_len, seed, float_durations)
LR_CASES = {
    "small": (3, 7, 5, None, 1, False),
    "wide_padded": (4, 33, 256, 400, 2, False),
    "single": (1, 1, 8, 10, 3, False),
    "cropped": (3, 12, 16, 20, 4, False),          # max_len shorter than the longest utterance: pad() crops
    "float_dur": (2, 9, 4, None, 5, True),         # expand() truncates a float duration with int()
    "all_zero_row": (2, 6, 4, None, 6, False),     # one utterance expands to zero frames
}


def lr_case(name):
    """(x [B,S,D] float32, dur [B,S] int64 or float32, max_len) of one case, from its seed."""
    B, S, D, max_len, seed, fl = LR_CASES[name]
    g = np.random.Generator(np.random.PCG64(1000 + seed))
    x = g.standard_normal((B, S, D), dtype=np.float32)
    if fl:
        dur = (g.random((B, S), dtype=np.float32) * 7.0 - 1.5).astype(np.float32)
    else:
        dur = g.integers(-2, 9, (B, S)).astype(np.int64)
    if name == "all_zero_row":
        dur[1] = np.minimum(dur[1], 0)
    return x, dur, max_len


def logd_case():
    """log-durations around every rounding boundary k + 0.5 plus random values, and the d_control factors tried."""
    g = np.random.Generator(np.random.PCG64(77))
    ks = np.arange(0, 12, dtype=np.float64)
    near = np.log(ks + 1.5)[:, None] + np.array([-3e-3, -1e-4, 1e-4, 3e-3])[None, :]
    rnd = g.standard_normal(400) * 1.2 + 0.8
    log_d = np.concatenate([near.reshape(-1), rnd, [-20.0, -1.0, 0.0, 5.0]]).astype(np.float32)
    return log_d, (1.0, 1.3, 0.75)


def main():
    le = ref_loader.load_module("model.linguistic_encoder")
    tools = ref_loader.load_module("utils.tools")
    lr = le.LengthRegulator()
    out = {"torch_version": torch.__version__, "numpy_version": np.__version__}
    for name in LR_CASES:
        x, dur, max_len = lr_case(name)
        o, ml = lr(torch.from_numpy(x), torch.from_numpy(dur), max_len)
        out[f"{name}/out"] = o.numpy()
        out[f"{name}/mel_len"] = ml.numpy()
        if int(ml.max()) > 0:
            out[f"{name}/mask"] = tools.get_mask_from_lengths(ml).numpy()
            out[f"{name}/mask_w"] = tools.get_mask_from_lengths(ml, int(o.shape[1])).numpy()
    log_d, controls = logd_case()
    for c in controls:
        ld = torch.from_numpy(log_d)
        d = torch.clamp((torch.round(torch.exp(ld) - 1) * c), min=0).long()      # linguistic_encoder.py:310-314
        out[f"dur_from_log/{c}"] = d.numpy()
    np.savez_compressed(os.path.join(HERE, "length_regulator.npz"), **out)
    print("wrote length_regulator.npz:", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
