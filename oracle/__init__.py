"""CPU oracle for the MixGAN-TTS diffusion-decoder reverse process.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (torch fp32 on the
host, numpy float64 for the schedule constants) of the reference's
``GaussianDiffusion`` sampling loop and ``Denoiser`` network.  It exists to
check the CUDA path; it is never on the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product package ``mixgan_tts_b200`` does not
import it and fails loudly when its CUDA library is missing.

Parity pin: the reference ships no tests, golden vectors or known-answer values
for this path, so the oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF, run in the build container by ``tests/golden/make_golden.py``
(which imports ``/root/reference`` through ``oracle/ref_loader.py``) and committed
as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` replays them anywhere;
``tests/test_oracle_vs_reference.py`` re-checks against the live reference when
``/root/reference`` is present.

Third-party arithmetic: the path's arithmetic lives in PyTorch (``conv1d``,
``linear``, ``tanh``, ``sigmoid``, ``softplus``, ``exp``); the reference does not
pin a torch version (``requirements.txt``).  The goldens were made with the
version recorded inside each ``.npz`` (``torch_version``).
"""
