"""CPU oracle: a NumPy / torch-CPU float64 restatement of the reference's hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mlx_audio_b200/`` (the product) may
import this package; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs do, and only as the
checker / the reported CPU baseline.

Every function cites the reference file:line it restates (paths relative to
``/root/reference/mlx_audio``).  The reference is pure Python on Apple MLX
(pinned ``mlx 0.31.2``), which cannot be imported here, so the oracle encodes the
public MLX op semantics and is pinned against every
known-answer vector the reference's own tests hold for this path
(``tests/test_oracle_pins.py``).  Quantities for which the reference's tests
hold no numeric pin (Kokoro / codec waveforms, Whisper activations and decoding,
Qwen3 logits and generated codes) are pinned by OUTPUTS OF THE REFERENCE'S OWN
SOURCE, executed in the build container with NumPy standing in for MLX
(``tests/golden/make_*_golden.py`` + ``numpy_mlx_nn.py`` -> ``tests/golden/*.npz``;
agreement 1e-12 or better, integer results identical) -- see DESIGN.md section 2.
"""
