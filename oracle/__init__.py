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
(``tests/test_oracle_pins.py``).  Quantities for which the reference holds no
numeric pin (Kokoro / codec waveforms, Whisper activations) are "parity
unpinned by the reference; pinned by this oracle" -- see DESIGN.md.
"""
