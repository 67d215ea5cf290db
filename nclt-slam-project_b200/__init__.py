"""B200-native landmark matcher + teach-map builder for the nclt-slam-project T&R stack.

Only the hot path of BASELINE.json / SURVEY.md section 8 lives here: csrc/ (sm_100a kernels +
the C ABI of include/nclt_b200.h) and the host-side mirror of the reference's call surface.
Importing the compute modules without the built extension raises; there is no CPU fallback.
"""
from . import synth  # noqa: F401  (pure numpy; importable without the extension)

__all__ = ['synth']
