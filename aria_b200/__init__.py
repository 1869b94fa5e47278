"""aria_b200 — B200-native (sm_100a) implementation of the Aria MoE-transformer hot path behind a C ABI.

Layout:
  csrc/            hand-written CUDA (tcgen05 / TMA / TMEM) + the extern "C" entry points (libaria_b200.so)
  _lib.py, ops.py  ctypes binding and tensor-level wrappers (no CPU fallback)
  moe_lm.py, vision_encoder.py, projector.py, modeling_aria.py
                   host-side mirrors of the reference's module interface (same names / HF weight layout)
  install.py       drop-in installer for an importable reference (`aria.model.*`)
"""
__all__ = ["ops", "moe_lm", "vision_encoder", "projector", "modeling_aria"]
