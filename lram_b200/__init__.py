"""lram_b200 — B200-native recurrent-inference path for LRAM's xLSTM policy.

Python host code over a C-ABI CUDA library (include/xlstm_b200.h, lram_b200/csrc). No CPU fallback.
"""
from .config import XLSTMPolicyConfig, preset  # noqa: F401

__all__ = ["XLSTMPolicyConfig", "preset"]
