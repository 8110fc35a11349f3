"""The reference's ``torchcde.misc`` names that sit on the hot path (misc.py:70-126)."""
from .coeffs import forward_fill, validate_input_path

__all__ = ["forward_fill", "validate_input_path"]
