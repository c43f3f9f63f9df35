"""lorads_b200 -- B200-native implementation of the LoRADS per-iteration low-rank kernel layer.

``lorads_b200.capi`` binds the CUDA library (include/lorads_b200.h); ``lorads_b200.sdpa`` builds synthetic SDPA
instances in the reference reader's in-memory format.  The CUDA library is the only compute path.
"""
from . import sdpa  # noqa: F401

__all__ = ["sdpa"]
