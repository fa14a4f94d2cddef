"""Host-side mirror of federatedml/secureprotol for the FLASHE path (same class and method names,
argument meaning and error behaviour as the reference), backed by the CUDA library."""
from .aciq import ACIQ  # noqa: F401
from .encrypt import Encrypt  # noqa: F401
from .flashe import FlasheCipher  # noqa: F401
from .quantize import QuantizingClient  # noqa: F401
