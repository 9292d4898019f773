"""Exception types at the provider and framing boundary.

When the reference package is importable its own classes are re-exported, so that
``main.py`` / ``api.py`` catch what this provider raises (src/neuralstego/exceptions.py:8-46);
otherwise equivalent local classes are defined.
"""

from dataclasses import dataclass
from typing import List

try:  # pragma: no cover - depends on the host installation
    from neuralstego.exceptions import (  # type: ignore
        ConfigurationError, FramingError, MissingChunksError, NeuralStegoError, PacketCRCError, PacketECCError)
except Exception:  # reference not installed

    class NeuralStegoError(Exception):
        """Base class for all neural-steganography errors."""

    class ConfigurationError(NeuralStegoError):
        """Raised when user-supplied configuration is invalid."""

    class FramingError(NeuralStegoError):
        """Raised when packet framing or chunk assembly fails."""

    class PacketECCError(FramingError):
        """Raised when ECC decoding fails irrecoverably."""

    class PacketCRCError(FramingError):
        """Raised when CRC verification fails."""

    @dataclass
    class MissingChunksError(FramingError):
        missing_indices: List[int]
        partial_payload: bytes

        def __str__(self) -> str:
            return "Missing chunks at indices: " + ", ".join(str(i) for i in self.missing_indices)


__all__ = ["ConfigurationError", "FramingError", "MissingChunksError", "NeuralStegoError", "PacketCRCError",
           "PacketECCError"]
