"""Exception types at the provider boundary.

When the reference package is importable its own classes are re-exported, so that
``main.py`` / ``api.py`` catch what this provider raises (src/neuralstego/exceptions.py:8-13);
otherwise equivalent local classes are defined.
"""

try:  # pragma: no cover - depends on the host installation
    from neuralstego.exceptions import ConfigurationError, NeuralStegoError  # type: ignore
except Exception:  # reference not installed

    class NeuralStegoError(Exception):
        """Base class for all neural-steganography errors."""

    class ConfigurationError(NeuralStegoError):
        """Raised when user-supplied configuration is invalid."""


__all__ = ["ConfigurationError", "NeuralStegoError"]
