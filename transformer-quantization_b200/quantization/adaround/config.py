"""Import location of the AdaRound option container and its defaults for code written against the
reference (``quantization/adaround/config.py`` there); both live in ``quantization.adaround.utils``."""
from quantization.adaround.utils import DEFAULT_ADAROUND_CONFIG, AdaRoundConfig  # noqa: F401

__all__ = ['AdaRoundConfig', 'DEFAULT_ADAROUND_CONFIG']
