from quantization.adaround.adaround import apply_adaround_to_layer
from quantization.adaround.utils import (
    AdaRoundInitMode,
    AdaRoundMode,
    AdaRoundActQuantMode,
    AdaRoundLossType,
    AdaRoundTempDecayType,
)
