"""AdaRound (learned weight rounding) on top of the sm_100a quantizer kernels: package surface of the
reference's ``quantization.adaround``."""
from quantization.adaround import utils as _options
from quantization.adaround.adaround import apply_adaround_to_layer

__all__ = ['apply_adaround_to_layer']
for _name in ('AdaRoundInitMode', 'AdaRoundMode', 'AdaRoundActQuantMode', 'AdaRoundLossType', 'AdaRoundTempDecayType'):
    globals()[_name] = getattr(_options, _name)
    __all__.append(_name)
