"""Host-side helpers under the reference's ``utils`` package name, restricted to what the hot path
and the unchanged ``models/quantized_*.py`` import (the CLI / dataset / HF
glue modules of the reference's ``utils`` are out of scope)."""
from utils.adaround_utils import apply_adaround_to_model
from utils.per_embd_quant_utils import (
    hijack_act_quant,
    hijack_weight_quant,
    hijack_act_quant_modules,
    set_act_quant_axis_and_groups,
)
from utils.qat_utils import prepare_model_for_quantization
from utils.quant_options import make_qparams, quant_config
from utils.tb_utils import _tb_advance_global_step, _tb_advance_token_counters, _tb_hist
from utils.utils import (
    seed_all,
    count_params,
    count_embedding_params,
    pass_data_for_range_estimation,
    DotDict,
    Stopwatch,
    StopForwardException,
)
