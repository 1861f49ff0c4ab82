"""Quantization-aware-training set-up: calibrate, then choose how every quantizer's range behaves while
the weights train (the reference's utils/qat_utils.py entry point, same name and arguments).

  qat.learn_ranges           ranges become nn.Parameters; their gradients come from tq_qdq_bwd_f32
  otherwise                  ranges keep following the data in train mode (estimate_ranges_train), except
                             the kinds frozen by qat.fix_weight_ranges / qat.fix_act_ranges
"""
import logging

from utils.utils import pass_data_for_range_estimation

logger = logging.getLogger('QAT')
logger.setLevel('INFO')


def prepare_model_for_quantization(config, model, loader):
    quant, qat, act = config.quant, config.qat, config.act_quant
    pass_data_for_range_estimation(loader=loader, model=model, act_quant=quant.act_quant,
                                   weight_quant=quant.weight_quant, max_num_batches=act.num_batches,
                                   cross_entropy_layer=act.cross_entropy_layer)
    if qat.learn_ranges:
        logger.info('Make quantizers learnable')
        model.learn_ranges()
    else:
        logger.info(f'Fix quantizer ranges to fixW={qat.fix_weight_ranges} and fixA={qat.fix_act_ranges}')
        model.estimate_ranges_train()
        for frozen, freeze in ((qat.fix_weight_ranges, model.fix_weight_ranges),
                               (qat.fix_act_ranges, model.fix_act_ranges)):
            if frozen:
                freeze()
    model.set_quant_state(quant.weight_quant, quant.act_quant)
    return model
