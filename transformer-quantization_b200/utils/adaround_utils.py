"""Model-level AdaRound driver (mirror of the reference's utils/adaround_utils.py)."""
import logging

import torch

from quantization.base_quantized_classes import QuantizedModule
from utils.utils import Stopwatch, pass_data_for_range_estimation

logger = logging.getLogger('AdaRound')
logger.setLevel(logging.INFO)


def get_train_samples(data_loader, num_samples, return_labels=False, inp_idx=0, lbl_idx=1):
    """first ``num_samples`` inputs (and labels) of a loader as one tensor"""
    xs, ys, seen = [], [], 0
    for batch in data_loader:
        xs.append(batch[inp_idx])
        if return_labels:
            ys.append(batch[lbl_idx])
        seen += batch[inp_idx].size(0)
        if seen >= num_samples:
            break
    x = torch.cat(xs, dim=0)[:num_samples]
    if return_labels:
        return x, torch.cat(ys, dim=0)[:num_samples]
    return x


def _weight_layers(model):
    """names of the quantized modules that own a weight, in module order"""
    return [name for name, m in model.named_modules() if isinstance(m, QuantizedModule) and hasattr(m, 'weight')]


def _select_layers(requested, available):
    if 'all' in requested:
        return list(available)
    chosen = [name for name in requested if name in available]
    for name in requested:
        if name not in available:
            logger.warning(f'skipping unknown layer {name}')
    return chosen


def apply_adaround_to_model(config, model, data_loader, range_est_data_loader, batch_size, driver=None,
                            get_samples_fn=get_train_samples, inp_idx=0):
    """AdaRound every selected weight layer of ``model`` in module order (each against the FP32 output of
    that layer, inputs recorded with the preceding layers already rounded when ``asym``), then -- in
    ``post_adaround`` mode -- calibrate and fix the activation quantizers.  ``config.adaround`` /
    ``config.quant`` / ``config.act_quant`` as in the reference (utils/adaround_utils.py:35-146)."""
    # imported here: quantization.adaround itself imports utils.utils (package import cycle otherwise)
    from quantization.adaround import apply_adaround_to_layer
    from quantization.adaround.utils import AdaRoundActQuantMode as ActMode
    ada = config.adaround
    selected = _select_layers(ada.layers, _weight_layers(model))
    if not selected:
        logger.warning('No layers to apply AdaRound for, exiting...')
        return
    if ada.act_quant_mode not in (ActMode.no_act_quant, ActMode.post_adaround):
        raise NotImplementedError(f"act mode '{ada.act_quant_mode}' is not implemented")

    samples = get_samples_fn(data_loader, num_samples=ada.num_samples).to(next(model.parameters()).device)
    config.quant.act_quant = False
    model.reset_act_ranges()
    model.full_precision_acts()

    total = Stopwatch(verbose=False)
    for name, module in model.named_modules():
        if name not in selected:
            continue
        logger.info(f'Started AdaRound for layer {name}')
        model.full_precision()
        module.quantized_weights()
        total.start()
        with Stopwatch(verbose=False) as per_layer:
            apply_adaround_to_layer(model, module, samples, batch_size=batch_size, act_quant=config.quant.act_quant,
                                    adaround_config=ada)
        total.stop()
        logger.info(f'Done AdaRound for layer {name}. {per_layer.format()}\n')
    logger.info(f'Done optimizing all layers. {total.format()}')

    if ada.act_quant_mode == ActMode.post_adaround:
        if driver is not None:                      # accuracy report only
            model.quantized_weights()
            acc = driver.validate().metrics['top_1_accuracy']
            logger.info(f'FINAL res (without acts quant):\t{acc * 100:.2f}%')
        config.quant.act_quant = True
        model.estimate_act_ranges()
        pass_data_for_range_estimation(loader=range_est_data_loader, model=model, act_quant=True, weight_quant=True,
                                       max_num_batches=config.act_quant.num_batches,
                                       cross_entropy_layer=config.act_quant.cross_entropy_layer, inp_idx=inp_idx)
        model.fix_act_ranges()
    model.set_quant_state(weight_quant=True, act_quant=config.quant.act_quant)
