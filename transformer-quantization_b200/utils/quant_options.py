"""Quantization options without the command line: the configuration object and ``make_qparams`` of the
reference's utils/quant_click_options.py for programmatic callers (the click decorators there are CLI plumbing).

``quant_config(**overrides)`` builds the ``config.quant`` / ``config.act_quant`` / ``config.qat`` / ``config.adaround``
DotDicts with the reference's option names and defaults (quant_click_options.py:50-108, 134-198, 200-229,
231-352; ``--qmethod`` is a required option there, here it defaults to ``symmetric_uniform``); ``make_qparams(config)`` turns them into the keyword arguments every ``Quantized*`` class of this
package takes (reference :356-380)."""
import copy

from quantization.quantizers import QMethods
from quantization.range_estimators import OptMethod, RangeEstimators
from utils.utils import DotDict

_QUANT_DEFAULTS = dict(qmethod='symmetric_uniform', qmethod_act=None, weight_quant_method='current_minmax',
                       weight_opt_method='grid', num_candidates=None, n_bits=8, n_bits_act=None, per_channel=False,
                       percentile=None, act_quant=True, weight_quant=True, quant_setup='all', quant_dict={})
_ACT_DEFAULTS = dict(act_quant_method='running_minmax', act_opt_method='grid', act_num_candidates=None,
                     act_momentum=None, cross_entropy_layer=None, num_est_batches=1)
_QAT_DEFAULTS = dict(learn_ranges=False, fix_act_ranges=False, fix_weight_ranges=False)


def quant_config(**overrides):
    """config with the reference's defaults; keyword names are the CLI option names with underscores
    (``n_bits=4, qmethod_act='asymmetric_uniform', act_quant_method='MSE', act_num_candidates=50, ...``)"""
    known = set(_QUANT_DEFAULTS) | set(_ACT_DEFAULTS) | set(_QAT_DEFAULTS) | {'adaround'}
    unknown = set(overrides) - known
    if unknown:
        raise TypeError(f'unknown quantization option(s): {sorted(unknown)}')
    pick = lambda defaults: {k: overrides.get(k, copy.deepcopy(v)) for k, v in defaults.items()}     # noqa: E731 (own copy of mutable defaults)
    config = DotDict()
    config.quant = DotDict(pick(_QUANT_DEFAULTS))
    config.quant.qmethod_act = config.quant.qmethod_act or config.quant.qmethod
    a = pick(_ACT_DEFAULTS)
    options = {}
    if a['act_num_candidates'] is not None:
        if a['act_quant_method'] != 'MSE':
            raise ValueError('Wrong option num_candidates passed')
        options['num_candidates'] = a['act_num_candidates']
    if a['act_momentum'] is not None:
        if a['act_quant_method'] != 'running_minmax':
            raise ValueError('Wrong option momentum passed')
        options['momentum'] = a['act_momentum']
    if a['act_opt_method'] != 'grid':
        options['opt_method'] = OptMethod[a['act_opt_method']]
    config.act_quant = DotDict(quant_method=a['act_quant_method'], cross_entropy_layer=a['cross_entropy_layer'],
                               num_batches=a['num_est_batches'], options=options)
    config.qat = DotDict(pick(_QAT_DEFAULTS))
    # imported here: quantization.adaround.utils itself imports utils.utils (package import cycle otherwise)
    from quantization.adaround.utils import DEFAULT_ADAROUND_CONFIG, AdaRoundConfig
    config.adaround = AdaRoundConfig(DEFAULT_ADAROUND_CONFIG)
    config.adaround.update(overrides.get('adaround') or {})
    return config


def make_qparams(config):
    """constructor keywords of the quantized model / layer classes from a config (reference :356-380)"""
    q, act = config.quant, config.act_quant
    weight_range_options = {}
    if q.weight_quant_method in ('MSE', 'cross_entropy'):
        weight_range_options['opt_method'] = OptMethod[q.weight_opt_method]
    if q.num_candidates is not None:
        weight_range_options['num_candidates'] = q.num_candidates
    if q.percentile is not None:
        act.options['percentile'] = q.percentile          # sic: the reference files it under the activation options
    return dict(method=QMethods[q.qmethod], act_method=QMethods[q.qmethod_act], n_bits=q.n_bits,
                n_bits_act=q.n_bits_act, per_channel_weights=q.per_channel, percentile=q.percentile,
                quant_setup=q.quant_setup, weight_range_method=RangeEstimators[q.weight_quant_method],
                weight_range_options=weight_range_options, act_range_method=RangeEstimators[act.quant_method],
                act_range_options=act.options)
