"""Per-embedding / per-embedding-group (PEG) and mixed-precision switches for existing quantizer
sites.  Mirror of the reference's utils/per_embd_quant_utils.py (same function names and
``quant_dict`` value grammar): an int sets n_bits, 'fp32' disables the site, 'per_embd' quantizes
per hidden dim, 'ngK' uses K contiguous groups of hidden dims, 'ngpK' K groups after sorting the
dims by their dynamic range.

With these settings the quantizer parameters become [d]-vectors with K distinct values; the QDQ
kernel stages the resolved per-dim table in shared memory and indexes it by hidden dim
(csrc/tq_qdq.cu, qdq_cols_vec_kernel), the statistics come from tq_minmax_axis_f32 +
tq_group_minmax_f32.
"""
from quantization.base_quantized_classes import FP32Acts


def set_act_quant_axis_and_groups(module, axis, n_groups, permute=False):
    """Turn a per-tensor activation quantizer into a per-axis / per-group one (reference :54-68).
    With ``permute`` the next FP32 pass only collects per-dim ranges for the group permutation."""
    mgr = module.activation_quantizer if hasattr(module, 'activation_quantizer') else module
    for owner in (mgr, mgr.quantizer, mgr.range_estimator):
        owner.axis = axis
    for owner in (mgr, mgr.range_estimator):
        owner.n_groups = n_groups
    if permute:
        mgr.range_estimator.per_group_range_estimation = True
    return mgr


def _parse_groups(value):
    """'per_embd' | 'ng<K>' | 'ngp<K>'  ->  (n_groups | None, permute) or None if ``value`` is not a PEG spec"""
    if value == 'per_embd':
        return None, False
    for prefix, permute in (('ngp', True), ('ng', False)):          # 'ngp' first: 'ng' is its prefix
        if value.startswith(prefix):
            return int(value[len(prefix):]), permute
    return None


def _hijack(module, attr, value, allow_groups):
    """apply one ``quant_dict`` value to ``module.<attr>`` (the activation or the weight QuantizationManager)"""
    if value is None:
        return
    if isinstance(value, int):
        getattr(module, attr).quantizer.n_bits = value
        return
    if value == 'fp32':
        setattr(module, attr, FP32Acts())
        return
    groups = _parse_groups(value) if allow_groups else None
    if groups is None:
        raise NotImplementedError(f'Unknown value "{value}" in quant_dict')
    set_act_quant_axis_and_groups(module, axis=2, n_groups=groups[0], permute=groups[1])


def _hijack_act_quant(module, value):
    _hijack(module, 'activation_quantizer', value, allow_groups=True)


def _hijack_weight_quant(module, value):
    _hijack(module, 'weight_quantizer', value, allow_groups=False)


def hijack_act_quant(quant_dict, name, m):
    _hijack_act_quant(m, quant_dict.get(name, None))


def hijack_weight_quant(quant_dict, name, m):
    _hijack_weight_quant(m, quant_dict.get(name, None))


def hijack_act_quant_modules(quant_dict, name, m):
    """the same value for every module below ``m`` that owns an activation quantizer"""
    value = quant_dict.get(name, None)
    for sub in m.modules():
        if hasattr(sub, 'activation_quantizer'):
            _hijack_act_quant(sub, value)
