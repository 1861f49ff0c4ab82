"""TensorBoard hooks the reference's model forwards call (utils/tb_utils.py there).  They stay
no-ops unless a caller attaches ``global_step`` / ``tb_token_count`` / ``tb_writer`` attributes to
a module -- observability is outside the hot path, the names only need to exist."""


def _tb_advance_global_step(module):
    if hasattr(module, 'global_step'):
        module.global_step += 1
    return module


def _tb_advance_token_counters(module, tensor, verbose=False):
    tc = getattr(module, 'tb_token_count', None)
    if tc is not None:
        T = tensor.size(1)
        if tc.last != T:
            if tc.last != 0:
                tc.total += tc.last
                tc.sample_idx += 1
            tc.last = T
        if verbose:
            print(f'>>> T={T}\tlast_T={tc.last}\tcumsum_T={tc.total}')
    return module


def _tb_hist(module, tensor, name, verbose=False):
    writer = getattr(module, 'tb_writer', None)
    if writer is None:
        return
    if module.layer_idx == module.num_layers - 1:
        tensor = tensor[:, 0]
    layer_s = str(1 + module.layer_idx).zfill(2)
    writer.add_histogram(f'{layer_s}/layer/{name}', tensor, global_step=module.global_step, bins='auto')
    sample_s = str(module.tb_token_count.sample_idx + 1).zfill(2)
    for i in range(tensor.size(1)):
        writer.add_histogram(f'{layer_s}/token/{sample_s}/{name}', tensor[0, i], global_step=i, bins='auto')
