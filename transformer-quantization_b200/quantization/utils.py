"""Host helpers (mirror of the reference's quantization/utils.py)."""
import numpy as np


def to_numpy(tensor):
    """Tensor / array-like -> numpy array.  For CUDA tensors this is a device->host copy and a
    host sync: nothing on the kernel path calls it (the reference does, hijacker.py:81-85 and
    range_estimators.py:254-256)."""
    if isinstance(tensor, np.ndarray):
        return tensor
    if hasattr(tensor, 'detach'):
        tensor = tensor.detach()
        if getattr(tensor, 'is_cuda', False):
            tensor = tensor.cpu()
        return tensor.numpy()
    if hasattr(tensor, 'numpy'):
        return tensor.numpy()
    return np.array(tensor)
