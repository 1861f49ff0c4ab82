"""Where the UNMODIFIED reference lives (test / bench infrastructure only).

Order: ``$TQ_REFERENCE``; the read-only checkout of the build container (``/root/reference``); the copy
``tools/install_reference.sh`` installed under the git-ignored ``baseline/_ref/`` -- that one travels to the
GPU box with the snapshot, so the reference's own model files can run there on the CUDA back-end."""
import os

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    env = os.environ.get('TQ_REFERENCE')
    if env:
        return env
    for cand in ('/root/reference', os.path.join(_ROOT, 'baseline', '_ref')):
        if os.path.isdir(os.path.join(cand, 'models')) and os.path.isdir(os.path.join(cand, 'quantization')):
            return cand
    return '/root/reference'
