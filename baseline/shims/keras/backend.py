"""``keras.backend`` stand-in: backend name, base directory attribute, device selection."""
import os

_keras_base_dir = os.path.expanduser("~/.keras")
_EPSILON = 1e-7
_session = None


def backend():
    return "torch"


def epsilon():
    return _EPSILON


def set_session(session):  # TensorFlow-only call sites are guarded by ``K.backend() == 'tensorflow'``
    global _session
    _session = session


def clear_session():
    pass


def device():
    """Where replicas live: ``KERAS_SHIM_DEVICE`` if set, else the (first visible) GPU, else the CPU.
    The Spark shim gives every executor process its own GPU through CUDA_VISIBLE_DEVICES."""
    import torch

    name = os.environ.get("KERAS_SHIM_DEVICE")
    if name:
        return torch.device(name)
    if torch.cuda.is_available():
        # TensorFlow enables TF32 matmuls / convolutions by default on Ampere and later
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        return torch.device("cuda", 0)
    return torch.device("cpu")
