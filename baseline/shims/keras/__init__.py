"""Substrate shim: the slice of the Keras 2 API that cerndb/dist-keras touches, on plain PyTorch.

This package exists ONLY so the unmodified reference under ``baseline/_ref`` can run on an image
that has neither Keras nor TensorFlow/Theano (there is no network to install them).  It imports
nothing from ``distkeras_b200``: models are ``torch.nn`` modules driven by ``torch.optim``
optimizers through the stock autograd path (cuBLAS / cuDNN when a GPU is visible), exposed with the
Keras call surface the reference uses -- ``Sequential`` / ``model_from_json`` / ``to_json`` /
``get_weights`` / ``set_weights`` / ``compile`` / ``train_on_batch`` / ``predict`` and
``optimizers.serialize`` / ``deserialize``.
"""
from . import backend, layers, models, optimizers  # noqa: F401

__version__ = "2.0.8+torchshim"
