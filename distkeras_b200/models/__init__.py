"""Model specs (Keras-shaped ``Sequential`` + layers) and the reference architectures."""
from .core import (LAYER_CLASSES, Activation, BatchNormalization, Conv2D, Convolution2D, Dense, Dropout,
                   Flatten, GlobalAveragePooling2D, Layer, MaxPooling2D, Reshape, ResidualBlock, Sequential,
                   apply_deferred, compute_accuracy, compute_loss, load_model, model_from_config, model_from_json,
                   prepare_input)
from .functional import Add, Concatenate, Input, Model
from .zoo import ZOO, cifar10_cnn, higgs_mlp, mnist_convnet, mnist_mlp, resnet18

__all__ = [n for n in dir() if not n.startswith("_")]
