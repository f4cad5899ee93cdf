"""Reference architectures.

The MNIST MLP / ConvNet and the Higgs MLP are the models the reference's examples train
(``examples/mnist_analysis.ipynb:247-252``, ``examples/mnist.py:150-162``,
``examples/example_1_analysis.ipynb:302-305``).  The CIFAR-10 CNN and ResNet-18 are named by
BASELINE.json but defined nowhere in the reference (only a CIFAR-10 preprocessing notebook
exists), so their architectures are specified here.
"""
from __future__ import annotations

from .core import (Activation, BatchNormalization, Conv2D, Dense, Dropout, Flatten, GlobalAveragePooling2D,
                   MaxPooling2D, ResidualBlock, Sequential)


def mnist_mlp(hidden2: int = 200, dropout: bool = True, seed=None) -> Sequential:
    """784 -> 1000 ReLU -> 200 ReLU -> 10 softmax (987,210 parameters)."""
    layers = [Dense(1000, activation="relu", input_shape=(784,))]
    if dropout:
        layers.append(Dropout(0.2))
    layers.append(Dense(hidden2, activation="relu"))
    if dropout:
        layers.append(Dropout(0.2))
    layers.append(Dense(10, activation="softmax"))
    return Sequential(layers, name="mnist_mlp", seed=seed)


def mnist_convnet(seed=None) -> Sequential:
    """28x28x1 -> conv3x3x32 -> conv3x3x32 -> maxpool2 -> 225 -> 10 (1,048,853 parameters)."""
    return Sequential([
        Conv2D(32, 3, padding="valid", activation="relu", input_shape=(28, 28, 1)),
        Conv2D(32, 3, padding="valid", activation="relu"),
        MaxPooling2D(2),
        Flatten(),
        Dense(225, activation="relu"),
        Dense(10, activation="softmax"),
    ], name="mnist_convnet", seed=seed)


def higgs_mlp(hidden_layers: int = 3, dropout: bool = True, seed=None) -> Sequential:
    """30 -> 500 -> 500 -> 500 -> 2 softmax (517,502 parameters with three hidden layers)."""
    layers = [Dense(500, activation="relu", input_shape=(30,))]
    if dropout:
        layers.append(Dropout(0.4))
    if hidden_layers >= 2:
        layers.append(Dense(500, activation="relu"))
        if dropout:
            layers.append(Dropout(0.6))
    if hidden_layers >= 3:
        layers.append(Dense(500, activation="relu"))
    layers.append(Dense(2, activation="softmax"))
    return Sequential(layers, name="higgs_mlp", seed=seed)


def cifar10_cnn(seed=None) -> Sequential:
    """32x32x3 -> [conv3x3x32 same, conv3x3x32, pool] -> [conv3x3x64 same, conv3x3x64, pool]
    -> 512 -> 10: the classic Keras CIFAR-10 example network (1,250,858 parameters)."""
    return Sequential([
        Conv2D(32, 3, padding="same", activation="relu", input_shape=(32, 32, 3)),
        Conv2D(32, 3, padding="valid", activation="relu"),
        MaxPooling2D(2),
        Conv2D(64, 3, padding="same", activation="relu"),
        Conv2D(64, 3, padding="valid", activation="relu"),
        MaxPooling2D(2),
        Flatten(),
        Dense(512, activation="relu"),
        Dense(10, activation="softmax"),
    ], name="cifar10_cnn", seed=seed)


def resnet18(input_shape=(224, 224, 3), classes: int = 1000, seed=None) -> Sequential:
    """ResNet-18 (He et al. 2016): 7x7/2 stem, 3x3/2 max-pool, 4 stages x 2 basic blocks."""
    layers = [
        Conv2D(64, 7, strides=2, padding="same", use_bias=False, input_shape=tuple(input_shape)),
        BatchNormalization(),
        Activation("relu"),
        MaxPooling2D(2, 2),
    ]
    for stage, filters in enumerate((64, 128, 256, 512)):
        layers.append(ResidualBlock(filters, strides=1 if stage == 0 else 2))
        layers.append(ResidualBlock(filters, strides=1))
    layers += [GlobalAveragePooling2D(), Dense(classes, activation="softmax")]
    return Sequential(layers, name="resnet18", seed=seed)


ZOO = {
    "mnist_mlp": mnist_mlp,
    "mnist_convnet": mnist_convnet,
    "higgs_mlp": higgs_mlp,
    "cifar10_cnn": cifar10_cnn,
    "resnet18": resnet18,
}
