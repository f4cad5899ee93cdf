"""Hot ops: fused flat-buffer optimizers, parameter-server algebra, tcgen05 GEMM wrappers."""
from .flat_optim import SGD, Adadelta, Adagrad, Adam, Adamax, FlatOptimizer, Nadam, OptimizerSpec, RMSprop

__all__ = ["SGD", "Adadelta", "Adagrad", "Adam", "Adamax", "FlatOptimizer", "Nadam", "OptimizerSpec", "RMSprop"]
