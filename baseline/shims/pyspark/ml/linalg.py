"""Dense vector / matrix stand-ins (numpy arrays with the pyspark spelling)."""
import numpy as np


class DenseVector(np.ndarray):
    def __new__(cls, values):
        return np.asarray(values, dtype=np.float64).reshape(-1).view(cls)

    def toArray(self):
        return np.asarray(self)

    def __reduce__(self):
        return (DenseVector, (np.asarray(self),))


class SparseVector(object):
    def __init__(self, size, indices, values=None):
        if values is None and isinstance(indices, dict):
            indices, values = zip(*sorted(indices.items())) if indices else ((), ())
        self.size, self.indices, self.values = int(size), np.asarray(indices, dtype=np.int64), np.asarray(values, dtype=np.float64)

    def toArray(self):
        out = np.zeros(self.size)
        out[self.indices] = self.values
        return out


class DenseMatrix(object):
    def __init__(self, numRows, numCols, values):
        self.numRows, self.numCols = int(numRows), int(numCols)
        self.values = np.asarray(values, dtype=np.float64)

    def toArray(self):
        return self.values.reshape(self.numCols, self.numRows).T


class Vectors(object):
    @staticmethod
    def dense(*values):
        if len(values) == 1 and not np.isscalar(values[0]):
            values = values[0]
        return DenseVector(values)

    @staticmethod
    def sparse(size, *args):
        return SparseVector(size, *args)
