"""``pyspark.sql`` stand-in: ``Row``, a columnar local ``DataFrame``, ``SparkSession`` / ``SQLContext``."""
import numpy as np

from pyspark import RDD, ColumnarPartition, ListPartition, SparkContext


class Row(tuple):
    """``Row(a=1, b=2)`` builds a row; ``Row("a", "b")`` builds a factory -- both as in pyspark."""

    def __new__(cls, *args, **kwargs):
        if args and kwargs:
            raise ValueError("Can not use both args and kwargs to create Row")
        if kwargs:
            row = tuple.__new__(cls, list(kwargs.values()))
            row.__fields__ = list(kwargs.keys())
            return row
        return tuple.__new__(cls, args)

    def __call__(self, *values):
        if len(values) > len(self):
            raise ValueError("Can not create Row with fields %s, expected %d values but got %s" % (self, len(self), values))
        row = tuple.__new__(Row, values)
        row.__fields__ = list(self)
        return row

    def __getitem__(self, item):
        if isinstance(item, (int, slice)):
            return tuple.__getitem__(self, item)
        try:
            return tuple.__getitem__(self, self.__fields__.index(item))
        except (AttributeError, ValueError):
            raise KeyError(item)

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        try:
            return tuple.__getitem__(self, self.__fields__.index(item))
        except (AttributeError, ValueError):
            raise AttributeError(item)

    def __contains__(self, item):
        return item in getattr(self, "__fields__", ()) or tuple.__contains__(self, item)

    def asDict(self, recursive=False):
        return dict(zip(self.__fields__, self))

    def __reduce__(self):
        if hasattr(self, "__fields__"):
            return (_restore_row, (self.__fields__, tuple(self)))
        return tuple.__reduce__(self)

    def __repr__(self):
        if hasattr(self, "__fields__"):
            return "Row(%s)" % ", ".join("%s=%r" % kv for kv in zip(self.__fields__, tuple(self)))
        return "<Row(%s)>" % ", ".join(repr(f) for f in self)


def _restore_row(fields, values):
    row = tuple.__new__(Row, values)
    row.__fields__ = list(fields)
    return row


class DataFrame(object):
    """Rows held as partitions of an RDD.  ``from_columns`` keeps the data columnar (one numpy array
    per column) so re-partitioning is a re-slice and executors receive contiguous buffers."""

    def __init__(self, rdd, columns=None):
        self._rdd = rdd
        self._columns = columns  # dict name -> ndarray when the frame is still columnar

    @classmethod
    def from_columns(cls, sc, columns, num_partitions=1):
        n = len(next(iter(columns.values())))
        parts = _slice_parts(columns, n, num_partitions)
        return cls(RDD(sc, parts), dict(columns))

    @property
    def rdd(self):
        return self._rdd

    @property
    def columns(self):
        if self._columns is not None:
            return list(self._columns.keys())
        first = next(self._rdd._partitions[0].iter_rows())
        return list(first.__fields__)

    def _n(self):
        return sum(len(p) for p in self._rdd._partitions)

    def count(self):
        return self._n()

    def cache(self):
        return self

    persist = cache

    def unpersist(self):
        return self

    def repartition(self, n):
        if self._columns is not None:
            return DataFrame(RDD(self._rdd.ctx, _slice_parts(self._columns, self._n(), n)), self._columns)
        return DataFrame(self._rdd.repartition(n))

    def coalesce(self, n):
        return self if n >= self._rdd.getNumPartitions() else self.repartition(n)

    def orderBy(self, *cols, **kwargs):
        from pyspark.sql.functions import Column

        if self._columns is not None and cols and isinstance(cols[0], Column) and cols[0].op == "rand":
            perm = np.random.RandomState(cols[0].arg).permutation(self._n())
            shuffled = {k: v[perm] for k, v in self._columns.items()}
            return DataFrame(RDD(self._rdd.ctx, _slice_parts(shuffled, len(perm), self._rdd.getNumPartitions())), shuffled)
        raise NotImplementedError("orderBy is only implemented for rand() on columnar frames")

    def select(self, *names):
        if self._columns is None:
            raise NotImplementedError("select needs a columnar frame")
        cols = {k: self._columns[k] for k in names}
        return DataFrame(RDD(self._rdd.ctx, _slice_parts(cols, self._n(), self._rdd.getNumPartitions())), cols)

    def collect(self):
        return [r for p in self._rdd._partitions for r in p.iter_rows()]

    def take(self, n):
        out = []
        for p in self._rdd._partitions:
            for r in p.iter_rows():
                out.append(r)
                if len(out) >= n:
                    return out
        return out

    def first(self):
        return self.take(1)[0]

    def printSchema(self):
        print("root\n" + "\n".join(" |-- %s" % c for c in self.columns))


def _slice_parts(columns, n, num_partitions):
    num_partitions = max(1, int(num_partitions))
    per = -(-n // num_partitions)
    return [ColumnarPartition(columns, min(n, i * per), min(n, (i + 1) * per)) for i in range(num_partitions)]


class SparkSession(object):
    class Builder(object):
        def __init__(self):
            self._master, self._name = None, None

        def master(self, m):
            self._master = m
            return self

        def appName(self, n):
            self._name = n
            return self

        def config(self, *a, **k):
            return self

        def getOrCreate(self):
            return SparkSession(SparkContext._active or SparkContext(self._master, self._name))

    builder = Builder()

    def __init__(self, sc):
        self.sparkContext = sc

    def createDataFrame(self, data, schema=None):
        if isinstance(data, dict):
            return DataFrame.from_columns(self.sparkContext, data, self.sparkContext.defaultParallelism)
        if isinstance(data, RDD):
            return data.toDF()
        return DataFrame(self.sparkContext.parallelize(list(data)))

    def stop(self):
        self.sparkContext.stop()


class SQLContext(SparkSession):
    pass
