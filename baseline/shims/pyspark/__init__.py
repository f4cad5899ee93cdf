"""Substrate shim: a local, Spark-shaped runtime (driver + executor processes) with the slice of the
pyspark API that cerndb/dist-keras touches.

It exists ONLY so the unmodified reference under ``baseline/_ref`` can run on an image without
pyspark / a JVM.  It imports nothing from ``distkeras_b200``.  The execution model mirrors Spark local
mode: the program that builds the trainer is the *driver*; ``rdd.mapPartitionsWithIndex(f).collect()``
pickles the closure, ships it with the partitions to ``local[N]`` *executor processes* (spawned Python
workers, like ``pyspark.daemon``'s), runs one task per partition and collects the results.  When
``SPARK_SHIM_GPUS`` lists GPU ids, each executor process is pinned to one of them through
``CUDA_VISIBLE_DEVICES`` (the analogue of Spark's GPU resource scheduling).
"""
import os
import pickle
import re
import sys
import threading

try:  # closures (lambdas) need cloudpickle, exactly like real pyspark
    import cloudpickle as _fnpickle
except ImportError:  # pragma: no cover
    _fnpickle = pickle

__version__ = "2.2.0+localshim"


class SparkConf(object):
    def __init__(self, loadDefaults=True):
        self._conf = {}

    def set(self, key, value):
        self._conf[key] = str(value)
        return self

    def setMaster(self, master):
        return self.set("spark.master", master)

    def setAppName(self, name):
        return self.set("spark.app.name", name)

    def get(self, key, default=None):
        return self._conf.get(key, default)

    def getAll(self):
        return list(self._conf.items())


# ------------------------------------------------------------------------------------------------
# executors
# ------------------------------------------------------------------------------------------------
def _executor_boot(counter, gpus, paths, cores):
    """Runs once in every spawned executor process, before any task."""
    # Spark (>= 3.0, SPARK-28843) exports OMP_NUM_THREADS = executor cores to its Python workers so that
    # co-located executors do not oversubscribe the host
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    os.environ.setdefault("MKL_NUM_THREADS", str(cores))
    with counter.get_lock():
        idx = counter.value
        counter.value += 1
    os.environ["SPARK_SHIM_EXECUTOR_ID"] = str(idx)
    if gpus:
        os.environ["CUDA_VISIBLE_DEVICES"] = gpus[idx % len(gpus)]
        os.environ.pop("KERAS_SHIM_DEVICE", None)
    for p in reversed(paths):
        if p not in sys.path:
            sys.path.insert(0, p)


def _run_task(payload):
    func, index, part = _fnpickle.loads(payload)
    return pickle.dumps(list(func(index, part.iter_rows())), -1)


class _Executors(object):
    def __init__(self, n):
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor

        ctx = mp.get_context("spawn")
        gpus = [g for g in os.environ.get("SPARK_SHIM_GPUS", "").split(",") if g != ""]
        self.n = n
        self.pool = ProcessPoolExecutor(max_workers=n, mp_context=ctx, initializer=_executor_boot,
                                        initargs=(ctx.Value("i", 0), gpus, list(sys.path), max(1, (os.cpu_count() or 1) // max(1, n))))

    def run(self, func, partitions):
        futures = [self.pool.submit(_run_task, _fnpickle.dumps((func, i, p), -1)) for i, p in enumerate(partitions)]
        return [pickle.loads(f.result()) for f in futures]

    def shutdown(self):
        self.pool.shutdown(wait=True, cancel_futures=True)


class SparkContext(object):
    _active = None
    _lock = threading.Lock()

    def __init__(self, master=None, appName=None, conf=None, **kwargs):
        self._conf = conf or SparkConf()
        master = master or self._conf.get("spark.master", "local[*]")
        m = re.match(r"local\[(\d+|\*)\]", master or "")
        if m and m.group(1) != "*":
            n = int(m.group(1))
        elif self._conf.get("spark.executor.instances"):
            n = int(self._conf.get("spark.executor.instances")) * int(self._conf.get("spark.executor.cores", "1"))
        else:
            n = os.cpu_count() or 1
        self.master, self.appName = master, appName
        self.defaultParallelism = max(1, n)
        self._executors = None
        SparkContext._active = self

    @classmethod
    def getOrCreate(cls, conf=None):
        with cls._lock:
            return cls._active or cls(conf=conf)

    def executors(self):
        if self._executors is None:
            self._executors = _Executors(self.defaultParallelism)
        return self._executors

    def parallelize(self, data, numSlices=None):
        data = list(data)
        n = max(1, min(numSlices or self.defaultParallelism, max(1, len(data))))
        per = -(-len(data) // n)
        return RDD(self, [ListPartition(data[i * per:(i + 1) * per]) for i in range(n)])

    def stop(self):
        if self._executors is not None:
            self._executors.shutdown()
            self._executors = None
        if SparkContext._active is self:
            SparkContext._active = None


# ------------------------------------------------------------------------------------------------
# partitions and RDDs
# ------------------------------------------------------------------------------------------------
class ListPartition(object):
    def __init__(self, rows):
        self.rows = rows

    def iter_rows(self):
        return iter(self.rows)

    def __len__(self):
        return len(self.rows)


class ColumnarPartition(object):
    """Row range ``[lo, hi)`` of named numpy columns; rows are materialised lazily by the iterator
    (Spark streams pickled Rows from the JVM to the Python worker; the columns are the analogue of the
    cached JVM-side partition)."""

    def __init__(self, columns, lo, hi):
        self.columns, self.lo, self.hi = columns, int(lo), int(hi)

    def __getstate__(self):  # ship only this partition's slice to the executor
        return {"columns": {k: v[self.lo:self.hi] for k, v in self.columns.items()}, "lo": 0, "hi": self.hi - self.lo}

    def iter_rows(self):
        from pyspark.sql import Row

        names = list(self.columns.keys())
        cols = [self.columns[k] for k in names]
        make = Row(*names)
        for i in range(self.lo, self.hi):
            yield make(*[c[i] for c in cols])

    def __len__(self):
        return self.hi - self.lo


def _identity(index, it):
    return it


class RDD(object):
    def __init__(self, sc, partitions, func=None):
        self.ctx = sc
        self._partitions = list(partitions)
        self._func = func or _identity

    def getNumPartitions(self):
        return len(self._partitions)

    def cache(self):
        return self

    persist = cache

    def mapPartitionsWithIndex(self, f, preservesPartitioning=False):
        prev = self._func

        def chained(index, it, _prev=prev, _f=f):
            return _f(index, _prev(index, it))

        return RDD(self.ctx, self._partitions, chained)

    def mapPartitions(self, f, preservesPartitioning=False):
        return self.mapPartitionsWithIndex(lambda index, it, _f=f: _f(it))

    def map(self, f):
        return self.mapPartitionsWithIndex(lambda index, it, _f=f: (_f(x) for x in it))

    def filter(self, f):
        return self.mapPartitionsWithIndex(lambda index, it, _f=f: (x for x in it if _f(x)))

    def _compute(self):
        return self.ctx.executors().run(self._func, self._partitions)

    def collect(self):
        return [x for part in self._compute() for x in part]

    def count(self):
        return sum(len(p) for p in self._compute())

    def first(self):
        return self.collect()[0]

    def take(self, n):
        return self.collect()[:n]

    def _materialise(self):
        return [ListPartition(rows) for rows in self._compute()] if self._func is not _identity else self._partitions

    def repartition(self, n):
        rows = [x for p in self._materialise() for x in p.iter_rows()]
        per = -(-len(rows) // max(1, n))
        return RDD(self.ctx, [ListPartition(rows[i * per:(i + 1) * per]) for i in range(n)])

    def coalesce(self, n):
        return self if n >= self.getNumPartitions() else self.repartition(n)

    def toDF(self, schema=None):
        from pyspark.sql import DataFrame

        return DataFrame(RDD(self.ctx, self._materialise()))
