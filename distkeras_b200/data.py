"""Columnar in-memory dataset: the Spark DataFrame stand-in.

The reference moves data as Spark DataFrames whose columns hold vectors (``features``,
``label``; see ``distkeras/trainers.py:517-526`` and ``distkeras/workers.py:140-147`` where each
worker turns partition Rows into numpy mini-batches).  On a single B200 node there is no JVM: a
:class:`Dataset` is a dict of equally long, contiguous (optionally pinned) torch tensors, one per
column, plus a partition count.  Partitions are row ranges, so "repartition" is free and a
worker's mini-batch is a zero-copy slice that can be DMA'd straight from pinned host memory.

The small DataFrame surface the reference's users touch (``select``, ``repartition``, ``coalesce``,
``cache``, ``count``, ``randomSplit``, ``unionAll``, ``rdd.mapPartitionsWithIndex`` ...) is
kept so the examples read the same.
"""
from __future__ import annotations

import json
import math
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence

import numpy as np
import torch


class Row(dict):
    """A single record; attribute and item access like ``pyspark.sql.Row``."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:  # pragma: no cover - defensive
            raise AttributeError(name) from exc

    def asDict(self):
        return dict(self)


def _as_tensor(values) -> torch.Tensor:
    if isinstance(values, torch.Tensor):
        return values
    if isinstance(values, np.ndarray):
        if values.dtype == np.float64:
            values = values.astype(np.float32)
        return torch.from_numpy(np.ascontiguousarray(values))
    arr = np.asarray(values)
    if arr.dtype == np.float64:
        arr = arr.astype(np.float32)
    if arr.dtype == object:
        arr = np.stack([np.asarray(v, dtype=np.float32) for v in values])
    return torch.from_numpy(np.ascontiguousarray(arr))


class Partition:
    """A contiguous row range of a :class:`Dataset` (what one worker task consumes)."""

    def __init__(self, dataset: "Dataset", index: int, start: int, stop: int):
        self.dataset = dataset
        self.index = index
        self.start = start
        self.stop = stop

    def __len__(self) -> int:
        return self.stop - self.start

    def column(self, name: str) -> torch.Tensor:
        return self.dataset.columns_data[name][self.start:self.stop]

    def __iter__(self) -> Iterator[Row]:
        cols = self.dataset.columns
        data = [self.dataset.columns_data[c] for c in cols]
        for i in range(self.start, self.stop):
            yield Row((c, _cell(d[i])) for c, d in zip(cols, data))

    def batches(self, columns: Sequence[str], batch_size: int, drop_last: bool = True):
        """Yield tuples of tensor slices (zero-copy) of ``batch_size`` rows."""
        data = [self.dataset.columns_data[c] for c in columns]
        n = len(self)
        end = n - (n % batch_size) if drop_last else n
        for off in range(0, end, batch_size):
            lo, hi = self.start + off, min(self.start + off + batch_size, self.stop)
            yield tuple(d[lo:hi] for d in data)


def _cell(v: torch.Tensor):
    if v.dim() == 0:
        return v.item()
    return v.numpy()


class _RDD:
    """Minimal RDD facade: ``mapPartitionsWithIndex(f).collect()`` and friends."""

    def __init__(self, dataset: "Dataset"):
        self._dataset = dataset

    def mapPartitionsWithIndex(self, f: Callable[[int, Iterable], Iterable]) -> "_Collected":
        return _Collected(lambda: [list(f(p.index, p)) for p in self._dataset.partitions()])

    def mapPartitions(self, f: Callable[[Iterable], Iterable]) -> "_Collected":
        return _Collected(lambda: [list(f(p)) for p in self._dataset.partitions()])

    def map(self, f: Callable[[Row], Row]) -> "_Collected":
        return _Collected(lambda: [[f(r) for r in p] for p in self._dataset.partitions()])

    def getNumPartitions(self) -> int:
        return self._dataset.num_partitions


class _Collected:
    def __init__(self, thunk):
        self._thunk = thunk

    def collect(self) -> list:
        out = []
        for part in self._thunk():
            out.extend(part)
        return out

    def toDF(self) -> "Dataset":
        return Dataset.from_rows(self.collect())


class Dataset:
    """Dict of column tensors + a partition count."""

    def __init__(self, columns: Dict[str, object], num_partitions: int = 1):
        self.columns_data: Dict[str, torch.Tensor] = {k: _as_tensor(v) for k, v in columns.items()}
        lengths = {int(v.shape[0]) for v in self.columns_data.values()}
        if len(lengths) > 1:
            raise ValueError(f"columns have different lengths: {sorted(lengths)}")
        self._n = lengths.pop() if lengths else 0
        self.num_partitions = max(1, int(num_partitions))
        self._pinned = False

    # ------------------------------------------------------------------ constructors
    @classmethod
    def from_rows(cls, rows: Sequence[dict], num_partitions: int = 1) -> "Dataset":
        if not rows:
            return cls({}, num_partitions)
        cols = list(rows[0].keys())
        return cls({c: np.stack([np.asarray(r[c]) for r in rows]) for c in cols}, num_partitions)

    @classmethod
    def from_csv(cls, path: str, label_col: Optional[str] = None, features_col: str = "features",
                 header: bool = True, dtype=np.float32, num_partitions: int = 1, drop_cols: Sequence[str] = (),
                 label_map: Optional[Dict[str, int]] = None) -> "Dataset":
        """Read a CSV; all non-label columns are assembled into one vector column (the VectorAssembler
        step of ``examples/mnist.py:99-107``).  ``drop_cols`` removes bookkeeping columns (``EventId``,
        ``Weight``); ``label_map`` indexes a string label (``{"b": 0, "s": 1}`` -- the StringIndexer step of
        ``examples/workflow.ipynb``)."""
        with open(path, "r") as f:
            names = f.readline().strip().split(",") if header else None
        if label_map is not None or drop_cols:
            import pandas as pd

            df = pd.read_csv(path, header=0 if header else None)
            if names is None:
                df.columns = names = [f"c{i}" for i in range(df.shape[1])]
            df = df.drop(columns=[c for c in drop_cols if c in df.columns])
            cols: Dict[str, object] = {}
            if label_col is not None and label_col in df.columns:
                lab = df.pop(label_col)
                cols[label_col] = (np.array(lab.map(label_map), dtype=np.int64) if label_map is not None
                                   else np.array(lab, dtype=dtype))
            cols[features_col] = np.array(df.to_numpy(dtype=dtype), order="C")
            return cls(cols, num_partitions)
        arr = np.loadtxt(path, delimiter=",", skiprows=1 if header else 0, dtype=dtype, ndmin=2)
        if names is None:
            names = [f"c{i}" for i in range(arr.shape[1])]
        cols = {}
        if label_col is not None and label_col in names:
            li = names.index(label_col)
            cols[label_col] = arr[:, li].copy()
            arr = np.delete(arr, li, axis=1)
        cols[features_col] = np.ascontiguousarray(arr)
        return cls(cols, num_partitions)

    @classmethod
    def from_numpy_shards(cls, paths: Sequence[str], column: str = "features") -> "Dataset":
        """Concatenate ``.npy`` shards (``examples/distributed_numpy_parsing.ipynb:514-525``)."""
        return cls({column: np.concatenate([np.load(p) for p in paths], axis=0)}, len(paths))

    @classmethod
    def from_parquet(cls, path: str, num_partitions: int = 1) -> "Dataset":
        import pyarrow.parquet as pq

        table = pq.read_table(path)
        cols = {}
        for name in table.column_names:
            col = table.column(name).to_pylist()
            cols[name] = np.asarray(col, dtype=np.float32 if isinstance(col[0], (float, list)) else None)
        return cls(cols, num_partitions)

    def to_parquet(self, path: str) -> None:
        import pyarrow as pa
        import pyarrow.parquet as pq

        arrays = {}
        for k, v in self.columns_data.items():
            a = v.numpy()
            arrays[k] = pa.array(a.tolist())
        pq.write_table(pa.table(arrays), path)

    # ------------------------------------------------------------------ DataFrame-like surface
    @property
    def columns(self) -> List[str]:
        return list(self.columns_data.keys())

    @property
    def rdd(self) -> _RDD:
        return _RDD(self)

    def __len__(self) -> int:
        return self._n

    def count(self) -> int:
        return self._n

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.columns_data[name]

    def select(self, *names: str) -> "Dataset":
        flat = [n for group in names for n in ([group] if isinstance(group, str) else group)]
        return Dataset({n: self.columns_data[n] for n in flat}, self.num_partitions)

    def drop(self, *names: str) -> "Dataset":
        return Dataset({k: v for k, v in self.columns_data.items() if k not in names}, self.num_partitions)

    def with_column(self, name: str, values) -> "Dataset":
        cols = dict(self.columns_data)
        cols[name] = _as_tensor(values)
        out = Dataset(cols, self.num_partitions)
        return out

    withColumn = with_column

    def withColumnRenamed(self, old: str, new: str) -> "Dataset":
        return Dataset({(new if k == old else k): v for k, v in self.columns_data.items()},
                       self.num_partitions)

    def repartition(self, n: int) -> "Dataset":
        out = Dataset(self.columns_data, n)
        out._pinned = self._pinned
        return out

    def coalesce(self, n: int) -> "Dataset":
        return self.repartition(min(n, self.num_partitions) if self.num_partitions > n else n)

    def cache(self) -> "Dataset":
        return self

    persist = cache

    def unpersist(self) -> "Dataset":
        return self

    def take(self, n: int) -> List[Row]:
        return list(Partition(self, 0, 0, min(n, self._n)))

    def first(self) -> Row:
        return self.take(1)[0]

    def collect(self) -> List[Row]:
        return list(Partition(self, 0, 0, self._n))

    def limit(self, n: int) -> "Dataset":
        return Dataset({k: v[:n] for k, v in self.columns_data.items()}, self.num_partitions)

    def show(self, n: int = 20) -> None:
        """Print the first ``n`` rows (vectors abbreviated), like ``DataFrame.show``."""
        names = self.columns
        print(" | ".join(names))
        for row in self.take(n):
            cells = []
            for k in names:
                v = np.asarray(row[k].cpu() if torch.is_tensor(row[k]) else row[k]).reshape(-1)
                if v.size > 6:
                    cells.append(f"[{', '.join(f'{float(t):.3g}' for t in v[:3])}, ... x{v.size}]")
                else:
                    cells.append(str(v.tolist() if v.size > 1 else v.item()))
            print(" | ".join(cells))

    def printSchema(self) -> None:
        for k, v in self.columns_data.items():
            print(f" |-- {k}: {str(v.dtype).replace('torch.', '')}{list(v.shape[1:])}")

    def filter(self, predicate: Callable[["Dataset"], torch.Tensor]) -> "Dataset":
        mask = predicate(self)
        idx = torch.nonzero(_as_tensor(mask).bool(), as_tuple=False).flatten()
        return self.take_rows(idx)

    where = filter

    def take_rows(self, idx: torch.Tensor) -> "Dataset":
        return Dataset({k: v[idx] for k, v in self.columns_data.items()}, self.num_partitions)

    def union(self, other: "Dataset") -> "Dataset":
        return Dataset({k: torch.cat([v, other.columns_data[k]], dim=0)
                        for k, v in self.columns_data.items()}, self.num_partitions)

    unionAll = union

    def shuffle(self, seed: Optional[int] = None) -> "Dataset":
        g = torch.Generator()
        if seed is not None:
            g.manual_seed(int(seed))
        perm = torch.randperm(self._n, generator=g)
        return self.take_rows(perm)

    def randomSplit(self, weights: Sequence[float], seed: Optional[int] = None) -> List["Dataset"]:
        g = torch.Generator()
        if seed is not None:
            g.manual_seed(int(seed))
        perm = torch.randperm(self._n, generator=g)
        total = float(sum(weights))
        bounds, acc = [0], 0.0
        for w in weights:
            acc += w / total
            bounds.append(int(round(acc * self._n)))
        bounds[-1] = self._n
        return [self.take_rows(perm[bounds[i]:bounds[i + 1]]) for i in range(len(weights))]

    split = randomSplit

    def sample(self, fraction: float, seed: Optional[int] = None) -> "Dataset":
        return self.randomSplit([fraction, 1.0 - fraction], seed)[0]

    # ------------------------------------------------------------------ partitions
    def partitions(self, n: Optional[int] = None) -> List[Partition]:
        n = self.num_partitions if n is None else max(1, int(n))
        size = math.ceil(self._n / n) if self._n else 0
        parts = []
        for i in range(n):
            lo = min(i * size, self._n)
            hi = min(lo + size, self._n)
            parts.append(Partition(self, i, lo, hi))
        return parts

    def pin_memory(self) -> "Dataset":
        """Page-lock every column so mini-batches can be DMA'd without a staging copy."""
        if self._pinned or not torch.cuda.is_available():
            return self
        self.columns_data = {k: v.contiguous().pin_memory() for k, v in self.columns_data.items()}
        self._pinned = True
        return self

    def share_memory_(self) -> "Dataset":
        """Move columns to shared memory so spawned worker processes see them without a copy."""
        for v in self.columns_data.values():
            if not v.is_shared():
                v.share_memory_()
        return self

    def __repr__(self) -> str:
        cols = ", ".join(f"{k}:{tuple(v.shape[1:])}" for k, v in self.columns_data.items())
        return f"Dataset(rows={self._n}, partitions={self.num_partitions}, columns=[{cols}])"


def json_rows_to_dataset(lines: Iterable[str], num_partitions: int = 1) -> Dataset:
    """Streaming-inference helper: JSON records -> Dataset (Kafka notebook, cell 14)."""
    return Dataset.from_rows([json.loads(s) for s in lines], num_partitions)


# ---------------------------------------------------------------------------------------------
# synthetic generators for the benchmark configs (no network: shapes only, random content)
# ---------------------------------------------------------------------------------------------
def synthetic_mnist(n: int = 60000, seed: int = 0, as_uint8: bool = True, flat: bool = True,
                    learnable: bool = True) -> Dataset:
    """MNIST-shaped data: 784 uint8 pixels + int label.  With ``learnable`` the label is a
    function of the pixels (class-dependent blob) so training curves are meaningful."""
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, 10, (n,), generator=g)
    x = torch.randint(0, 64, (n, 784), generator=g, dtype=torch.int32)
    if learnable:
        proto = torch.randint(0, 192, (10, 784), generator=g, dtype=torch.int32)
        x = x + proto[labels]
    x = x.clamp_(0, 255)
    feats = x.to(torch.uint8) if as_uint8 else x.to(torch.float32)
    if not flat:
        feats = feats.reshape(n, 28, 28, 1)
    return Dataset({"features": feats, "label": labels.to(torch.int32)})


def synthetic_cifar10(n: int = 50000, seed: int = 0, as_uint8: bool = True, learnable: bool = True) -> Dataset:
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, 10, (n,), generator=g)
    x = torch.randint(0, 64, (n, 32 * 32 * 3), generator=g, dtype=torch.int32)
    if learnable:
        proto = torch.randint(0, 192, (10, 32 * 32 * 3), generator=g, dtype=torch.int32)
        x = x + proto[labels]
    x = x.clamp_(0, 255)
    feats = (x.to(torch.uint8) if as_uint8 else x.to(torch.float32)).reshape(n, 32, 32, 3)
    return Dataset({"features": feats, "label": labels.to(torch.int32)})


def synthetic_higgs(n: int = 100000, seed: int = 0) -> Dataset:
    """ATLAS-Higgs-shaped data: 30 float features, binary label (``examples/workflow.ipynb``)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 30, generator=g)
    w = torch.randn(30, generator=g)
    labels = ((x @ w + 0.5 * torch.randn(n, generator=g)) > 0).to(torch.int32)
    return Dataset({"features": x, "label": labels})


def synthetic_imagenet(n: int = 1024, size: int = 224, classes: int = 1000, seed: int = 0) -> Dataset:
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 256, (n, size, size, 3), generator=g, dtype=torch.uint8)
    labels = torch.randint(0, classes, (n,), generator=g).to(torch.int32)
    return Dataset({"features": x, "label": labels})
