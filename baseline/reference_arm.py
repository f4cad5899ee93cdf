"""Reference arm of ``bench.py --impl reference``: drives the UNMODIFIED cerndb/dist-keras installed
under ``baseline/_ref`` through its own public API and stock code path.

    distkeras.trainers.ADAG(keras_model, 'adam', 'categorical_crossentropy', num_workers=N, batch_size=B,
                            communication_window=tau, ...).train(dataframe)

i.e. the reference's Spark job: a parameter-server thread in the driver (``SocketParameterServer``,
TCP + pickle, one mutex), one ``ADAGWorker`` task per partition running ``train_on_batch`` with a
commit + pull every ``tau`` mini-batches (``distkeras/workers.py:327-342``).

The image has neither pyspark nor Keras/TensorFlow and no network, so the two *substrates* the
reference sits on are supplied by ``baseline/shims`` (a local driver/executor-process Spark runtime
and a Keras API on plain torch; both import nothing from ``distkeras_b200``).  Nothing under
``baseline/_ref`` is patched.  Every executor process gets its own GPU (replicas run through torch's
stock cuBLAS path with TF32 enabled, TensorFlow's default), the center variable crosses host memory,
pickle and a loopback TCP socket, exactly as in the reference.

Metric: same as the native arm -- W untimed warm-up steps, then exactly K steps, max over workers --
taken from the reference's OWN per-batch history timestamps (``workers.py:268-275``), which excludes
Spark/worker start-up, model compilation and the 10 s queue-timeout drain every worker pays at the end
(``workers.py:121-123``); the wall-clock ``trainer.get_training_time()`` is reported next to it.
"""
from __future__ import annotations

import json
import os
import socket
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def _prepare_paths() -> None:
    for p in (os.path.join(HERE, "_ref"), os.path.join(HERE, "shims")):
        if p not in sys.path:
            sys.path.insert(0, p)


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ensure_hostname_resolves() -> str:
    """``DistributedTrainer.__init__`` calls ``gethostbyname(gethostname())`` (``networking.py:11-15``);
    container hostnames do not always resolve.  Environment fix only: fall back to the loopback name."""
    try:
        socket.gethostbyname(socket.gethostname())
        return "resolved"
    except OSError:
        socket.gethostname = lambda: "localhost"  # type: ignore[assignment]
        return "gethostname() -> localhost (container hostname does not resolve)"


def build_keras_model(name: str):
    from keras.layers import Activation, Conv2D, Dense, Dropout, Flatten, MaxPooling2D
    from keras.models import Sequential

    m = Sequential()
    if name == "mnist_mlp":      # examples/mnist_analysis.ipynb:247-252 (987,210 parameters)
        m.add(Dense(1000, activation="relu", input_shape=(784,)))
        m.add(Dropout(0.2))
        m.add(Dense(200, activation="relu"))
        m.add(Dropout(0.2))
        m.add(Dense(10, activation="softmax"))
        return m, (784,), 10
    if name == "higgs_mlp":      # examples/example_1_analysis.ipynb:302-305
        m.add(Dense(500, activation="relu", input_shape=(30,)))
        m.add(Dropout(0.4))
        m.add(Dense(500, activation="relu"))
        m.add(Dropout(0.6))
        m.add(Dense(500, activation="relu"))
        m.add(Dense(2, activation="softmax"))
        return m, (30,), 2
    if name == "mnist_convnet":  # examples/mnist.py:150-162
        m.add(Conv2D(32, 3, padding="valid", activation="relu", input_shape=(28, 28, 1)))
        m.add(Conv2D(32, 3, padding="valid", activation="relu"))
        m.add(MaxPooling2D((2, 2)))
        m.add(Flatten())
        m.add(Dense(225, activation="relu"))
        m.add(Dense(10, activation="softmax"))
        return m, (28, 28, 1), 10
    if name == "cifar10_cnn":
        m.add(Conv2D(32, 3, padding="same", activation="relu", input_shape=(32, 32, 3)))
        m.add(Conv2D(32, 3, padding="valid", activation="relu"))
        m.add(MaxPooling2D((2, 2)))
        m.add(Conv2D(64, 3, padding="same", activation="relu"))
        m.add(Conv2D(64, 3, padding="valid", activation="relu"))
        m.add(MaxPooling2D((2, 2)))
        m.add(Flatten())
        m.add(Dense(512, activation="relu"))
        m.add(Dense(10, activation="softmax"))
        return m, (32, 32, 3), 10
    raise ValueError(name)


def run(algo: str, model_name: str, n_gpus: int, steps: int, warmup: int, batch: int, window: int, optimizer: str,
        dedicated_ps: bool = False) -> dict:
    """Run the reference job and return the bench JSON record (rank 0 only; other torchrun ranks
    return None -- the reference's own launcher is its Spark driver, which spawns the executors)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return None
    _prepare_paths()
    host_note = _ensure_hostname_resolves()
    os.environ.setdefault("KERAS_SHIM_DEVICE", "cpu")  # the driver only holds the PS copy of the model
    import numpy as np

    n_gpu_visible = 0
    try:
        import torch

        n_gpu_visible = torch.cuda.device_count()
    except Exception:
        pass
    workers = max(1, n_gpus - 1) if (dedicated_ps and n_gpus > 1) else max(1, n_gpus)
    if n_gpu_visible:
        first = 1 if (dedicated_ps and n_gpu_visible > workers) else 0
        os.environ["SPARK_SHIM_GPUS"] = ",".join(str((first + i) % n_gpu_visible) for i in range(workers))
    from distkeras import trainers as ref_trainers
    from pyspark import SparkConf, SparkContext
    from pyspark.sql import DataFrame

    K, W, B, tau = int(steps), max(1, int(warmup)), int(batch), int(window)
    model, in_shape, classes = build_keras_model(model_name)
    sc = SparkContext(conf=SparkConf().setMaster("local[%d]" % workers).setAppName("dist-keras reference arm"))
    rows = workers * (W + K) * B
    rng = np.random.RandomState(1234)
    feat = int(np.prod(in_shape))
    x = rng.randint(0, 256, size=(rows, feat)).astype(np.float32) / np.float32(255.0)
    if len(in_shape) > 1:
        x = x.reshape((rows,) + tuple(in_shape))
    labels = rng.randint(0, classes, size=rows)
    y = np.zeros((rows, classes), dtype=np.float32)
    y[np.arange(rows), labels] = 1.0
    df = DataFrame.from_columns(sc, {"features": x, "label_encoded": y}, workers)

    cls = {"adag": ref_trainers.ADAG, "downpour": ref_trainers.DOWNPOUR, "aeasgd": ref_trainers.AEASGD,
           "dynsgd": ref_trainers.DynSGD, "eamsgd": ref_trainers.EAMSGD}[algo]
    kw = dict(keras_model=model, worker_optimizer=optimizer, loss="categorical_crossentropy", num_workers=workers,
              batch_size=B, features_col="features", label_col="label_encoded", num_epoch=1,
              communication_window=tau, master_port=_free_port())
    if algo in ("aeasgd", "eamsgd"):
        kw.update(rho=0.1, learning_rate=0.1)
    trainer = cls(**kw)
    t0 = time.time()
    trained = trainer.train(df)
    wall = time.time() - t0
    sc.stop()

    history = trainer.get_history()
    per_worker = {}
    for h in history:
        per_worker.setdefault(h["worker_id"], []).append((h["iteration"], h["timestamp"], h["history"]))
    elapsed, done_steps, ms_list = [], [], []
    for wid, recs in sorted(per_worker.items()):
        recs.sort()
        stamps = {it: ts for it, ts, _ in recs}
        done_steps.append(len(recs))
        if W in stamps and (W + K) in stamps:
            elapsed.append(stamps[W + K] - stamps[W])
            ms_list.append(1e3 * elapsed[-1] / K)
    ok = len(elapsed) == workers
    out = {
        "metric": f"{model_name} {algo.upper()} training throughput (samples/s, whole job)",
        "unit": "samples/s", "n_gpus": n_gpus, "steps": K, "warmup": W, "higher_is_better": True, "scaling": "weak",
        "dtype": "fp32 storage, TF32 matmul on GPU (TensorFlow's default)" if n_gpu_visible else "fp32 (CPU)",
        "data": "synthetic", "impl": "reference",
        "reference_class": "unmodified cerndb/dist-keras (baseline/_ref) + shimmed Keras/Spark substrate "
                           "(baseline/shims: Keras API on torch, local driver/executor-process Spark)",
        "config": {"model": model_name, "trainer": cls.__name__, "global_batch": workers * B, "batch_per_worker": B,
                   "seq_len": None, "num_workers": workers, "communication_window": tau, "worker_optimizer": optimizer,
                   "parallelism": f"async-ps(driver thread, TCP+pickle)+dp{workers}",
                   "ps_transport": "TCP loopback + pickle (distkeras/networking.py)",
                   "replica_device": "cuda (one GPU per executor process)" if n_gpu_visible else "cpu"},
        "gpu_launches": 0,
        "wallclock_training_time_s": trainer.get_training_time(), "wallclock_total_s": wall,
        "wallclock_value": workers * B * (W + K) / max(trainer.get_training_time(), 1e-9),
        "num_updates": int(trainer.parameter_server.num_updates) if trainer.parameter_server is not None else None,
        "steps_done_per_worker": done_steps, "hostname_note": host_note,
        "timing": "reference's own per-batch history timestamps: (t[W+K] - t[W]) max over workers; excludes executor "
                  "start-up, model compile and the 10 s queue-timeout drain",
    }
    if ok:
        worst = max(elapsed)
        value = workers * B * K / worst
        bytes_in = B * (feat + classes) * 4
        out.update({"value": value, "ms_per_step": 1e3 * worst / K, "per_worker_ms_per_step": ms_list,
                    "vs_baseline": None,
                    "e2e": {"value": value, "unit": "samples/s", "ms_per_step": 1e3 * worst / K, "steps": K,
                            "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": 8,
                            "api": f"distkeras.trainers.{cls.__name__}(...).train(dataframe)"}})
    else:
        out["unavailable"] = ("the reference ran but %d of %d workers did not finish %d steps (steps done: %s)"
                              % (workers - len(elapsed), workers, W + K, done_steps))
    del trained
    return out


def main(argv=None) -> None:
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--algo", default="adag")
    ap.add_argument("--model", default="mnist_mlp")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--window", type=int, default=12)
    ap.add_argument("--optimizer", default="adam")
    ap.add_argument("--dedicated-ps", action="store_true")
    a = ap.parse_args(argv)
    rec = run(a.algo, a.model, a.gpus, a.steps, a.warmup, a.batch, a.window, a.optimizer, a.dedicated_ps)
    if rec is not None:
        print(json.dumps(rec))


if __name__ == "__main__":
    main()
