"""Remote job deployment: submit a trainer to a daemon on the GPU box and fetch the result.

Capability parity with ``distkeras/job_deployment.py`` -- ``Job`` (client), ``Punchcard`` (REST
daemon with ``/api/submit | state | cancel | destroy`` and a ``secrets.json`` allow-list),
``PunchcardJob`` (one training run).  Re-designed for a single 8xB200 node:

* no Flask / urllib2 / generated Spark driver script: the daemon is a stdlib
  ``ThreadingHTTPServer``; a job runs the pickled trainer in a **child process** (so it can be
  cancelled for real -- the reference calls the non-existent ``thread.exit()``,
  ``job_deployment.py:184-185``) on a dataset loaded from ``data_path`` (``.pt`` / ``.npz`` /
  ``.parquet`` / ``.csv``);
* payloads are hex-encoded pickles exactly like the reference (``job_deployment.py:338-350``), so
  only submit jobs to daemons you trust -- the secret is an allow-list entry, not a sandbox.
"""
from __future__ import annotations

import binascii
import json
import multiprocessing as mp
import threading
import time
import urllib.error
import urllib.parse
import urllib.request
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Optional

import numpy as np

from .data import Dataset
from .utils import deserialize_keras_model, pickle_object, serialize_keras_model, unpickle_object


def _hex(obj) -> str:
    return binascii.hexlify(pickle_object(obj)).decode("ascii")


def _unhex(text: str):
    return unpickle_object(binascii.unhexlify(text.encode("ascii")))


def load_dataset(path: str) -> Dataset:
    """Load a job's training data (the reference reads Parquet from HDFS, ``job_deployment.py:243``)."""
    import torch

    if path.endswith(".pt"):
        obj = torch.load(path, weights_only=False)
        return obj if isinstance(obj, Dataset) else Dataset(obj)
    if path.endswith(".npz"):
        with np.load(path) as z:
            return Dataset({k: z[k] for k in z.files})
    if path.endswith(".parquet"):
        return Dataset.from_parquet(path)
    if path.endswith(".csv"):
        return Dataset.from_csv(path, label_col="label")
    raise ValueError(f"unsupported data_path {path!r}")


def _job_main(trainer_bytes: bytes, data_path: str, num_workers: Optional[int], conn) -> None:
    try:
        trainer = unpickle_object(trainer_bytes)
        if num_workers and hasattr(trainer, "set_num_workers"):
            trainer.set_num_workers(num_workers)
        model = trainer.train(load_dataset(data_path))
        conn.send({"ok": True, "model": serialize_keras_model(model), "history": trainer.get_history(),
                   "training_time": trainer.get_training_time()})
    except BaseException as exc:  # reported to the client, not swallowed
        conn.send({"ok": False, "error": repr(exc)})
    finally:
        conn.close()


class PunchcardJob:
    """One submitted training run (``job_deployment.py:152-281``)."""

    def __init__(self, secret, job_name, data_path, num_executors, num_processes, trainer):
        self.secret = secret
        self.job_name = job_name
        self.data_path = data_path
        self.num_executors = int(num_executors)
        self.num_processes = int(num_processes)
        self.trainer = trainer if isinstance(trainer, bytes) else pickle_object(trainer)
        self.trained_model = None
        self.history = None
        self.error: Optional[str] = None
        self.is_running = False
        self.process: Optional[mp.Process] = None
        self._thread: Optional[threading.Thread] = None

    def get_job_name(self):
        return self.job_name

    def get_secret(self):
        return self.secret

    def get_history(self):
        return self.history

    def get_trained_model(self):
        return self.trained_model

    def running(self) -> bool:
        return self.is_running

    def start(self) -> None:
        self.is_running = True
        self._thread = threading.Thread(target=self.run, daemon=True)
        self._thread.start()

    def join(self, timeout=None) -> None:
        if self._thread is not None:
            self._thread.join(timeout)

    def cancel(self) -> None:
        if self.process is not None and self.process.is_alive():
            self.process.terminate()
        self.is_running = False

    def run(self) -> None:
        ctx = mp.get_context("spawn")
        parent, child = ctx.Pipe(duplex=False)
        workers = self.num_executors * self.num_processes
        self.process = ctx.Process(target=_job_main, args=(self.trainer, self.data_path, workers or None, child))
        self.process.start()
        child.close()
        try:
            result = parent.recv()
            if result.get("ok"):
                self.trained_model = deserialize_keras_model(result["model"])
                self.history = result["history"]
            else:
                self.error = result.get("error")
        except EOFError:
            self.error = "job process exited without a result (cancelled?)"
        finally:
            self.process.join(timeout=30)
            self.is_running = False


class Punchcard:
    """REST daemon (``job_deployment.py:37-149``)."""

    def __init__(self, secrets_path="secrets.json", port=80, host="0.0.0.0"):
        self.secrets_path = secrets_path
        self.port = int(port)
        self.host = host
        self.mutex = threading.Lock()
        self.jobs = {}
        self.server: Optional[ThreadingHTTPServer] = None

    def read_secrets(self):
        with open(self.secrets_path) as f:
            return json.load(f)

    def valid_secret(self, secret, secrets) -> bool:
        return any(d.get("secret") == secret for d in secrets)

    def secret_in_use(self, secret) -> bool:
        return secret in self.jobs

    def get_submitted_job(self, secret) -> Optional[PunchcardJob]:
        with self.mutex:
            return self.jobs.get(secret)

    # -- route bodies -------------------------------------------------------------------------
    def submit_job(self, data: dict):
        secret = data["secret"]
        secrets = self.read_secrets()
        with self.mutex:
            if self.valid_secret(secret, secrets) and not self.secret_in_use(secret):
                job = PunchcardJob(secret, data["job_name"], data["data_path"], data.get("num_executors", 1),
                                   data.get("num_processes", 1), binascii.unhexlify(data["trainer"]))
                self.jobs[secret] = job
                job.start()
                return 200, ""
        return 403, ""

    def job_state(self, secret):
        job = self.get_submitted_job(secret)
        if job is None:
            return 404, ""
        return 200, json.dumps({"job_name": job.get_job_name(), "running": job.running(), "error": job.error})

    def cancel(self, secret):
        job = self.get_submitted_job(secret)
        if job is not None and job.running():
            with self.mutex:
                job.cancel()
                del self.jobs[secret]
        return 200, ""

    def destroy_job(self, secret):
        job = self.get_submitted_job(secret)
        if job is None or job.running():
            return 400, ""
        with self.mutex:
            model = job.get_trained_model()
            d = {"model": _hex(serialize_keras_model(model)) if model is not None else None,
                 "history": _hex(job.get_history()), "error": job.error}
            del self.jobs[secret]
        return 200, json.dumps(d)

    # -- server ---------------------------------------------------------------------------------
    def _make_handler(self):
        daemon = self

        class Handler(BaseHTTPRequestHandler):
            def log_message(self, fmt, *args):  # quiet
                pass

            def _reply(self, code, body):
                payload = body.encode("utf-8")
                self.send_response(code)
                self.send_header("Content-Type", "application/json")
                self.send_header("Content-Length", str(len(payload)))
                self.end_headers()
                self.wfile.write(payload)

            def do_POST(self):
                if urllib.parse.urlparse(self.path).path != "/api/submit":
                    return self._reply(404, "")
                n = int(self.headers.get("Content-Length", "0"))
                try:
                    code, body = daemon.submit_job(json.loads(self.rfile.read(n)))
                except (KeyError, ValueError) as exc:
                    code, body = 400, json.dumps({"error": repr(exc)})
                self._reply(code, body)

            def do_GET(self):
                url = urllib.parse.urlparse(self.path)
                secret = urllib.parse.parse_qs(url.query).get("secret", [None])[0]
                routes = {"/api/state": daemon.job_state, "/api/cancel": daemon.cancel,
                          "/api/destroy": daemon.destroy_job}
                fn = routes.get(url.path)
                if fn is None:
                    return self._reply(404, "")
                self._reply(*fn(secret))

        return Handler

    def start(self) -> int:
        """Start serving in a background thread; returns the bound port (0 -> OS-assigned)."""
        self.server = ThreadingHTTPServer((self.host, self.port), self._make_handler())
        self.port = self.server.server_address[1]
        threading.Thread(target=self.server.serve_forever, daemon=True).start()
        return self.port

    def run(self) -> None:
        self.server = ThreadingHTTPServer((self.host, self.port), self._make_handler())
        self.port = self.server.server_address[1]
        self.server.serve_forever()

    def shutdown(self) -> None:
        if self.server is not None:
            self.server.shutdown()
            self.server.server_close()
            self.server = None


class Job:
    """Client side of a remote job (``job_deployment.py:284-356``)."""

    def __init__(self, secret, job_name, data_path, num_executors=1, num_processes=1, trainer=None):
        self.secret = secret
        self.job_name = job_name
        self.num_executors = int(num_executors)
        self.num_processes = int(num_processes)
        self.data_path = data_path
        self.trainer = trainer
        self.trained_model = None
        self.history = None
        self.error = None
        self.address: Optional[str] = None
        self.poll_interval = 10.0
        self._thread: Optional[threading.Thread] = None

    def set_num_executors(self, num_executors: int) -> None:
        self.num_executors = int(num_executors)

    def set_num_processes(self, num_processes: int) -> None:
        self.num_processes = int(num_processes)

    def get_trained_model(self):
        return self.trained_model

    def get_history(self):
        return self.history

    def _get(self, route: str):
        url = f"{self.address}/api/{route}?secret={urllib.parse.quote(self.secret)}"
        try:
            with urllib.request.urlopen(url, timeout=30) as r:
                return r.status, r.read().decode("utf-8")
        except urllib.error.HTTPError as e:
            return e.code, ""

    def is_finished(self) -> bool:
        code, body = self._get("state")
        if code != 200:
            return True
        return not json.loads(body)["running"]

    def destroy_remote_job(self) -> None:
        code, body = self._get("destroy")
        if code == 200:
            d = json.loads(body)
            self.error = d.get("error")
            if d.get("model"):
                self.trained_model = deserialize_keras_model(_unhex(d["model"]))
            self.history = _unhex(d["history"]) if d.get("history") else None

    def start(self) -> None:
        self._thread = threading.Thread(target=self.run, daemon=True)
        self._thread.start()

    def wait_completion(self) -> None:
        if self._thread is not None:
            self._thread.join()

    def cancel(self) -> None:
        self._get("cancel")

    def send(self, address: str) -> None:
        """Submit to ``http://host:port`` and start polling (``job_deployment.py:338-350``)."""
        self.address = address.rstrip("/")
        data = {"secret": self.secret, "job_name": self.job_name, "num_executors": self.num_executors,
                "num_processes": self.num_processes, "data_path": self.data_path,
                "trainer": binascii.hexlify(self.trainer.serialize()).decode("ascii")}
        req = urllib.request.Request(self.address + "/api/submit", data=json.dumps(data).encode("utf-8"),
                                     headers={"Content-Type": "application/json"}, method="POST")
        try:
            with urllib.request.urlopen(req, timeout=60) as r:
                if r.status != 200:
                    raise RuntimeError(f"submit rejected with HTTP {r.status}")
        except urllib.error.HTTPError as e:
            raise RuntimeError(f"submit rejected with HTTP {e.code}") from e
        self.start()

    def run(self) -> None:
        while not self.is_finished():
            time.sleep(self.poll_interval)
        self.destroy_remote_job()
