#!/usr/bin/env python
"""Event producer of the streaming pipeline (counterpart of the reference's ``examples/kafka_producer.py``: rows of the
ATLAS Higgs CSV pushed as JSON messages to the topic ``Machine_Learning`` every few seconds).

Sinks: ``--sink kafka`` (needs the ``kafka`` package and a broker: ``--bootstrap``), ``--sink spool`` (default here, no
broker offline: one JSON-lines file per burst in ``--dir``, renamed into place so a consumer never sees half a file) or
``--sink stdout``.

    python examples/kafka_producer.py --csv /tmp/atlas_higgs_sample.csv --dir /tmp/topic --bursts 5 --rows 1000 --interval 0.2
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

TOPIC = "Machine_Learning"


def read_rows(csv_path, drop=("EventId", "Weight", "Label")):
    """Feature vectors of the CSV (bookkeeping and label columns dropped), as lists of floats."""
    with open(csv_path) as f:
        header = f.readline().strip().split(",")
        keep = [i for i, name in enumerate(header) if name not in drop]
        for line in f:
            cells = line.rstrip("\n").split(",")
            yield [float(cells[i]) for i in keep]


def synthetic_rows(n, seed=3):
    from distkeras_b200.data import synthetic_higgs

    for r in synthetic_higgs(n, seed=seed)["features"].numpy():
        yield r.tolist()


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--csv", default=None, help="ATLAS Higgs CSV (examples/data/make_sample_data.py writes one); synthetic rows if omitted")
    ap.add_argument("--sink", default="spool", choices=["spool", "stdout", "kafka"])
    ap.add_argument("--dir", default=os.path.join("/tmp", TOPIC), help="spool directory (the topic)")
    ap.add_argument("--bootstrap", default="localhost:9092")
    ap.add_argument("--bursts", type=int, default=5)
    ap.add_argument("--rows", type=int, default=1000, help="messages per burst")
    ap.add_argument("--interval", type=float, default=5.0, help="seconds between bursts (the reference uses 5)")
    a = ap.parse_args(argv)
    rows = read_rows(a.csv) if a.csv else synthetic_rows(a.bursts * a.rows)
    producer = None
    if a.sink == "kafka":
        from kafka import KafkaProducer   # not available in this sandbox; kept for a real deployment

        producer = KafkaProducer(bootstrap_servers=a.bootstrap, value_serializer=lambda v: json.dumps(v).encode())
    elif a.sink == "spool":
        os.makedirs(a.dir, exist_ok=True)
    sent = 0
    for burst in range(a.bursts):
        messages = []
        for features in rows:
            messages.append({"features": features})
            if len(messages) == a.rows:
                break
        if not messages:
            break
        if producer is not None:
            for m in messages:
                producer.send(TOPIC, m)
            producer.flush()
        elif a.sink == "spool":
            tmp = os.path.join(a.dir, f".burst_{burst:05d}.jsonl.tmp")
            with open(tmp, "w") as f:
                f.write("\n".join(json.dumps(m) for m in messages) + "\n")
            os.replace(tmp, os.path.join(a.dir, f"burst_{burst:05d}.jsonl"))
        else:
            for m in messages:
                print(json.dumps(m))
        sent += len(messages)
        time.sleep(a.interval)
    if a.sink == "spool":   # end-of-stream marker for the example consumer
        open(os.path.join(a.dir, "_DONE"), "w").close()
    print(f"sent {sent} messages in {burst + 1} bursts to {a.sink}", file=sys.stderr)


if __name__ == "__main__":
    main()
