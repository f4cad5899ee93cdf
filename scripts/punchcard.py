#!/usr/bin/env python
"""Start the Punchcard job daemon (``scripts/punchcard.py`` of the reference).

    python scripts/punchcard.py --port 8000 --secrets secrets.json
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from distkeras_b200.job_deployment import Punchcard  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description="distkeras_b200 job daemon")
    ap.add_argument("-p", "--port", type=int, default=8000, help="port to listen on")
    ap.add_argument("-s", "--secrets", default="secrets.json", help="path of the secrets allow-list")
    ap.add_argument("--host", default="0.0.0.0")
    a = ap.parse_args()
    daemon = Punchcard(secrets_path=a.secrets, port=a.port, host=a.host)
    print(f"punchcard listening on {a.host}:{a.port} (secrets: {a.secrets})")
    daemon.run()


if __name__ == "__main__":
    main()
