#!/usr/bin/env python
"""Print a ``secrets.json`` entry for an identity (``scripts/generate_secret.py`` of the reference).

    python scripts/generate_secret.py --identity alice
"""
import argparse
import json
import secrets
import string


def generate_secret(identity: str, length: int = 64) -> dict:
    alphabet = string.ascii_uppercase + string.digits
    return {"secret": "".join(secrets.choice(alphabet) for _ in range(length)), "identity": identity}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-i", "--identity", required=True, help="who the secret is issued to")
    a = ap.parse_args()
    print(json.dumps(generate_secret(a.identity)))


if __name__ == "__main__":
    main()
