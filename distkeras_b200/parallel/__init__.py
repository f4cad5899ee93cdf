"""Parallel runtime: NVLink fabric, native step engine, replicas, process launcher."""
