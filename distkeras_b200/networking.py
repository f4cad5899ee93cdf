"""Wire protocol of the socket parameter server (CPU / multi-host control path).

Capability parity with ``distkeras/networking.py``: ``determine_host_address``, ``recvall``,
``recv_data``, ``send_data``, ``connect``.  The reference frames every message as a 20-byte
zero-padded ASCII length + a pickle (``networking.py:42-86``) and re-allocates the receive
buffer on every chunk.  Here a message is an 16-byte binary header (magic, payload length) followed
by a copy-free encoding: a small JSON metadata document plus the raw bytes of every
tensor, received straight into a preallocated ``bytearray`` with ``recv_into``.  The metadata is JSON (dtype /
shape / scalars only), never a pickle: a peer that can reach the port cannot make the server execute code.

On the GPU path none of this is used: commits and pulls are loads / atomics issued from CUDA
kernels over NVLink (``csrc/ps_kernels.cu``).
"""
from __future__ import annotations

import json
import socket
import struct
from typing import Any

import numpy as np

_MAGIC = 0x444B3230  # "DK20"
_HEADER = struct.Struct("!IIQ")  # magic, metadata length, raw payload length


def determine_host_address() -> str:
    """Address other processes can reach this host on (``networking.py:11-15``); falls back to
    loopback when the hostname does not resolve (containers)."""
    try:
        return socket.gethostbyname(socket.gethostname())
    except OSError:
        return "127.0.0.1"


def recvall(connection: socket.socket, num_bytes: int) -> bytearray:
    """Read exactly ``num_bytes`` (``networking.py:18-39``) without quadratic re-allocation."""
    buf = bytearray(num_bytes)
    view = memoryview(buf)
    got = 0
    while got < num_bytes:
        n = connection.recv_into(view[got:], num_bytes - got)
        if n == 0:
            raise ConnectionError("socket closed while receiving")
        got += n
    return buf


def _encode(data: Any):
    """Split ``data`` into (metadata, [raw buffers]); numpy arrays travel as raw bytes."""
    buffers = []

    def walk(obj):
        if isinstance(obj, np.ndarray):
            flat = np.ascontiguousarray(obj).reshape(-1)  # (ascontiguousarray alone would turn 0-d into 1-d)
            buffers.append(memoryview(flat.view(np.uint8)))
            return {"__nd__": len(buffers) - 1, "dtype": flat.dtype.str, "shape": list(obj.shape)}
        if isinstance(obj, dict):
            return {"__map__": [[walk(k), walk(v)] for k, v in obj.items()]}
        if isinstance(obj, tuple):
            return {"__tuple__": [walk(v) for v in obj]}
        if isinstance(obj, (np.integer, np.floating, np.bool_)):
            return obj.item()
        if isinstance(obj, bytes):
            return {"__bytes__": obj.hex()}
        if isinstance(obj, (list, tuple)):
            return [walk(v) for v in obj]
        try:
            import torch

            if isinstance(obj, torch.Tensor):
                return walk(obj.detach().cpu().numpy())
        except ImportError:  # pragma: no cover
            pass
        return obj

    return walk(data), buffers


def _hashable(k):
    return tuple(_hashable(x) for x in k) if isinstance(k, list) else k


def _decode(meta: Any, raw: memoryview, offsets):
    def walk(obj):
        if isinstance(obj, dict):
            if "__nd__" in obj:
                lo, hi = offsets[obj["__nd__"]]
                dt = np.dtype(obj["dtype"])
                if dt.hasobject:
                    raise ConnectionError("object arrays are not accepted on the wire")
                return np.frombuffer(raw[lo:hi], dtype=dt).reshape(tuple(obj["shape"]))
            if "__map__" in obj:
                return {_hashable(walk(k)): walk(v) for k, v in obj["__map__"]}
            if "__tuple__" in obj:
                return tuple(walk(v) for v in obj["__tuple__"])
            if "__bytes__" in obj:
                return bytes.fromhex(obj["__bytes__"])
            raise ConnectionError("malformed frame metadata")
        if isinstance(obj, list):
            return [walk(v) for v in obj]
        return obj

    return walk(meta)


def send_data(connection: socket.socket, data: Any) -> None:
    """Send one framed message (``networking.py:65-86``)."""
    meta, buffers = _encode(data)
    sizes = [len(b) for b in buffers]
    meta_bytes = json.dumps({"meta": meta, "sizes": sizes}, allow_nan=True).encode("utf-8")
    connection.sendall(_HEADER.pack(_MAGIC, len(meta_bytes), sum(sizes)))
    connection.sendall(meta_bytes)
    for b in buffers:
        connection.sendall(b)


def recv_data(connection: socket.socket) -> Any:
    """Receive one framed message (``networking.py:42-62``)."""
    magic, meta_len, raw_len = _HEADER.unpack(bytes(recvall(connection, _HEADER.size)))
    if magic != _MAGIC:
        raise ConnectionError("bad frame magic")
    info = json.loads(bytes(recvall(connection, meta_len)).decode("utf-8"))
    raw = memoryview(recvall(connection, raw_len)) if raw_len else memoryview(b"")
    offsets, pos = [], 0
    for s in info["sizes"]:
        offsets.append((pos, pos + s))
        pos += s
    return _decode(info["meta"], raw, offsets)


def connect(host: str, port: int, disable_nagle: bool = True) -> socket.socket:
    """Open a TCP connection to the parameter server (``networking.py:89-99``)."""
    fd = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    if disable_nagle:
        fd.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
    fd.connect((host, port))
    return fd
