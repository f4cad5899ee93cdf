"""``python -m distkeras_b200 {info,build,bench}`` -- small operator CLI.

* ``info``   devices, native library status, available backends and the environment switches in effect
* ``build``  compile the sm_100a kernels (same as ``python build_native.py``)
* ``bench``  forward the remaining arguments to the repository's ``bench.py``
"""
from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SWITCHES = ("DK_CONV_LDGSTS", "DK_BACKEND", "DK_COMM", "DK_STRICT", "DK_DEDICATED_PS", "DK_LOG", "DK_NVTX", "DK_NUMA", "DK_FAULT",
            "DK_FAULT_KILL", "DK_PERSISTENT", "DK_PAIR", "DK_PDL", "DK_SIDE_STREAMS", "DK_FUSED_HEAD", "DK_IMPLICIT_CONV",
            "DK_IMPLICIT_WGRAD", "DK_EXPERIMENTAL", "DK_COMPACT", "DK_COMPACT_MAX_BATCH", "DK_HEAD_IN_FWD", "DK_GEMM_KCH",
            "DK_SPLIT_M", "DK_GEMM_MCAST", "DK_CONV_TMA", "DK_CONV_WRES", "DK_PAD_INPUT_CHANNELS", "DK_BARRIER_TIMEOUT_MS",
            "DK_TRACE")


def info() -> dict:
    import torch

    from . import _native

    lib_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libdistkeras_b200.so")
    out = {"torch": torch.__version__, "cuda_available": torch.cuda.is_available(),
           "native_library": lib_path if os.path.exists(lib_path) else None,
           "backends": ["thread", "socket", "nccl"] + (["fabric"] if torch.cuda.is_available() else []),
           "switches": {k: os.environ[k] for k in SWITCHES if k in os.environ}, "devices": []}
    if torch.cuda.is_available():
        for i in range(torch.cuda.device_count()):
            p = torch.cuda.get_device_properties(i)
            out["devices"].append({"index": i, "name": p.name, "sm": f"{p.major}.{p.minor}", "sms": p.multi_processor_count,
                                   "memory_gib": round(p.total_memory / 2**30, 1)})
        try:
            _native.lib()
            out["native_loaded"] = True
        except Exception as exc:  # report, do not hide
            out["native_loaded"] = f"failed: {exc}"
    return out


def main(argv) -> int:
    cmd = argv[0] if argv else "info"
    if cmd == "info":
        print(json.dumps(info(), indent=1))
        return 0
    if cmd == "build":
        return subprocess.call([sys.executable, os.path.join(ROOT, "build_native.py"), *argv[1:]])
    if cmd == "bench":
        return subprocess.call([sys.executable, os.path.join(ROOT, "bench.py"), *argv[1:]])
    print(__doc__)
    return 2


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
