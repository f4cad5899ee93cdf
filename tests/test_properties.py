"""Property-based tests (hypothesis): wire protocol round trips, transformer formulas, history utilities and
the parameter-server algebra on arbitrary inputs."""
import socket
import threading

import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st
from hypothesis.extra import numpy as hnp

from distkeras_b200 import networking
from distkeras_b200.data import Dataset
from distkeras_b200.models import Dense, Sequential
from distkeras_b200.parameter_servers import DeltaParameterServer, DynSGDParameterServer
from distkeras_b200.transformers import LabelIndexTransformer, MinMaxTransformer, OneHotTransformer
from distkeras_b200.utils import history_executors_average

FAST = settings(max_examples=40, deadline=None, derandomize=True)  # deterministic: the driver runs with -x
floats = st.floats(-1e3, 1e3, allow_nan=False, width=32)
arrays = hnp.arrays(dtype=st.sampled_from([np.float32, np.int32, np.uint8, np.float64]),
                    shape=hnp.array_shapes(min_dims=0, max_dims=3, min_side=0, max_side=5))
payloads = st.recursive(st.one_of(arrays, st.integers(-2**31, 2**31), floats, st.text(max_size=8), st.none()),
                        lambda kids: st.one_of(st.lists(kids, max_size=3), st.dictionaries(st.text(max_size=4), kids, max_size=3)),
                        max_leaves=8)


def _same(a, b) -> bool:
    if isinstance(a, np.ndarray):
        return isinstance(b, np.ndarray) and a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=a.dtype.kind == "f")
    if isinstance(a, dict):
        return isinstance(b, dict) and a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return isinstance(b, (list, tuple)) and len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    if isinstance(a, float):
        return a == b or (np.isnan(a) and np.isnan(b))
    return a == b


@FAST
@given(payloads)
def test_wire_protocol_roundtrips_any_payload(data):
    a, b = socket.socketpair()
    try:
        t = threading.Thread(target=networking.send_data, args=(a, data))
        t.start()
        got = networking.recv_data(b)
        t.join()
        assert _same(data, got)
    finally:
        a.close()
        b.close()


@FAST
@given(hnp.arrays(np.float32, hnp.array_shapes(min_dims=2, max_dims=2, min_side=1, max_side=6), elements=floats),
       st.floats(-10, 10), st.floats(0.5, 100), st.floats(-5, 5), st.floats(0.5, 50))
def test_minmax_matches_reference_formula(x, o_min, o_span, n_min, n_span):
    o_max, n_max = o_min + o_span, n_min + n_span
    out = MinMaxTransformer(o_min, o_max, n_min, n_max, "f", "g").transform(Dataset({"f": x}))["g"].numpy()
    scale = (n_max - n_min) / (o_max - o_min)
    assert np.allclose(out, scale * (x - o_max) + n_max, rtol=1e-4, atol=1e-3)  # transformers.py:58-74


@FAST
@given(st.lists(st.integers(0, 6), min_size=1, max_size=12))
def test_onehot_then_label_index_is_identity(labels):
    ds = Dataset({"label": np.asarray(labels, dtype=np.int64)})
    enc = OneHotTransformer(7, "label", "enc").transform(ds)
    assert enc["enc"].sum(1).tolist() == [1.0] * len(labels)
    idx = LabelIndexTransformer(7, input_col="enc", output_col="idx").transform(enc)
    assert idx["idx"].long().tolist() == labels


@FAST
@given(st.integers(1, 4), st.integers(1, 6), st.data())
def test_history_average_is_per_iteration_mean(workers, iters, data):
    vals = data.draw(hnp.arrays(np.float32, (workers, iters, 2), elements=st.floats(0, 10, width=32)))
    hist = [{"history": vals[w, i].tolist(), "worker_id": w, "iteration": i + 1} for w in range(workers) for i in range(iters)]
    avg = history_executors_average(hist)
    assert len(avg) == iters
    assert np.allclose(np.stack(avg), vals.mean(0), atol=1e-5)


def _ps(cls):
    m = Sequential([Dense(3, input_shape=(2,))], seed=0)
    ps = cls(m, None)
    ps.initialize_inproc()
    return ps


P = Sequential([Dense(3, input_shape=(2,))], seed=0).num_params  # flat length (segments are 8-element aligned)


@FAST
@given(st.lists(hnp.arrays(np.float32, (P,), elements=st.floats(-4, 4, width=32)), min_size=1, max_size=6))
def test_delta_server_center_is_initial_plus_sum_of_commits(deltas):
    ps = _ps(DeltaParameterServer)
    c0 = ps.center_variable.clone()
    for i, d in enumerate(deltas):
        ps.apply_commit({"worker_id": i % 2, "delta": d})
    assert torch.allclose(ps.center_variable, c0 + torch.from_numpy(np.sum(deltas, axis=0)), atol=1e-4)
    assert ps.get_num_updates() == 1 + len(deltas)


@FAST
@given(st.lists(st.tuples(hnp.arrays(np.float32, (P,), elements=st.floats(-4, 4, width=32)), st.integers(0, 3)),
                min_size=1, max_size=6))
def test_dynsgd_scales_by_staleness(commits):
    ps = _ps(DynSGDParameterServer)
    want = ps.center_variable.clone()
    n = 1
    for r, lag in commits:
        last = max(1, n - lag)  # the update counter this worker saw at its pull
        want += torch.from_numpy(r) / float((n - last) + 1)
        ps.apply_commit({"worker_id": 0, "residual": r, "last_update": last})
        n += 1
    assert torch.allclose(ps.center_variable, want, atol=1e-4)


@FAST
@given(st.integers(1, 40000), st.integers(1, 5000), st.integers(1, 300000))
def test_gemm_tile_and_split_heuristics_fill_one_wave(M, Nn, K):
    """Host-side scheduling rules of the tcgen05 GEMM (csrc/gemm_tcgen05.cu): legal tile widths, and a split-K
    factor whose tiles x splits never spill into a second, nearly empty wave."""
    from distkeras_b200 import _native

    try:
        lib = _native.lib()
    except RuntimeError:
        import pytest

        pytest.skip("native library not built")
    bn = lib.dk_gemm_pick_bn(Nn)
    assert bn in (16, 32, 64, 128, 256) and (bn >= min(Nn, 128) or bn == 128)
    bn2 = lib.dk_gemm_pick_bn2(M, Nn)
    assert bn2 in (16, 32, 64, 128)
    bk = lib.dk_gemm_pick_bn_splitk(M, Nn, K)
    assert bk in (64, 128, 256)
    splits = lib.dk_gemm_pick_splits(M, Nn, K, bk, 0)
    tiles = ((M + 127) // 128) * ((Nn + bk - 1) // bk)
    slots = 148 if bk > 128 else 296
    assert 1 <= splits <= 32
    assert splits == 1 or tiles * splits <= slots          # single wave
    assert splits == 1 or splits <= max(1, ((K + 63) // 64) // 4)  # at least 4 k-blocks per slice
    ps = lib.dk_gemm_pick_splits_pair(M, Nn, K, 256)
    pair_tiles = ((M + 255) // 256) * ((Nn + 255) // 256)
    assert 1 <= ps <= 32 and (ps == 1 or pair_tiles * ps <= 74)
