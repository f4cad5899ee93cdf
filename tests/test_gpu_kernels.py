"""Numerics of every native sm_100a kernel against a plain PyTorch fp32 reference (needs a B200)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def N():
    from distkeras_b200 import _native

    _native.lib()
    torch.cuda.set_device(0)
    return _native


def st():
    return C.c_void_p(int(torch.cuda.current_stream().cuda_stream))


def bf(x):
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("M,Nn,K", [(128, 128, 64), (1024, 1000, 784), (37, 10, 200), (300, 77, 136), (4096, 200, 1000)])
def test_gemm_k_major(N, M, Nn, K):
    from distkeras_b200.ops.gemm import gemm_tn

    torch.manual_seed(0)
    a, b = bf(torch.randn(M, K, device="cuda")), bf(torch.randn(Nn, K, device="cuda"))
    bias = torch.randn(Nn, device="cuda")
    ref = torch.relu(a.float() @ b.float().t() + bias)
    out = gemm_tn(a, b, bias=bias, relu=True, out_fp32=True)
    assert torch.allclose(out, ref, atol=2e-3 * K ** 0.5, rtol=1e-3)
    out16 = gemm_tn(a, b, bias=bias, relu=True)
    assert torch.allclose(out16.float(), ref, atol=0.05 * K ** 0.5, rtol=2e-2)


@pytest.mark.parametrize("M,Nn,K", [(128, 128, 64), (1000, 784, 1024), (16, 200, 4096), (200, 1000, 512)])
def test_gemm_mn_major_operands(N, M, Nn, K):
    """wgrad form (A and B MN-major) and dgrad form (B MN-major) of the tcgen05 kernel."""
    from distkeras_b200.ops.gemm import gemm_tn

    torch.manual_seed(1)
    a, b = bf(torch.randn(M, K, device="cuda")), bf(torch.randn(Nn, K, device="cuda"))
    ref = a.float() @ b.float().t()
    tol = dict(atol=2e-3 * K ** 0.5, rtol=1e-3)
    assert torch.allclose(gemm_tn(a.t().contiguous(), b.t().contiguous(), a_mn=True, b_mn=True, out_fp32=True), ref, **tol)
    assert torch.allclose(gemm_tn(a, b.t().contiguous(), b_mn=True, out_fp32=True), ref, **tol)
    mask = bf(torch.randn(M, Nn, device="cuda"))
    got = gemm_tn(a, b.t().contiguous(), b_mn=True, mask=mask, out_fp32=True, alpha=1.25)
    assert torch.allclose(got, torch.where(mask.float() > 0, 1.25 * ref, torch.zeros_like(ref)), **tol)


def test_gemm_tf32(N):
    from distkeras_b200.ops.gemm import gemm_tn

    torch.manual_seed(2)
    a, b = torch.randn(1000, 784, device="cuda"), torch.randn(256, 784, device="cuda")
    out = gemm_tn(a, b, tf32=True, out_fp32=True)
    assert torch.allclose(out, a @ b.t(), atol=0.15, rtol=1e-2)


@pytest.mark.parametrize("name", ["sgd", "momentum", "adagrad", "rmsprop", "adam", "adadelta", "adamax", "nadam"])
def test_fused_optimizer_matches_reference(N, name):
    from distkeras_b200.ops.flat_optim import FlatOptimizer

    spec = {"class_name": "sgd", "config": {"lr": 0.05, "momentum": 0.9, "nesterov": True}} if name == "momentum" \
        else {"class_name": name, "config": {}}
    n = 100003
    torch.manual_seed(3)
    w0 = torch.randn(n)
    gpu, cpu = FlatOptimizer(spec, n, "cuda"), FlatOptimizer(spec, n, "cpu")
    wg, wc = w0.cuda(), w0.clone()
    wb = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
    for i in range(4):
        g = torch.randn(n)
        gpu.step(wg, g.cuda(), wb)
        cpu.step(wc, g)
    assert torch.allclose(wg.cpu(), wc, atol=1e-5, rtol=1e-5), float((wg.cpu() - wc).abs().max())
    assert torch.equal(wb, wg.to(torch.bfloat16))


@pytest.mark.parametrize("B,Cc,dense", [(256, 10, False), (100, 2, True), (64, 1000, False), (33, 40, True)])
def test_softmax_xent(N, B, Cc, dense):
    torch.manual_seed(4)
    z = torch.randn(B, Cc, device="cuda") * 3
    labels = torch.randint(0, Cc, (B,), device="cuda")
    y = F.one_hot(labels, Cc).float()
    ldz = (Cc + 7) // 8 * 8
    dz = torch.zeros(B, ldz, dtype=torch.bfloat16, device="cuda")
    probs = torch.zeros(B, Cc, device="cuda")
    hist = torch.zeros(4, 2, device="cuda")
    step = torch.tensor([3], dtype=torch.int32, device="cuda")
    li = labels.to(torch.int32)
    N.check(N.lib().dk_softmax_xent(z.data_ptr(), Cc, None if dense else li.data_ptr(), y.data_ptr() if dense else None,
                                    B, Cc, dz.data_ptr(), ldz, None, 0, probs.data_ptr(), hist.data_ptr(),
                                    step.data_ptr(), 4, st()))
    p = torch.softmax(z, 1)
    assert torch.allclose(probs, p, atol=1e-5)
    assert abs(float(hist[2, 0]) - float(F.cross_entropy(z, labels))) < 1e-3
    assert abs(float(hist[2, 1]) - float((z.argmax(1) == labels).float().mean())) < 1e-6
    assert torch.allclose(dz[:, :Cc].float(), (p - y) / B, atol=2e-3 / B + 1e-5, rtol=1e-2)


def test_input_stage_and_colsum(N):
    torch.manual_seed(5)
    B, Fdim = 100, 30
    x = torch.randint(0, 256, (B, Fdim), dtype=torch.uint8, device="cuda")
    ld = 32
    xb = torch.zeros(B, ld, dtype=torch.bfloat16, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    N.check(N.lib().dk_input_stage(x.data_ptr(), N.IN_U8, B, Fdim, 1 / 255.0, -0.5, xb.data_ptr(), ld, None, 0,
                                   cnt.data_ptr(), None, 0, st()))
    assert int(cnt) == 1
    assert torch.allclose(xb[:, :Fdim].float(), x.float() / 255.0 - 0.5, atol=4e-3)
    assert float(xb[:, Fdim:].abs().max()) == 0.0
    src = bf(torch.randn(1000, 200, device="cuda"))
    out = torch.zeros(200, device="cuda")
    N.check(N.lib().dk_colsum_bf16(src.data_ptr(), 1000, 200, 200, out.data_ptr(), 1.0, st()))
    assert torch.allclose(out, src.float().sum(0), atol=1e-2, rtol=1e-4)


@pytest.mark.parametrize("H,Cin,k,stride,pad", [(28, 1, 3, 1, 0), (32, 8, 3, 1, 1), (16, 16, 3, 2, 1), (8, 8, 1, 1, 0)])
def test_im2col_col2im(N, H, Cin, k, stride, pad):
    torch.manual_seed(6)
    B = 3
    x = bf(torch.randn(B, H, H, Cin, device="cuda"))
    OH = (H + 2 * pad - k) // stride + 1
    K = k * k * Cin
    ld = (K + 7) // 8 * 8
    col = torch.zeros(B * OH * OH, ld, dtype=torch.bfloat16, device="cuda")
    N.check(N.lib().dk_im2col(x.data_ptr(), B, H, H, Cin, k, k, stride, pad, OH, OH, col.data_ptr(), ld, st()))
    # reference: unfold gives (c, kh, kw) order -> reorder to (kh, kw, c)
    u = F.unfold(x.float().permute(0, 3, 1, 2), k, padding=pad, stride=stride)  # [B, C*k*k, L]
    u = u.view(B, Cin, k, k, OH * OH).permute(0, 4, 2, 3, 1).reshape(B * OH * OH, K)
    assert torch.equal(col[:, :K].float(), u)
    dcol = bf(torch.randn(B * OH * OH, ld, device="cuda"))
    dx = torch.zeros(B, H, H, Cin, dtype=torch.bfloat16, device="cuda")
    N.check(N.lib().dk_col2im(dcol.data_ptr(), ld, B, H, H, Cin, k, k, stride, pad, OH, OH, dx.data_ptr(), st()))
    d = dcol[:, :K].float().view(B, OH * OH, k, k, Cin).permute(0, 4, 2, 3, 1).reshape(B, K, OH * OH)
    ref = F.fold(d, (H, H), k, padding=pad, stride=stride).permute(0, 2, 3, 1)
    assert torch.allclose(dx.float(), ref, atol=0.05, rtol=2e-2)


def test_maxpool(N):
    torch.manual_seed(7)
    B, H, Cc = 4, 24, 32
    x = bf(torch.randn(B, H, H, Cc, device="cuda"))
    y = torch.zeros(B, H // 2, H // 2, Cc, dtype=torch.bfloat16, device="cuda")
    N.check(N.lib().dk_maxpool_fwd(x.data_ptr(), B, H, H, Cc, 2, 2, y.data_ptr(), st()))
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.max_pool2d(xr, 2)
    assert torch.equal(y.float(), yr.permute(0, 2, 3, 1))
    dy = bf(torch.randn_like(y.float()))
    dx = torch.zeros_like(x)
    N.check(N.lib().dk_maxpool_bwd(x.data_ptr(), y.data_ptr(), dy.data_ptr(), B, H, H, Cc, 2, 2, dx.data_ptr(), st()))
    yr.backward(dy.float().permute(0, 3, 1, 2))
    assert torch.allclose(dx.float(), xr.grad.permute(0, 2, 3, 1), atol=1e-6)


def test_ps_kernels_same_device(N):
    """commit / pull / exchange / elastic / damped / ticket against the formulas of SURVEY 2.6."""
    lib = N.lib()
    n = 100003
    torch.manual_seed(8)
    c = torch.randn(n, device="cuda")
    ctrl = torch.zeros(N.CTRL_WORDS, dtype=torch.int32, device="cuda")
    w, w1 = torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
    wb = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
    c0 = c.clone()
    N.check(lib.dk_ps_commit(c.data_ptr(), w.data_ptr(), w1.data_ptr(), n, 0.25, None, ctrl.data_ptr(), 2, 7, st()))
    assert torch.allclose(c, c0 + 0.25 * (w - w1), atol=1e-6)
    assert int(ctrl[N.CTRL_NUM_UPDATES]) == 1 and int(ctrl[N.CTRL_HEARTBEAT + 2]) == 1
    last = torch.zeros(1, dtype=torch.int32, device="cuda")
    N.check(lib.dk_ps_pull(c.data_ptr(), w.data_ptr(), w1.data_ptr(), wb.data_ptr(), n, ctrl.data_ptr(), last.data_ptr(), st()))
    assert torch.equal(w, c) and torch.equal(w1, c) and torch.equal(wb, c.to(torch.bfloat16)) and int(last) == 1
    # exchange == commit + pull
    w = c + torch.randn(n, device="cuda") * 0.1
    c0, w0 = c.clone(), w.clone()
    N.check(lib.dk_ps_exchange(c.data_ptr(), w.data_ptr(), w1.data_ptr(), wb.data_ptr(), n, 0.5, None, ctrl.data_ptr(), 0, 1,
                               last.data_ptr(), st()))
    want = c0 + 0.5 * (w0 - c0)
    assert torch.allclose(c, want, atol=1e-6) and torch.allclose(w, want, atol=1e-6) and torch.equal(w, w1)
    assert int(last) == 2
    # elastic: W + C conserved, E = alpha (W - C)
    w = c + torch.randn(n, device="cuda")
    c0, w0 = c.clone(), w.clone()
    N.check(lib.dk_ps_elastic(c.data_ptr(), w.data_ptr(), wb.data_ptr(), n, 0.3, ctrl.data_ptr(), 0, 1, st()))
    e = 0.3 * (w0 - c0)
    assert torch.allclose(w, w0 - e, atol=1e-6) and torch.allclose(c, c0 + e, atol=1e-6)
    # damped exchange (Experimental PS)
    w1 = c - 0.2 * torch.rand(n, device="cuda")
    w = w1 + torch.randn(n, device="cuda") * 0.1
    c0, stale, w0 = c.clone(), w1.clone(), w.clone()
    N.check(lib.dk_ps_damped_exchange(c.data_ptr(), w.data_ptr(), w1.data_ptr(), wb.data_ptr(), n, 0.2, 2.0,
                                      ctrl.data_ptr(), 0, 1, st()))
    r = 0.2 * (w0 - stale) / (2.0 * (c0 - stale) ** 2 + 1.0)
    assert torch.allclose(c, c0 + r, atol=1e-5) and torch.allclose(w, c, atol=1e-6)
    # DynSGD ticket: scale = 1 / (num_updates - last_update + 1)
    nu = int(ctrl[N.CTRL_NUM_UPDATES])
    last.fill_(nu - 2)
    scale = torch.zeros(1, device="cuda")
    N.check(lib.dk_ps_ticket(ctrl.data_ptr(), last.data_ptr(), scale.data_ptr(), st()))
    assert abs(float(scale) - 1.0 / 3.0) < 1e-6 and int(ctrl[N.CTRL_NUM_UPDATES]) == nu + 1
    assert int(ctrl[N.CTRL_STALENESS_HIST + 3]) == 1
    # strict-mode ticket lock round trip
    t = torch.zeros(1, dtype=torch.int32, device="cuda")
    N.check(lib.dk_ps_lock_acquire(ctrl.data_ptr(), t.data_ptr(), st()))
    N.check(lib.dk_ps_lock_release(ctrl.data_ptr(), t.data_ptr(), st()))
    torch.cuda.synchronize()
    assert int(ctrl[N.CTRL_LOCK_SERVING]) == 1
    # averaging kernel on views of one device
    reps = [torch.randn(1024, device="cuda") for _ in range(3)]
    mean = torch.stack(reps).mean(0)
    arr = (C.c_void_p * 3)(*[r.data_ptr() for r in reps])
    N.check(lib.dk_ps_average(arr, 3, 0, 1024, st()))
    for r in reps:
        assert torch.allclose(r, mean, atol=1e-6)


def test_label_index_kernel(N):
    from distkeras_b200.transformers import LabelIndexTransformer

    torch.manual_seed(9)
    p = torch.softmax(torch.randn(500, 10, device="cuda") * 2, 1)
    labels = torch.randint(0, 10, (500,), device="cuda", dtype=torch.int32)
    idx = torch.zeros(500, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    N.check(N.lib().dk_label_index(p.data_ptr(), 500, 10, 0.55, 0, idx.data_ptr(), labels.data_ptr(), cnt.data_ptr(), st()))
    want = LabelIndexTransformer(10).indices(p.cpu(), device="cpu")
    assert torch.equal(idx.cpu().long(), want)
    assert int(cnt) == int((want == labels.cpu().long()).sum())


def test_label_index_transformer_runs_the_kernel_and_matches_the_cpu_rule(N):
    from distkeras_b200.data import Dataset
    from distkeras_b200.transformers import LabelIndexTransformer

    torch.manual_seed(10)
    p = torch.softmax(torch.randn(1000, 7) * 3, 1)
    p[:50] = 0.0                      # no positive entry: default index
    p[50:100] = 0.01
    p[50:100, 3] = 0.6                # two entries over the threshold: the first wins
    p[50:100, 5] = 0.9
    t = LabelIndexTransformer(7, default_index=2, activation_threshold=0.55)
    got = t.transform(Dataset({"prediction": p}))["prediction_index"]
    assert torch.equal(got.long(), t.indices(p, device="cpu"))
    assert set(got[:50].tolist()) == {2.0} and set(got[50:100].tolist()) == {3.0}


def test_gemm_pull_fused_kernel(N):
    """First-layer forward with the weight pull fused in: B operand read by TMA from the (here:
    same-device) center buffer as tf32, local W / W1 / bf16 shadow refreshed by the kernel."""
    torch.manual_seed(11)
    B, K, Nout = 256, 784, 1000
    x = torch.rand(B, K, device="cuda")
    center = torch.randn(Nout, K, device="cuda") * 0.05
    bias = torch.randn(Nout, device="cuda")
    w = torch.zeros(Nout, K, device="cuda")
    w1 = torch.zeros_like(w)
    wb = torch.zeros(Nout, K, dtype=torch.bfloat16, device="cuda")
    out = torch.zeros(B, Nout, dtype=torch.bfloat16, device="cuda")
    ep = N.GemmEpilogue()
    ep.bias, ep.act, ep.d, ep.ldd, ep.alpha = bias.data_ptr(), 1, out.data_ptr(), Nout, 1.0
    N.check(N.lib().dk_gemm_pull(x.data_ptr(), K, center.data_ptr(), K, C.byref(ep), B, Nout, K, w.data_ptr(),
                                 w1.data_ptr(), wb.data_ptr(), st()))
    torch.cuda.synchronize()
    assert torch.equal(w, center) and torch.equal(w1, center) and torch.equal(wb, center.to(torch.bfloat16))
    ref = torch.relu(x @ center.t() + bias)
    assert torch.allclose(out.float(), ref, atol=0.06, rtol=2e-2)


@pytest.mark.parametrize("M,Nn,K,bn", [(128, 256, 64, 256), (1024, 1000, 784, 256), (4096, 200, 1000, 256),
                                        (300, 130, 136, 256), (2048, 128, 512, 128), (37, 1000, 200, 256), (4096, 32, 288, 64), (1000, 64, 32, 64),
                                        (16384, 1000, 784, 256)])
def test_gemm_persistent(N, M, Nn, K, bn):
    """Persistent kernel (double-buffered TMEM accumulator): forward form and dgrad form with mask."""
    from distkeras_b200.ops.gemm import gemm_tn

    torch.manual_seed(12)
    a, b = bf(torch.randn(M, K, device="cuda")), bf(torch.randn(Nn, K, device="cuda"))
    bias = torch.randn(Nn, device="cuda")
    ref = torch.relu(a.float() @ b.float().t() + bias)
    ld = (Nn + 7) // 8 * 8
    if ld == Nn:
        out = gemm_tn(a, b, bias=bias, relu=True, persistent=True, bn=bn)
        assert torch.allclose(out.float(), ref, atol=0.05 * K ** 0.5, rtol=2e-2)
        mask = bf(torch.randn(M, Nn, device="cuda"))
        got = gemm_tn(a, b.t().contiguous(), b_mn=True, mask=mask, alpha=1.5, persistent=True, bn=bn)
        want = torch.where(mask.float() > 0, 1.5 * (a.float() @ b.float().t()), torch.zeros_like(ref))
        assert torch.allclose(got.float(), want, atol=0.05 * K ** 0.5, rtol=2e-2)


@pytest.mark.parametrize("M,Nn,K,bn", [(256, 256, 64, 256), (1000, 1000, 784, 256), (512, 384, 320, 128), (4096, 2048, 1024, 256)])
def test_gemm_cta_pair(N, M, Nn, K, bn):
    """cta_group::2 kernel: a CTA pair computes one 256 x bn tile (each CTA stages half of B)."""
    from distkeras_b200.ops.gemm import gemm_tn

    torch.manual_seed(13)
    a, b = bf(torch.randn(M, K, device="cuda")), bf(torch.randn(Nn, K, device="cuda"))
    bias = torch.randn(Nn, device="cuda")
    out = gemm_tn(a, b, bias=bias, relu=True, pair=True, bn=bn, out_fp32=True)
    ref = torch.relu(a.float() @ b.float().t() + bias)
    assert torch.allclose(out, ref, atol=2e-3 * K ** 0.5, rtol=1e-3)


@pytest.mark.parametrize("M,Nn,K,bn,splits", [(1000, 784, 4096, 256, 4), (512, 256, 1024, 128, 2), (304, 520, 2048, 256, 1)])
def test_gemm_cta_pair_wgrad_form(N, M, Nn, K, bn, splits):
    """cta_group::2 with both operands MN-major and split-K (the weight-gradient GEMM of wide layers)."""
    from distkeras_b200.ops.gemm import gemm_tn

    torch.manual_seed(14)
    at, bt = bf(torch.randn(K, M, device="cuda")), bf(torch.randn(K, Nn, device="cuda"))  # dZ [B, out], X [B, in]
    out = gemm_tn(at, bt, a_mn=True, b_mn=True, pair=True, bn=bn, out_fp32=True, splits=splits)
    ref = at.float().t() @ bt.float()
    assert torch.allclose(out, ref, atol=2e-3 * K ** 0.5, rtol=1e-3)


@pytest.mark.parametrize("H,Cc", [(13, 64), (7, 8), (13, 3)])
def test_maxpool_bwd_odd_size_and_fused_relu(N, H, Cc):
    """Floor-mode pooling of an odd image (uncovered last row / column gets zero gradient) with the
    producer's dReLU fused in (x is a post-ReLU activation)."""
    torch.manual_seed(15)
    B = 3
    x = bf(torch.relu(torch.randn(B, H, H, Cc, device="cuda")))
    OH = H // 2
    y = torch.zeros(B, OH, OH, Cc, dtype=torch.bfloat16, device="cuda")
    N.check(N.lib().dk_maxpool_fwd(x.data_ptr(), B, H, H, Cc, 2, 2, y.data_ptr(), st()))
    dy = bf(torch.randn(B, OH, OH, Cc, device="cuda"))
    dx = torch.full_like(x, 7.0)  # poison: every element must be written
    N.check(N.lib().dk_maxpool_bwd_ex(x.data_ptr(), y.data_ptr(), dy.data_ptr(), B, H, H, Cc, 2, 2, dx.data_ptr(), 1, st()))
    pre = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    F.max_pool2d(torch.relu(pre), 2).backward(dy.float().permute(0, 3, 1, 2))
    ref = pre.grad.permute(0, 2, 3, 1)
    # ties inside a window (several zeros) route differently in torch; compare where the max is unique / positive
    assert torch.allclose(dx.float() * (x.float() > 0), ref * (x.float() > 0), atol=1e-6)
    assert float((dx.float() * (x.float() <= 0)).abs().max()) == 0.0


def test_col2im_fused_mask(N):
    torch.manual_seed(16)
    B, H, Cin, k = 2, 12, 16, 3
    OH = H - k + 1
    K = k * k * Cin
    dcol = bf(torch.randn(B * OH * OH, K, device="cuda"))
    mask = bf(torch.randn(B, H, H, Cin, device="cuda"))
    plain, masked = torch.zeros(B, H, H, Cin, dtype=torch.bfloat16, device="cuda"), torch.zeros(
        B, H, H, Cin, dtype=torch.bfloat16, device="cuda")
    N.check(N.lib().dk_col2im(dcol.data_ptr(), K, B, H, H, Cin, k, k, 1, 0, OH, OH, plain.data_ptr(), st()))
    N.check(N.lib().dk_col2im_ex(dcol.data_ptr(), K, B, H, H, Cin, k, k, 1, 0, OH, OH, masked.data_ptr(), mask.data_ptr(), st()))
    assert torch.equal(masked, torch.where(mask.float() > 0, plain, torch.zeros_like(plain)))


@pytest.mark.parametrize("B,Cc,K,dense,use_mask", [(256, 10, 200, False, True), (100, 2, 504, True, False), (64, 10, 200, False, True),
                                                  (33, 16, 64, False, True), (4096, 10, 1024, False, True)])
def test_fused_classifier_head(N, B, Cc, K, dense, use_mask):
    """logits GEMM + softmax cross-entropy + dgrad (+ dReLU mask) in one kernel vs PyTorch fp32."""
    torch.manual_seed(17)
    h = bf(torch.relu(torch.randn(B, K, device="cuda")) if use_mask else torch.randn(B, K, device="cuda"))
    w = bf(torch.randn(Cc, K, device="cuda") * 0.1)
    bias = torch.randn(Cc, device="cuda")
    labels = torch.randint(0, Cc, (B,), device="cuda")
    y = F.one_hot(labels, Cc).float()
    ldz = (Cc + 7) // 8 * 8
    dz = torch.full((B, ldz), 9.0, dtype=torch.bfloat16, device="cuda")
    dh = torch.full((B, K), 9.0, dtype=torch.bfloat16, device="cuda")
    hist = torch.zeros(4, 2, device="cuda")
    step = torch.tensor([2], dtype=torch.int32, device="cuda")
    li = labels.to(torch.int32)
    alpha = 1.25
    N.check(N.lib().dk_dense_softmax_head(h.data_ptr(), K, w.data_ptr(), K, bias.data_ptr(),
                                          None if dense else li.data_ptr(), y.data_ptr() if dense else None, B, Cc, K,
                                          dz.data_ptr(), ldz, dh.data_ptr(), K, alpha, int(use_mask), hist.data_ptr(),
                                          step.data_ptr(), 4, st()))
    z = h.float() @ w.float().t() + bias
    p = torch.softmax(z, 1)
    assert abs(float(hist[1, 0]) - float(F.cross_entropy(z, labels))) < 2e-3
    assert abs(float(hist[1, 1]) - float((z.argmax(1) == labels).float().mean())) < 1e-6
    g = (p - y) / B
    assert torch.allclose(dz[:, :Cc].float(), g, atol=2e-3 / B + 1e-6, rtol=1e-2)
    assert float(dz[:, Cc:].float().abs().max()) == 0.0 if ldz > Cc else True
    ref = alpha * (dz[:, :Cc].float() @ w.float())
    if use_mask:
        ref = torch.where(h.float() > 0, ref, torch.zeros_like(ref))
    assert torch.allclose(dh.float(), ref, atol=1e-2 * float(ref.abs().max()) + 1e-8, rtol=2e-2)


@pytest.fixture
def gather_mode(request, N):
    """0 = register-pipelined gather, 1 = cp.async (LDGSTS) ring completing on the stage mbarrier."""
    N.lib().dk_conv_gather_mode(request.param)
    yield request.param
    N.lib().dk_conv_gather_mode(0)


@pytest.mark.parametrize("gather_mode", [0, 1], indirect=True)
@pytest.mark.parametrize("B,H,Cin,Cout,k,stride,pad", [(4, 16, 32, 32, 3, 1, 1), (3, 14, 32, 64, 3, 1, 0),
                                                       (2, 12, 64, 64, 3, 2, 1), (5, 9, 16, 24, 1, 1, 0)])
def test_implicit_gemm_conv_forward_and_dgrad(N, B, H, Cin, Cout, k, stride, pad, gather_mode):
    """conv_gemm_kernel (A operand gathered from the NHWC activation) vs torch conv2d and its input gradient."""
    torch.manual_seed(18)
    OH = (H + 2 * pad - k) // stride + 1
    x = bf(torch.randn(B, H, H, Cin, device="cuda"))
    w = bf(torch.randn(Cout, k, k, Cin, device="cuda") * 0.1)  # [Cout, (kh, kw, c)]
    bias = torch.randn(Cout, device="cuda")
    K = k * k * Cin
    M = B * OH * OH
    out = torch.zeros(M, Cout, dtype=torch.bfloat16, device="cuda")
    ep = N.GemmEpilogue()
    ep.bias, ep.act, ep.d, ep.ldd, ep.alpha = bias.data_ptr(), 1, out.data_ptr(), Cout, 1.0
    N.check(N.lib().dk_conv_gemm(x.data_ptr(), H, H, Cin, OH, OH, k, k, stride, pad, 1, w.data_ptr(), K, C.byref(ep), M, Cout,
                                 K, st()), "conv fwd")
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.float().permute(0, 3, 1, 2)
    pre = F.conv2d(xr, wr, bias, stride=stride, padding=pad)
    ref = torch.relu(pre).permute(0, 2, 3, 1).reshape(M, Cout)
    assert torch.allclose(out.float(), ref, atol=0.03 * K ** 0.5 * 0.1 + 0.02, rtol=2e-2)
    # dgrad: dX = gather(dZ) x flipped weights, with a dReLU-style mask fused
    dz = bf(torch.randn(B, OH, OH, Cout, device="cuda"))
    wd = torch.zeros(Cin, k * k * Cout, dtype=torch.bfloat16, device="cuda")
    N.check(N.lib().dk_conv_weight_flip(w.data_ptr(), K, wd.data_ptr(), k * k * Cout, Cout, Cin, k, k, st()), "flip")
    mask = bf(torch.randn(B * H * H, Cin, device="cuda"))
    dx = torch.zeros(B * H * H, Cin, dtype=torch.bfloat16, device="cuda")
    ep2 = N.GemmEpilogue()
    ep2.d, ep2.ldd, ep2.alpha, ep2.mask, ep2.ld_mask = dx.data_ptr(), Cin, 1.0, mask.data_ptr(), Cin
    N.check(N.lib().dk_conv_gemm(dz.data_ptr(), OH, OH, Cout, H, H, k, k, 1, k - 1 - pad, stride, wd.data_ptr(), k * k * Cout,
                                 C.byref(ep2), B * H * H, Cin, k * k * Cout, st()), "conv dgrad")
    pre.backward(dz.float().permute(0, 3, 1, 2))
    gref = xr.grad.permute(0, 2, 3, 1).reshape(B * H * H, Cin)
    gref = torch.where(mask.float() > 0, gref, torch.zeros_like(gref))
    assert torch.allclose(dx.float(), gref, atol=0.03 * float(gref.abs().max()) + 1e-3, rtol=3e-2)


@pytest.mark.parametrize("B,H,Cin,Cout,k,stride,pad,splits", [(4, 16, 32, 32, 3, 1, 1, 4), (3, 14, 32, 64, 3, 1, 0, 1),
                                                              (2, 12, 64, 64, 3, 2, 1, 3)])
def test_implicit_gemm_conv_wgrad_experimental(N, B, H, Cin, Cout, k, stride, pad, splits):
    """Implicit weight gradient (cp.async gather of the MN-major B operand) vs autograd."""
    torch.manual_seed(19)
    OH = (H + 2 * pad - k) // stride + 1
    x = bf(torch.randn(B, H, H, Cin, device="cuda"))
    dz = bf(torch.randn(B, OH, OH, Cout, device="cuda"))
    K = k * k * Cin
    rows = B * OH * OH
    dw = torch.zeros(Cout, K, dtype=torch.float32, device="cuda")
    N.check(N.lib().dk_conv_wgrad(x.data_ptr(), H, H, Cin, OH, OH, k, k, stride, pad, dz.data_ptr(), Cout, dw.data_ptr(), K,
                                  Cout, rows, splits, st()), "conv wgrad")
    w = torch.zeros(Cout, Cin, k, k, device="cuda", requires_grad=True)
    F.conv2d(x.float().permute(0, 3, 1, 2), w, None, stride=stride, padding=pad).backward(dz.float().permute(0, 3, 1, 2))
    ref = w.grad.permute(0, 2, 3, 1).reshape(Cout, K)  # [Cout, (kh, kw, c)]
    assert torch.allclose(dw, ref, atol=2e-3 * rows ** 0.5, rtol=1e-3)


@pytest.mark.parametrize("B,comm", [(64, 0), (96, 0), (16, 1), (128, 1), (64, 2)])
@pytest.mark.parametrize("opt", ["adam", "sgd"])
def test_fused_dense_backward_update_kernel(N, B, comm, opt):
    """dk_bwd_update: dW = dZ^T X (tcgen05, MN-major operands), db via the ones-tile MMA, optimizer rule in the
    epilogue, bf16 shadow, and (comm != 0) the PS exchange -- against plain fp32 PyTorch."""
    torch.manual_seed(5)
    dims = [(1000, 784), (200, 1000), (10, 200)]              # (n_out, k_in) of the MNIST MLP
    P = sum(o * i + o for o, i in dims)
    W = torch.randn(P, device="cuda") * 0.05
    W1 = W - torch.randn(P, device="cuda") * 0.01            # last pulled center
    center = W1 + torch.randn(P, device="cuda") * 0.02       # other workers moved it meanwhile
    s0, s1 = torch.rand(P, device="cuda") * 0.01, torch.rand(P, device="cuda") * 0.01
    Wb = torch.zeros(P, dtype=torch.bfloat16, device="cuda")
    step = torch.tensor([3], dtype=torch.int32, device="cuda")
    done = torch.zeros(1, dtype=torch.int32, device="cuda")
    ctrl = torch.zeros(N.CTRL_WORDS, dtype=torch.int32, device="cuda")
    last = torch.zeros(1, dtype=torch.int32, device="cuda")
    kind = N.OPT_KINDS[opt]
    lr, p0, p1, eps = (0.001, 0.9, 0.999, 1e-7) if opt == "adam" else (0.05, 0.0, 0.0, 0.0)
    d = N.BwdUpdateDesc()
    d.nlayers, d.batch = len(dims), B
    keep, off, ref = [], 0, {}
    Wr, s0r, s1r = W.clone(), s0.clone(), s1.clone()
    for i, (o, k) in enumerate(dims):
        ldz = (o + 7) // 8 * 8
        dz = torch.zeros(B, ldz, dtype=torch.bfloat16, device="cuda")
        dz[:, :o] = bf(torch.randn(B, o, device="cuda") * 0.1)
        x = bf(torch.randn(B, k, device="cuda"))
        keep += [dz, x]
        L = d.layer[i]
        L.dz, L.lddz, L.x, L.ldx, L.x_slot = dz.data_ptr(), ldz, x.data_ptr(), k, -1
        L.n_out, L.k_in, L.w_off, L.b_off, L.wb_pad, L.ldwb_pad = o, k, off, off + o * k, None, 0
        ref[i] = (dz[:, :o].float().t() @ x.float(), dz[:, :o].float().sum(0), off, o, k)
        off += o * k + o
    d.w, d.s0, d.s1, d.w1, d.wb = W.data_ptr(), s0.data_ptr() if opt == "adam" else None, \
        s1.data_ptr() if opt == "adam" else None, W1.data_ptr(), Wb.data_ptr()
    d.opt_kind, d.lr, d.p0, d.p1, d.eps, d.decay, d.nesterov = kind, lr, p0, p1, eps, 0.0, 0
    d.step, d.done_counter, d.step_inc = step.data_ptr(), done.data_ptr(), 1
    d.comm_mode, d.comm_scale, d.alpha, d.nshards, d.shard_per = comm, 0.25, 0.3, 1, P
    d.shard_center[0] = center.data_ptr()
    d.ctrl, d.worker, d.last_update = ctrl.data_ptr(), 2, last.data_ptr()
    W1_0, center_0 = W1.clone(), center.clone()
    N.check(N.lib().dk_bwd_update(C.byref(d), st()), "dk_bwd_update")
    torch.cuda.synchronize()
    # reference: gradient, optimizer rule at t = 3, then the exchange
    G = torch.zeros(P, device="cuda")
    for i, (gw, gb, o_, o, k) in ref.items():
        G[o_:o_ + o * k] = gw.reshape(-1)
        G[o_ + o * k:o_ + o * k + o] = gb
    if opt == "adam":
        s0r = p0 * s0r + (1 - p0) * G
        s1r = p1 * s1r + (1 - p1) * G * G
        corr = (1 - p1 ** 3) ** 0.5 / (1 - p0 ** 3)
        Wr = Wr - lr * corr * s0r / (s1r.sqrt() + eps)
    else:
        Wr = Wr - lr * G
    if comm == 1:
        r = (Wr - W1_0) * 0.25
        center_r = center_0 + r
        Wr = center_r.clone()
        assert torch.allclose(center, center_r, atol=1e-5) and torch.allclose(W1, center_r, atol=1e-5)
        assert int(ctrl[N.CTRL_NUM_UPDATES]) == 1 and int(ctrl[N.CTRL_HEARTBEAT + 2]) == 1 and int(last) == 1
    elif comm == 2:
        E = 0.3 * (Wr - center_0)
        Wr = Wr - E
        assert torch.allclose(center, center_0 + E, atol=1e-5)
    else:
        assert torch.equal(center, center_0)
    # bf16 operands, fp32 accumulation over B <= 128 terms: tight tolerances
    tol = 5e-3 if opt == "adam" else 2e-4    # Adam normalises tiny gradients: rounding of g shows in the step
    assert torch.allclose(W, Wr, atol=tol, rtol=1e-3), float((W - Wr).abs().max())
    if opt == "adam":
        assert torch.allclose(s0, s0r, atol=1e-4, rtol=1e-3) and torch.allclose(s1, s1r, atol=1e-5, rtol=1e-2)
    assert torch.allclose(Wb.float(), W, atol=1e-2, rtol=1e-2)
    assert int(step) == 4 and int(done) == 0


@pytest.mark.parametrize("B,H,Cin,Cout,k,stride,pad", [(4, 16, 32, 32, 3, 1, 1), (3, 14, 32, 64, 3, 1, 0), (2, 12, 64, 64, 3, 2, 1),
                                                       (5, 9, 64, 24, 1, 1, 0), (2, 30, 32, 32, 3, 1, 0), (3, 10, 128, 160, 3, 1, 1)])
def test_tma_im2col_conv_forward_and_dgrad(N, B, H, Cin, Cout, k, stride, pad):
    """Persistent implicit-GEMM convolution whose A operand is produced by TMA im2col descriptors
    (cp.async.bulk.tensor.4d...im2col): forward with bias + ReLU, and (stride 1) the input gradient with a fused
    dReLU mask, against F.conv2d / autograd."""
    torch.manual_seed(23)
    OH = (H + 2 * pad - k) // stride + 1
    x = bf(torch.randn(B, H, H, Cin, device="cuda"))
    w = bf(torch.randn(Cout, k, k, Cin, device="cuda") * 0.1)
    bias = torch.randn(Cout, device="cuda")
    K, M = k * k * Cin, B * OH * OH
    ldo = (Cout + 7) // 8 * 8
    out = torch.zeros(M, ldo, dtype=torch.bfloat16, device="cuda")
    ep = N.GemmEpilogue()
    ep.d, ep.ldd, ep.alpha, ep.bias, ep.act = out.data_ptr(), ldo, 1.0, bias.data_ptr(), 1
    N.check(N.lib().dk_conv_tma(x.data_ptr(), B, H, H, Cin, OH, OH, k, k, stride, pad, w.data_ptr(), K, C.byref(ep), M, Cout,
                                st()), "conv tma fwd")
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.float().permute(0, 3, 1, 2)
    pre = F.conv2d(xr, wr, bias, stride=stride, padding=pad)
    ref = torch.relu(pre).permute(0, 2, 3, 1).reshape(M, Cout)
    assert torch.allclose(out[:, :Cout].float(), ref, atol=0.03 * K ** 0.5 * 0.1 + 0.02, rtol=2e-2)
    if stride != 1 or not N.lib().dk_conv_tma_supported(Cout, 1, Cin, Cin, 0):
        return
    dz = bf(torch.randn(B, OH, OH, Cout, device="cuda"))
    wd = torch.zeros(Cin, k * k * Cout, dtype=torch.bfloat16, device="cuda")
    N.check(N.lib().dk_conv_weight_flip(w.data_ptr(), K, wd.data_ptr(), k * k * Cout, Cout, Cin, k, k, st()), "flip")
    mask = bf(torch.randn(B * H * H, Cin, device="cuda"))
    dx = torch.zeros(B * H * H, Cin, dtype=torch.bfloat16, device="cuda")
    ep2 = N.GemmEpilogue()
    ep2.d, ep2.ldd, ep2.alpha, ep2.mask, ep2.ld_mask = dx.data_ptr(), Cin, 1.0, mask.data_ptr(), Cin
    N.check(N.lib().dk_conv_tma(dz.data_ptr(), B, OH, OH, Cout, H, H, k, k, 1, k - 1 - pad, wd.data_ptr(), k * k * Cout,
                                C.byref(ep2), B * H * H, Cin, st()), "conv tma dgrad")
    pre.backward(dz.float().permute(0, 3, 1, 2))
    gref = xr.grad.permute(0, 2, 3, 1).reshape(B * H * H, Cin)
    gref = torch.where(mask.float() > 0, gref, torch.zeros_like(gref))
    assert torch.allclose(dx.float(), gref, atol=0.03 * float(gref.abs().max()) + 1e-3, rtol=3e-2)


@pytest.mark.parametrize("B,H,Cin,Cout,k,stride,pad", [(4, 16, 32, 32, 3, 1, 1), (3, 14, 32, 64, 3, 1, 0), (2, 12, 64, 64, 3, 2, 1),
                                                       (6, 9, 64, 24, 1, 1, 0), (3, 10, 128, 128, 3, 1, 1)])
def test_tma_im2col_conv_wgrad(N, B, H, Cin, Cout, k, stride, pad):
    """Weight gradient with the im2col operand produced by TMA (split-K over a persistent grid, the gradient block
    resident in TMEM, TMA reduce-add) + bias gradient from the ones-tile MMA, against autograd."""
    torch.manual_seed(29)
    OH = (H + 2 * pad - k) // stride + 1
    x = bf(torch.randn(B, H, H, Cin, device="cuda"))
    dz = bf(torch.randn(B, OH, OH, Cout, device="cuda"))
    K = k * k * Cin
    dw = torch.zeros(Cout, K, dtype=torch.float32, device="cuda")
    db = torch.zeros(Cout, dtype=torch.float32, device="cuda")
    assert N.lib().dk_conv_wgrad_tma_supported(Cin, Cout, Cout, K)
    N.check(N.lib().dk_conv_wgrad_tma(x.data_ptr(), B, H, H, Cin, OH, OH, k, k, stride, pad, dz.data_ptr(), Cout, dw.data_ptr(),
                                      K, Cout, db.data_ptr(), st()), "conv wgrad tma")
    w = torch.zeros(Cout, Cin, k, k, device="cuda", requires_grad=True)
    F.conv2d(x.float().permute(0, 3, 1, 2), w, None, stride=stride, padding=pad).backward(dz.float().permute(0, 3, 1, 2))
    ref = w.grad.permute(0, 2, 3, 1).reshape(Cout, K)  # [Cout, (kh, kw, c)]
    rows = B * OH * OH
    assert torch.allclose(dw, ref, atol=2e-3 * rows ** 0.5, rtol=1e-3), float((dw - ref).abs().max())
    assert torch.allclose(db, dz.float().sum(dim=(0, 1, 2)), atol=2e-3 * rows ** 0.5, rtol=1e-3)
