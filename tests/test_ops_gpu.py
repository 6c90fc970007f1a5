"""GPU parity of every hand-written kernel, called through the C ABI, against plain torch fp32 on the same inputs.
Tolerances: inputs/outputs are bf16 with fp32 accumulation, so one op carries ~2^-9 relative rounding on its output
(rel-L2 <= 4e-3 asserted; measured ~1.7e-3 for GEMM/conv, ~2.2e-3 for attention)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 4e-3


@pytest.fixture(scope="module")
def lib():
    from lightdiffusion_next_b200 import _lib as L
    return L, L.load()


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


@pytest.mark.parametrize("M,N,K,bias,res,BN", [(128, 128, 64, False, False, 0), (256, 320, 320, True, False, 0),
                                               (1000, 640, 1280, True, True, 0), (4096, 1280, 2560, True, True, 256),
                                               (512, 64, 128, False, False, 64), (77, 768, 768, True, False, 0),
                                               (2048, 960, 1280, False, True, 0)])
def test_gemm(lib, M, N, K, bias, res, BN):
    L, l = lib
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda") if bias else None
    R = torch.randn(M, N, device="cuda").bfloat16() if res else None
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    L.check(l.ldn_gemm_bf16(A.data_ptr(), K, K, None, 0, 0, W.data_ptr(), M, N, L.ptr(b), None, 0, 0, L.ptr(R), N,
                            out.data_ptr(), N, None, 0, 0, 0, BN, L.cur_stream()))
    ref = A.float() @ W.float().t()
    if bias:
        ref += b
    if res:
        ref += R.float()
    assert rel(out, ref) < TOL


def test_gemm_virtual_concat_rowbias_and_head_slots(lib):
    L, l = lib
    torch.manual_seed(1)
    M, K0, K1, N, d, slot = 512, 640, 320, 320, 40, 64
    A0 = torch.randn(M, K0, device="cuda").bfloat16()
    A1 = torch.randn(M, K1, device="cuda").bfloat16()
    W = (torch.randn(N, K0 + K1, device="cuda") / (K0 + K1) ** 0.5).bfloat16()
    rb = torch.randn(2, N, device="cuda")
    out = torch.zeros(M, (N // d) * slot, device="cuda", dtype=torch.bfloat16)
    L.check(l.ldn_gemm_bf16(A0.data_ptr(), K0, K0, A1.data_ptr(), K1, K1, W.data_ptr(), M, N, None, rb.data_ptr(), N,
                            M // 2, None, 0, out.data_ptr(), out.shape[1], None, 0, d, slot, 0, L.cur_stream()))
    ref = torch.cat([A0, A1], 1).float() @ W.float().t() + rb.repeat_interleave(M // 2, 0)
    got = out.view(M, N // d, slot)
    assert rel(got[:, :, :d].reshape(M, N), ref) < TOL
    assert float(got[:, :, d:].abs().max()) == 0.0  # slot padding untouched


def test_gemm_geglu(lib):
    L, l = lib
    torch.manual_seed(2)
    M, C = 384, 320
    A = torch.randn(M, C, device="cuda").bfloat16()
    W = (torch.randn(8 * C, C, device="cuda") / C ** 0.5).bfloat16()
    b = torch.randn(8 * C, device="cuda") * 0.1
    BN, half = 160, 80
    idx = []
    for t in range(8 * C // BN):
        idx += list(range(t * half, (t + 1) * half)) + list(range(4 * C + t * half, 4 * C + (t + 1) * half))
    idx = torch.tensor(idx, device="cuda")
    Wi, bi = W[idx].contiguous(), b[idx].contiguous()
    out = torch.zeros(M, 4 * C, device="cuda", dtype=torch.bfloat16)
    L.check(l.ldn_gemm_bf16(A.data_ptr(), C, C, None, 0, 0, Wi.data_ptr(), M, 8 * C, bi.data_ptr(), None, 0, 0, None, 0,
                            out.data_ptr(), 4 * C, None, 1, 0, 0, BN, L.cur_stream()))
    h = A.float() @ W.float().t() + b
    a, g = h.chunk(2, dim=-1)
    assert rel(out, a * F.gelu(g)) < TOL


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 64, 64), (2, 8, 8, 128, 320), (1, 32, 32, 320, 320),
                                            (2, 128, 128, 320, 320), (2, 4, 4, 1280, 1280), (3, 24, 40, 64, 128),
                                            (2, 2, 2, 2560, 1280)])
def test_conv3x3(lib, B, H, W, Cin, Cout):
    L, l = lib
    torch.manual_seed(B * H + Cin)
    x = torch.randn(B, H, W, Cin, device="cuda").bfloat16()
    w = (torch.randn(Cout, 3, 3, Cin, device="cuda") / (9 * Cin) ** 0.5).bfloat16()
    b = torch.randn(Cout, device="cuda")
    rb = torch.randn(B, Cout, device="cuda")
    R = torch.randn(B, H, W, Cout, device="cuda").bfloat16()
    out = torch.zeros(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    L.check(l.ldn_conv3x3_bf16(x.data_ptr(), w.data_ptr(), B, H, W, Cin, Cout, b.data_ptr(), rb.data_ptr(), Cout,
                               R.data_ptr(), out.data_ptr(), L.cur_stream()))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b, padding=1)
    ref = ref + rb[:, :, None, None] + R.float().permute(0, 3, 1, 2)
    assert rel(out.permute(0, 3, 1, 2), ref) < TOL


@pytest.mark.parametrize("B,H,W,Cin,Cout,res,expect_fused", [
    (2, 64, 64, 320, 320, False, True),     # conv1 -> gn2 at level 0 of a 512^2 image: 10 channels per group
    (2, 64, 32, 640, 640, True, True),      # conv2 (+ residual) -> transformer.norm, 20 channels per group
    (2, 100, 24, 320, 320, True, True),     # tiles hang over the right image edge: masked pixels count for nothing
    (2, 50, 48, 320, 320, False, True),     # ... and over the bottom edge
    (2, 32, 32, 640, 1280, False, True),    # 40 channels per group
    (3, 32, 32, 960, 640, False, True),     # odd batch, Cin != Cout
    (1, 40, 24, 320, 320, True, False),     # few tiles -> split-K: the statistics kernel runs
    (2, 8, 8, 1280, 1280, False, False),    # split-K: the statistics kernel runs
    (2, 16, 16, 320, 256, False, False),    # not a UNet width: the statistics kernel runs
])
def test_conv3x3_groupnorm_statistics_in_the_epilogue(lib, B, H, W, Cin, Cout, res, expect_fused):
    """SURVEY K4 / ResBlock.py:251-292: conv3x3 (+ time-embedding row bias, + residual) -> GroupNorm(32) + SiLU with the
    statistics taken in the conv's epilogue, against torch fp32 (conv2d -> group_norm -> silu on the bf16-rounded conv
    output, as the engine's two-kernel path and the reference's fp16 path do).  Bit-deterministic."""
    import ctypes
    L, l = lib
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H + Cout)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, 3, 3, Cin, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    b = torch.randn(Cout, device="cuda", generator=g)
    rb = torch.randn(B, Cout, device="cuda", generator=g) * 2
    R = (torch.randn(B, H, W, Cout, device="cuda", generator=g) * 3 + 1).bfloat16() if res else None
    gamma = torch.randn(Cout, device="cuda", generator=g)
    beta = torch.randn(Cout, device="cuda", generator=g)
    conv_out = torch.zeros(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    gn_out = torch.zeros_like(conv_out)
    fused = ctypes.c_int(-1)

    def run():
        L.check(l.ldn_conv3x3_groupnorm_bf16(x.data_ptr(), w.data_ptr(), B, H, W, Cin, Cout, b.data_ptr(), rb.data_ptr(), Cout,
                                             R.data_ptr() if res else 0, 1e-5, gamma.data_ptr(), beta.data_ptr(), 1,
                                             conv_out.data_ptr(), gn_out.data_ptr(), ctypes.addressof(fused), L.cur_stream()))
        torch.cuda.synchronize()
    run()
    assert bool(fused.value) == expect_fused, fused.value
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b, padding=1) + rb[:, :, None, None]
    if res:
        ref = ref + R.float().permute(0, 3, 1, 2)
    assert rel(conv_out.permute(0, 3, 1, 2), ref) < TOL
    gn_ref = F.silu(F.group_norm(conv_out.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-5))
    assert rel(gn_out.permute(0, 3, 1, 2), gn_ref) < TOL
    first = gn_out.clone()
    run()
    assert torch.equal(first, gn_out)  # integer accumulators: bit-deterministic


@pytest.mark.parametrize("B,H,Nq,Nk,d,causal,ones", [(1, 1, 128, 128, 64, False, False), (2, 8, 1024, 1024, 40, False, False),
                                                     (2, 8, 1024, 77, 40, False, False), (2, 8, 256, 256, 160, False, False),
                                                     (2, 8, 1024, 1024, 80, False, False), (2, 8, 256, 77, 160, False, False),
                                                     (2, 8, 64, 64, 160, False, False), (2, 8, 16, 16, 160, False, False),
                                                     (3, 12, 77, 77, 64, True, False), (2, 8, 4096, 154, 40, False, False),
                                                     (1, 8, 4096, 4096, 40, False, False),
                                                     # head dim 40 with the ones-row V^T layout (attention5.cu: row sums on the
                                                     # tensor core, O resident in TMEM, lazy rescale)
                                                     (2, 8, 1024, 1024, 40, False, True), (2, 8, 1024, 77, 40, False, True),
                                                     (1, 8, 4096, 4096, 40, False, True), (2, 8, 64, 64, 40, False, True),
                                                     (2, 8, 4096, 154, 40, False, True), (1, 2, 200, 1000, 40, False, True),
                                                     # head dim 80, ones-row V^T with 96 rows per head (attention6.cu: P aliased over S in TMEM)
                                                     (2, 8, 1024, 1024, 80, False, True), (1, 8, 4096, 4096, 80, False, True),
                                                     (2, 8, 256, 77, 80, False, True), (1, 2, 200, 1000, 80, False, True),
                                                     (2, 3, 64, 64, 80, False, True),
                                                     # head dim 128 (Flux joint attention): attention6.cu with the row sum in registers
                                                     (2, 8, 1024, 1024, 128, False, False), (1, 3, 300, 520, 128, False, False),
                                                     (1, 24, 4352, 4352, 128, False, False)])
def test_attention(lib, B, H, Nq, Nk, d, causal, ones):
    L, l = lib
    torch.manual_seed(Nq + Nk + d)
    slot = (d + 63) // 64 * 64
    nk_pad = (Nk + 127) // 128 * 128 if Nk % 8 else Nk
    hs = {40: 48, 80: 96}[d] if ones else d
    q = torch.randn(B, H, Nq, d, device="cuda").bfloat16()
    k = torch.randn(B, H, Nk, d, device="cuda").bfloat16()
    v = torch.randn(B, H, Nk, d, device="cuda").bfloat16()
    if ones:  # make the running max grow along the keys so the lazy-rescale path is exercised
        k = (k.float() * torch.linspace(0.2, 3.0, Nk, device="cuda")[None, None, :, None]).bfloat16()
    Qb = torch.zeros(B * Nq, H * slot, device="cuda", dtype=torch.bfloat16)
    Kb = torch.zeros(B * nk_pad, H * slot, device="cuda", dtype=torch.bfloat16)
    Qb.view(B, Nq, H, slot)[..., :d] = q.permute(0, 2, 1, 3)
    Kb.view(B, nk_pad, H, slot)[:, :Nk, :, :d] = k.permute(0, 2, 1, 3)
    Vt = torch.zeros(H * hs, B * nk_pad, device="cuda", dtype=torch.bfloat16)
    Vt.view(H, hs, B, nk_pad)[:, :d, :, :Nk] = v.permute(1, 3, 0, 2)
    if ones:
        Vt.view(H, hs, B, nk_pad)[:, d] = 1.0
    out = torch.zeros(B * Nq, H * d, device="cuda", dtype=torch.bfloat16)
    L.check(l.ldn_attention_bf16(Qb.data_ptr(), H * slot, Kb.data_ptr(), H * slot, Vt.data_ptr(), B * nk_pad, H * hs,
                                 hs if ones else 0, B, H, Nq, Nk, nk_pad, d, slot, int(causal), d ** -0.5, out.data_ptr(),
                                 H * d, L.cur_stream()))
    ref = F.scaled_dot_product_attention(q.float(), k.float(), v.float(), is_causal=causal)
    ref = ref.permute(0, 2, 1, 3).reshape(B * Nq, H * d)
    assert rel(out, ref) < TOL

@pytest.mark.parametrize("B,H,Nq,Nk,grow", [(2, 8, 1024, 1024, True), (1, 8, 4096, 4096, False), (1, 2, 200, 1090, True),
                                            (1, 2, 300, 520, True), (2, 8, 1024, 77, False)])
def test_attention_folded_operands(lib, B, H, Nq, Nk, grow):
    """`causal` bit 1 of ldn_attention_bf16 (d = 40, ones-row V^T): Q arrives pre-scaled by scale * log2(e) and column 40 of
    every K head slot holds 1.0; the long-sequence kernel (attention9.cu, folded variant) keeps -m in column 40 of its Q
    tile so that the tensor core delivers offset scores.  This is how the UNet program calls its level-0 self-attention.
    The fp32 reference is built from the SAME rounded, pre-scaled Q (the UNet rounds once, after the scale).  Nk = 77 takes
    the short-sequence kernel, which must ignore the ones column (column 40 of Q is zero in global memory)."""
    L, l = lib
    torch.manual_seed(Nq + Nk)
    d, slot, hs = 40, 64, 48
    qs_scale = d ** -0.5 * 1.4426950408889634
    nk_pad = (Nk + 127) // 128 * 128 if Nk % 8 else Nk
    qs = (torch.randn(B, H, Nq, d, device="cuda") * qs_scale).bfloat16()  # what the projection epilogue would store
    k = torch.randn(B, H, Nk, d, device="cuda").bfloat16()
    v = torch.randn(B, H, Nk, d, device="cuda").bfloat16()
    if grow:  # the running offset has to move several times along the keys
        k = (k.float() * torch.linspace(0.2, 3.0, Nk, device="cuda")[None, None, :, None]).bfloat16()
    Qb = torch.zeros(B * Nq, H * slot, device="cuda", dtype=torch.bfloat16)
    Kb = torch.zeros(B * nk_pad, H * slot, device="cuda", dtype=torch.bfloat16)
    Qb.view(B, Nq, H, slot)[..., :d] = qs.permute(0, 2, 1, 3)
    Kb.view(B, nk_pad, H, slot)[:, :Nk, :, :d] = k.permute(0, 2, 1, 3)
    Kb.view(B, nk_pad, H, slot)[:, :, :, d] = 1.0
    Vt = torch.zeros(H * hs, B * nk_pad, device="cuda", dtype=torch.bfloat16)
    Vt.view(H, hs, B, nk_pad)[:, :d, :, :Nk] = v.permute(1, 3, 0, 2)
    Vt.view(H, hs, B, nk_pad)[:, d] = 1.0
    out = torch.zeros(B * Nq, H * d, device="cuda", dtype=torch.bfloat16)

    def run():
        L.check(l.ldn_attention_bf16(Qb.data_ptr(), H * slot, Kb.data_ptr(), H * slot, Vt.data_ptr(), B * nk_pad, H * hs, hs,
                                     B, H, Nq, Nk, nk_pad, d, slot, 2, 0.0, out.data_ptr(), H * d, L.cur_stream()))
    run()
    ref = F.scaled_dot_product_attention(qs.float() / 1.4426950408889634, k.float(), v.float(), scale=1.0)
    ref = ref.permute(0, 2, 1, 3).reshape(B * Nq, H * d)
    assert rel(out, ref) < TOL
    first = out.clone()
    run()
    torch.cuda.synchronize()
    assert torch.equal(first, out)  # bit-deterministic


@pytest.mark.parametrize("B,HW,C0,C1,silu,eps", [(2, 256, 320, 0, True, 1e-5), (2, 1024, 640, 320, True, 1e-5),
                                                 (2, 64, 1280, 640, True, 1e-5), (1, 4096, 320, 0, False, 1e-6),
                                                 (2, 4, 1280, 1280, True, 1e-5), (3, 100, 128, 0, True, 1e-6)])
def test_groupnorm_virtual_concat(lib, B, HW, C0, C1, silu, eps):
    L, l = lib
    torch.manual_seed(HW + C0)
    C = C0 + C1
    x0 = (torch.randn(B, HW, C0, device="cuda") * 2 + 0.5).bfloat16()
    x1 = (torch.randn(B, HW, C1, device="cuda") - 1.0).bfloat16() if C1 else None
    gamma = torch.randn(C, device="cuda")
    beta = torch.randn(C, device="cuda")
    out = torch.zeros(B, HW, C, device="cuda", dtype=torch.bfloat16)
    L.check(l.ldn_groupnorm_bf16(x0.data_ptr(), C0, L.ptr(x1), C1, B, HW, 32, eps, gamma.data_ptr(), beta.data_ptr(),
                                 int(silu), out.data_ptr(), L.cur_stream()))
    xin = torch.cat([x0, x1], -1) if C1 else x0
    ref = F.group_norm(xin.float().permute(0, 2, 1), 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    assert rel(out.permute(0, 2, 1), ref) < TOL


@pytest.mark.parametrize("rows,C", [(1000, 320), (512, 640), (77, 768), (300, 1280)])
def test_layernorm(lib, rows, C):
    L, l = lib
    torch.manual_seed(rows)
    x = (torch.randn(rows, C, device="cuda") * 3 + 1).bfloat16()
    gamma = torch.randn(C, device="cuda")
    beta = torch.randn(C, device="cuda")
    out = torch.zeros_like(x)
    L.check(l.ldn_layernorm_bf16(x.data_ptr(), rows, C, 1e-5, gamma.data_ptr(), beta.data_ptr(), out.data_ptr(),
                                 L.cur_stream()))
    assert rel(out, F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)) < TOL


def test_cfg_step_matches_torch(lib):
    L, l = lib
    torch.manual_seed(0)
    n = 4 * 128 * 128
    x, u, c, nz = (torch.randn(n, device="cuda") for _ in range(4))
    xo, do = torch.empty_like(x), torch.empty_like(x)
    L.check(l.ldn_cfg_step(x.data_ptr(), u.data_ptr(), c.data_ptr(), 7.0, 0, 0.8, -0.3, 0.0, None, xo.data_ptr(),
                           do.data_ptr(), n, L.cur_stream()))
    den = torch.lerp(u, c, 7.0)
    assert torch.allclose(do, den, rtol=1e-6, atol=1e-6)
    assert torch.allclose(xo, 0.8 * x - (-0.3) * den, rtol=1e-5, atol=1e-5)
    L.check(l.ldn_cfg_step(x.data_ptr(), u.data_ptr(), c.data_ptr(), 7.0, 1, -0.4, 0.25, 2.0, nz.data_ptr(), xo.data_ptr(),
                           do.data_ptr(), n, L.cur_stream()))
    assert torch.allclose(xo, x + ((x - den) / 2.0) * (-0.4) + nz * 0.25, rtol=1e-5, atol=1e-5)


def test_gemm_cta_pair_kernel_subprocess():
    """The opt-in CTA-pair GEMM (gemm_pair.cu: cluster of 2, tcgen05.mma.cta_group::2, LDN_GEMM_PAIR=1) computes the same
    results as the default kernels. The mode is latched per process, hence the subprocess."""
    import subprocess, sys, textwrap
    code = textwrap.dedent('''
        import torch, torch.nn.functional as F
        from lightdiffusion_next_b200 import _lib as L
        lib = L.load(); torch.manual_seed(0)
        def rel(a, b): return float((a.float() - b.float()).norm() / b.float().norm())
        for (M, N, K) in [(1000, 320, 320), (4096, 640, 1280), (300, 2560, 320)]:
            A = torch.randn(M, K, device="cuda").bfloat16(); W = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
            b = torch.randn(N, device="cuda"); R = torch.randn(M, N, device="cuda").bfloat16()
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            L.check(lib.ldn_gemm_bf16(A.data_ptr(), K, K, 0, 0, 0, W.data_ptr(), M, N, b.data_ptr(), 0, 0, 0, R.data_ptr(), N,
                                      out.data_ptr(), N, 0, 0, 0, 0, 0, L.cur_stream()))
            ref = A.float() @ W.float().t() + b + R.float()
            assert rel(out, ref) < 4e-3, (M, N, K, rel(out, ref))
        for (B, H, Wd, Cin, Cout) in [(2, 16, 16, 64, 64), (1, 32, 32, 320, 320), (3, 8, 8, 128, 320)]:
            x = torch.randn(B, H, Wd, Cin, device="cuda").bfloat16()
            w = (torch.randn(Cout, 3, 3, Cin, device="cuda") / (9 * Cin) ** 0.5).bfloat16()
            b = torch.randn(Cout, device="cuda")
            out = torch.empty(B, H, Wd, Cout, device="cuda", dtype=torch.bfloat16)
            L.check(lib.ldn_conv3x3_bf16(x.data_ptr(), w.data_ptr(), B, H, Wd, Cin, Cout, b.data_ptr(), 0, 0, 0,
                                         out.data_ptr(), L.cur_stream()))
            ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b, padding=1)
            assert rel(out.permute(0, 3, 1, 2), ref) < 4e-3, (B, H, Wd, Cin, Cout)
        print("PAIR_OK")
    ''')
    env = dict(os.environ, LDN_GEMM_PAIR="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PAIR_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("env", [{"LDN_GEMM_PAIR": "0"}, {"LDN_GEMM_PAIR": "0", "LDN_GEMM_BIAS_SMEM": "0"}, {"LDN_GEMM_BIAS_SMEM": "0"}])
def test_one_cta_long_k_kernel_stays_green(env):
    """Since round 2 the long-K problems (3x3 convs) run on the CTA-pair kernel by default.  The one-CTA kernel (gemm_tc_kernel:
    LDN_GEMM_PAIR=0) carries the same lean epilogue, GroupNorm statistics and bias staging; the switches are read once per
    process, so the conv tests of this file run again in a child process under each setting."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_ops_gpu.py"), "-m", "gpu", "-q", "-x",
                        "-p", "no:cacheprovider", "-k", "conv3x3"], cwd=root, env=dict(os.environ, **env), capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and " passed" in r.stdout, (r.stdout + r.stderr)[-3000:]
