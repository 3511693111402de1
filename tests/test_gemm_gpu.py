"""GPU parity of the tcgen05/TMA implicit-GEMM engine against (a) its SIMT restatement and
(b) a plain PyTorch fp32 reference of the same op on the fp16-rounded operands."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _setup():
    from gemm_cases import _lib
    _lib.init(0)


def _cmp(name, got, ref, tol=2e-3):
    got = got.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert math.isfinite(err) and err <= tol * max(1.0, scale), f"{name}: max err {err} (scale {scale})"


def _both(fn):
    """run fn(impl) for tc and simt, return both outputs"""
    return fn(0), fn(1)


def test_linear_bias_gelu():
    _setup()
    from gemm_cases import run_gemm
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    M, N, K = 300, 200, 192
    A = (torch.randn(M, K, device=dev, generator=g)).half()
    Wt = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).half()
    bias = torch.randn(N, device=dev, generator=g)
    ref = F.gelu(A.float() @ Wt.float().t() + bias)

    def run(impl):
        o32 = torch.zeros(M, N, device=dev)
        o16 = torch.zeros(M, N, device=dev, dtype=torch.half)
        run_gemm(impl, A, Wt, M, N, [(0, 0, 0, K // 64)], bias=bias, act1="gelu", out32=o32, ld32=N, out16=o16, ld16=N)
        return o32, o16
    (t32, t16), (s32, s16) = _both(run)
    _cmp("tc32", t32, ref); _cmp("simt32", s32, ref); _cmp("tc16", t16, ref, 4e-3)
    _cmp("tc-vs-simt", t32, s32, 1e-4)


@pytest.mark.parametrize("cin,cout,k,dil,T,bk", [(128, 128, 7, 3, 1000, 64), (32, 32, 11, 5, 2000, 32),
                                                 (192, 384, 5, 1, 333, 64), (64, 64, 3, 1, 130, 64),
                                                 (512, 256, 3, 1, 97, 64)])
def test_conv1d_taps(cin, cout, k, dil, T, bk):
    _setup()
    from gemm_cases import run_gemm, pack_conv1d
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(2)
    x = torch.randn(T, cin, device=dev, generator=g).half()
    w = (torch.randn(cout, cin, k, device=dev, generator=g) / math.sqrt(cin * k)).half()
    bias = torch.randn(cout, device=dev, generator=g)
    res = torch.randn(T, cout, device=dev, generator=g)
    acc = torch.randn(T, cout, device=dev, generator=g)
    pad = (k - 1) * dil // 2
    conv = F.conv1d(x.float().t()[None], w.float(), bias, dilation=dil, padding=pad)[0].t()
    ref = (conv + res) / 3.0 + acc
    B = pack_conv1d(w, bk)
    segs = [(j * dil - pad, 0, 0, (cin + bk - 1) // bk) for j in range(k)]

    def run(impl):
        o32 = torch.zeros(T, cout, device=dev)
        o16 = torch.zeros(T, cout, device=dev, dtype=torch.half)
        run_gemm(impl, x, B, T, cout, segs, block_k=bk, bias=bias, res1=res, res2=acc, alpha=1.0 / 3.0,
                 act2="lrelu", act2_p=0.1, out32=o32, ld32=cout, out16=o16, ld16=cout)
        return o32, o16
    (t32, t16), (s32, _) = _both(run)
    _cmp("tc32", t32, ref); _cmp("simt32", s32, ref)
    _cmp("tc16", t16, F.leaky_relu(ref, 0.1), 4e-3)
    _cmp("tc-vs-simt", t32, s32, 1e-4)


@pytest.mark.parametrize("cin,cout,H,W,bk", [(16, 16, 96, 128, 16), (64, 128, 24, 32, 64), (128, 64, 3, 4, 64),
                                            (32, 16, 40, 64, 32), (512, 512, 51, 4, 64), (256, 256, 102, 8, 64), (128, 128, 204, 16, 64)])
def test_conv2d_3x3(cin, cout, H, W, bk):
    _setup()
    from gemm_cases import run_gemm, pack_conv2d
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(H, W, cin, device=dev, generator=g).half()
    w = (torch.randn(cout, cin, 3, 3, device=dev, generator=g) / math.sqrt(cin * 9)).half()
    bias = torch.randn(cout, device=dev, generator=g)
    res = torch.randn(H * W, cout, device=dev, generator=g)
    conv = F.conv2d(x.float().permute(2, 0, 1)[None], w.float(), bias, padding=1)[0].permute(1, 2, 0).reshape(H * W, cout)
    ref = F.relu(conv) + res
    B = pack_conv2d(w, bk)
    segs = [(dh - 1, 0, dw - 1, (cin + bk - 1) // bk) for dh in range(3) for dw in range(3)]

    def run(impl):
        o32 = torch.zeros(H * W, cout, device=dev)
        run_gemm(impl, x, B, H * W, cout, segs, block_k=bk, a_rows=H, a_cols=cin, lda=cin, conv2d_W=W, bias=bias,
                 act1="relu", res2=res, out32=o32, ld32=cout)
        return o32
    t, s = _both(run)
    _cmp("tc", t, ref); _cmp("simt", s, ref); _cmp("tc-vs-simt", t, s, 1e-4)


def test_convT2d_up2_into_concat():
    _setup()
    from gemm_cases import run_gemm, pack_convT2d_up2
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(4)
    cin, cout, H, W, bk = 32, 16, 10, 8, 32
    x = torch.randn(H, W, cin, device=dev, generator=g).half()
    w = (torch.randn(cin, cout, 3, 3, device=dev, generator=g) / math.sqrt(cin * 9 / 4)).half()
    bias = torch.randn(cout, device=dev, generator=g)
    ref = F.relu(F.conv_transpose2d(x.float().permute(2, 0, 1)[None], w.float(), None, stride=2, padding=1,
                                    output_padding=1)[0] + bias[:, None, None]).permute(1, 2, 0)      # [2H,2W,C]
    B = pack_convT2d_up2(w, bk)
    bias4 = bias.repeat(4).contiguous()
    segs = [(dh, 0, dw, 1) for dh, dw in ((0, 0), (0, 1), (1, 0), (1, 1))]

    def run(impl):
        o16 = torch.zeros(2 * H, 2 * W, 2 * cout, device=dev, dtype=torch.half)    # concat buffer, we fill [:cout]
        run_gemm(impl, x, B, H * W, 4 * cout, segs, block_k=bk, a_rows=H, a_cols=cin, lda=cin, conv2d_W=W, bias=bias4,
                 act1="relu", out16=o16, ld16=2 * cout, up2_C=cout)
        return o16
    t, s = _both(run)
    _cmp("tc", t[:, :, :cout], ref, 4e-3); _cmp("simt", s[:, :, :cout], ref, 4e-3)
    assert (t[:, :, cout:] == 0).all()


def test_attention_pieces_batched():
    _setup()
    from gemm_cases import run_gemm
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(5)
    T, Hh, D = 203, 3, 64
    Tp = (T + 7) // 8 * 8
    x = torch.randn(T, Hh * D, device=dev, generator=g).half()
    qk = torch.randn(T, 2 * Hh * D, device=dev, generator=g).half() * 0.3
    # scores S[h] = q_h k_h^T
    ref_s = torch.einsum("thd,shd->hts", qk[:, :Hh * D].float().view(T, Hh, D), qk[:, Hh * D:].float().view(T, Hh, D))

    def run_s(impl):
        S = torch.zeros(Hh, T, Tp, device=dev)
        run_gemm(impl, qk, qk, T, T, [(0, 0, 0, 1)], a_cols=Hh * D, batch=Hh, a_col_z=D, b_col0=Hh * D, b_col_z=D,
                 c_z=T * Tp, out32=S, ld32=Tp)
        return S[:, :, :T]
    t, s = _both(run_s)
    _cmp("scores tc", t, ref_s); _cmp("scores simt", s, ref_s)
    # V^T = Wv x^T + bv (bias per row)
    Wv = (torch.randn(Hh * D, Hh * D, device=dev, generator=g) / math.sqrt(Hh * D)).half()
    bv = torch.randn(Hh * D, device=dev, generator=g)
    ref_vt = Wv.float() @ x.float().t() + bv[:, None]

    def run_vt(impl):
        vt = torch.full((Hh * D, Tp), float("nan"), device=dev, dtype=torch.half)
        run_gemm(impl, Wv, x, Hh * D, T, [(0, 0, 0, Hh * D // 64)], bias=bv, bias_per_row=1, out16=vt, ld16=Tp)
        return vt
    tvt, svt = _both(run_vt)
    _cmp("vT tc", tvt[:, :T], ref_vt, 4e-3); _cmp("vT simt", svt[:, :T], ref_vt, 4e-3)
    # O = P V with P [Hh*T, Tp] (pad column garbage must be ignored via b_cols / a_cols = T)
    P = torch.softmax(ref_s, -1)
    Pp = torch.full((Hh * T, Tp), 7.0, device=dev, dtype=torch.half)
    Pp[:, :T] = P.reshape(Hh * T, T).half()
    vt = tvt.clone()
    ref_o = torch.einsum("hts,hds->thd", Pp[:, :T].float().view(Hh, T, T), vt[:, :T].float().view(Hh, D, T)).reshape(T, Hh * D)

    def run_o(impl):
        o = torch.zeros(T, Hh * D, device=dev, dtype=torch.half)
        run_gemm(impl, Pp, vt, T, D, [(0, 0, 0, (T + 63) // 64)], a_rows=Hh * T, a_cols=T, b_cols=T, batch=Hh, a_row_z=T,
                 b_row_z=D, c_z=D, out16=o, ld16=Hh * D)
        return o
    to, so = _both(run_o)
    _cmp("PV tc", to, ref_o, 4e-3); _cmp("PV simt", so, ref_o, 4e-3)


def test_stride2_conv_view():
    _setup()
    from gemm_cases import run_gemm
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(6)
    C, T = 64, 401
    xbuf = torch.zeros(T + 1, C, device=dev, dtype=torch.half)
    xbuf[:T] = torch.randn(T, C, device=dev, generator=g).half()
    w = (torch.randn(96, C, 3, device=dev, generator=g) / math.sqrt(3 * C)).half()
    ref = F.gelu(F.conv1d(xbuf[:T].float().t()[None], w.float(), None, stride=2)[0].t())
    To = (T - 3) // 2 + 1
    # B = [W0 | W1 | W2], A viewed as [(T+1)/2, 2C]
    B = torch.cat([w[:, :, 0], w[:, :, 1], w[:, :, 2]], dim=1).contiguous()
    view = xbuf.view((T + 1) // 2, 2 * C)

    def run(impl):
        o = torch.zeros(To, 96, device=dev, dtype=torch.half)
        run_gemm(impl, view, B, To, 96, [(0, 0, 0, 2), (1, 0, 0, 1)], act1="gelu", out16=o, ld16=96)
        return o
    t, s = _both(run)
    _cmp("tc", t, ref, 4e-3); _cmp("simt", s, ref, 4e-3)


@pytest.mark.parametrize("cin,cout,k,s,T", [(64, 32, 4, 2, 300), (128, 64, 24, 12, 77), (96, 48, 16, 10, 50)])
def test_convT1d_polyphase(cin, cout, k, s, T):
    _setup()
    from gemm_cases import run_gemm, pack_convT1d
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(7)
    pad = (k - s) // 2
    x = torch.randn(T, cin, device=dev, generator=g).half()
    w = (torch.randn(cin, cout, k, device=dev, generator=g) / math.sqrt(cin * k / s)).half()
    bias = torch.randn(cout, device=dev, generator=g)
    ref = F.conv_transpose1d(x.float().t()[None], w.float(), bias, stride=s, padding=pad)[0].t()     # [T*s, cout]
    assert ref.shape[0] == T * s
    bk = 64 if cin % 64 == 0 else 32
    B = pack_convT1d(w, s, pad, bk)
    nk = (cin + bk - 1) // bk
    biasr = bias.repeat(s).contiguous()

    def run(impl):
        o = torch.zeros(T * s, cout, device=dev)
        run_gemm(impl, x, B, T, s * cout, [(-1, 0, 0, nk), (0, 0, 0, nk), (1, 0, 0, nk)], block_k=bk, bias=biasr, out32=o,
                 ld32=s * cout)
        return o
    t, sm = _both(run)
    _cmp("tc", t, ref); _cmp("simt", sm, ref)


def test_gate_and_groups():
    _setup()
    from gemm_cases import run_gemm, pack_conv1d
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(8)
    T, C = 257, 64
    x = torch.randn(T, C, device=dev, generator=g).half()
    w = (torch.randn(2 * C, C, 5, device=dev, generator=g) / math.sqrt(C * 5)).half()
    bias = torch.randn(2 * C, device=dev, generator=g)
    a = F.conv1d(x.float().t()[None], w.float(), bias, padding=2)[0].t()
    ref = torch.tanh(a[:, :C]) * torch.sigmoid(a[:, C:])
    perm = torch.stack([torch.arange(C), torch.arange(C) + C], 1).reshape(-1).to(dev)   # interleave (tanh_i, sig_i)
    B = pack_conv1d(w[perm], 64)
    bp = bias[perm].contiguous()

    def run(impl):
        o = torch.zeros(T, C, device=dev, dtype=torch.half)
        run_gemm(impl, x, B, T, 2 * C, [(j - 2, 0, 0, 1) for j in range(5)], bias=bp, gate=1, out16=o, ld16=C)
        return o
    t, s = _both(run)
    _cmp("gate tc", t, ref, 4e-3); _cmp("gate simt", s, ref, 4e-3)
    # grouped conv (pos_conv style): 2 groups of 48 channels, k = 8, pad 4, drop last
    G, cg, k = 2, 48, 8
    xg = torch.randn(T, G * cg, device=dev, generator=g).half()
    wg = (torch.randn(G * cg, cg, k, device=dev, generator=g) / math.sqrt(cg * k)).half()
    bg = torch.randn(G * cg, device=dev, generator=g)
    resg = torch.randn(T, G * cg, device=dev, generator=g)
    refg = resg + F.gelu(F.conv1d(xg.float().t()[None], wg.float(), bg, padding=k // 2, groups=G)[0, :, :-1].t())
    Bg = torch.zeros(G, 64, k, 64, device=dev, dtype=torch.half)       # per group: 64 rows (48 used) x k taps x 64 ch (48 used)
    Bg[:, :cg, :, :cg] = wg.view(G, cg, cg, k).permute(0, 1, 3, 2)
    Bg = Bg.reshape(G * 64, k * 64).contiguous()

    def rung(impl):
        o = torch.zeros(T, G * cg, device=dev)
        run_gemm(impl, xg, Bg, T, cg, [(j - k // 2, 0, 0, 1) for j in range(k)], batch=G, a_col_z=cg, b_row_z=64, c_z=cg,
                 bias=bg, bias_z=cg, act1="gelu", res2=resg, out32=o, ld32=G * cg)
        return o
    t, s = _both(rung)
    _cmp("group tc", t, refg); _cmp("group simt", s, refg)


@pytest.mark.parametrize("cin,cout,k,dil,T,bk", [(64, 64, 7, 3, 60001, 64), (32, 32, 11, 5, 150000, 32), (128, 128, 3, 1, 90000, 64),
                                                 (128, 128, 11, 5, 80000, 64), (128, 128, 7, 1, 70000, 64), (32, 1, 7, 1, 120000, 32)])
def test_conv1d_weight_stationary_halo_kernel(cin, cout, k, dil, T, bk):
    """Long stride-1 convolutions dispatch to the weight-stationary halo kernel (gemm_ws.cu): same contract,
    checked against the SIMT restatement and the torch fp32 reference."""
    _setup()
    from gemm_cases import run_gemm, pack_conv1d
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(12)
    x = torch.randn(T, cin, device=dev, generator=g).half()
    w = (torch.randn(cout, cin, k, device=dev, generator=g) / math.sqrt(cin * k)).half()
    bias = torch.randn(max(cout, 16), device=dev, generator=g)[:cout].contiguous() if cout >= 16 else None
    res = torch.randn(T, cout, device=dev, generator=g)
    acc = torch.randn(T, cout, device=dev, generator=g)
    pad = (k - 1) * dil // 2
    conv = F.conv1d(x.float().t()[None], w.float(), bias, dilation=dil, padding=pad)[0].t()
    ref = (conv + res) / 3.0 + acc
    B = torch.zeros(max(cout, 16), k * ((cin + bk - 1) // bk * bk), device=dev, dtype=torch.half)
    B[:cout] = pack_conv1d(w, bk)
    segs = [(j * dil - pad, 0, 0, (cin + bk - 1) // bk) for j in range(k)]

    def run(impl):
        o32 = torch.zeros(T, cout, device=dev)
        o16 = torch.zeros(T, cout, device=dev, dtype=torch.half)
        run_gemm(impl, x, B, T, cout, segs, block_k=bk, bias=bias, res1=res, res2=acc, alpha=1.0 / 3.0,
                 act2="lrelu", act2_p=0.1, out32=o32, ld32=cout, out16=o16, ld16=cout)
        return o32, o16
    (t32, t16), (s32, _) = _both(run)
    _cmp("ws32", t32, ref); _cmp("simt32", s32, ref)
    _cmp("ws16", t16, F.leaky_relu(ref, 0.1), 4e-3)
    _cmp("ws-vs-simt", t32, s32, 1e-4)


@pytest.mark.parametrize("epi", [0, 1, 2])
@pytest.mark.parametrize("cin,k,dil,T,bk", [(64, 11, 5, 80001, 64), (64, 3, 1, 76800, 64), (32, 7, 3, 150017, 32), (128, 3, 1, 90000, 64),
                                            (128, 11, 1, 80001, 64)])
def test_weight_stationary_tma_epilogue_variants(epi, cin, k, dil, T, bk):
    """The three vocoder epilogue patterns of the weight-stationary kernel's TMA-staged epilogue (gemm_ws.cu, v2):
    0: out16 = lrelu(conv + b);  1: y = conv + b + res -> out32, out16 = lrelu(y);  2: y = (conv + b + res)/3 -> out32 only
    (first branch of the 3-branch mean: no running sum, no fp16 hand-off).  Ragged M tails exercise the TMA store clipping;
    the fp16 output uses a wider leading dimension, as the stage hand-off buffers do."""
    _setup()
    from gemm_cases import run_gemm, pack_conv1d
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(40 + epi)
    cout = cin
    x = torch.randn(T, cin, device=dev, generator=g).half()
    w = (torch.randn(cout, cin, k, device=dev, generator=g) / math.sqrt(cin * k)).half()
    bias = torch.randn(cout, device=dev, generator=g)
    res = torch.randn(T, cout, device=dev, generator=g)
    pad = (k - 1) * dil // 2
    conv = F.conv1d(x.float().t()[None], w.float(), bias, dilation=dil, padding=pad)[0].t()
    B = pack_conv1d(w, bk).contiguous()
    segs = [(j * dil - pad, 0, 0, cin // bk) for j in range(k)]
    ld16 = cout + 64

    def run(impl):
        o32 = torch.full((T, cout), 7.0, device=dev)
        o16 = torch.full((T, ld16), 7.0, device=dev, dtype=torch.half)
        if epi == 0:
            run_gemm(impl, x, B, T, cout, segs, block_k=bk, bias=bias, act2="lrelu", act2_p=0.1, out16=o16, ld16=ld16)
        elif epi == 1:
            run_gemm(impl, x, B, T, cout, segs, block_k=bk, bias=bias, res1=res, act2="lrelu", act2_p=0.1, out32=o32, ld32=cout,
                     out16=o16, ld16=ld16)
        else:
            run_gemm(impl, x, B, T, cout, segs, block_k=bk, bias=bias, res1=res, alpha=1.0 / 3.0, out32=o32, ld32=cout)
        return o32, o16
    (t32, t16), (s32, s16) = _both(run)
    if epi == 0:
        _cmp("ws2 c1", t16[:, :cout], F.leaky_relu(conv, 0.1), 4e-3)
        _cmp("ws2 c1 vs simt", t16[:, :cout], s16[:, :cout], 2e-3)
    else:
        y = conv + res if epi == 1 else (conv + res) / 3.0
        _cmp("ws2 y", t32, y); _cmp("ws2 vs simt", t32, s32, 1e-4)
        if epi == 1:
            _cmp("ws2 y16", t16[:, :cout], F.leaky_relu(y, 0.1), 4e-3)
    assert (t16[:, cout:] == 7.0).all(), "columns past N of the fp16 hand-off buffer must not be touched"
    if epi == 0:
        assert (t32 == 7.0).all()


@pytest.mark.parametrize("cin,cout,k,dil,T", [(128, 128, 11, 3, 50001), (256, 256, 3, 1, 19176), (64, 192, 5, 1, 9000)])
def test_streaming_kernel_with_tma_staged_residual(cin, cout, k, dil, T):
    """Large launches with an fp32 residual on the streaming kernel, with and without the opt-in TMA prefetch of the residual
    tile into swizzled shared memory (gemm_tc.cu, RS variant; the env var is read once per process, so this test runs its RS
    half in a child process): same contract, checked against the SIMT restatement and torch; ragged M tail, 1 / 2 / 3 N-tiles,
    with and without the second residual."""
    import os, subprocess, sys
    if os.environ.get("RVCB_RS") != "1":
        env = dict(os.environ, RVCB_RS="1")
        node = f"{__file__}::test_streaming_kernel_with_tma_staged_residual[{cin}-{cout}-{k}-{dil}-{T}]"
        r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", node], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    _setup()
    from gemm_cases import run_gemm, pack_conv1d
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(77)
    x = torch.randn(T, cin, device=dev, generator=g).half()
    w = (torch.randn(cout, cin, k, device=dev, generator=g) / math.sqrt(cin * k)).half()
    bias = torch.randn(cout, device=dev, generator=g)
    res = torch.randn(T, cout, device=dev, generator=g)
    acc = torch.randn(T, cout, device=dev, generator=g)
    pad = (k - 1) * dil // 2
    conv = F.conv1d(x.float().t()[None], w.float(), bias, dilation=dil, padding=pad)[0].t()
    B = pack_conv1d(w, 64).contiguous()
    segs = [(j * dil - pad, 0, 0, cin // 64) for j in range(k)]
    for with_res2 in (False, True):
        ref = conv + res + (acc if with_res2 else 0)

        def run(impl):
            o32 = torch.zeros(T, cout, device=dev)
            o16 = torch.zeros(T, cout, device=dev, dtype=torch.half)
            run_gemm(impl, x, B, T, cout, segs, bias=bias, res1=res, res2=acc if with_res2 else None, act2="lrelu", act2_p=0.1,
                     out32=o32, ld32=cout, out16=o16, ld16=cout)
            return o32, o16
        (t32, t16), (s32, _) = _both(run)
        _cmp("rs32", t32, ref); _cmp("rs-vs-simt", t32, s32, 1e-4)
        _cmp("rs16", t16, F.leaky_relu(ref, 0.1), 4e-3)


@pytest.mark.parametrize("M,N,K", [(1598, 192, 2304), (799, 96, 1024), (204, 512, 4608), (333, 64, 768)])
def test_cluster_split_k_small_m_long_k(M, N, K):
    """Small-M, long-K launches run on the cluster split-K kernel (gemm_sk.cu): 2 or 4 CTAs per output tile, partial tiles
    reduced over distributed shared memory in a fixed order, fused epilogue on the reduced slab.  Same contract: checked
    against the SIMT restatement and torch with both residuals, both outputs and an activation."""
    _setup()
    from gemm_cases import run_gemm
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(91)
    x = (torch.randn(M, K, device=dev, generator=g)).half()
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).half()
    bias = torch.randn(N, device=dev, generator=g)
    r1 = torch.randn(M, N, device=dev, generator=g)
    r2 = torch.randn(M, N, device=dev, generator=g)
    ref = F.gelu(x.float() @ w.float().t() + bias + r1) * 0.5 + r2

    def run(impl):
        o32 = torch.zeros(M, N, device=dev)
        o16 = torch.zeros(M, N + 8, device=dev, dtype=torch.half)
        run_gemm(impl, x, w.contiguous(), M, N, [(0, 0, 0, K // 64)], bias=bias, res1=r1, act1="gelu", alpha=0.5, res2=r2,
                 act2="lrelu", act2_p=0.1, out32=o32, ld32=N, out16=o16, ld16=N + 8)
        return o32, o16
    (t32, t16), (s32, _) = _both(run)
    _cmp("sk32", t32, ref); _cmp("sk-vs-simt", t32, s32, 1e-4)
    _cmp("sk16", t16[:, :N], F.leaky_relu(ref, 0.1), 4e-3)
    a, b = run(0), run(0)
    assert torch.equal(a[0], b[0]), "the cluster reduction order is fixed: results are run-to-run identical"
