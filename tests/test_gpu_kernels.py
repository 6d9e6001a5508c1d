"""Per-kernel parity tests (GPU): every libb200rl kernel against a plain torch fp32/fp64
restatement of the same op on identical bf16-rounded inputs. Tolerances are stated per test."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, device, scale=1.0, seed=0):
    n = 1
    for s in shape:
        n *= s
    if n > (1 << 27):   # the K = 152064 operands: generate on the device
        g = torch.Generator(device=device).manual_seed(seed)
        return (torch.randn(*shape, generator=g, device=device, dtype=torch.bfloat16) * scale)
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(device=device, dtype=torch.bfloat16)


def _rel_err(a, b):
    a = a.double()
    b = b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


# ------------------------------------------------------------------------------------------------
# G1: tcgen05 GEMM
# ------------------------------------------------------------------------------------------------
GEMM_SHAPES = [
    # (M, N, K1, K2, force_bn)
    (128, 64, 64, 0, 0),
    (128, 128, 128, 0, 0),
    (256, 256, 512, 0, 0),
    (300, 200, 192, 0, 0),        # ragged M, N (multiple of 8), K
    (77, 72, 40, 0, 0),           # K tail (zero-filled by TMA), tiny
    (1000, 512, 1024, 64, 0),     # LoRA side segment
    (512, 384, 256, 64, 128),
    (512, 384, 256, 128, 192),
    (640, 4608, 3584, 64, 256),   # QKV-shaped
    (4096, 1024, 2048, 0, 256),   # many tiles per CTA (pipeline wrap-around)
    (4100, 3584, 1024, 64, 0),
    (600, 3584, 512, 64, 224),    # 224-wide pair tile (N = 16 x 224)
    (900, 1000, 320, 64, 224),    # ragged N with the 224 tile
    (4446, 3584, 1024, 64, 192),
    # wide 256 x 512 pair tiles (force_bn = 512): full tiles, ragged N with a partly and a fully out-of-range sub-tile
    (4446, 3584, 1024, 64, 512),
    (8892, 4608, 3584, 64, 512),
    (600, 1024, 512, 64, 512),
    (900, 1000, 320, 64, 512),
    (700, 600, 192, 0, 512),
]


@pytest.fixture(params=[1, 0], ids=["cta_pair", "single_cta"])
def gemm_mode(request, cuda):
    """Run the GEMM tests through both kernel families (cta_group::2 pair tiles / single-CTA tiles)."""
    from distrl_llm_b200 import _capi
    _capi.lib().b200rl_gemm_set_cta_pair(request.param)
    yield request.param
    _capi.lib().b200rl_gemm_set_cta_pair(1)


@pytest.mark.parametrize("M,N,K1,K2,bn", GEMM_SHAPES)
def test_gemm_tn(cuda, gemm_mode, M, N, K1, K2, bn):
    if bn in (224, 512) and not gemm_mode:
        pytest.skip("224-wide and 512-wide tiles exist only in the CTA-pair kernel")
    from distrl_llm_b200 import ops
    a1 = _rand((M, K1), cuda, seed=1)
    b1 = _rand((N, K1), cuda, seed=2)
    a2 = _rand((M, K2), cuda, seed=3) if K2 else None
    b2 = _rand((N, K2), cuda, seed=4) if K2 else None
    out = ops.gemm(a1, b1, a2, b2, force_bn=bn)
    torch.cuda.synchronize()
    ref = a1.float() @ b1.float().T
    if K2:
        ref = ref + a2.float() @ b2.float().T
    # bf16 output rounding: 2^-9 relative per element; accumulate in fp32 like the reference's autocast matmul
    err = _rel_err(out, ref)
    assert err < 4e-3, f"rel err {err}"
    assert torch.isfinite(out.float()).all()


@pytest.mark.parametrize("M,N,K1,b_mn", [(4446, 3584, 8192, False), (4446, 3584, 8192, True),    # rem 30 of 74 -> 2 K-ranges
                                        (1536, 3584, 16384, False), (1536, 3584, 16384, True),  # rem 10 -> 4 K-ranges
                                        (2100, 4608, 12288, False),
                                        # the hot dX shapes of BASELINE config 2 (two fused micro-batches, packed rows):
                                        (8192, 3584, 152064, True),    # lm_head dX: K = vocab
                                        (8892, 3584, 152064, True),    # same with the ragged last m-block
                                        (8892, 3584, 37888, True),     # gate|up dX: K = 2 x inter (+ 64 LoRA)
                                        (4446, 3584, 37888, True)])    # one micro-batch per pass
@pytest.mark.parametrize("bn", [256, 512], ids=["tile256", "wide512"])
def test_gemm_tail_split(cuda, M, N, K1, b_mn, bn):
    """CTA-pair GEMM with the last partial wave split along K (gemm2_tcgen05.cu): same result as with the split
    disabled, up to the fp32 summation order of the K-ranges; fp32 output compared tightly against torch."""
    from distrl_llm_b200 import _capi, ops
    K2 = 64
    a1 = _rand((M, K1), cuda, seed=1)
    b1 = _rand((K1, N) if b_mn else (N, K1), cuda, seed=2)
    a2 = _rand((M, K2), cuda, seed=3)
    b2 = _rand((K2, N) if b_mn else (N, K2), cuda, seed=4)
    bias = _rand((N,), cuda, seed=5)
    res = _rand((M, N), cuda, seed=6)
    ref = a2.float() @ (b2.float() if b_mn else b2.float().T)
    for k0 in range(0, K1, 32768):   # chunked over K: the fp32 copies of the K = 152064 operands would be 13 GB
        ak = a1[:, k0:k0 + 32768].float()
        ref += ak @ (b1[k0:k0 + 32768].float() if b_mn else b1[:, k0:k0 + 32768].float().T)
    del ak
    ref = 0.25 * ref + bias.float()[None] + res.float()
    outs = []
    try:
        for en in (1, 0):
            _capi.lib().b200rl_gemm_set_tail_split(en)
            for _ in range(3):   # repeated launches reuse the workspace with a new epoch
                o32 = ops.gemm(a1, b1, a2, b2, bias=bias, residual=res, alpha=0.25, out_fp32=True, force_bn=bn, b_mn=b_mn)
            o16 = ops.gemm(a1, b1, a2, b2, bias=bias, residual=res, alpha=0.25, force_bn=bn, b_mn=b_mn)
            torch.cuda.synchronize()
            # fp32 accumulation over K: 2e-5 up to K = 16k.  The tensor core adds each K = 16 partial product into the
            # TMEM accumulator with truncation (the results are systematically SMALLER in magnitude than torch's:
            # -86.9969 vs -87.0132), so the error grows linearly with the number of accumulations: measured 1.8e-4 at
            # K = 152064 (9504 accumulations), i.e. 1/20 of the bf16 rounding of the output.  Allowed: 4e-4.
            assert _rel_err(o32, ref) < (2e-5 if K1 <= 16384 else 4e-4) and _rel_err(o16, ref) < 4e-3
            outs.append((o32, o16))
    finally:
        _capi.lib().b200rl_gemm_set_tail_split(1)
    assert (outs[0][0] - outs[1][0]).abs().max().item() < 1e-3 * ref.abs().max().item()


@pytest.mark.parametrize("wide", [1, 0], ids=["wide512", "tile256"])
@pytest.mark.parametrize("M,I,H", [(300, 256, 192), (4446, 1024, 512), (700, 18944, 256), (4446, 1152, 512), (8892, 18944, 128)])
def test_gemm_swiglu_fused_is_bit_identical(cuda, M, I, H, wide):
    """G1+G5 fusion (gemm2_tcgen05.cu FUSE 1/2): the fused epilogues give exactly the bits of GEMM + row kernel, with the
    256 x 512 wide tiles (I / 128 = 9 exercises the skipped second sub-tile of the last wide tile) and without."""
    from distrl_llm_b200 import _capi, ops
    lib = _capi.lib()
    K2 = 64
    st = _capi.stream()
    try:
        lib.b200rl_gemm_set_wide(wide)
        lib.b200rl_gemm_set_tail_split(0)      # same K summation order in both paths
        # forward: gu = h.Wgu^T + u.Bgu^T ; act = silu(gate)*up
        h, u = _rand((M, H), cuda, seed=1), _rand((M, K2), cuda, seed=2)
        Wgu, Bgu = _rand((2 * I, H), cuda, seed=3, scale=0.2), _rand((2 * I, K2), cuda, seed=4, scale=0.2)
        gu_ref = ops.gemm(h, Wgu, u, Bgu, force_bn=256)
        act_ref = ops.swiglu_fwd(gu_ref)
        gu = torch.empty_like(gu_ref)
        act = torch.empty_like(act_ref)
        _capi.check(lib.b200rl_gemm_swiglu(1, h.data_ptr(), H, Wgu.data_ptr(), H, H, u.data_ptr(), K2, Bgu.data_ptr(), K2, K2,
                                           gu.data_ptr(), 2 * I, act.data_ptr(), I, M, I, st), "gemm_swiglu fwd")
        torch.cuda.synchronize()
        assert torch.equal(gu.view(torch.int16), gu_ref.view(torch.int16))
        assert torch.equal(act.view(torch.int16), act_ref.view(torch.int16))
        # backward: dact = dx.Wd + du.Acat (Wd stored [H, I]) ; dgu = swiglu_bwd(gu, dact)
        dx, du = _rand((M, H), cuda, seed=5), _rand((M, K2), cuda, seed=6)
        Wd, Ad = _rand((H, I), cuda, seed=7, scale=0.2), _rand((K2, I), cuda, seed=8, scale=0.2)
        dact_ref = ops.gemm(dx, Wd, du, Ad, force_bn=256, b_mn=True)
        dgu_ref = ops.swiglu_bwd(gu_ref, dact_ref)
        dgu = torch.empty_like(dgu_ref)
        _capi.check(lib.b200rl_gemm_swiglu(2, dx.data_ptr(), H, Wd.data_ptr(), I, H, du.data_ptr(), K2, Ad.data_ptr(), I, K2,
                                           dgu.data_ptr(), 2 * I, gu_ref.data_ptr(), 2 * I, M, I, st), "gemm_swiglu bwd")
        torch.cuda.synchronize()
        assert torch.equal(dgu.view(torch.int16), dgu_ref.view(torch.int16))
    finally:
        lib.b200rl_gemm_set_tail_split(1)
        lib.b200rl_gemm_set_wide(2)    # library default (gemm2_tcgen05.cu: wide tiles for K-long GEMMs only)


@pytest.mark.parametrize("b_mn", [False, True], ids=["tn", "dx"])
def test_gemm_wide_tile_is_bit_identical_to_256(cuda, b_mn):
    """256 x 512 pair tiles (two N = 256 UMMAs per k-step sharing the A stage) vs 256 x 256 tiles: every output element is
    the same K-ordered fp32 accumulation, so the results agree bit for bit (tail split off: it changes the K order)."""
    from distrl_llm_b200 import _capi, ops
    M, N, K1, K2 = 2300, 3584, 2048, 64
    a1, a2 = _rand((M, K1), cuda, seed=1), _rand((M, K2), cuda, seed=3)
    b1 = _rand((K1, N) if b_mn else (N, K1), cuda, seed=2)
    b2 = _rand((K2, N) if b_mn else (N, K2), cuda, seed=4)
    res = _rand((M, N), cuda, seed=6)
    try:
        _capi.lib().b200rl_gemm_set_tail_split(0)
        o256 = ops.gemm(a1, b1, a2, b2, residual=res, force_bn=256, b_mn=b_mn)
        o512 = ops.gemm(a1, b1, a2, b2, residual=res, force_bn=512, b_mn=b_mn)
        torch.cuda.synchronize()
    finally:
        _capi.lib().b200rl_gemm_set_tail_split(1)
    assert torch.equal(o256.view(torch.int16), o512.view(torch.int16))


@pytest.mark.parametrize("bn", [256, 512], ids=["tile256", "wide512"])
@pytest.mark.parametrize("M,N,K1,K2,b_mn", [(4446, 3584, 3584, 64, False), (8892, 4608, 3584, 64, False), (300, 512, 256, 64, False),
                                            (4446, 3584, 4608, 64, True), (8892, 3584, 37888, 64, True), (1000, 768, 640, 128, False),
                                            (1000, 768, 640, 128, True), (33000, 512, 256, 64, False)])
def test_gemm_lora_in_kernel(cuda, M, N, K1, K2, b_mn, bn):
    """LoRA intermediate produced inside the launch (ext units + flags, gemm2_tcgen05.cu) vs the two-GEMM formulation:
    U = bf16(s * a1 @ Bext^T) must equal the separately computed skinny GEMM up to its fp32 summation order, and
    C = a1 @ B1^T + U @ B2^T is then checked against torch using the U the kernel actually produced."""
    from distrl_llm_b200 import ops
    s_ = 0.5
    a1 = _rand((M, K1), cuda, seed=1)
    b1 = _rand((K1, N) if b_mn else (N, K1), cuda, seed=2)
    bext = _rand((K1, K2) if b_mn else (K2, K1), cuda, seed=3, scale=0.1)
    b2 = _rand((K2, N) if b_mn else (N, K2), cuda, seed=4)
    res = _rand((M, N), cuda, seed=5)
    for _ in range(3):      # repeated launches: the flag epoch advances
        out, u = ops.gemm_lora(a1, b1, bext, b2, scale=s_, residual=res, b_mn=b_mn, force_bn=bn)
    torch.cuda.synchronize()
    u_ref = s_ * (a1.float() @ (bext.float() if b_mn else bext.float().T))
    assert _rel_err(u, u_ref) < 4e-3
    ref = a1.float() @ (b1.float() if b_mn else b1.float().T) + u.float() @ (b2.float() if b_mn else b2.float().T) + res.float()
    assert _rel_err(out, ref) < 4e-3
    assert torch.isfinite(out.float()).all()
    # same numbers as the separate skinny GEMM + K-extension path (bf16 rounding of U may differ by one ulp at most)
    u2 = ops.gemm(a1, bext, alpha=s_, b_mn=b_mn)
    assert (u.float() - u2.float()).abs().max().item() <= 2.0 ** -7 * u_ref.abs().max().item()


@pytest.mark.parametrize("bn", [256, 512], ids=["tile256", "wide512"])
@pytest.mark.parametrize("M,N,K1,K2,b_mn", [(4446, 3584, 3584, 64, False), (8892, 4608, 3584, 0, False), (300, 512, 256, 64, False),
                                            (4446, 3584, 4608, 64, True), (2100, 3584, 18944, 64, True), (700, 832, 320, 0, False),
                                            (700, 832, 320, 64, True)])
def test_gemm_nf4_in_mainloop_is_bit_identical(cuda, M, N, K1, K2, b_mn, bn):
    """NF4 base weight expanded by the GEMM's own producer warps (b200rl_gemm_nf4) vs nf4_dequant + the bf16 GEMM: the
    shared-memory tiles hold the same bf16 values and the K order is the same, so the outputs agree bit for bit
    (tail split off: it re-orders K)."""
    from distrl_llm_b200 import _capi, ops
    w = _rand((K1, N) if b_mn else (N, K1), cuda, seed=2, scale=0.05)
    packed, absmax = ops.nf4_quantize(w)
    a1 = _rand((M, K1), cuda, seed=1)
    a2 = _rand((M, K2), cuda, seed=3) if K2 else None
    b2 = (_rand((K2, N) if b_mn else (N, K2), cuda, seed=4)) if K2 else None
    res = _rand((M, N), cuda, seed=5)
    dense = ops.nf4_dequant(packed, absmax, *w.shape)
    try:
        _capi.lib().b200rl_gemm_set_tail_split(0)
        ref = ops.gemm(a1, dense, a2, b2, residual=res, force_bn=bn, b_mn=b_mn)
        for _ in range(2):
            out = ops.gemm_nf4(a1, packed, absmax, N, a2, b2, residual=res, b_mn=b_mn, force_bn=bn)
        torch.cuda.synchronize()
    finally:
        _capi.lib().b200rl_gemm_set_tail_split(1)
    assert _rel_err(ref, a1.float() @ (dense.float() if b_mn else dense.float().T)
                    + (a2.float() @ (b2.float() if b_mn else b2.float().T) if K2 else 0) + res.float()) < 4e-3
    assert torch.equal(out.view(torch.int16), ref.view(torch.int16))


@pytest.mark.parametrize("tokens,splits", [(4446, 2), (700, 1), (1000, 4)])
def test_gemm_dw_grouped(cuda, tokens, splits):
    """All dB / dA products of a layer in one persistent launch (gemm_dw_grouped.cu) vs torch fp32."""
    import ctypes as C
    from distrl_llm_b200 import _capi
    rows = [3584, 18944, 1024, 256, 4608, 512, 72, 136]
    Y = [_rand((tokens, r), cuda, seed=10 + i) for i, r in enumerate(rows)]
    U = [_rand((tokens, 64), cuda, seed=30 + i) for i in range(len(rows))]
    out = [torch.full((splits, r, 64), float("nan"), device=cuda, dtype=torch.float32) for r in rows]
    n = len(rows)
    arr_p = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])
    arr_ll = lambda xs: (C.c_longlong * n)(*xs)
    yp, up, cp = arr_p(Y), arr_p(U), arr_p(out)
    ldy, ldu, strides = arr_ll(rows), arr_ll([64] * n), arr_ll([r * 64 for r in rows])
    rws = (C.c_int * n)(*rows)
    used = _capi.lib().b200rl_gemm_dw_grouped(n, yp, ldy, rws, up, ldu, cp, strides, tokens, splits, _capi.stream())
    assert used >= 1, _capi.lib().b200rl_last_error()
    torch.cuda.synchronize()
    for y, u, o in zip(Y, U, out):
        ref = y.float().T @ u.float()
        got = o[:used].sum(0)
        assert torch.isfinite(got).all()
        assert _rel_err(got, ref) < 2e-5


def test_gemm_epilogues(cuda, gemm_mode):
    from distrl_llm_b200 import ops
    M, N, K = 384, 320, 256
    a = _rand((M, K), cuda, seed=1)
    b = _rand((N, K), cuda, seed=2)
    bias = _rand((N,), cuda, seed=3)
    res = _rand((M, N), cuda, seed=4)
    ref = 0.5 * (a.float() @ b.float().T) + bias.float()[None] + res.float()
    out = ops.gemm(a, b, bias=bias, residual=res, alpha=0.5)
    assert _rel_err(out, ref) < 4e-3
    out32 = ops.gemm(a, b, bias=bias, residual=res, alpha=0.5, out_fp32=True)
    assert out32.dtype == torch.float32
    assert _rel_err(out32, ref) < 1e-5
    # in-place residual (C aliases the residual) — used for x += o_proj(...)
    res2 = res.clone()
    ops.gemm(a, b, residual=res2, out=res2)
    assert _rel_err(res2, a.float() @ b.float().T + res.float()) < 4e-3


@pytest.mark.parametrize("M,N,K1,K2,bn", [(128, 64, 64, 0, 0), (300, 200, 192, 0, 0), (1000, 512, 1024, 64, 0),
                                           (640, 3584, 4608, 64, 256), (4100, 1024, 2048, 64, 192),
                                           (512, 384, 256, 128, 128), (6896, 3584, 1024, 64, 0),
                                           (600, 3584, 512, 64, 224), (900, 1000, 320, 64, 224),
                                           (700, 1000, 320, 64, 192),
                                           (4446, 3584, 4608, 64, 512), (900, 1000, 320, 64, 512), (700, 600, 192, 0, 512),
                                           (8892, 3584, 18944, 64, 512)])
def test_gemm_dx_form(cuda, gemm_mode, M, N, K1, K2, bn):
    if bn in (224, 512) and not gemm_mode:
        pytest.skip("224-wide and 512-wide tiles exist only in the CTA-pair kernel")
    """dX form: C = A1 @ B1 + A2 @ B2 with the B operands stored [K, N] (weights as stored [out, in])."""
    from distrl_llm_b200 import ops
    a1 = _rand((M, K1), cuda, seed=1)
    b1 = _rand((K1, N), cuda, seed=2)
    a2 = _rand((M, K2), cuda, seed=3) if K2 else None
    b2 = _rand((K2, N), cuda, seed=4) if K2 else None
    out = ops.gemm(a1, b1, a2, b2, force_bn=bn, b_mn=True)
    ref = a1.float() @ b1.float()
    if K2:
        ref = ref + a2.float() @ b2.float()
    assert _rel_err(out, ref) < 4e-3


@pytest.mark.parametrize("tokens,Ny,Nu,splits", [
    (64, 128, 64, 1), (512, 256, 64, 1), (1000, 384, 64, 4), (6896, 512, 64, 8), (2048, 1024, 128, 3),
    (200, 136, 64, 2),
])
def test_gemm_dw(cuda, tokens, Ny, Nu, splits):
    """dW form (both operands MN-major): slabs.sum(0) == y.T @ u."""
    from distrl_llm_b200 import ops
    y = _rand((tokens, Ny), cuda, seed=5)
    u = _rand((tokens, Nu), cuda, seed=6)
    slabs = ops.gemm_dw(y, u, splits=splits)
    got = slabs.sum(0)
    ref = y.float().T @ u.float()
    err = _rel_err(got, ref)
    assert err < 1e-4, f"rel err {err}"


# ------------------------------------------------------------------------------------------------
# row kernels
# ------------------------------------------------------------------------------------------------
def test_embed(cuda):
    from distrl_llm_b200 import ops
    table = _rand((512, 128), cuda, seed=1)
    ids = torch.randint(0, 512, (300,), device=cuda, dtype=torch.int32)
    out = ops.embed(ids, table)
    assert torch.equal(out, table[ids.long()])


@pytest.mark.parametrize("M,H", [(7, 128), (300, 3584), (64, 1024)])
def test_rmsnorm(cuda, M, H):
    from distrl_llm_b200 import ops
    x = _rand((M, H), cuda, seed=1, scale=2.0)
    w = (1 + 0.1 * torch.randn(H)).to(cuda, torch.bfloat16)
    eps = 1e-6
    y, rstd = ops.rmsnorm_fwd(x, w, eps)
    xf = x.float().requires_grad_(True)
    var = xf.pow(2).mean(-1, keepdim=True)
    ref = w.float() * (xf * torch.rsqrt(var + eps))
    assert _rel_err(y, ref) < 6e-3
    assert _rel_err(rstd, torch.rsqrt(var + eps).squeeze(-1)) < 1e-5
    dy = _rand((M, H), cuda, seed=2)
    dres = _rand((M, H), cuda, seed=3)
    ref.backward(dy.float())
    dx = ops.rmsnorm_bwd(dy, x, w, rstd, dres)
    assert _rel_err(dx, xf.grad + dres.float()) < 6e-3


def _rope_ref(x, L, hd, theta):
    # HF apply_rotary_pos_emb / rotate_half in fp32
    M, nh, _ = x.shape
    pos = (torch.arange(M, device=x.device) % L).float()
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2, device=x.device).float() / hd))
    fr = pos[:, None] * inv[None]
    emb = torch.cat([fr, fr], -1)
    cos, sin = emb.cos()[:, None], emb.sin()[:, None]
    x1, x2 = x[..., : hd // 2], x[..., hd // 2:]
    rot = torch.cat([-x2, x1], -1)
    return x * cos + rot * sin


def test_rope(cuda):
    from distrl_llm_b200 import ops
    B, L, nq, nkv, hd, theta = 2, 48, 4, 2, 128, 1e6
    qkv = _rand((B * L, (nq + 2 * nkv) * hd), cuda, seed=1)
    orig = qkv.clone()
    cs = ops.rope_table(L, hd, theta, cuda)
    ops.rope_(qkv, cs, L, nq + nkv, hd)
    ref = _rope_ref(orig[:, : (nq + nkv) * hd].float().view(B * L, nq + nkv, hd), L, hd, theta)
    assert _rel_err(qkv[:, : (nq + nkv) * hd].float().view(B * L, nq + nkv, hd), ref) < 6e-3
    assert torch.equal(qkv[:, (nq + nkv) * hd:], orig[:, (nq + nkv) * hd:])  # v untouched
    # backward = transpose of the rotation: <R x, y> == <x, R^T y>
    y = _rand((B * L, (nq + 2 * nkv) * hd), cuda, seed=2)
    yt = y.clone()
    ops.rope_(yt, cs, L, nq + nkv, hd, backward=True)
    lhs = (qkv.float() * y.float())[:, : (nq + nkv) * hd].sum()
    rhs = (orig.float() * yt.float())[:, : (nq + nkv) * hd].sum()
    assert abs(lhs - rhs) / abs(lhs) < 2e-2


def test_swiglu(cuda):
    from distrl_llm_b200 import ops
    M, I = 50, 256
    gu = _rand((M, 2 * I), cuda, seed=1, scale=2.0)
    act = ops.swiglu_fwd(gu)
    g = gu[:, :I].float().requires_grad_(True)
    u = gu[:, I:].float().requires_grad_(True)
    ref = torch.nn.functional.silu(g) * u
    assert _rel_err(act, ref) < 6e-3
    dact = _rand((M, I), cuda, seed=2)
    ref.backward(dact.float())
    dgu = ops.swiglu_bwd(gu, dact)
    assert _rel_err(dgu[:, :I], g.grad) < 6e-3
    assert _rel_err(dgu[:, I:], u.grad) < 6e-3


# ------------------------------------------------------------------------------------------------
# G7 fused log-softmax / gather / grad
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,V", [(9, 512), (64, 152064), (33, 8192)])
def test_logprob(cuda, rows, V):
    from distrl_llm_b200 import ops
    logits = _rand((rows, V), cuda, seed=1, scale=3.0)
    tg = torch.randint(0, V, (rows,), device=cuda, dtype=torch.int32)
    coef = torch.randn(rows, device=cuda)
    coef[0] = 0.0
    z = logits.float().requires_grad_(True)
    lp_ref = torch.log_softmax(z, -1).gather(1, tg.long()[:, None]).squeeze(1)
    lp = ops.logprob(logits.clone(), tg)
    # tolerance: fp32 log-softmax of identical bf16 logits, |dlp| <= 1e-5 * scale
    assert (lp - lp_ref).abs().max().item() < 2e-5 * max(1.0, lp_ref.abs().max().item())
    (lp_ref * coef).sum().backward()
    work = logits.clone()
    lp2 = ops.logprob(work, tg, coef, write_grad=True)
    assert torch.equal(lp, lp2)
    # dlogits stored in bf16
    assert _rel_err(work, z.grad) < 6e-3
    assert (work[0] == 0).all()


def test_loss_coef_and_value(cuda):
    from distrl_llm_b200 import ops
    Bm, T, nb = 5, 40, 3
    mask = torch.zeros(Bm, T, dtype=torch.int32)
    lens = [40, 17, 1, 0, 23]
    for i, n in enumerate(lens):
        mask[i, :n] = 1
    mask = mask.to(cuda)
    adv = torch.tensor([0.5, -1.25, 2.0, 3.0, -0.1], dtype=torch.float64, device=cuda)
    coef, ln = ops.loss_coef(mask, adv, nb)
    assert ln.tolist() == lens
    for i, n in enumerate(lens):
        exp = 0.0 if n == 0 else -adv[i].item() / (n * Bm * nb)
        assert torch.allclose(coef[i, :n], torch.full((n,), exp, device=cuda, dtype=torch.float32))
        assert (coef[i, n:] == 0).all()
    lp = -torch.rand(Bm, T, device=cuda)
    acc = torch.zeros(1, dtype=torch.float64, device=cuda)
    ops.loss_value(lp, mask, adv, acc, grpo=False)
    ref = 0.0
    for i, n in enumerate(lens):
        if n:
            ref += adv[i].item() * lp[i, :n].double().sum().item() / n
    assert abs(acc.item() - (-ref / Bm)) < 1e-9
    ops.loss_value(lp, mask, adv, acc, grpo=True)  # accumulates (reference returns the SUM, quirk Q2)
    ref2 = -sum(adv[i].item() for i, n in enumerate(lens) if n) / Bm
    assert abs(acc.item() - (-ref / Bm + ref2)) < 1e-9


# ------------------------------------------------------------------------------------------------
# G9 advantages / top-k : bit-exact against numpy (reference distributed_trainer.py:262-294)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("G,C,k", [(3, 8, 8), (5, 16, 4), (2, 256, 128), (4, 7, 16), (3, 130, 50)])
def test_group_advantage_topk(cuda, G, C, k):
    from distrl_llm_b200 import ops
    rng = np.random.default_rng(C * 7 + k)
    fmt = rng.choice([0.0, 0.1, 0.2, 0.35], size=(G, C))
    acc = (rng.random((G, C)) < 0.3).astype(np.float64) + rng.random((G, C)) * 1e-3
    rewards = np.stack([fmt, acc], -1)
    vals, base, idx, val = ops.group_advantage_topk(torch.from_numpy(rewards).to(cuda), k, grpo=True)
    for g in range(G):
        s = rewards[g].sum(axis=1)
        adv = (s - np.mean(s)) / (np.std(s) + 1e-8)
        assert np.array_equal(vals[g].cpu().numpy(), adv), "advantages must be bit-identical to numpy"
        assert base[g].item() == np.mean(s)
        top = np.argsort(adv, kind="stable")[-k:]
        assert np.array_equal(idx[g].cpu().numpy(), top)
        assert np.array_equal(val[g].cpu().numpy(), adv[top])
    vals_pg, base_pg, _, _ = ops.group_advantage_topk(torch.from_numpy(rewards).to(cuda), k, grpo=False)
    assert np.array_equal(vals_pg.cpu().numpy(), rewards.sum(-1))


# ------------------------------------------------------------------------------------------------
# NF4
# ------------------------------------------------------------------------------------------------
NF4 = np.array([-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
                -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
                0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941,
                0.7229568362236023, 1.0], dtype=np.float32)


def test_nf4_roundtrip(cuda):
    from distrl_llm_b200 import ops
    rows, cols = 136, 256
    w = _rand((rows, cols), cuda, seed=1, scale=0.02)
    packed, absmax = ops.nf4_quantize(w)
    wf = w.float().cpu().numpy().reshape(-1, 64)
    am = np.abs(wf).max(1)
    assert np.array_equal(absmax.cpu().numpy(), am)
    codes = np.abs((wf / am[:, None])[..., None] - NF4[None, None]).argmin(-1)
    ref_packed = (codes[:, 0::2] << 4 | codes[:, 1::2]).astype(np.uint8).reshape(-1)
    got = packed.cpu().numpy()
    # ties at exact midpoints are measure-zero for random data
    assert (got != ref_packed).mean() < 1e-4
    deq = ops.nf4_dequant(packed, absmax, rows, cols)
    lo, hi = got & 15, got >> 4
    vals = np.stack([NF4[hi], NF4[lo]], -1).reshape(-1, 64) * am[:, None]
    ref = torch.from_numpy(vals.reshape(rows, cols)).to(torch.bfloat16)
    assert torch.equal(deq.cpu(), ref), "dequant must be bit-exact: bf16(code * absmax)"
    deq_t = ops.nf4_dequant(packed, absmax, rows, cols, transpose=True)
    assert torch.equal(deq_t.cpu(), ref.T.contiguous())


# ------------------------------------------------------------------------------------------------
# G8 Adam (single learner) vs torch.optim.Adam
# ------------------------------------------------------------------------------------------------
def test_adam_matches_torch(cuda):
    from distrl_llm_b200 import ops
    n = 4096 * 3 + 4
    torch.manual_seed(0)
    p0 = torch.randn(n, device=cuda) * 0.1
    p = p0.clone()
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=2e-5, foreach=False, fused=False)
    for step in range(1, 6):
        g = torch.randn(n, device=cuda) * 10 ** (step - 3)
        ref_p.grad = g.clone()
        opt.step()
        gbuf = g.clone()
        ops.adamw_step(p, m, v, gbuf, step, 2e-5)
        assert (gbuf == 0).all()
        # fp32, same operation order as torch's single-tensor Adam: <= 2 ulp
        assert torch.allclose(p, ref_p.data, rtol=3e-7, atol=1e-9), (p - ref_p.data).abs().max()


# ------------------------------------------------------------------------------------------------
# G4 attention fwd/bwd vs a torch fp32 restatement (causal AND key-padding mask, GQA)
# ------------------------------------------------------------------------------------------------
def _attn_ref(qkv, key_mask, B, L, nq, nkv, hd):
    x = qkv.float().view(B, L, nq + 2 * nkv, hd)
    q = x[:, :, :nq].permute(0, 2, 1, 3)
    k = x[:, :, nq:nq + nkv].permute(0, 2, 1, 3).repeat_interleave(nq // nkv, dim=1)
    v = x[:, :, nq + nkv:].permute(0, 2, 1, 3).repeat_interleave(nq // nkv, dim=1)
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    causal = torch.tril(torch.ones(L, L, dtype=torch.bool, device=qkv.device))
    ok = causal[None, None] & (key_mask.bool()[:, None, None, :])
    s = s.masked_fill(~ok, float("-inf"))
    p = torch.softmax(s, -1)
    p = torch.nan_to_num(p, nan=0.0)  # rows without any visible key -> zeros (library convention)
    o = p @ v
    return o.permute(0, 2, 1, 3).reshape(B * L, nq * hd)


@pytest.fixture(params=[1, 0], ids=["tcgen05", "mma_sync"])
def attn_mode(request, cuda):
    from distrl_llm_b200 import _capi
    _capi.lib().b200rl_attn_set_tc(request.param)
    yield request.param
    _capi.lib().b200rl_attn_set_tc(1)


@pytest.mark.parametrize("B,L,nq,nkv,hd", [(2, 48, 4, 2, 32), (2, 200, 4, 2, 64), (1, 333, 14, 2, 128),
                                            (3, 64, 4, 4, 128), (2, 130, 7, 1, 128), (2, 862, 4, 2, 128),
                                            (1, 1024, 2, 1, 128)])
def test_attention(cuda, attn_mode, B, L, nq, nkv, hd):
    from distrl_llm_b200 import ops
    qkv = _rand((B * L, (nq + 2 * nkv) * hd), cuda, seed=1)
    key_mask = torch.ones(B, L, dtype=torch.int32, device=cuda)
    key_mask[0, :5] = 0             # left padding (prompt side)
    key_mask[-1, L - 7:] = 0        # right padding (completion side)
    out, lse = ops.attn_fwd(qkv, key_mask, B, L, nq, nkv, hd)
    qr = qkv.float().requires_grad_(True)
    ref = _attn_ref(qr, key_mask, B, L, nq, nkv, hd)
    valid_q = torch.ones(B, L, dtype=torch.bool, device=cuda)
    valid_q[0, :5] = False          # queries that see no key: undefined in HF, zeros here
    vq = valid_q.view(-1)
    assert torch.isfinite(out.float()).all()
    assert _rel_err(out[vq], ref[vq]) < 8e-3
    assert (out[~vq] == 0).all()
    dout = _rand((B * L, nq * hd), cuda, seed=2)
    (ref * dout.float()).sum().backward()
    dqkv = ops.attn_bwd(qkv, key_mask, out, dout, lse, B, L, nq, nkv, hd)
    assert torch.isfinite(dqkv.float()).all()
    g = qr.grad
    nqh = nq * hd
    for name, sl in (("dq", slice(0, nqh)), ("dk", slice(nqh, nqh + nkv * hd)), ("dv", slice(nqh + nkv * hd, None))):
        err = _rel_err(dqkv[:, sl], g[:, sl])
        assert err < 2e-2, f"{name} rel err {err}"


@pytest.mark.parametrize("B,L,nq,nkv", [(2, 700, 4, 2), (1, 1024, 2, 1)])
def test_attention_lazy_rescale(cuda, B, L, nq, nkv):
    """The tcgen05 forward keeps O in TMEM against a LAZY reference maximum and rescales O in place only when a key block
    exceeds it by more than 2^8.  Keys whose magnitude grows along the sequence make the row maximum climb by far more
    than that, several times per row: output, lse (through the backward) and gradients must still match fp32."""
    from distrl_llm_b200 import ops
    hd = 128
    qkv = _rand((B * L, (nq + 2 * nkv) * hd), cuda, seed=5).float()
    ramp = (1.0 + 40.0 * torch.arange(L, device=cuda, dtype=torch.float32) / L).repeat(B).view(-1, 1)
    qkv[:, nq * hd:(nq + nkv) * hd] *= ramp
    qkv = qkv.to(torch.bfloat16)
    key_mask = torch.ones(B, L, dtype=torch.int32, device=cuda)
    key_mask[-1, L - 9:] = 0
    out, lse = ops.attn_fwd(qkv, key_mask, B, L, nq, nkv, hd)
    qr = qkv.float().requires_grad_(True)
    ref = _attn_ref(qr, key_mask, B, L, nq, nkv, hd)
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse[lse < float("inf")]).all()
    assert _rel_err(out, ref) < 8e-3
    # the scores really do climb: the fp32 row maxima of the last queries are far above those of the first block
    with torch.no_grad():
        q = qr[:L, :hd]
        k = qr[:L, nq * hd:nq * hd + hd]
        sc = (q @ k.t()) * hd ** -0.5 * 1.4426950408889634
        sc = sc.masked_fill(torch.triu(torch.ones(L, L, device=cuda, dtype=torch.bool), 1), float("-inf"))
        first = sc[-1, :128].max()
        assert sc[-1].max() - first > 16.0
    dout = _rand((B * L, nq * hd), cuda, seed=6)
    (ref * dout.float()).sum().backward()
    dqkv = ops.attn_bwd(qkv, key_mask, out, dout, lse, B, L, nq, nkv, hd)
    g = qr.grad
    nqh = nq * hd
    for name, sl in (("dq", slice(0, nqh)), ("dk", slice(nqh, nqh + nkv * hd)), ("dv", slice(nqh + nkv * hd, None))):
        err = _rel_err(dqkv[:, sl], g[:, sl])
        assert err < 2e-2, f"{name} rel err {err}"


# ------------------------------------------------------------------------------------------------
# packed "shared-prompt" attention (tcgen05, head_dim 128): each distinct prompt stored once
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ragged", [True, False], ids=["ragged", "padded"])
@pytest.mark.parametrize("groups,per_group,P,T", [(2, 3, 70, 130), (1, 8, 350, 512), (3, 1, 40, 64)])
def test_attention_packed_shared_prompt(cuda, groups, per_group, P, T, ragged):
    import ctypes as C
    from distrl_llm_b200 import _capi, packing
    nq, nkv, hd = 4, 2, 128
    B = groups * per_group
    rng = np.random.default_rng(P + T)
    ids = np.zeros((B, P + T), np.int32)
    am = np.zeros_like(ids)
    for g in range(groups):
        plen = int(rng.integers(P // 2, P + 1))
        pr = rng.integers(1, 1000, size=plen)
        for j in range(per_group):
            i = g * per_group + j
            ids[i, P - plen:P] = pr
            am[i, P - plen:P] = 1
            n = int(rng.integers(T // 4, T + 1))
            ids[i, P:P + n] = rng.integers(1, 1000, size=n)
            am[i, P:P + n] = 1
    host = packing.pack_microbatch(ids, am, P, T, ragged=ragged)
    assert host.n_groups == groups
    assert host.rows == (int(am.sum() - am[:, :P].sum() + sum(am[np.argmax(host.seq_group == g), :P].sum() for g in range(groups)))
                         if ragged else groups * P + B * T)
    pk = packing.PackedDevice(host, cuda)
    rows, cols = host.rows, (nq + 2 * nkv) * hd
    qkv = _rand((rows, cols), cuda, seed=1)
    dout = _rand((rows, nq * hd), cuda, seed=2)
    out = torch.empty(rows, nq * hd, device=cuda, dtype=torch.bfloat16)
    lse = torch.empty(nq, rows, device=cuda, dtype=torch.float32)
    delta = torch.empty_like(lse)
    dqkv = torch.zeros_like(qkv)
    kvpart = torch.empty(host.part_rows, 2 * nkv * hd, device=cuda, dtype=torch.float32)
    scale = hd ** -0.5
    lib = _capi.lib()
    c = pk.c
    _capi.check(lib.b200rl_attn_seg_fwd(qkv.data_ptr(), c.key_mask, out.data_ptr(), lse.data_ptr(), rows, nq, nkv, scale,
                                        c.qblocks, c.n_qblocks, _capi.stream()), "attn_seg_fwd")
    _capi.check(lib.b200rl_attn_seg_bwd(qkv.data_ptr(), c.key_mask, out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                                        delta.data_ptr(), dqkv.data_ptr(), kvpart.data_ptr(), rows, nq, nkv, scale,
                                        c.qblocks, c.n_qblocks, c.kblocks, c.n_kblocks, c.red_start, c.red_list,
                                        _capi.stream()), "attn_seg_bwd")
    torch.cuda.synchronize()
    # reference: expand the packed rows to the per-sequence [B, L] layout (gather; dropped pad positions read a zero
    # row and are masked as keys), plain fp32 attention with autograd
    idx = np.full((B, P + T), rows, np.int64)                       # `rows` = index of the appended zero row
    for i in range(B):
        g = int(host.seq_group[i])
        st, n = host.prompt_ext[g]
        idx[i, st:st + n] = host.prompt_row0[g] + np.arange(n)
        idx[i, P:P + host.comp_len[i]] = host.comp_row0[i] + np.arange(host.comp_len[i])
    idx_t = torch.from_numpy(idx).to(cuda)
    leaf = qkv.float().requires_grad_(True)
    full = torch.cat([leaf, torch.zeros(1, cols, device=cuda)], 0)[idx_t.view(-1)]      # [B*L, cols]
    key_mask = torch.from_numpy(am).to(cuda)
    ref_full = _attn_ref(full, key_mask, B, P + T, nq, nkv, hd)     # [B*L, nq*hd]
    got_full = torch.cat([out.float(), torch.zeros(1, nq * hd, device=cuda)], 0)[idx_t.view(-1)]
    vq = key_mask.view(-1).bool()
    assert torch.isfinite(out.float()).all()
    assert _rel_err(got_full[vq], ref_full[vq]) < 8e-3
    # backward: dout per packed row; expand the same way but count each shared prompt row ONCE (first sequence of group)
    w = torch.zeros(B, P + T, device=cuda)
    first = [int(np.argmax(host.seq_group == g)) for g in range(groups)]
    for i in range(B):
        w[i, P:] = 1.0
        if i in first:
            w[i, :P] = 1.0
    dfull = torch.cat([dout.float(), torch.zeros(1, nq * hd, device=cuda)], 0)[idx_t.view(-1)] * w.view(-1, 1)
    (ref_full * dfull).sum().backward()
    g_ref = leaf.grad                                               # already summed over the group through the gather
    valid = torch.from_numpy(host.arrays["key_mask"]).to(cuda).bool()
    nqh = nq * hd
    for name, sl in (("dq", slice(0, nqh)), ("dk", slice(nqh, nqh + nkv * hd)), ("dv", slice(nqh + nkv * hd, None))):
        err = _rel_err(dqkv[:, sl][valid], g_ref[:, sl][valid])
        assert err < 2e-2, f"{name} rel err {err}"