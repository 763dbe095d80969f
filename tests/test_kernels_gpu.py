"""GPU: every kernel of libgenconvit_b200.so, called through the C ABI, against a plain torch
fp32 restatement of the same operator on the same inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# the torch restatements must be true fp32 (cuDNN convs default to TF32)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

DEV = "cuda"
DTYPES = [torch.float32, torch.bfloat16, torch.float16]
TOL = {torch.float32: 2e-5, torch.bfloat16: 2e-2, torch.float16: 3e-3}


def _lib():
    from genconvit_b200 import lib
    lib.load()
    return lib


def _rand(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def _close(got, want, tol, what=""):
    err = (got.float() - want.float()).abs().max().item()
    ref = want.float().abs().max().item()
    assert err <= tol * max(1.0, ref), f"{what}: max|d|={err:.3e} (ref max {ref:.3e}, tol {tol})"


# --------------------------------------------------------------------------- GEMM
GEMM_SHAPES = [(128, 128, 64), (300, 200, 96), (1000, 96, 384), (257, 1000, 768), (64, 12, 16), (100, 96, 48),
               (4096, 1536, 384), (15, 500, 2000), (777, 384, 1536), (50, 256, 32)]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", GEMM_SHAPES)
def test_gemm_tcgen05_plain(shape, dtype):
    L = _lib()
    M, N, K = shape
    a, b = _rand(M, K, dtype=dtype, seed=1), _rand(N, K, dtype=dtype, seed=2, scale=K ** -0.5)
    d = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float32)
    L.gemm(a, b, d, M, N, K, out_f32=True, backend=L.GEMM_TCGEN05)
    torch.cuda.synchronize()
    _close(d, a.float() @ b.float().t(), 1e-4, f"tcgen05 {shape}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(128 * 300 + 7, 384, 1536), (128 * 296, 768, 512), (128 * 311 + 100, 384, 768)])
def test_gemm_tcgen05_wide_pair_tiles(shape, dtype):
    """256 x 384 cta_group::2 tiles (two N=192 MMAs side by side, one TMEM stage).  The auto-tuner prefers two 192-wide
    tiles, so the path is FORCED here with backend = 1000 + 384 (the tile-width test hook of gcv_gemm)."""
    L = _lib()
    M, N, K = shape
    a, b = _rand(M, K, dtype=dtype, seed=1), _rand(N, K, dtype=dtype, seed=2, scale=K ** -0.5)
    bias, gamma = _rand(N, seed=3, scale=0.3), _rand(N, seed=4).abs()
    res = _rand(M, N, dtype=dtype, seed=5)
    want = res.float() + gamma * (a.float() @ b.float().t() + bias)
    d = res.clone()
    L.gemm(a, b, d, M, N, K, bias=bias, gamma=gamma, residual=d, ldr=N, backend=1000 + 384)
    torch.cuda.synchronize()
    _close(d, want, TOL[dtype], f"wide tiles {shape}")
    d32 = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float32)
    L.gemm(a, b, d32, M, N, K, out_f32=True, backend=1000 + 384)
    torch.cuda.synchronize()
    _close(d32, a.float() @ b.float().t(), 1e-4, f"wide tiles fp32 out {shape}")


@pytest.mark.parametrize("block_n", [32, 64, 96, 128, 192, 256])
def test_gemm_tcgen05_every_tile_width(block_n):
    L = _lib()
    M, N, K = 1000, 700, 448
    a, b = _rand(M, K, dtype=torch.bfloat16, seed=3), _rand(N, K, dtype=torch.bfloat16, seed=4, scale=K ** -0.5)
    d = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float32)
    L.gemm(a, b, d, M, N, K, out_f32=True, backend=1000 + block_n)
    torch.cuda.synchronize()
    _close(d, a.float() @ b.float().t(), 1e-4, f"block_n={block_n}")


def test_gemm_tcgen05_many_tiles_persistent():
    """More tiles than SMs: every CTA loops over several tiles and both TMEM stages wrap."""
    L = _lib()
    M, N, K = 148 * 128 * 3 + 17, 384, 96
    a, b = _rand(M, K, dtype=torch.bfloat16, seed=5), _rand(N, K, dtype=torch.bfloat16, seed=6, scale=K ** -0.5)
    d = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
    L.gemm(a, b, d, M, N, K, backend=L.GEMM_TCGEN05)
    torch.cuda.synchronize()
    _close(d, a.float() @ b.float().t(), 1e-2, "persistent")


@pytest.mark.parametrize("shape", [(40000, 512, 512), (148 * 128 * 2 + 128 * 3 + 5, 384, 1536), (20000, 1536, 384)])
def test_gemm_tcgen05_cta_pair_tiles(shape):
    """cta_group::2 pair tiles (256 x block_n, each CTA stages its own A rows and half of B) -- the mode every large
    contraction takes (>= 2 waves of tiles, K >= 256) -- including an odd number of 128-row tiles, where the second
    CTA of the last pair works on a tile that lies entirely past M."""
    L = _lib()
    M, N, K = shape
    a, b = _rand(M, K, dtype=torch.bfloat16, seed=7), _rand(N, K, dtype=torch.bfloat16, seed=8, scale=K ** -0.5)
    bias = _rand(N, seed=9)
    res = _rand(M, N, dtype=torch.bfloat16, seed=10)
    want = res.float() + (a.float() @ b.float().t() + bias)
    L.gemm(a, b, res, M, N, K, bias=bias, residual=res, ldr=N, backend=L.GEMM_TCGEN05)
    torch.cuda.synchronize()
    _close(res, want, 1e-2, f"pair {shape}")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("backend", ["simt", "auto"])
def test_gemm_epilogues(dtype, backend):
    L = _lib()
    be = L.GEMM_SIMT if backend == "simt" else L.GEMM_AUTO
    tol = TOL[dtype]
    M, N, K = 333, 192, 96
    a, b = _rand(M, K, dtype=dtype, seed=1), _rand(N, K, dtype=dtype, seed=2, scale=K ** -0.5)
    bias, gamma = _rand(N, seed=3), _rand(N, seed=4)
    acc = a.float() @ b.float().t() + bias
    # bias + GELU(erf)
    d = torch.empty((M, N), device=DEV, dtype=dtype)
    L.gemm(a, b, d, M, N, K, bias=bias, act=L.ACT_GELU, backend=be)
    _close(d, F.gelu(acc), tol, "gelu")
    # ReLU / LeakyReLU(0.01)
    L.gemm(a, b, d, M, N, K, bias=bias, act=L.ACT_RELU, backend=be)
    _close(d, F.relu(acc), tol, "relu")
    L.gemm(a, b, d, M, N, K, bias=bias, act=L.ACT_LEAKY, backend=be)
    _close(d, F.leaky_relu(acc, 0.01), tol, "leaky")
    # residual + gamma * (acc + bias), in place on the residual
    res = _rand(M, N, dtype=dtype, seed=5)
    want = res.float() + gamma * acc
    L.gemm(a, b, res, M, N, K, bias=bias, gamma=gamma, residual=res, ldr=N, backend=be)
    _close(res, want, tol, "gamma+residual")
    # column window of a wider buffer (ldd) + fp32 output
    wide = torch.zeros((M, 2 * N + 8), device=DEV, dtype=torch.float32)
    L.gemm(a, b, wide[:, N:], M, N, K, bias=bias, ldd=2 * N + 8, out_f32=dtype != torch.float32, backend=be)
    _close(wide[:, N:2 * N], acc, tol if dtype == torch.float32 else 1e-4, "ldd window")
    assert wide[:, :N].abs().max().item() == 0 and wide[:, 2 * N:].abs().max().item() == 0


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("C", [96, 192])
@pytest.mark.parametrize("M", [1000, 148 * 128 * 2 + 77, 128])
def test_mlp_fused_matches_two_linear_layers(M, C, dtype):
    """x += gamma * (GELU(y W1^T + b1) W2^T + b2): the fused kernel vs fp32 torch (hidden never leaves the SM)."""
    L = _lib()
    assert L.mlp_fused_supported(dtype, C) and not L.mlp_fused_supported(dtype, 384)
    y, x = _rand(M, C, dtype=dtype, seed=1), _rand(M, C, dtype=dtype, seed=2)
    w1 = _rand(4 * C, C, dtype=dtype, seed=3, scale=C ** -0.5)
    w2 = _rand(C, 4 * C, dtype=dtype, seed=4, scale=(4 * C) ** -0.5)
    b1, b2, gamma = _rand(4 * C, seed=5, scale=0.3), _rand(C, seed=6, scale=0.3), _rand(C, seed=7).abs()
    hid = F.gelu(y.float() @ w1.float().t() + b1).to(dtype).float()       # the kernel keeps the hidden in 16-bit too
    want = x.float() + gamma * (hid @ w2.float().t() + b2)
    L.mlp_fused(y, w1, b1, w2, b2, gamma, x, M, C)
    torch.cuda.synchronize()
    _close(x, want, 2e-2 if dtype == torch.bfloat16 else 4e-3, f"mlp_fused M={M} C={C}")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cfg", [(2, 7, 7, 256, 128), (1, 14, 14, 64, 32), (2, 5, 3, 16, 3), (1, 28, 28, 32, 16)])
def test_gemm_pixel_shuffle_is_conv_transpose(cfg, dtype):
    """k2 s2 ConvTranspose2d as GEMM + pixel-shuffle store (reference genconvit_ed.py:44-56)."""
    L = _lib()
    from genconvit_b200.engine import _pack_convt
    B, H, W, ci, co = cfg
    x = _rand(B, ci, H, W, seed=1)
    w, bias = _rand(ci, co, 2, 2, seed=2, scale=ci ** -0.5), _rand(co, seed=3)
    want = F.leaky_relu(F.conv_transpose2d(x, w, bias, stride=2), 0.01).permute(0, 2, 3, 1)
    wp, bp, _ = _pack_convt(w, bias, DEV, dtype)
    tokens = x.permute(0, 2, 3, 1).reshape(B * H * W, ci).to(dtype).contiguous()
    out = torch.empty((B, 2 * H, 2 * W, co), device=DEV, dtype=dtype)
    L.gemm(tokens, wp, out, B * H * W, 4 * co, ci, bias=bp, act=L.ACT_LEAKY, store=L.STORE_PIXEL_SHUFFLE2, ps=(H, W, co))
    _close(out, want, TOL[dtype], f"convT {cfg}")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("act", ["relu", "leaky"])
def test_convt2x2_small_output_layer(dtype, act):
    """The decoders' 16 -> 3 output layer on the streaming kernel (genconvit_ed.py:56-57, genconvit_vae.py:77-78)."""
    L = _lib()
    from genconvit_b200.engine import _pack_convt
    B, H, W = 3, 9, 21
    x = _rand(B, 16, H, W, seed=1)
    w, bias = _rand(16, 3, 2, 2, seed=2, scale=0.25), _rand(3, seed=3)
    y = F.conv_transpose2d(x, w, bias, stride=2)
    want = (F.relu(y) if act == "relu" else F.leaky_relu(y, 0.01)).permute(0, 2, 3, 1)
    _, bp, wp = _pack_convt(w, bias, DEV, dtype)
    assert wp.dtype == torch.float32
    tokens = x.permute(0, 2, 3, 1).reshape(B * H * W, 16).to(dtype).contiguous()
    out = torch.full((B, 2 * H, 2 * W, 3), float("nan"), device=DEV, dtype=dtype)
    L.convt2x2_small(tokens, out, wp, bp, L.ACT_RELU if act == "relu" else L.ACT_LEAKY, B, H, W, 16, 3)
    _close(out, want, TOL[dtype], "convt2x2_small")


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("cfg", [(2, 8, 8, 64, "relu"), (3, 28, 28, 64, "leaky"), (1, 4, 12, 32, "relu"), (5, 56, 56, 32, "leaky")])
def test_convt2x2_mma_small_channel_layers(cfg, dtype):
    """ConvTranspose2d(64 -> 32) / (32 -> 16), k2 s2, + act on the HMMA kernel (genconvit_ed.py:51-55,
    genconvit_vae.py:56-62) against torch's conv_transpose2d."""
    L = _lib()
    from genconvit_b200.engine import _pack_convt
    B, H, W, ci, act = cfg
    co = ci // 2
    x = _rand(B, ci, H, W, seed=1).to(dtype).float()
    w, bias = _rand(ci, co, 2, 2, seed=2, scale=ci ** -0.5).to(dtype).float(), _rand(co, seed=3)
    y = F.conv_transpose2d(x, w, bias, stride=2)
    want = (F.relu(y) if act == "relu" else F.leaky_relu(y, 0.01)).permute(0, 2, 3, 1)
    wp, bp, _ = _pack_convt(w, bias, DEV, dtype)
    tokens = x.permute(0, 2, 3, 1).reshape(B * H * W, ci).to(dtype).contiguous()
    out = torch.full((B, 2 * H, 2 * W, co), float("nan"), device=DEV, dtype=dtype)
    L.convt2x2_mma(tokens, out, wp, bp, L.ACT_RELU if act == "relu" else L.ACT_LEAKY, B, H, W, ci)
    _close(out, want, TOL[dtype], f"convt2x2_mma {cfg}")


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("cfg", [(1, 8, 8, "relu"), (3, 28, 28, "leaky"), (1, 4, 12, "leaky"), (6, 56, 56, "relu")])
def test_convt2x2_mma_fused_output_tail(cfg, dtype):
    """The fused decoder tail ConvT(32 -> 16) + act + ConvT(16 -> 3) + act (genconvit_ed.py:53-57,
    genconvit_vae.py:60-64): the 16-channel intermediate is rounded to the activation type like the stored tensor
    it replaces, so the comparison target rounds it too."""
    L = _lib()
    from genconvit_b200.engine import _pack_convt
    B, H, W, act = cfg
    f = F.relu if act == "relu" else (lambda v: F.leaky_relu(v, 0.01))
    x = _rand(B, 32, H, W, seed=1).to(dtype).float()
    w1, b1 = _rand(32, 16, 2, 2, seed=2, scale=32 ** -0.5).to(dtype).float(), _rand(16, seed=3)
    w2, b2 = _rand(16, 3, 2, 2, seed=4, scale=0.25).to(dtype).float(), _rand(3, seed=5)
    mid = f(F.conv_transpose2d(x, w1, b1, stride=2)).to(dtype).float()
    want = f(F.conv_transpose2d(mid, w2, b2, stride=2)).permute(0, 2, 3, 1)
    p1, q1, _ = _pack_convt(w1, b1, DEV, dtype)
    p2, q2, _ = _pack_convt(w2, b2, DEV, dtype)
    tokens = x.permute(0, 2, 3, 1).reshape(B * H * W, 32).to(dtype).contiguous()
    out = torch.full((B, 4 * H, 4 * W, 3), float("nan"), device=DEV, dtype=dtype)
    L.convt2x2_mma(tokens, out, p1, q1, L.ACT_RELU if act == "relu" else L.ACT_LEAKY, B, H, W, 32, w2=p2, b2=q2)
    _close(out, want, TOL[dtype], f"convt2x2_mma tail {cfg}")


def test_convt2x2_mma_rejects_unsupported_shapes():
    L = _lib()
    x = torch.zeros(2 * 7 * 7, 32, device=DEV, dtype=torch.float16)
    y = torch.zeros(2 * 14 * 14, 16, device=DEV, dtype=torch.float16)
    w, b = torch.zeros(64, 32, device=DEV, dtype=torch.float16), torch.zeros(64, device=DEV)
    with pytest.raises(L.GcvError):
        L.convt2x2_mma(x, y, w, b, L.ACT_RELU, 2, 7, 7, 32)            # 98 tokens: not whole 16-token tiles (GEMM path)
    with pytest.raises(L.GcvError):
        L.convt2x2_mma(torch.zeros(128, 32, device=DEV), torch.zeros(512, 16, device=DEV), w.float(), b, L.ACT_RELU,
                       2, 8, 8, 32)                                    # fp32 mode keeps the GEMM path


@pytest.mark.parametrize("dtype", DTYPES)
def test_gemm_reparam_epilogue(dtype):
    """z = eps*exp(0.5*mu) + mu with eps given in the reference's (c*hw + hw) latent order."""
    L = _lib()
    M, K, C, HW = 5, 64, 16, 4
    N = C * HW
    a, b = _rand(M, K, dtype=dtype, seed=1), _rand(N, K, dtype=dtype, seed=2, scale=K ** -0.5)
    bias, eps = _rand(N, seed=3), _rand(M, N, seed=4)
    mu = a.float() @ b.float().t() + bias                       # columns n = hw*C + c
    eps_nhwc = eps.view(M, C, HW).transpose(1, 2).reshape(M, N)  # reference index c*HW + hw -> n
    want = eps_nhwc * torch.exp(0.5 * mu) + mu
    z = torch.empty((M, N), device=DEV, dtype=dtype)
    mu_out = torch.empty((M, N), device=DEV, dtype=torch.float32)
    L.gemm(a, b, z, M, N, K, bias=bias, eps=eps, eps_c=C, eps_hw=HW, mu_out=mu_out)
    _close(z, want, TOL[dtype], "reparam")
    _close(mu_out, mu, 1e-4, "mu_out")


# --------------------------------------------------------------------------- ConvNeXt kernels
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cfg", [(2, 56, 56, 96), (1, 28, 28, 192), (2, 14, 14, 384), (2, 7, 7, 768), (1, 3, 3, 768),
                                 (1, 9, 13, 96), (3, 28, 28, 96), (2, 5, 20, 192), (1, 9, 13, 64), (2, 8, 8, 160)])
def test_dwconv7_ln(cfg, dtype):
    L = _lib()
    B, H, W, C = cfg
    x = _rand(B, H, W, C, dtype=dtype, seed=1)
    w, bias = _rand(C, 1, 7, 7, seed=2, scale=1 / 7), _rand(C, seed=3, scale=0.1)
    lw, lb = _rand(C, seed=4).abs() + 0.5, _rand(C, seed=5, scale=0.1)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, bias, padding=3, groups=C).permute(0, 2, 3, 1)
    want = F.layer_norm(y, (C,), lw, lb, 1e-6)
    out = torch.full_like(x, float("nan"))
    L.dwconv7_ln(x, out, w.reshape(C, 49).t().contiguous(), bias, lw, lb, 1e-6, B, H, W, C)
    _close(out, want, TOL[dtype] * 2, f"dwconv {cfg}")


def _fold_ln_into_fc1(w1, b1, lw, lb, dtype):
    """W1' = W1 diag(lw) rounded to `dtype`, column sums of the rounded matrix, b1' = b1 + W1 lb (what engine.py packs)."""
    w1f = (w1.float() * lw[None, :]).to(dtype).contiguous()
    return w1f, w1f.float().sum(dim=1).contiguous(), (b1 + w1.float() @ lb).contiguous()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cfg", [(2, 56, 56, 96), (17, 28, 28, 192), (33, 14, 14, 384), (16, 7, 7, 768), (1, 3, 3, 768),
                                 (1, 9, 13, 96), (3, 5, 20, 64), (40, 7, 7, 384), (2, 8, 8, 160)])
def test_dwconv7_stats(cfg, dtype):
    """Tensor-core depthwise conv: un-normalised output + LayerNorm partial sums (sum, sumsq per 32-channel chunk)."""
    L = _lib()
    B, H, W, C = cfg
    x = _rand(B, H, W, C, dtype=dtype, seed=1)
    w, bias = _rand(C, 1, 7, 7, seed=2, scale=1 / 7), _rand(C, seed=3, scale=0.1)
    # the kernel rounds the taps to the activation type, like every other 16-bit weight
    wq = w.to(dtype).float()
    want = F.conv2d(x.float().permute(0, 3, 1, 2), wq, bias, padding=3, groups=C).permute(0, 2, 3, 1)
    out = torch.full_like(x, float("nan"))
    stats = torch.full((B * H * W, C // 32, 2), float("nan"), device=DEV)
    L.dwconv7_stats(x, out, stats, w.reshape(C, 49).t().contiguous(), bias, B, H, W, C)
    torch.cuda.synchronize()
    _close(out, want, TOL[dtype], f"dwconv7_stats {cfg}")
    # the statistics describe the ROUNDED output exactly (up to fp32 summation order)
    o = out.float().view(B * H * W, C // 32, 32)
    _close(stats[..., 0], o.sum(-1), 1e-5, "chunk sums")
    _close(stats[..., 1], (o * o).sum(-1), 1e-5, "chunk sums of squares")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(1000, 384, 96), (4096 + 13, 1536, 384), (300, 3072, 768), (128 * 148 * 2 + 5, 768, 192),
                                   (128 * 148 * 3 + 50, 64, 64), (128 * 148 * 5 + 1, 128, 128)])
def test_gemm_folded_layernorm(shape, dtype):
    """fc1(LayerNorm(v)) computed as GELU(rstd * (v W'^T - mean * colsum) + b') from per-chunk partial sums."""
    L = _lib()
    M, N, K = shape
    v = (_rand(M, K, seed=1) * 1.7 + 0.4).to(dtype)
    w1, b1 = _rand(N, K, seed=2, scale=K ** -0.5), _rand(N, seed=3, scale=0.3)
    lw, lb = _rand(K, seed=4).abs() + 0.5, _rand(K, seed=5, scale=0.2)
    w1f, cs, b1f = _fold_ln_into_fc1(w1, b1, lw, lb, dtype)
    vc = v.float().view(M, K // 32, 32)
    stats = torch.stack([vc.sum(-1), (vc * vc).sum(-1)], dim=-1).contiguous()
    want = F.gelu(F.layer_norm(v.float(), (K,), lw, lb, 1e-6) @ w1.float().t() + b1)
    d = torch.full((M, N), float("nan"), device=DEV, dtype=dtype)
    L.gemm(v, w1f, d, M, N, K, bias=b1f, act=L.ACT_GELU, ln_stats=stats, ln_colsum=cs, ln_eps=1e-6)
    torch.cuda.synchronize()
    _close(d, want, TOL[dtype], f"folded LN gemm {shape}")
    # the same with the partial sums reduced once by ln_finalize ([M, 2] rows of (rstd, -mean * rstd), ln_chunks = 0)
    rowstat = torch.full((M, 2), float("nan"), device=DEV)
    L.ln_finalize(stats, rowstat, M, K, 1e-6)
    mean, var = v.float().mean(1), v.float().var(1, unbiased=False)
    _close(rowstat[:, 0], (var + 1e-6).rsqrt(), 1e-4, "ln_finalize rstd")
    _close(rowstat[:, 1], -mean * (var + 1e-6).rsqrt(), 1e-4, "ln_finalize -mean*rstd")
    d1 = torch.full((M, N), float("nan"), device=DEV, dtype=dtype)
    L.gemm(v, w1f, d1, M, N, K, bias=b1f, act=L.ACT_GELU, ln_stats=rowstat, ln_colsum=cs, ln_eps=1e-6)
    _close(d1, want, TOL[dtype], f"folded LN gemm, finalised stats {shape}")
    # the GEMM's statistics warp reduces the partial sums in the same order as ln_finalize: identical outputs
    assert torch.equal(d, d1)
    # the SIMT back end implements the same epilogue
    if M <= 1000:
        d2 = torch.full((M, N), float("nan"), device=DEV, dtype=dtype)
        L.gemm(v, w1f, d2, M, N, K, bias=b1f, act=L.ACT_GELU, ln_stats=stats, ln_colsum=cs, ln_eps=1e-6, backend=L.GEMM_SIMT)
        _close(d2, want, TOL[dtype], f"folded LN simt {shape}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("C", [96, 192])
@pytest.mark.parametrize("M", [1000, 148 * 128 + 77])
def test_mlp_fused_folded_layernorm(M, C, dtype):
    """x += gamma * (GELU(LN(v) W1^T + b1) W2^T + b2) with the LayerNorm folded into the fused kernel's fc1 epilogue."""
    L = _lib()
    v, x = (_rand(M, C, seed=1) * 1.3 - 0.2).to(dtype), _rand(M, C, dtype=dtype, seed=2)
    w1, b1 = _rand(4 * C, C, seed=3, scale=C ** -0.5), _rand(4 * C, seed=5, scale=0.3)
    w2 = _rand(C, 4 * C, dtype=dtype, seed=4, scale=(4 * C) ** -0.5)
    b2, gamma = _rand(C, seed=6, scale=0.3), _rand(C, seed=7).abs()
    lw, lb = _rand(C, seed=8).abs() + 0.5, _rand(C, seed=9, scale=0.2)
    w1f, cs, b1f = _fold_ln_into_fc1(w1, b1, lw, lb, dtype)
    vc = v.float().view(M, C // 32, 32)
    stats = torch.stack([vc.sum(-1), (vc * vc).sum(-1)], dim=-1).contiguous()
    hid = F.gelu(F.layer_norm(v.float(), (C,), lw, lb, 1e-6) @ w1.float().t() + b1).to(dtype).float()
    want = x.float() + gamma * (hid @ w2.float().t() + b2)
    L.mlp_fused_ln(v, stats, 1e-6, w1f, b1f, cs, w2, b2, gamma, x, M, C)
    torch.cuda.synchronize()
    _close(x, want, 2e-2 if dtype == torch.bfloat16 else 4e-3, f"mlp_fused_ln M={M} C={C}")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cfg", [(2, 56, 56, 96), (1, 7, 7, 384), (2, 14, 14, 192)])
def test_ln_patchify2(cfg, dtype):
    L = _lib()
    B, H, W, C = cfg
    x = _rand(B, H, W, C, dtype=dtype, seed=1)
    lw, lb = _rand(C, seed=4).abs() + 0.5, _rand(C, seed=5, scale=0.1)
    ln = F.layer_norm(x.float(), (C,), lw, lb, 1e-6)
    Ho, Wo = H // 2, W // 2
    want = ln[:, :2 * Ho, :2 * Wo].reshape(B, Ho, 2, Wo, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(B * Ho * Wo, 4 * C)
    out = torch.full((B * Ho * Wo, 4 * C), float("nan"), device=DEV, dtype=dtype)
    L.ln_patchify2(x, out, lw, lb, 1e-6, B, H, W, C)
    _close(out, want, TOL[dtype], f"ln_patchify2 {cfg}")
    # and it is the im2col of the reference's 2x2 s2 conv: GEMM with the packed weight == F.conv2d
    wc = _rand(2 * C, C, 2, 2, seed=6, scale=(4 * C) ** -0.5)
    conv = F.conv2d(ln[:, :2 * Ho, :2 * Wo].permute(0, 3, 1, 2), wc, stride=2).permute(0, 2, 3, 1).reshape(-1, 2 * C)
    _close(want @ wc.permute(0, 2, 3, 1).reshape(2 * C, -1).t(), conv, 1e-4, "patchify order")


@pytest.mark.parametrize("dtype", DTYPES)
def test_stem_patchify_both_layouts(dtype):
    L = _lib()
    B, H, W = 2, 32, 24
    x = _rand(B, 3, H, W, seed=1)
    want = F.unfold(x, 4, stride=4).view(B, 3, 4, 4, -1).permute(0, 4, 2, 3, 1).reshape(B * (H // 4) * (W // 4), 48)
    out = torch.full((want.shape[0], 48), float("nan"), device=DEV, dtype=dtype)
    L.stem_patchify_nchw(x, out, B, H, W)
    _close(out, want, TOL[dtype], "stem nchw")
    xn = x.permute(0, 2, 3, 1).contiguous().to(dtype)
    out2 = torch.full_like(out, float("nan"))
    L.stem_patchify_nhwc(xn, out2, B, H, W)
    _close(out2, want, TOL[dtype], "stem nhwc")
    # column order matches the packed stem weight: GEMM == conv 4x4 s4
    w = _rand(96, 3, 4, 4, seed=2, scale=48 ** -0.5)
    conv = F.conv2d(x, w, stride=4).permute(0, 2, 3, 1).reshape(-1, 96)
    _close(want @ w.permute(0, 2, 3, 1).reshape(96, 48).t(), conv, 1e-4, "stem order")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C", [96, 768, 1536])
def test_layernorm_rows_and_pool_ln(C, dtype):
    L = _lib()
    rows = 77
    x = _rand(rows, C, dtype=dtype, seed=1) + 0.5
    lw, lb = _rand(C, seed=4).abs() + 0.5, _rand(C, seed=5, scale=0.1)
    out = torch.full_like(x, float("nan"))
    L.layernorm_rows(x, out, lw, lb, 1e-6, rows, C)
    _close(out, F.layer_norm(x.float(), (C,), lw, lb, 1e-6), TOL[dtype], "layernorm_rows")
    B, HW = 7, 11
    xp = _rand(B, HW, C, dtype=dtype, seed=2)
    pooled = torch.full((B, C), float("nan"), device=DEV, dtype=dtype)
    L.pool_ln(xp, pooled, lw, lb, 1e-6, B, HW, C)
    _close(pooled, F.layer_norm(xp.float().mean(1), (C,), lw, lb, 1e-6), TOL[dtype], "pool_ln")


# --------------------------------------------------------------------------- autoencoder kernels
@pytest.mark.parametrize("dtype", DTYPES)
def test_conv3x3_first_variants(dtype):
    L = _lib()
    B, H, W = 2, 20, 28
    x = _rand(B, 3, H, W, seed=1)
    w, b = _rand(16, 3, 3, 3, seed=2, scale=27 ** -0.5), _rand(16, seed=3, scale=0.1)
    # ED: stride 1 + ReLU + maxpool2
    want = F.max_pool2d(F.relu(F.conv2d(x, w, b, padding=1)), 2).permute(0, 2, 3, 1)
    out = torch.full((B, H // 2, W // 2, 16), float("nan"), device=DEV, dtype=dtype)
    L.conv3x3_first(x, out, w, b, 1, L.ACT_RELU, True, B, H, W)
    _close(out, want, TOL[dtype], "conv3x3 s1 relu pool")
    # VAE: stride 2 + LeakyReLU
    want = F.leaky_relu(F.conv2d(x, w, b, stride=2, padding=1), 0.01).permute(0, 2, 3, 1)
    out = torch.full((B, H // 2, W // 2, 16), float("nan"), device=DEV, dtype=dtype)
    L.conv3x3_first(x, out, w, b, 2, L.ACT_LEAKY, False, B, H, W)
    _close(out, want, TOL[dtype], "conv3x3 s2 leaky")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("stride", [1, 2])
def test_im2col3x3_is_conv(stride, dtype):
    L = _lib()
    from genconvit_b200.engine import _pack_conv3x3
    B, H, W, C, Co = 2, 14, 10, 16, 32
    x = _rand(B, H, W, C, dtype=dtype, seed=1)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    a = torch.full((B * Ho * Wo, 9 * C), float("nan"), device=DEV, dtype=dtype)
    L.im2col3x3(x, a, B, H, W, C, stride)
    w = _rand(Co, C, 3, 3, seed=2, scale=(9 * C) ** -0.5)
    want = F.conv2d(x.float().permute(0, 3, 1, 2), w, stride=stride, padding=1).permute(0, 2, 3, 1).reshape(-1, Co)
    _close(a.float() @ _pack_conv3x3(w, DEV, torch.float32).t(), want, 1e-4, "im2col order")


@pytest.mark.parametrize("dtype", DTYPES)
def test_maxpool2_resize_and_layout(dtype):
    L = _lib()
    B, H, W, C = 2, 12, 10, 16
    x = _rand(B, H, W, C, dtype=dtype, seed=1)
    out = torch.empty((B, H // 2, W // 2, C), device=DEV, dtype=dtype)
    L.maxpool2(x, out, B, H, W, C)
    assert torch.equal(out.float(), F.max_pool2d(x.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1))
    img = _rand(B, 16, 12, 3, dtype=dtype, seed=2)
    up = torch.empty((B, 3, 32, 24), device=DEV, dtype=torch.float32)
    L.resize2x_to_nchw(img, up, B, 16, 12, 3)
    want = F.interpolate(img.float().permute(0, 3, 1, 2), size=(32, 24), mode="bilinear", align_corners=False,
                         antialias=True)
    _close(up, want, 1e-5, "resize2x")
    nchw = torch.empty((B, 3, 16, 12), device=DEV, dtype=torch.float32)
    L.nhwc_to_nchw_f32(img, nchw, B, 16, 12, 3)
    assert torch.equal(nchw, img.float().permute(0, 3, 1, 2))


def test_score_videos_matches_pred_func_semantics():
    from genconvit_b200 import engine
    from oracle import nets
    n_nets, fpv, V = 2, 15, 5
    logits = _rand(n_nets * V * fpv, 2, seed=9)
    mean, cls, val = engine.score_videos(logits, n_nets, V * fpv, fpv)
    lg = logits.cpu().view(n_nets, V, fpv, 2)
    for v in range(V):
        want_cls, want_val = nets.pred_vid(lg[:, v].reshape(-1, 2))
        assert int(cls[v]) == want_cls and abs(float(val[v]) - want_val) < 1e-6
    # tie -> else branch, class 0
    tie = torch.zeros(4, 2, device=DEV)
    _, cls, val = engine.score_videos(tie, 1, 4, 4)
    assert int(cls[0]) == 0 and abs(float(val[0]) - 0.5) < 1e-7


def test_bad_arguments_return_errors_not_crashes():
    L = _lib()
    x = torch.zeros(4, 10, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(L.GcvError):
        L.gemm(x, x, x, 4, 4, 10, backend=L.GEMM_TCGEN05)          # K % 8 != 0
    with pytest.raises(L.GcvError):
        L.dwconv7_ln(x, x, x, x, x, x, 1e-6, 1, 2, 2, 10)           # C % 32 != 0
    with pytest.raises(L.GcvError):
        L.gemm(torch.zeros(4, 8), torch.zeros(4, 8), torch.zeros(4, 4), 4, 4, 8)   # CPU tensors: no fallback


# --------------------------------------------------------------------------- Swin kernels / frame ingest
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cfg", [(2, 56, 96, 3, 0), (2, 56, 96, 3, 3), (3, 28, 192, 6, 3), (2, 14, 384, 12, 3),
                                 (2, 14, 384, 12, 0), (5, 7, 768, 24, 0)])
def test_swin_window_attention(cfg, dtype):
    """softmax(q k^T / sqrt(32) + rel-pos bias + shifted-window mask) v per 7x7 window and head, with the cyclic shift,
    window partition and reverse folded into the indexing -- vs the oracle's roll / partition / mask restatement."""
    from oracle import backbones as OB
    L = _lib()
    B, res, C, heads, shift = cfg
    qkv = _rand(B * res * res, 3 * C, dtype=dtype, seed=1)
    table = _rand(169, heads, seed=2, scale=0.5)
    out = torch.full((B * res * res, C), float("nan"), device=DEV, dtype=dtype)
    L.swin_window_attention(qkv, out, table, B, res, C, heads, shift)
    torch.cuda.synchronize()
    # torch restatement on the same (rounded) qkv, following oracle/backbones.py swin_block / swin_window_attention
    h = qkv.float().cpu().view(B, res, res, 3 * C)
    if shift:
        h = torch.roll(h, (-shift, -shift), (1, 2))
    hw = OB._window_partition(h, 7).view(-1, 49, 3, heads, 32).permute(2, 0, 3, 1, 4)
    q, k, v = hw[0] * 32 ** -0.5, hw[1], hw[2]
    attn = q @ k.transpose(-2, -1)
    idx = OB.swin_relative_position_index(7)
    attn = attn + table.cpu()[idx.view(-1)].view(49, 49, heads).permute(2, 0, 1).unsqueeze(0)
    if shift:
        mask = OB.swin_attn_mask(res, 7, shift)
        nw = mask.shape[0]
        attn = (attn.view(B, nw, heads, 49, 49) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, 49, 49)
    o = (attn.softmax(-1) @ v).transpose(1, 2).reshape(-1, 7, 7, C)
    o = OB._window_reverse(o, 7, res, res)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    _close(out.cpu(), o.reshape(B * res * res, C), TOL[dtype], f"swin attention {cfg}")


@pytest.mark.parametrize("dtype", DTYPES)
def test_swin_patch_merge_and_mean_tokens(dtype):
    L = _lib()
    B, res, C = 3, 14, 192
    x = _rand(B, res, res, C, dtype=dtype, seed=1)
    out = torch.empty(B, res // 2, res // 2, 4 * C, device=DEV, dtype=dtype)
    L.swin_patch_merge(x, out, B, res, C)
    want = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    assert torch.equal(out, want)
    y = torch.empty(B, C, device=DEV, dtype=dtype)
    L.mean_tokens(x, y, B, res * res, C)
    _close(y, x.float().view(B, -1, C).mean(1), TOL[dtype], "mean_tokens")


@pytest.mark.parametrize("shape", [(1, 224, 224), (15, 224, 224), (4, 6, 10)])
def test_preprocess_frames_bit_exact(shape):
    """uint8 NHWC -> ((x / 255) - mean) / std fp32 NCHW: bit-identical to the reference's CPU arithmetic
    (model/pred_func.py:95-108 + dataset/loader.py:63-77)."""
    L = _lib()
    N, H, W = shape
    g = torch.Generator().manual_seed(5)
    frames = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, generator=g)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    want = frames.float().permute(0, 3, 1, 2) / 255.0
    want = (want - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    y = torch.full((N, 3, H, W), float("nan"), device=DEV)
    L.preprocess_frames(frames.to(DEV), y, N, H, W, mean, std)
    assert torch.equal(y.cpu(), want)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(2, 32, 32), (3, 20, 28), (1, 112, 112), (2, 18, 50)])
def test_conv3x3_c16_direct(shape, dtype):
    """Direct 16 -> 32 conv on mma.sync vs torch conv2d (fp32) on the same 16-bit-rounded inputs:
    stride 1 + ReLU + 2x2 max-pool (genconvit_ed.py:18-20) and stride 2 + LeakyReLU (genconvit_vae.py:19-21)."""
    L = _lib()
    from genconvit_b200.engine import _pack_conv3x3
    B, H, W = shape
    x = _rand(B, H, W, 16, dtype=dtype, seed=1)
    w = _rand(32, 16, 3, 3, seed=2, scale=144 ** -0.5)
    b = _rand(32, seed=3, scale=0.1)
    wp = _pack_conv3x3(w, DEV, dtype)
    wr = wp.float().reshape(32, 3, 3, 16).permute(0, 3, 1, 2)          # the rounded weights the kernel sees
    xin = x.float().permute(0, 3, 1, 2)
    want = F.max_pool2d(F.relu(F.conv2d(xin, wr, b, padding=1)), 2).permute(0, 2, 3, 1)
    out = torch.full((B, H // 2, W // 2, 32), float("nan"), device=DEV, dtype=dtype)
    L.conv3x3_c16(x, out, wp, b, 1, L.ACT_RELU, True, B, H, W)
    _close(out, want, TOL[dtype], "conv3x3_c16 s1 relu pool")
    want = F.leaky_relu(F.conv2d(xin, wr, b, stride=2, padding=1), 0.01).permute(0, 2, 3, 1)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.full((B, Ho, Wo, 32), float("nan"), device=DEV, dtype=dtype)
    L.conv3x3_c16(x, out, wp, b, 2, L.ACT_LEAKY, False, B, H, W)
    _close(out, want, TOL[dtype], "conv3x3_c16 s2 leaky")
    want = F.relu(F.conv2d(xin, wr, b, padding=1)).permute(0, 2, 3, 1)
    out = torch.full((B, H, W, 32), float("nan"), device=DEV, dtype=dtype)
    L.conv3x3_c16(x, out, wp, b, 1, L.ACT_RELU, False, B, H, W)
    _close(out, want, TOL[dtype], "conv3x3_c16 s1 relu")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(2, 224, 224), (3, 112, 112), (5, 32, 48), (3, 20, 28)])
def test_stem_fused(shape, dtype):
    """Fused stem (Conv2d(3,96,4,4) + LayerNorm2d) vs torch conv2d + layer_norm in fp32 on the same 16-bit-rounded
    operands, for both input layouts (fp32 NCHW frames, 16-bit NHWC reconstructions)."""
    L = _lib()
    B, H, W = shape
    x = _rand(B, 3, H, W, seed=1)
    w = _rand(96, 3, 4, 4, seed=2, scale=48 ** -0.5)
    b, lw, lb = _rand(96, seed=3, scale=0.1), 1 + _rand(96, seed=4, scale=0.1), _rand(96, seed=5, scale=0.1)
    wr = w.to(dtype).float()
    M = B * (H // 4) * (W // 4)

    def ref(xin):
        t = F.conv2d(xin, wr, b, stride=4).permute(0, 2, 3, 1)
        return F.layer_norm(t, (96,), lw, lb, 1e-6).reshape(M, 96)
    out = torch.full((M, 96), float("nan"), device=DEV, dtype=dtype)
    L.stem_fused(x, out, w.reshape(96, 48).to(dtype).contiguous(), b, lw, lb, 1e-6, B, H, W, True)
    _close(out, ref(x.to(dtype).float()), TOL[dtype], "stem_fused nchw")
    xh = x.permute(0, 2, 3, 1).contiguous().to(dtype)
    out = torch.full((M, 96), float("nan"), device=DEV, dtype=dtype)
    L.stem_fused(xh, out, w.permute(0, 2, 3, 1).reshape(96, 48).to(dtype).contiguous(), b, lw, lb, 1e-6, B, H, W, False)
    _close(out, ref(xh.float().permute(0, 3, 1, 2)), TOL[dtype], "stem_fused nhwc")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(2, 32, 32), (3, 20, 28), (1, 56, 56), (2, 18, 50)])
def test_conv3x3_c32_direct(shape, dtype):
    """Direct 32 -> 64 conv (weights in shared memory, ldmatrix operands) vs torch conv2d (fp32) on the same rounded
    inputs: stride 1 + ReLU + 2x2 max-pool (genconvit_ed.py:22-24), stride 2 + LeakyReLU (genconvit_vae.py:22-24)."""
    L = _lib()
    from genconvit_b200.engine import _pack_conv3x3
    B, H, W = shape
    x = _rand(B, H, W, 32, dtype=dtype, seed=1)
    w = _rand(64, 32, 3, 3, seed=2, scale=288 ** -0.5)
    b = _rand(64, seed=3, scale=0.1)
    wp = _pack_conv3x3(w, DEV, dtype)
    wr = wp.float().reshape(64, 3, 3, 32).permute(0, 3, 1, 2)
    xin = x.float().permute(0, 3, 1, 2)
    want = F.max_pool2d(F.relu(F.conv2d(xin, wr, b, padding=1)), 2).permute(0, 2, 3, 1)
    out = torch.full((B, H // 2, W // 2, 64), float("nan"), device=DEV, dtype=dtype)
    L.conv3x3_c32(x, out, wp, b, 1, L.ACT_RELU, True, B, H, W)
    _close(out, want, TOL[dtype], "conv3x3_c32 s1 relu pool")
    want = F.leaky_relu(F.conv2d(xin, wr, b, stride=2, padding=1), 0.01).permute(0, 2, 3, 1)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.full((B, Ho, Wo, 64), float("nan"), device=DEV, dtype=dtype)
    L.conv3x3_c32(x, out, wp, b, 2, L.ACT_LEAKY, False, B, H, W)
    _close(out, want, TOL[dtype], "conv3x3_c32 s2 leaky")
    want = F.relu(F.conv2d(xin, wr, b, padding=1)).permute(0, 2, 3, 1)
    out = torch.full((B, H, W, 64), float("nan"), device=DEV, dtype=dtype)
    L.conv3x3_c32(x, out, wp, b, 1, L.ACT_RELU, False, B, H, W)
    _close(out, want, TOL[dtype], "conv3x3_c32 s1 relu")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cfg", [(2, 28, 28, 64, 128), (3, 14, 14, 128, 256), (5, 20, 12, 64, 128), (1, 8, 8, 128, 256),
                                 (170, 28, 28, 64, 128), (7, 6, 6, 128, 64)])
def test_conv3x3_tc_implicit_gemm(cfg, dtype):
    """tcgen05 implicit-GEMM 3x3 conv (TMA-gathered taps, zero-filled padding) vs torch conv2d (fp32) on the same rounded
    inputs: stride 1 + ReLU + 2x2 max-pool (genconvit_ed.py:26-32), stride 2 + LeakyReLU (genconvit_vae.py:25-27), and
    stride 1 without pooling.  (170 frames x 8 tiles > one round of the 148 persistent CTAs: both accumulator stages
    and the smem ring wrap.)"""
    L = _lib()
    from genconvit_b200.engine import _pack_conv3x3
    B, H, W, ci, co = cfg
    x = _rand(B, H, W, ci, dtype=dtype, seed=1)
    w = _rand(co, ci, 3, 3, seed=2, scale=(9 * ci) ** -0.5)
    b = _rand(co, seed=3, scale=0.1)
    wp = _pack_conv3x3(w, DEV, dtype)
    wr = wp.float().reshape(co, 3, 3, ci).permute(0, 3, 1, 2)
    xin = x.float().permute(0, 3, 1, 2)
    want = F.max_pool2d(F.relu(F.conv2d(xin, wr, b, padding=1)), 2).permute(0, 2, 3, 1)
    out = torch.full((B, H // 2, W // 2, co), float("nan"), device=DEV, dtype=dtype)
    L.conv3x3_tc(x, out, wp, b, 1, L.ACT_RELU, True, B, H, W, ci, co)
    _close(out, want, TOL[dtype], "conv3x3_tc s1 relu pool")
    want = F.leaky_relu(F.conv2d(xin, wr, b, stride=2, padding=1), 0.01).permute(0, 2, 3, 1)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.full((B, Ho, Wo, co), float("nan"), device=DEV, dtype=dtype)
    L.conv3x3_tc(x, out, wp, b, 2, L.ACT_LEAKY, False, B, H, W, ci, co)
    _close(out, want, TOL[dtype], "conv3x3_tc s2 leaky")
    want = F.relu(F.conv2d(xin, wr, b, padding=1)).permute(0, 2, 3, 1)
    out = torch.full((B, H, W, co), float("nan"), device=DEV, dtype=dtype)
    L.conv3x3_tc(x, out, wp, b, 1, L.ACT_RELU, False, B, H, W, ci, co)
    _close(out, want, TOL[dtype], "conv3x3_tc s1 relu")


def test_conv3x3_tc_rejects_unsupported():
    L = _lib()
    assert not L.conv3x3_tc_supported(torch.float32, 64, 128)
    assert not L.conv3x3_tc_supported(torch.float16, 32, 64)
    x = torch.zeros(1, 7, 7, 64, device=DEV, dtype=torch.float16)
    w, b = torch.zeros(128, 576, device=DEV, dtype=torch.float16), torch.zeros(128, device=DEV)
    with pytest.raises(L.GcvError):                    # pooling needs even conv output sizes
        L.conv3x3_tc(x, torch.zeros(1, 3, 3, 128, device=DEV, dtype=torch.float16), w, b, 1, L.ACT_RELU, True, 1, 7, 7, 64, 128)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_uint8_first_touch_kernels_equal_preprocess_then_fp32_kernels(dtype):
    """gcv_conv3x3_first_u8 / gcv_stem_fused_u8 on raw uint8 NHWC crops == gcv_preprocess_frames followed by the fp32-input
    kernels, bit for bit (reference model/pred_func.py:95-108 then genconvit_ed.py:15-16 / genconvit_vae.py:16-18 /
    the ConvNeXt stem): same operand bits, same K order."""
    L = _lib()
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    B, H, W = 3, 64, 96
    g = torch.Generator().manual_seed(11)
    u8 = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8).to(DEV)
    u8[0, :2] = 0
    u8[0, 2:4] = 255
    x = torch.empty((B, 3, H, W), device=DEV)
    L.preprocess_frames(u8, x, B, H, W, mean, std)
    w1, b1 = _rand(16, 3, 3, 3, seed=2, scale=0.2), _rand(16, seed=3, scale=0.1)
    for stride, act, pool in ((1, L.ACT_RELU, True), (2, L.ACT_LEAKY, False), (1, L.ACT_RELU, False)):
        ho, wo = ((H - 1) // stride + 1) // (2 if pool else 1), ((W - 1) // stride + 1) // (2 if pool else 1)
        a = torch.full((B, ho, wo, 16), float("nan"), device=DEV, dtype=dtype)
        b = torch.full_like(a, float("nan"))
        L.conv3x3_first(x, a, w1, b1, stride, act, pool, B, H, W)
        L.conv3x3_first_u8(u8, b, w1, b1, stride, act, pool, B, H, W, mean, std)
        assert torch.equal(a, b), (stride, pool)
    ws = _rand(96, 3, 4, 4, seed=4, scale=0.15).reshape(96, 48).to(dtype).contiguous()
    bs, lw, lb = _rand(96, seed=5, scale=0.1), _rand(96, seed=6).abs() + 0.5, _rand(96, seed=7, scale=0.2)
    M = B * (H // 4) * (W // 4)
    a = torch.full((M, 96), float("nan"), device=DEV, dtype=dtype)
    b = torch.full_like(a, float("nan"))
    L.stem_fused(x, a, ws, bs, lw, lb, 1e-6, B, H, W, True)
    L.stem_fused_u8(u8, b, ws, bs, lw, lb, 1e-6, B, H, W, mean, std)
    assert torch.equal(a, b)


def _guarded(shape, dtype, pad=4096):
    """An output tensor carved out of a larger buffer whose margins hold a sentinel (compute-sanitizer is not available on
    the GPU pool: stray writes are caught by comparing the margins afterwards)."""
    n = 1
    for s in shape:
        n *= s
    buf = torch.full((n + 2 * pad,), 7.0, device=DEV, dtype=dtype)
    return buf, buf[pad:pad + n].view(shape), pad


def _margins_intact(buf, pad):
    return bool((buf[:pad] == 7).all()) and bool((buf[-pad:] == 7).all())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_round2_kernels_do_not_write_outside_their_outputs(dtype):
    """conv3x3_tc (ragged tiles: 20 x 12 and 7 x 7 maps, odd batch), convt2x2_mma (both forms), the uint8 first-touch kernels
    and ln_patchify2 write exactly their output tensors."""
    L = _lib()
    from genconvit_b200.engine import _pack_conv3x3, _pack_convt
    # conv3x3_tc: pooled, stride-2 (8 x 8 x 2 tile boxes with an odd image count), plain
    for (B, H, W, ci, co, stride, pool) in ((3, 20, 12, 64, 128, 1, True), (3, 14, 14, 128, 256, 2, False), (1, 6, 10, 64, 64, 1, False)):
        x = _rand(B, H, W, ci, dtype=dtype, seed=1)
        wp = _pack_conv3x3(_rand(co, ci, 3, 3, seed=2, scale=(9 * ci) ** -0.5), DEV, dtype)
        ho, wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        if pool:
            ho, wo = ho // 2, wo // 2
        buf, out, pad = _guarded((B, ho, wo, co), dtype)
        L.conv3x3_tc(x, out, wp, _rand(co, seed=3), stride, L.ACT_RELU, pool, B, H, W, ci, co)
        torch.cuda.synchronize()
        assert _margins_intact(buf, pad) and not bool((out == 7).all()), ("conv3x3_tc", B, H, W, ci, co, stride, pool)
    # convt2x2_mma: 64 -> 32 and the fused tail
    B, H, W = 3, 4, 12
    p1, q1, _ = _pack_convt(_rand(64, 32, 2, 2, seed=4, scale=0.1), _rand(32, seed=5), DEV, dtype)
    buf, out, pad = _guarded((B, 2 * H, 2 * W, 32), dtype)
    L.convt2x2_mma(_rand(B * H * W, 64, dtype=dtype, seed=6), out, p1, q1, L.ACT_RELU, B, H, W, 64)
    torch.cuda.synchronize()
    assert _margins_intact(buf, pad)
    p1, q1, _ = _pack_convt(_rand(32, 16, 2, 2, seed=7, scale=0.1), _rand(16, seed=8), DEV, dtype)
    p2, q2, _ = _pack_convt(_rand(16, 3, 2, 2, seed=9, scale=0.2), _rand(3, seed=10), DEV, dtype)
    buf, out, pad = _guarded((B, 4 * H, 4 * W, 3), dtype)
    L.convt2x2_mma(_rand(B * H * W, 32, dtype=dtype, seed=11), out, p1, q1, L.ACT_LEAKY, B, H, W, 32, w2=p2, b2=q2)
    torch.cuda.synchronize()
    assert _margins_intact(buf, pad)
    # uint8 first-touch kernels
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    B, H, W = 2, 32, 40
    u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=DEV)
    buf, out, pad = _guarded((B, H // 2, W // 2, 16), dtype)
    L.conv3x3_first_u8(u8, out, _rand(16, 3, 3, 3, seed=12, scale=0.2), _rand(16, seed=13), 1, L.ACT_RELU, True, B, H, W, mean, std)
    torch.cuda.synchronize()
    assert _margins_intact(buf, pad)
    buf, out, pad = _guarded((B * (H // 4) * (W // 4), 96), dtype)
    L.stem_fused_u8(u8, out, _rand(96, 48, seed=14, scale=0.15).to(dtype), _rand(96, seed=15), _rand(96, seed=16).abs() + 0.5,
                    _rand(96, seed=17), 1e-6, B, H, W, mean, std)
    torch.cuda.synchronize()
    assert _margins_intact(buf, pad)
    # ln_patchify2 (packed arithmetic), odd height: the last row is dropped
    B, H, W, C = 3, 7, 6, 96
    buf, out, pad = _guarded((B * (H // 2) * (W // 2), 4 * C), dtype)
    L.ln_patchify2(_rand(B, H, W, C, dtype=dtype, seed=18), out, _rand(C, seed=19).abs() + 0.5, _rand(C, seed=20), 1e-6, B, H, W, C)
    torch.cuda.synchronize()
    assert _margins_intact(buf, pad) and bool(torch.isfinite(out.float()).all())
