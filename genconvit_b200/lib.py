"""ctypes binding of libgenconvit_b200.so (the C ABI in include/genconvit_b200.h).

The product path has no CPU or library fallback: if the shared library is
missing, or a kernel call fails, this raises.  torch is used only for device
memory (``tensor.data_ptr()``) and the current stream.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GCV_LIB") or os.path.join(_HERE, "libgenconvit_b200.so")   # GCV_LIB: A/B builds in tools/

F32, BF16, F16 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_RELU, ACT_LEAKY = 0, 1, 2, 3
STORE_ROWS, STORE_PIXEL_SHUFFLE2 = 0, 1
GEMM_AUTO, GEMM_TCGEN05, GEMM_SIMT = 0, 1, 2

DTYPE_CODE = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}

EXPORTS = [
    "gcv_abi_version", "gcv_last_error", "gcv_device_supported", "gcv_gemm", "gcv_mlp_fused_supported", "gcv_mlp_fused", "gcv_mlp_fused_ln", "gcv_dwconv7_ln", "gcv_dwconv7_stats", "gcv_ln_patchify2",
    "gcv_stem_patchify_nchw", "gcv_stem_patchify_nhwc", "gcv_layernorm_rows", "gcv_pool_ln", "gcv_conv3x3_first",
    "gcv_im2col3x3", "gcv_maxpool2", "gcv_conv3x3_c16", "gcv_conv3x3_c32", "gcv_stem_fused", "gcv_stem_fused_u8", "gcv_conv3x3_first_u8", "gcv_ln_finalize", "gcv_convt2x2_small", "gcv_convt2x2_mma", "gcv_conv3x3_tc_supported", "gcv_conv3x3_tc", "gcv_resize2x_to_nchw", "gcv_nhwc_to_nchw_f32", "gcv_score_videos", "gcv_score_videos_pair",
    "gcv_swin_window_attention", "gcv_swin_patch_merge", "gcv_mean_tokens", "gcv_preprocess_frames",
]


class Epilogue(C.Structure):
    """Mirror of ``gcv_epilogue``."""
    _fields_ = [
        ("bias", C.c_void_p), ("act", C.c_int32), ("gamma", C.c_void_p), ("residual", C.c_void_p),
        ("ldr", C.c_int64), ("eps", C.c_void_p), ("eps_c", C.c_int32), ("eps_hw", C.c_int32),
        ("mu_out", C.c_void_p), ("store", C.c_int32), ("ps_h", C.c_int32), ("ps_w", C.c_int32),
        ("ps_co", C.c_int32), ("ldd", C.c_int64), ("out_f32", C.c_int32),
        ("ln_stats", C.c_void_p), ("ln_colsum", C.c_void_p), ("ln_chunks", C.c_int32), ("ln_eps", C.c_float),
    ]


class GcvError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library once; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GcvError(f"{LIB_PATH} not found: build it with `python -m genconvit_b200.build` "
                       "(there is no CPU / eager fallback for the GenConViT forward)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    lib.gcv_abi_version.restype = C.c_int
    lib.gcv_last_error.restype = C.c_char_p
    lib.gcv_device_supported.argtypes = [i32]
    lib.gcv_gemm.argtypes = [i32, i32, vp, i64, vp, i64, vp, i64, i64, i64, C.POINTER(Epilogue), vp]
    lib.gcv_mlp_fused_supported.argtypes = [i32, i32]
    lib.gcv_mlp_fused.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, i64, i32, vp]
    lib.gcv_mlp_fused_ln.argtypes = [i32, vp, vp, f32, vp, vp, vp, vp, vp, vp, vp, i64, i32, vp]
    lib.gcv_dwconv7_ln.argtypes = [i32, vp, vp, vp, vp, vp, vp, f32, i32, i32, i32, i32, vp]
    lib.gcv_dwconv7_stats.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.gcv_ln_patchify2.argtypes = [i32, vp, vp, vp, vp, f32, i32, i32, i32, i32, vp]
    lib.gcv_stem_patchify_nchw.argtypes = [i32, vp, vp, i32, i32, i32, vp]
    lib.gcv_stem_patchify_nhwc.argtypes = [i32, vp, vp, i32, i32, i32, vp]
    lib.gcv_layernorm_rows.argtypes = [i32, vp, vp, vp, vp, f32, i64, i32, vp]
    lib.gcv_pool_ln.argtypes = [i32, vp, vp, vp, vp, f32, i32, i32, i32, vp]
    lib.gcv_conv3x3_first.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.gcv_im2col3x3.argtypes = [i32, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.gcv_maxpool2.argtypes = [i32, vp, vp, i32, i32, i32, i32, vp]
    lib.gcv_ln_finalize.argtypes = [vp, vp, i64, i32, i32, f32, vp]
    lib.gcv_stem_fused.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, f32, i32, i32, i32, vp]
    lib.gcv_stem_fused_u8.argtypes = [i32, vp, vp, vp, vp, vp, vp, f32, i32, i32, i32, vp, vp, vp]
    lib.gcv_conv3x3_first_u8.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.gcv_conv3x3_c32.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.gcv_conv3x3_c16.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.gcv_convt2x2_small.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.gcv_conv3x3_tc_supported.argtypes = [i32, i32, i32]
    lib.gcv_conv3x3_tc.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.gcv_convt2x2_mma.argtypes = [i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.gcv_resize2x_to_nchw.argtypes = [i32, vp, vp, i32, i32, i32, i32, vp]
    lib.gcv_nhwc_to_nchw_f32.argtypes = [i32, vp, vp, i32, i32, i32, i32, vp]
    lib.gcv_score_videos.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp]
    lib.gcv_score_videos_pair.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.gcv_swin_window_attention.argtypes = [i32, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.gcv_swin_patch_merge.argtypes = [i32, vp, vp, i32, i32, i32, vp]
    lib.gcv_mean_tokens.argtypes = [i32, vp, vp, i32, i32, i32, vp]
    lib.gcv_preprocess_frames.argtypes = [vp, vp, i32, i32, i32, C.POINTER(C.c_float), C.POINTER(C.c_float), vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("gcv_last_error",):
            fn.restype = C.c_int
    if lib.gcv_abi_version() != 3:
        raise GcvError("libgenconvit_b200.so ABI version mismatch")
    _lib = lib
    return lib


def _check(rc, what):
    if rc != 0:
        msg = load().gcv_last_error().decode(errors="replace")
        raise GcvError(f"{what} failed (status {rc}): {msg}")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise GcvError(f"{what}: tensor is on {t.device}; the GenConViT forward only runs on a CUDA (sm_100a) device "
                       "-- there is no CPU fallback")
    if t.device.index != torch.cuda.current_device():
        # kernels launch on the CURRENT device's stream (and per-device kernel attributes are set for it): foreign
        # pointers there fault the GPU.  The model forwards enter torch.cuda.device(x.device) themselves.
        raise GcvError(f"{what}: tensor is on {t.device} but the current CUDA device is {torch.cuda.current_device()}; "
                       "wrap the call in torch.cuda.device(tensor.device)")


def require_cuda_tensor(t: torch.Tensor, what: str):
    """Only the is-CUDA half of ``require_cuda`` (the caller switches the current device itself)."""
    if not t.is_cuda:
        raise GcvError(f"{what}: tensor is on {t.device}; the GenConViT forward only runs on a CUDA (sm_100a) device "
                       "-- there is no CPU fallback")


def _avail(t: torch.Tensor) -> int:
    """Elements addressable from ``t.data_ptr()`` to the end of its storage (views / tails of larger buffers)."""
    return t.untyped_storage().nbytes() // t.element_size() - t.storage_offset()


def _need(t, count, what):
    if _avail(t) < count:
        raise GcvError(f"{what}: needs {count} elements from the tensor's first element, its storage holds {_avail(t)}")


# ---- launch counter (bench.py reports how many of our kernels ran in the timed region) ----
launches = 0

# ---- optional per-launch timing: set ``profile`` to a list and every wrapped launch appends
# (kernel, work, start_event, end_event); work = FLOPs for GEMMs, algorithmic bytes otherwise ----
profile = None


class _Timed:
    def __init__(self, kernel, work, tag=""):
        self.kernel, self.work, self.tag = kernel, work, tag

    def __enter__(self):
        if profile is not None:
            self.s, self.e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.s.record()

    def __exit__(self, *exc):
        if profile is not None:
            self.e.record()
            profile.append((self.kernel, self.work, self.s, self.e, self.tag))
        return False


def gemm(a, b, d, M, N, K, *, lda=None, ldb=None, ldd=None, bias=None, act=ACT_NONE, gamma=None, residual=None,
         ldr=None, eps=None, eps_c=0, eps_hw=0, mu_out=None, store=STORE_ROWS, ps=(0, 0, 0), out_f32=False,
         ln_stats=None, ln_colsum=None, ln_eps=0.0, backend=GEMM_AUTO):
    """D[M,N] = epilogue(A[M,K] @ B[N,K]^T); see gcv_gemm / gcv_epilogue."""
    global launches
    require_cuda(a, "gemm")
    lda_, ldb_ = (lda if lda is not None else K), (ldb if ldb is not None else K)
    # the C ABI takes raw pointers: reject shapes that would read / write past the tensors (a TMA map over a too-short
    # buffer is an illegal-address fault, which on a shared box takes the GPU down for everybody)
    _need(a, (M - 1) * lda_ + K, f"gemm A [{M},{K}] lda={lda_}")
    _need(b, (N - 1) * ldb_ + K, f"gemm B [{N},{K}] ldb={ldb_}")
    if store == STORE_ROWS:
        _need(d, (M - 1) * (ldd if ldd is not None else N) + N, f"gemm D [{M},{N}]")
    else:
        _need(d, M * N, f"gemm D (pixel shuffle) [{M},{N}]")
    if residual is not None:
        _need(residual, (M - 1) * (ldr if ldr is not None else N) + N, "gemm residual")
    if eps is not None:
        _need(eps, M * N, "gemm eps")
    ep = Epilogue()
    ep.bias = bias.data_ptr() if bias is not None else None
    ep.act = act
    ep.gamma = gamma.data_ptr() if gamma is not None else None
    ep.residual = residual.data_ptr() if residual is not None else None
    ep.ldr = ldr if ldr is not None else N
    ep.eps = eps.data_ptr() if eps is not None else None
    ep.eps_c, ep.eps_hw = eps_c, eps_hw
    ep.mu_out = mu_out.data_ptr() if mu_out is not None else None
    ep.store = store
    ep.ps_h, ep.ps_w, ep.ps_co = ps
    ep.ldd = ldd if ldd is not None else N
    ep.out_f32 = 1 if out_f32 else 0
    ep.ln_stats = ln_stats.data_ptr() if ln_stats is not None else None
    ep.ln_colsum = ln_colsum.data_ptr() if ln_colsum is not None else None
    # [M, chunks, 2] partial sums, or [M, 2] rows already reduced by ln_finalize (ln_chunks = 0)
    ep.ln_chunks = (ln_stats.shape[1] if ln_stats.dim() == 3 else 0) if ln_stats is not None else 0
    ep.ln_eps = ln_eps
    tc = backend == GEMM_TCGEN05 or backend >= 1000 or (
        backend == GEMM_AUTO and a.dtype != torch.float32 and K % 8 == 0 and lda_ % 8 == 0 and ldb_ % 8 == 0
        and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0)
    tag = f"M{M} N{N} K{K} act{act}" + ("+res" if residual is not None else "") + ("+eps" if eps is not None else "") + \
        ("+ps" if store else "") + ("+ln" if ln_stats is not None else "")
    with _Timed("gemm_tcgen05" if tc else "gemm_simt", 2.0 * M * N * K, tag):
        rc = load().gcv_gemm(backend, DTYPE_CODE[a.dtype], _p(a), lda_, _p(b), ldb_, _p(d), M, N, K, C.byref(ep),
                             _stream())
    _check(rc, f"gcv_gemm(M={M},N={N},K={K})")
    launches += 1


def mlp_fused_supported(dtype, c):
    return dtype in DTYPE_CODE and bool(load().gcv_mlp_fused_supported(DTYPE_CODE[dtype], c))


def mlp_fused(y, w1, b1, w2, b2, gamma, x, M, c):
    """x += gamma * (GELU(y @ w1^T + b1) @ w2^T + b2), in place; see gcv_mlp_fused."""
    _run("mlp_fused", 16.0 * M * c * c, lambda: load().gcv_mlp_fused(
        DTYPE_CODE[y.dtype], _p(y), _p(w1), _p(b1), _p(w2), _p(b2), _p(gamma), _p(x), M, c, _stream()),
        f"M{M} C{c}")


def mlp_fused_ln(y, stats, ln_eps, w1, b1, colsum1, w2, b2, gamma, x, M, c):
    """mlp_fused with the block's LayerNorm folded into fc1 (y = un-normalised dwconv output); see gcv_mlp_fused_ln."""
    _run("mlp_fused", 16.0 * M * c * c, lambda: load().gcv_mlp_fused_ln(
        DTYPE_CODE[y.dtype], _p(y), _p(stats), ln_eps, _p(w1), _p(b1), _p(colsum1), _p(w2), _p(b2), _p(gamma), _p(x),
        M, c, _stream()), f"M{M} C{c} +ln")


def dwconv7_stats(x, y, stats, taps, bias, B, H, W, Cc):
    """y = dwconv7x7(x) + bias on the tensor cores, stats = LayerNorm partial sums; see gcv_dwconv7_stats."""
    _run("dwconv7_mma", 2.0 * B * H * W * Cc * x.element_size(), lambda: load().gcv_dwconv7_stats(
        DTYPE_CODE[x.dtype], _p(x), _p(y), _p(stats), _p(taps), _p(bias), B, H, W, Cc, _stream()),
        f"B{B} H{H} W{W} C{Cc}")


def dwconv7_ln(x, y, taps, bias, ln_w, ln_b, eps, B, H, W, Cc):
    global launches
    with _Timed("dwconv7_ln", 2.0 * B * H * W * Cc * x.element_size(), f"B{B} H{H} W{W} C{Cc}"):
        rc = load().gcv_dwconv7_ln(DTYPE_CODE[x.dtype], _p(x), _p(y), _p(taps), _p(bias), _p(ln_w), _p(ln_b), eps,
                                   B, H, W, Cc, _stream())
    _check(rc, "gcv_dwconv7_ln")
    launches += 1


def _run(kernel, work, call, tag=""):
    """Launch one kernel through the C ABI: count it, optionally time it, raise on a non-zero status."""
    global launches
    with _Timed(kernel, work, tag):
        rc = call()
    _check(rc, "gcv_" + kernel)
    launches += 1


def ln_patchify2(x, a, ln_w, ln_b, eps, B, H, W, Cc):
    es = x.element_size()
    _run("ln_patchify2", 2.0 * B * H * W * Cc * es, lambda: load().gcv_ln_patchify2(
        DTYPE_CODE[x.dtype], _p(x), _p(a), _p(ln_w), _p(ln_b), eps, B, H, W, Cc, _stream()))


def ln_finalize(stats, out, M, K, eps):
    """[M, chunks, 2] LayerNorm partial sums -> [M, 2] (rstd, -mean * rstd); see gcv_ln_finalize."""
    _run("ln_finalize", M * (stats.shape[1] + 1) * 8.0, lambda: load().gcv_ln_finalize(
        _p(stats), _p(out), M, stats.shape[1], K, eps, _stream()))


def stem_fused(x, y, w, bias, ln_w, ln_b, eps, B, H, W, nchw):
    """Conv2d(3,96,k4,s4) + bias + LayerNorm2d in one pass; see gcv_stem_fused."""
    if nchw:
        assert x.dtype == torch.float32
    _run("stem_fused", B * H * W * 3.0 * x.element_size() + B * (H // 4) * (W // 4) * 96.0 * y.element_size(),
         lambda: load().gcv_stem_fused(DTYPE_CODE[y.dtype], 1 if nchw else 0, _p(x), _p(y), _p(w), _p(bias), _p(ln_w),
                                       _p(ln_b), eps, B, H, W, _stream()), f"B{B} H{H} {'nchw' if nchw else 'nhwc'}")


def stem_patchify_nchw(x, a, B, H, W):
    assert x.dtype == torch.float32
    _run("stem_patchify_nchw", B * H * W * 3.0 * (4 + a.element_size()), lambda: load().gcv_stem_patchify_nchw(
        DTYPE_CODE[a.dtype], _p(x), _p(a), B, H, W, _stream()))


def stem_patchify_nhwc(x, a, B, H, W):
    _run("stem_patchify_nhwc", B * H * W * 3.0 * 2 * a.element_size(), lambda: load().gcv_stem_patchify_nhwc(
        DTYPE_CODE[a.dtype], _p(x), _p(a), B, H, W, _stream()))


def layernorm_rows(x, y, w, b, eps, rows, Cc):
    _run("layernorm_rows", 2.0 * rows * Cc * x.element_size(), lambda: load().gcv_layernorm_rows(
        DTYPE_CODE[x.dtype], _p(x), _p(y), _p(w), _p(b), eps, rows, Cc, _stream()))


def pool_ln(x, y, w, b, eps, B, HW, Cc):
    _run("pool_ln", 1.0 * B * (HW + 1) * Cc * x.element_size(), lambda: load().gcv_pool_ln(
        DTYPE_CODE[x.dtype], _p(x), _p(y), _p(w), _p(b), eps, B, HW, Cc, _stream()))


def conv3x3_first(x, y, w, b, stride, act, pool, B, H, W):
    assert x.dtype == torch.float32
    _run("conv3x3_first", B * H * W * 3.0 * 4 + y.numel() * y.element_size(), lambda: load().gcv_conv3x3_first(
        DTYPE_CODE[y.dtype], _p(x), _p(y), _p(w), _p(b), stride, act, 1 if pool else 0, B, H, W, _stream()))


def conv3x3_first_u8(x, y, w, b, stride, act, pool, B, H, W, mean, std):
    """The first encoder conv straight from raw uint8 NHWC crops (normalised in the kernel); see gcv_conv3x3_first_u8."""
    assert x.dtype == torch.uint8
    _need(x, B * H * W * 3, "conv3x3_first_u8 x")
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    _run("conv3x3_first", B * H * W * 3.0 + y.numel() * y.element_size(), lambda: load().gcv_conv3x3_first_u8(
        DTYPE_CODE[y.dtype], _p(x), _p(y), _p(w), _p(b), stride, act, 1 if pool else 0, B, H, W, m3, s3, _stream()), "u8")


def stem_fused_u8(x, y, w, bias, ln_w, ln_b, eps, B, H, W, mean, std):
    """The ConvNeXt stem straight from raw uint8 NHWC crops (normalised in the kernel); see gcv_stem_fused_u8."""
    assert x.dtype == torch.uint8
    _need(x, B * H * W * 3, "stem_fused_u8 x")
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    _run("stem_fused", B * H * W * 3.0 + B * (H // 4) * (W // 4) * 96.0 * y.element_size(),
         lambda: load().gcv_stem_fused_u8(DTYPE_CODE[y.dtype], _p(x), _p(y), _p(w), _p(bias), _p(ln_w), _p(ln_b), eps, B, H, W,
                                          m3, s3, _stream()), f"B{B} H{H} u8")


def im2col3x3(x, a, B, H, W, Cc, stride):
    es = x.element_size()
    _run("im2col3x3", (B * H * W * Cc + 9.0 * B * (H // stride) * (W // stride) * Cc) * es, lambda: load().gcv_im2col3x3(
        DTYPE_CODE[x.dtype], _p(x), _p(a), B, H, W, Cc, stride, _stream()))


def conv3x3_c16(x, y, w, bias, stride, act, pool, B, H, W):
    """Direct 16 -> 32 channel 3x3 conv (+ act, + 2x2 max-pool) on the tensor cores; see gcv_conv3x3_c16."""
    es = x.element_size()
    _run("conv3x3_c16", (B * H * W * 16.0 + y.numel()) * es, lambda: load().gcv_conv3x3_c16(
        DTYPE_CODE[x.dtype], _p(x), _p(y), _p(w), _p(bias), stride, act, 1 if pool else 0, B, H, W, _stream()),
        f"B{B} H{H} W{W} s{stride}")


def conv3x3_c32(x, y, w, bias, stride, act, pool, B, H, W):
    """Direct 32 -> 64 channel 3x3 conv (+ act, + 2x2 max-pool) on the tensor cores; see gcv_conv3x3_c32."""
    es = x.element_size()
    _run("conv3x3_c32", (B * H * W * 32.0 + y.numel()) * es, lambda: load().gcv_conv3x3_c32(
        DTYPE_CODE[x.dtype], _p(x), _p(y), _p(w), _p(bias), stride, act, 1 if pool else 0, B, H, W, _stream()),
        f"B{B} H{H} W{W} s{stride}")


def maxpool2(x, y, B, H, W, Cc):
    _run("maxpool2", 1.25 * B * H * W * Cc * x.element_size(), lambda: load().gcv_maxpool2(
        DTYPE_CODE[x.dtype], _p(x), _p(y), B, H, W, Cc, _stream()))


def conv3x3_tc_supported(dt, c, n):
    return dt in DTYPE_CODE and bool(load().gcv_conv3x3_tc_supported(DTYPE_CODE[dt], c, n))


def conv3x3_tc(x, y, w, bias, stride, act, pool, B, H, W, ci, co):
    """3x3 conv (+ act, + 2x2 max-pool) of the encoders' 64 -> 128 / 128 -> 256 layers as a tcgen05 implicit GEMM; see
    gcv_conv3x3_tc."""
    ho, wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    if pool:
        ho, wo = ho // 2, wo // 2
    _need(x, B * H * W * ci, "conv3x3_tc x")
    _need(y, B * ho * wo * co, "conv3x3_tc y")
    _need(w, 9 * ci * co, "conv3x3_tc w")
    _run("conv3x3_tc", 2.0 * B * ((H - 1) // stride + 1) * ((W - 1) // stride + 1) * co * 9 * ci,
         lambda: load().gcv_conv3x3_tc(DTYPE_CODE[x.dtype], _p(x), _p(y), _p(w), _p(bias), stride, act, int(bool(pool)),
                                       B, H, W, ci, co, _stream()), f"B{B} H{H} C{ci} s{stride}")


def convt2x2_small(x, y, w, bias, act, B, H, W, ci, co):
    _run("convt2x2_small", B * H * W * (ci + 4.0 * co) * x.element_size(), lambda: load().gcv_convt2x2_small(
        DTYPE_CODE[x.dtype], _p(x), _p(y), _p(w), _p(bias), act, B, H, W, ci, co, _stream()))


def convt2x2_mma(x, y, w1, b1, act, B, H, W, ci, w2=None, b2=None):
    """ConvTranspose2d(ci -> ci/2, k2 s2) + act on HMMA; with (w2, b2) also the fused 16 -> 3 output layer.  See
    gcv_convt2x2_mma."""
    tail = w2 is not None
    _need(x, B * H * W * ci, "convt2x2_mma x")
    _need(y, B * H * W * (48 if tail else 2 * ci), "convt2x2_mma y")
    _need(w1, 2 * ci * ci, "convt2x2_mma w1")
    _run("convt2x2_mma", B * H * W * (ci + (48.0 if tail else 2.0 * ci)) * x.element_size(),
         lambda: load().gcv_convt2x2_mma(DTYPE_CODE[x.dtype], _p(x), _p(y), _p(w1), _p(b1), _p(w2) if tail else None,
                                         _p(b2) if tail else None, act, B, H, W, ci, int(tail), _stream()),
         f"B{B} H{H} C{ci}" + ("+tail" if tail else ""))


def resize2x_to_nchw(x, y, B, H, W, Cc):
    _run("resize2x_to_nchw", B * H * W * Cc * (x.element_size() + 16.0), lambda: load().gcv_resize2x_to_nchw(
        DTYPE_CODE[x.dtype], _p(x), _p(y), B, H, W, Cc, _stream()))


def nhwc_to_nchw_f32(x, y, B, H, W, Cc):
    _run("nhwc_to_nchw_f32", B * H * W * Cc * (x.element_size() + 4.0), lambda: load().gcv_nhwc_to_nchw_f32(
        DTYPE_CODE[x.dtype], _p(x), _p(y), B, H, W, Cc, _stream()))


def swin_window_attention(qkv, out, bias_table, B, res, c, heads, shift):
    _run("swin_window_attention", 4.0 * B * res * res * 49 * c, lambda: load().gcv_swin_window_attention(
        DTYPE_CODE[qkv.dtype], _p(qkv), _p(out), _p(bias_table), B, res, c, heads, shift, _stream()),
        f"B{B} res{res} C{c} shift{shift}")


def swin_patch_merge(x, out, B, res, c):
    _run("swin_patch_merge", 2.0 * B * res * res * c * x.element_size(), lambda: load().gcv_swin_patch_merge(
        DTYPE_CODE[x.dtype], _p(x), _p(out), B, res, c, _stream()))


def mean_tokens(x, y, B, Lt, c):
    _run("mean_tokens", 1.0 * B * Lt * c * x.element_size(), lambda: load().gcv_mean_tokens(
        DTYPE_CODE[x.dtype], _p(x), _p(y), B, Lt, c, _stream()))


def preprocess_frames(x_u8, y, N, H, W, mean, std):
    """uint8 NHWC device frames -> ImageNet-normalised fp32 NCHW; see gcv_preprocess_frames."""
    require_cuda(x_u8, "preprocess_frames")
    assert x_u8.dtype == torch.uint8 and y.dtype == torch.float32
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    _run("preprocess_frames", 15.0 * N * H * W, lambda: load().gcv_preprocess_frames(
        _p(x_u8), _p(y), N, H, W, m3, s3, _stream()))


def score_videos(logits, n_nets, n_frames, fpv, mean_out, cls_out, val_out):
    require_cuda(logits, "score_videos")
    _run("score_videos", 8.0 * n_nets * n_frames, lambda: load().gcv_score_videos(
        _p(logits), n_nets, n_frames, fpv, _p(mean_out), _p(cls_out), _p(val_out), _stream()))


def score_videos_pair(logits_ed, logits_vae, n_frames, fpv, out):
    """pred_vid scoring straight from the ED / VAE logit buffers (either may be None) into fp32 out[2, V]."""
    ref = logits_ed if logits_ed is not None else logits_vae
    require_cuda(ref, "score_videos_pair")
    for t in (logits_ed, logits_vae):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous()
            _need(t, 2 * n_frames, "score_videos_pair logits")
    _need(out, 2 * (n_frames // fpv), "score_videos_pair out")
    _run("score_videos", 8.0 * n_frames * ((logits_ed is not None) + (logits_vae is not None)),
         lambda: load().gcv_score_videos_pair(_p(logits_ed), _p(logits_vae), n_frames, fpv, _p(out), _stream()))
