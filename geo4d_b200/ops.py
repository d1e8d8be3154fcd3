"""Thin Python wrappers over the C-ABI kernels (device tensors in, device tensors out).

Activation layout everywhere: bf16, frames-major channels-last, i.e. a feature
map of N frames is a 2-D matrix [N*H*W, C] (row = ((n*H)+y)*W+x).  All wrappers
enqueue on torch's current stream and never synchronise.
"""
from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache
from typing import Optional, Sequence, Tuple

import torch

from . import _cabi
from ._cabi import GemmDesc, check, lib

ACT_NONE, ACT_SILU, ACT_GEGLU, ACT_GELU = 0, 1, 2, 3

TAPS_1 = ((0, 0),)
TAPS_3x3 = tuple((dx, dy) for dy in (-1, 0, 1) for dx in (-1, 0, 1))  # weight[:, :, ky, kx] order
TAPS_T3 = ((0, -1), (0, 0), (0, 1))  # temporal taps act on the H (= frame) axis


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class capture_graph:
    """`with capture_graph(g): ...` records the launches of the block into the torch.cuda.CUDAGraph `g` on the
    CURRENT stream (which must be a side stream).  Unlike `torch.cuda.graph` it does not run gc.collect() and
    torch.cuda.empty_cache() first: those cost 50-1000 ms per capture once a 25 GB pipeline is resident and hand
    every cached block back to the driver, and the alignment captures two small graphs per call."""

    def __init__(self, graph: "torch.cuda.CUDAGraph"):
        self.graph = graph

    def __enter__(self):
        # thread_local: calls made by OTHER threads (NCCL's watchdog, the clock sampler) cannot invalidate the capture
        self.graph.capture_begin(capture_error_mode="thread_local")
        return self.graph

    def __exit__(self, exc_type, exc, tb):
        self.graph.capture_end()
        return False


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


@lru_cache(maxsize=None)
def pick_box(W: int, H: int, N: int) -> Tuple[int, int, int]:
    """Choose the (box_w, box_h, box_n) M-tile (<=128 rows) wasting the fewest MMA rows."""
    best, best_key = None, None
    bws = sorted({d for d in range(1, min(W, 128) + 1) if W % d == 0} | ({128} if W > 128 else set()))
    for bw in bws:
        for bh in range(1, min(H, 128 // bw) + 1):
            for bn in range(1, min(N, 128 // (bw * bh)) + 1):
                tiles = -(-W // bw) * -(-H // bh) * -(-N // bn)
                eff = (W * H * N) / (tiles * 128.0)
                key = (round(eff, 6), bw * bh * bn, bw, bh)
                if best_key is None or key > best_key:
                    best, best_key = (bw, bh, bn), key
    return best


def tap_gemm(a: torch.Tensor, K: int, W: int, H: int, N: int, strides: Tuple[int, int, int],
             taps: Sequence[Tuple[int, int]], b: torch.Tensor, n_out: int, out: torch.Tensor, ldc: int,
             *, bias: Optional[torch.Tensor] = None, row_bias: Optional[torch.Tensor] = None,
             rows_per_bias: int = 0, act: int = ACT_NONE, residual: Optional[torch.Tensor] = None,
             ldr: int = 0, alpha: float = 1.0, b_batched: bool = False,
             box: Optional[Tuple[int, int, int]] = None) -> torch.Tensor:
    d = GemmDesc()
    d.a = a.data_ptr()
    d.K, d.W, d.H, d.N = K, W, H, N
    d.a_stride_w, d.a_stride_h, d.a_stride_n = strides
    bw, bh, bn = box if box is not None else pick_box(W, H, 1 if b_batched else N)
    d.box_w, d.box_h, d.box_n = bw, bh, bn
    d.num_taps = len(taps)
    for i, (dx, dy) in enumerate(taps):
        d.tap_dx[i] = dx
        d.tap_dy[i] = dy
    d.b = b.data_ptr()
    d.n_out = n_out
    d.b_batched = 1 if b_batched else 0
    d.out = out.data_ptr()
    d.ldc = ldc
    d.out_fp32 = 1 if out.dtype == torch.float32 else 0
    d.alpha = alpha
    d.bias = _ptr(bias)
    d.row_bias = _ptr(row_bias)
    d.row_bias_ld = row_bias.stride(0) if row_bias is not None else 0
    d.rows_per_bias = rows_per_bias
    d.act = act
    d.residual = _ptr(residual)
    d.ldr = ldr
    ws = _splitk_workspace(a.device)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel() * 4
    if _FORCE_TILE is not None:
        d.tile_n, d.cta_pair = _FORCE_TILE
        d.split_k = 1 if _FORCE_SPLIT is None else _FORCE_SPLIT   # bitwise-equality tests keep the k-step order fixed
    elif _AUTOTUNE:
        key = (K, W, H, N, bw, bh, bn, len(taps), n_out, int(b_batched), d.out_fp32, act, bias is not None,
               row_bias is not None, residual is not None)
        cfg = _TUNED.get(key)
        if cfg is None and not torch.cuda.is_current_stream_capturing():
            cfg = _autotune(d, key, out, residual)
        if cfg is not None:
            d.tile_n, d.cta_pair, d.split_k = cfg
    rc = lib().geo4d_tap_gemm(C.byref(d), C.c_void_p(_stream()))
    if rc != 0 and _FORCE_TILE is not None:    # a forced configuration that is not legal for this shape: library's choice
        d.tile_n, d.cta_pair = 0, 0
        rc = lib().geo4d_tap_gemm(C.byref(d), C.c_void_p(_stream()))
    check(rc, "geo4d_tap_gemm")
    return out


# ---------------------------------------------------------------------------------------------- autotuning
# The tile shape (tile_n, single CTA vs CTA pair) does not change the result of a tap-GEMM, only its speed, and
# the best choice depends on the interplay of L2 bandwidth, epilogue cost and wave quantisation.  So the first
# time a shape is seen (outside CUDA-graph capture) every legal configuration is timed on scratch outputs with a
# small CUDA graph (GPU-bound, no launch gaps) and the winner is pinned for the rest of the process.
# GEO4D_AUTOTUNE=0 leaves the choice to the library's cost model.
_AUTOTUNE = os.environ.get("GEO4D_AUTOTUNE", "1") != "0"
_SPLITK = os.environ.get("GEO4D_SPLITK", "1") != "0"
_splitk_ws: dict = {}


def _splitk_workspace(device) -> torch.Tensor:
    """Caller-owned scratch for the partial planes of split-K tap-GEMMs (only small-M layers split: 64 MiB holds
    16 planes of a 640 x 1280 fp32 tile set with room to spare).  One buffer per device; launches on one stream
    are ordered, so consecutive GEMMs can share it."""
    ws = _splitk_ws.get(device)
    if ws is None:
        ws = torch.empty(16 << 20, device=device, dtype=torch.float32)
        _splitk_ws[device] = ws
    return ws
_TUNED: dict = {}
_TUNE_LOG: list = []
_TUNE_ERRORS: list = []   # configurations the library refused during tuning (visible, not swallowed)
_FORCE_TILE: Optional[Tuple[int, int]] = None   # (tile_n, cta_pair) for every launch; tests and experiments only
_FORCE_SPLIT: Optional[int] = None              # with _FORCE_TILE: number of K parts (tests)


def tuned_configs():
    """[(key, (tile_n, cta_pair), {candidate: microseconds})] for every shape tuned so far."""
    return list(_TUNE_LOG)


def _autotune(d: GemmDesc, key, out: torch.Tensor, residual: Optional[torch.Tensor]):
    dev = out.device
    scratch = torch.empty_strided(out.shape, out.stride(), device=dev, dtype=out.dtype)
    res = None
    if residual is not None:
        res = torch.zeros_like(residual)
    t = GemmDesc()
    C.memmove(C.byref(t), C.byref(d), C.sizeof(GemmDesc))
    t.out = scratch.data_ptr()
    if res is not None:
        t.residual = res.data_ptr()
        t.ldr = res.stride(0)
    L = lib()
    cur = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    results = {}
    n_rep = 6
    cands = [(tn, pair, 1) for pair in (1, 2) for tn in (256, 160, 128, 64, 32)]
    if _SPLITK and not d.b_batched and d.act != ACT_GEGLU:
        # split-K candidates only where the library would actually split: few output tiles, long reduction
        m_tiles = -(-d.W // d.box_w) * -(-d.H // d.box_h) * -(-d.N // d.box_n)
        k_iters = d.num_taps * (d.K // 64)
        for tn in (256, 160, 128, 64):
            if m_tiles * -(-d.n_out // tn) * 2 <= 148 and k_iters >= 16:
                cands.append((tn, 1, 0))   # 0: the library picks the number of K parts
    for tn, pair, sk in cands:
        if True:
            t.tile_n, t.cta_pair, t.split_k = tn, pair, sk
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                if L.geo4d_tap_gemm(C.byref(t), C.c_void_p(side.cuda_stream)) != 0:
                    continue   # configuration not legal for this shape
                try:
                    g = torch.cuda.CUDAGraph()
                    with capture_graph(g):
                        for _ in range(n_rep):
                            check(L.geo4d_tap_gemm(C.byref(t), C.c_void_p(side.cuda_stream)), "autotune")
                    g.replay()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    best = None
                    for _ in range(3):
                        e0.record(side)
                        g.replay()
                        e1.record(side)
                        e1.synchronize()
                        us = e0.elapsed_time(e1) * 1e3 / n_rep
                        best = us if best is None else min(best, us)
                    results[(tn, pair, sk)] = best
                    del g
                except _cabi.Geo4DError as ex:   # the library refused this configuration for this shape
                    _TUNE_ERRORS.append((key, (tn, pair, sk), str(ex)))
                    continue
    cur.wait_stream(side)
    if not results:
        _TUNED[key] = (0, 0, 1)
        return (0, 0, 1)
    cfg = min(results, key=results.get)
    _TUNED[key] = cfg
    _TUNE_LOG.append((key, cfg, results))
    if os.environ.get("GEO4D_AUTOTUNE_VERBOSE") == "1":
        print(f"[geo4d autotune] {key} -> tile_n={cfg[0]} {'pair' if cfg[1] == 2 else 'single'}{' split-K' if cfg[2] == 0 else ''} "
              f"{results[cfg]:.1f} us  (all: {dict((k, round(v, 1)) for k, v in sorted(results.items()))})", flush=True)
    return cfg


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
           out: Optional[torch.Tensor] = None, act: int = ACT_NONE,
           residual: Optional[torch.Tensor] = None, alpha: float = 1.0,
           out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """out[M, n] = epilogue(x[M, K] @ w[n, K]^T).  x may be a row-strided view."""
    assert x.dim() == 2 and w.dim() == 2 and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    M, K = x.shape
    n = w.shape[0]
    n_store = n // 2 if act == ACT_GEGLU else n
    if out is None:
        out = torch.empty((M, n_store), device=x.device, dtype=out_dtype)
    ldr = residual.stride(0) if residual is not None else 0
    lda = x.stride(0)
    return tap_gemm(x, K, M, 1, 1, (lda, lda * M, lda * M), TAPS_1, w, n, out, out.stride(0),
                    bias=bias, act=act, residual=residual, ldr=ldr, alpha=alpha)


def conv3x3(x: torch.Tensor, N: int, H: int, W: int, w9: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
            out: Optional[torch.Tensor] = None, row_bias: Optional[torch.Tensor] = None,
            rows_per_bias: int = 0, residual: Optional[torch.Tensor] = None,
            out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """3x3, stride 1, zero pad 1.  x [N*H*W, Cin] bf16; w9 [9, Cout, Cin] bf16 (tap = ky*3+kx)."""
    Cin = x.shape[1]
    Cout = w9.shape[1]
    if out is None:
        out = torch.empty((N * H * W, Cout), device=x.device, dtype=out_dtype)
    ld = x.stride(0)
    ldr = residual.stride(0) if residual is not None else 0
    return tap_gemm(x, Cin, W, H, N, (ld, ld * W, ld * W * H), TAPS_3x3, w9, Cout, out, out.stride(0),
                    bias=bias, row_bias=row_bias, rows_per_bias=rows_per_bias, residual=residual, ldr=ldr)


def temporal_conv3(x: torch.Tensor, B: int, T: int, HW: int, w3: torch.Tensor,
                   bias: Optional[torch.Tensor] = None, *, out: Optional[torch.Tensor] = None,
                   residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Conv3d kernel (3,1,1), pad (1,0,0).  x [(B*T*HW), C] bf16; w3 [3, Cout, Cin]."""
    Cin = x.shape[1]
    Cout = w3.shape[1]
    if out is None:
        out = torch.empty((B * T * HW, Cout), device=x.device, dtype=torch.bfloat16)
    ld = x.stride(0)
    ldr = residual.stride(0) if residual is not None else 0
    return tap_gemm(x, Cin, HW, T, B, (ld, ld * HW, ld * HW * T), TAPS_T3, w3, Cout, out, out.stride(0),
                    bias=bias, residual=residual, ldr=ldr)


def bmm_nt(a: torch.Tensor, b: torch.Tensor, *, out: Optional[torch.Tensor] = None, alpha: float = 1.0,
           out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """out[i] = alpha * a[i] @ b[i]^T with a [B, M, K], b [B, n, K] (both K-contiguous)."""
    Bt, M, K = a.shape
    n = b.shape[1]
    if out is None:
        out = torch.empty((Bt, M, n), device=a.device, dtype=out_dtype)
    return tap_gemm(a, K, M, 1, Bt, (a.stride(1), a.stride(1) * M, a.stride(0)), TAPS_1, b, n, out, out.stride(1),
                    alpha=alpha, b_batched=True, box=(min(M, 128), 1, 1))


# --------------------------------------------------------------------------- other kernels
def _vp(t):
    return C.c_void_p(None if t is None else t.data_ptr())


def _s():
    return C.c_void_p(_stream())


_gn_ws = {}


def _gn_workspace(device, num_stats: int, rows_per_stat: int, Cc: int) -> torch.Tensor:
    f = lib().geo4d_groupnorm_workspace_bytes
    f.restype = C.c_size_t
    need = int(f(num_stats, rows_per_stat, Cc)) // 4
    ws = _gn_ws.get(device)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 1 << 18), device=device, dtype=torch.float32)  # tickets start at zero
        _gn_ws[device] = ws
    return ws


def groupnorm(x: torch.Tensor, num_stats: int, rows_per_stat: int, gamma: torch.Tensor, beta: torch.Tensor,
              eps: float, silu: bool, out: Optional[torch.Tensor] = None,
              workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GroupNorm(32 groups)[+SiLU] on rows [num_stats*rows_per_stat, C] bf16; fp32 affine."""
    Cc = x.shape[1]
    if out is None:
        out = torch.empty((x.shape[0], Cc), device=x.device, dtype=torch.bfloat16)
    ws = workspace if workspace is not None else _gn_workspace(x.device, num_stats, rows_per_stat, Cc)
    check(lib().geo4d_groupnorm_silu(_vp(x), C.c_int64(x.stride(0)), _vp(out), C.c_int64(out.stride(0)),
                                     num_stats, rows_per_stat, Cc, _vp(gamma), _vp(beta), C.c_float(eps),
                                     1 if silu else 0, _vp(ws), C.c_size_t(ws.numel() * 4), _s()),
          "geo4d_groupnorm_silu")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    M, Cc = x.shape
    if out is None:
        out = torch.empty((M, Cc), device=x.device, dtype=torch.bfloat16)
    check(lib().geo4d_layernorm(_vp(x), C.c_int64(x.stride(0)), _vp(out), C.c_int64(out.stride(0)), M, Cc,
                                _vp(gamma), _vp(beta), C.c_float(eps), _s()), "geo4d_layernorm")
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, B: int, H: int, Lq: int,
              Lk: int, *, kv_batch_div: int = 1, accumulate: bool = False, scale: float = 0.125) -> torch.Tensor:
    """q [B*Lq, >=H*64] (row-strided view ok), k/v [ceil(B/kv_batch_div)*Lk, >=H*64] with equal strides; out [B*Lq, >=H*64]."""
    assert k.stride(0) == v.stride(0)
    check(lib().geo4d_attention(_vp(q), C.c_int64(q.stride(0)), _vp(k), _vp(v), C.c_int64(k.stride(0)), _vp(out),
                                C.c_int64(out.stride(0)), B, H, Lq, Lk, kv_batch_div,
                                1 if accumulate else 0, C.c_float(scale), _s()), "geo4d_attention")
    return out


def cross_attention2(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, Lk: int, div: int, k2: torch.Tensor,
                     v2: torch.Tensor, Lk2: int, div2: int, out: torch.Tensor, B: int, H: int, Lq: int,
                     scale: float = 0.125) -> torch.Tensor:
    """text + image cross-attention in one launch: out = softmax(q k^T) v + softmax(q k2^T) v2 (two softmaxes)."""
    assert k.stride(0) == v.stride(0) and k2.stride(0) == v2.stride(0)
    check(lib().geo4d_cross_attention2(_vp(q), C.c_int64(q.stride(0)), _vp(k), _vp(v), C.c_int64(k.stride(0)), Lk, div,
                                       _vp(k2), _vp(v2), C.c_int64(k2.stride(0)), Lk2, div2, _vp(out),
                                       C.c_int64(out.stride(0)), B, H, Lq, C.c_float(scale), _s()),
          "geo4d_cross_attention2")
    return out


def temporal_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, B: int, T: int,
                       HW: int, heads: int, scale: float = 0.125) -> torch.Tensor:
    assert q.stride(0) == k.stride(0) == v.stride(0)
    check(lib().geo4d_temporal_attention(_vp(q), _vp(k), _vp(v), C.c_int64(q.stride(0)), _vp(out),
                                         C.c_int64(out.stride(0)), B, T, HW, heads, C.c_float(scale), _s()),
          "geo4d_temporal_attention")
    return out


def bcthw_to_rows(src0: torch.Tensor, src1: Optional[torch.Tensor], Cpad: int,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, C0, T, H, W = src0.shape
    C1 = 0 if src1 is None else src1.shape[1]
    assert src0.is_contiguous() and src0.dtype == torch.float32
    if src1 is not None:
        assert src1.is_contiguous() and src1.dtype == torch.float32
    if out is None:
        out = torch.empty((B * T * H * W, Cpad), device=src0.device, dtype=torch.bfloat16)
    check(lib().geo4d_bcthw_to_rows(_vp(src0), C0, _vp(src1), C1, B, T, H, W, _vp(out), Cpad, _s()),
          "geo4d_bcthw_to_rows")
    return out


def rows_to_bcthw(rows: torch.Tensor, Cc: int, B: int, T: int, H: int, W: int,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    assert rows.dtype == torch.float32
    if out is None:
        out = torch.empty((B, Cc, T, H, W), device=rows.device, dtype=torch.float32)
    check(lib().geo4d_rows_to_bcthw(_vp(rows), C.c_int64(rows.stride(0)), Cc, B, T, H, W, _vp(out), _s()),
          "geo4d_rows_to_bcthw")
    return out


def concat_rows(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    rows = a.shape[0]
    if out is None:
        out = torch.empty((rows, a.shape[1] + b.shape[1]), device=a.device, dtype=torch.bfloat16)
    check(lib().geo4d_concat_rows(_vp(a), C.c_int64(a.stride(0)), a.shape[1], _vp(b), C.c_int64(b.stride(0)),
                                  b.shape[1], _vp(out), C.c_int64(rows), _s()), "geo4d_concat_rows")
    return out


def upsample2x(x: torch.Tensor, N: int, H: int, W: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    Cc = x.shape[1]
    assert x.is_contiguous()
    if out is None:
        out = torch.empty((N * 4 * H * W, Cc), device=x.device, dtype=torch.bfloat16)
    check(lib().geo4d_upsample_nearest2x(_vp(x), _vp(out), N, H, W, Cc, _s()), "geo4d_upsample_nearest2x")
    return out


def im2col_s2(x: torch.Tensor, N: int, H: int, W: int, pad_before: int, Ho: int, Wo: int,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    Cc = x.shape[1]
    assert x.is_contiguous()
    if out is None:
        out = torch.empty((N * Ho * Wo, 9 * Cc), device=x.device, dtype=torch.bfloat16)
    check(lib().geo4d_im2col_3x3_s2(_vp(x), _vp(out), N, H, W, Cc, pad_before, Ho, Wo, _s()),
          "geo4d_im2col_3x3_s2")
    return out


def ddim_step(x: torch.Tensor, v: torch.Tensor, coef: torch.Tensor, step_idx: Optional[torch.Tensor],
              pred_x0: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None) -> None:
    assert x.is_contiguous() and v.is_contiguous() and x.dtype == torch.float32 and v.dtype == torch.float32
    check(lib().geo4d_ddim_step(_vp(x), _vp(v), _vp(pred_x0), _vp(noise), _vp(coef), _vp(step_idx),
                                C.c_int64(x.numel()), _s()), "geo4d_ddim_step")


def advance_counter(counter: torch.Tensor, delta: int = 1, modulo: int = 0) -> None:
    check(lib().geo4d_advance_counter(_vp(counter), delta, modulo, _s()), "geo4d_advance_counter")


def gather_row(table: torch.Tensor, idx: torch.Tensor, out: torch.Tensor) -> None:
    check(lib().geo4d_gather_row(_vp(table), C.c_int64(table.stride(0)), _vp(idx), _vp(out), out.numel(), _s()),
          "geo4d_gather_row")


def softmax_rows(scores: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Row softmax of fp32 [rows, cols] -> bf16 probabilities."""
    rows, cols = scores.shape
    if out is None:
        out = torch.empty((rows, cols), device=scores.device, dtype=torch.bfloat16)
    check(lib().geo4d_softmax_rows(_vp(scores), C.c_int64(scores.stride(0)), _vp(out), C.c_int64(out.stride(0)),
                                   C.c_int64(rows), cols, _s()), "geo4d_softmax_rows")
    return out


def transpose_bf16(x: torch.Tensor, batch: int, R: int, Cc: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x rows [batch*R, >=Cc] (row stride ld) -> out [batch, Cc, R] contiguous."""
    if out is None:
        out = torch.empty((batch, Cc, R), device=x.device, dtype=torch.bfloat16)
    check(lib().geo4d_transpose_bf16(_vp(x), C.c_int64(x.stride(0)), _vp(out), batch, R, Cc, _s()),
          "geo4d_transpose_bf16")
    return out


# --------------------------------------------------------------------------- geometry kernels
def postprocess_window(maps: torch.Tensor, T: int, H: int, W: int, valid: Optional[torch.Tensor] = None,
                       sky_eps: float = 0.1, far_value: float = 1.99, has_conf: bool = True):
    """maps [11, T, H, W] fp32 contiguous -> (pts [T,H,W,3], inv_conf [T,H,W,1], invdepth [T,H,W,1])."""
    assert maps.is_contiguous() and maps.dtype == torch.float32 and maps.shape[0] == 11
    thw = T * H * W
    pts = torch.empty((T, H, W, 3), device=maps.device, dtype=torch.float32)
    conf = torch.empty((T, H, W, 1), device=maps.device, dtype=torch.float32)
    invd = torch.empty((T, H, W, 1), device=maps.device, dtype=torch.float32)
    check(lib().geo4d_postprocess_window(_vp(maps), C.c_int64(thw), _vp(pts), _vp(conf), _vp(invd), _vp(valid),
                                         C.c_float(1.05), C.c_float(sky_eps), C.c_float(far_value), C.c_float(2.0),
                                         C.c_float(2.0), 1 if has_conf else 0, _s()), "geo4d_postprocess_window")
    return pts, conf, invd


def raymap_moments(raydir: torch.Tensor, raymoment: torch.Tensor, T: int, H: int, W: int) -> torch.Tensor:
    """raydir / raymoment [3, T, H, W] fp32 contiguous -> [T, 18] fp64 moments."""
    assert raydir.is_contiguous() and raymoment.is_contiguous()
    out = torch.empty((T, 18), device=raydir.device, dtype=torch.float64)
    check(lib().geo4d_raymap_moments(_vp(raydir), _vp(raymoment), T, H, W, _vp(out), _s()), "geo4d_raymap_moments")
    return out


def transform_points(x: torch.Tensor, mats: torch.Tensor, depth_only: bool = False) -> torch.Tensor:
    """x [n_sets, P, 3] fp32 contiguous, mats [n_sets, 12] (row-major 3x4) -> A x + t ([n_sets, P, 3]) or its z row."""
    assert x.is_contiguous() and x.dtype == torch.float32 and x.dim() == 3 and x.shape[2] == 3
    n_sets, P = x.shape[0], x.shape[1]
    m = mats.reshape(n_sets, 12).to(device=x.device, dtype=torch.float32).contiguous()
    out = torch.empty((n_sets, P) if depth_only else (n_sets, P, 3), device=x.device, dtype=torch.float32)
    check(lib().geo4d_transform_points(_vp(x), n_sets, C.c_int64(P), _vp(m), _vp(out), 1 if depth_only else 0, _s()),
          "geo4d_transform_points")
    return out


def umeyama_moments(x: torch.Tensor, y: torch.Tensor, w1: torch.Tensor, w2: Optional[torch.Tensor], npts: int,
                    pass_: int, means: Optional[torch.Tensor]) -> torch.Tensor:
    out = torch.empty(10, device=x.device, dtype=torch.float64)
    check(lib().geo4d_umeyama_moments(_vp(x), _vp(y), _vp(w1), _vp(w2), C.c_int64(npts), pass_, _vp(means), _vp(out),
                                      _s()), "geo4d_umeyama_moments")
    return out


def lad_step(x, y, n_per_group: int, G: int, state, acc, lr: float, tol: float = 1e-6):
    check(lib().geo4d_lad_step(_vp(x), _vp(y), C.c_int64(n_per_group), G, _vp(state), _vp(acc), C.c_float(lr),
                               C.c_float(tol), _s()), "geo4d_lad_step")


def lad_fit_workspace_size(G: int) -> int:
    f = lib().geo4d_lad_fit_workspace_doubles
    f.restype = C.c_size_t
    return int(f(G))


def lad_fit_workspace(G: int, device) -> torch.Tensor:
    f = lib().geo4d_lad_fit_workspace_doubles
    f.restype = C.c_size_t
    return torch.zeros(int(f(G)), device=device, dtype=torch.float64)


def lad_fit(x, y, n_per_group: int, G: int, state, acc, lr: float, iters: int, tol: float = 1e-6):
    """acc: lad_fit_workspace(G, device) (a smaller buffer is replaced: older callers passed 4*G doubles)"""
    need = lad_fit_workspace_size(G)
    if acc is None or acc.numel() < need:
        acc = torch.zeros(need, device=x.device, dtype=torch.float64)
    check(lib().geo4d_lad_fit(_vp(x), _vp(y), C.c_int64(n_per_group), G, _vp(state), _vp(acc), C.c_float(lr),
                              C.c_float(tol), int(iters), _s()), "geo4d_lad_fit")


def delta125(x, y, w, n_per_group: int, G: int, st, st_stride: int) -> torch.Tensor:
    out = torch.empty((G, 2), device=x.device, dtype=torch.float64)
    check(lib().geo4d_delta125(_vp(x), _vp(y), _vp(w), C.c_int64(n_per_group), G, _vp(st), st_stride, _vp(out), _s()),
          "geo4d_delta125")
    return out


def pnp_moments(pts: torch.Tensor, conf: torch.Tensor, F: int, HW: int, W: int, cx: float, cy: float,
                gate: Optional[torch.Tensor] = None, ncand: int = 1, thr_px: float = 5.0) -> torch.Tensor:
    """pts [F, HW, 3], conf [F, HW] fp32 contiguous; gate [F, ncand, 13] (w2c 3x4 row-major + focal) or None.
    Returns [F, ncand, 41] fp64 moments."""
    out = torch.empty((F, ncand, 41), device=pts.device, dtype=torch.float64)
    check(lib().geo4d_pnp_moments(_vp(pts), _vp(conf), F, HW, W, C.c_float(cx), C.c_float(cy), _vp(gate), ncand,
                                  C.c_float(thr_px), _vp(out), _s()), "geo4d_pnp_moments")
    return out


def shift_focal_sums(pts: torch.Tensor, conf: torch.Tensor, G: int, HW: int, W: int, H: int, shift: torch.Tensor,
                     zoff: float) -> torch.Tensor:
    out = torch.empty((G, 6), device=pts.device, dtype=torch.float64)
    check(lib().geo4d_shift_focal_sums(_vp(pts), _vp(conf), G, HW, W, H, _vp(shift), C.c_float(zoff), _vp(out), _s()),
          "geo4d_shift_focal_sums")
    return out


# --------------------------------------------------------------------------- launch accounting
_replayed_kernels = 0


def launch_count() -> int:
    """Kernels launched by libgeo4d_b200 so far: direct launches + kernels re-issued by CUDA-graph replays."""
    f = lib().geo4d_launch_count
    f.restype = C.c_uint64
    return int(f()) + _replayed_kernels


def raw_launch_count() -> int:
    f = lib().geo4d_launch_count
    f.restype = C.c_uint64
    return int(f())


def note_replay(kernels_in_graph: int, times: int = 1) -> None:
    global _replayed_kernels
    _replayed_kernels += kernels_in_graph * times
