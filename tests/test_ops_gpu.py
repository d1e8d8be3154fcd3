"""GPU: every C-ABI kernel of the denoising / decode path against a plain PyTorch fp32 reference of the same op
on the same (bf16-rounded) inputs.  Tolerances: bf16 outputs 2.5e-3 relative L2 (one bf16 rounding of the
result), fp32 outputs 1e-5; integer / data-movement ops exact.  The tap-GEMM cases run in every kernel mode:
single CTA vs CTA pair (cta_group::2), TMA-store vs direct-store epilogue."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_TOL = 2.5e-3


def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))


@pytest.fixture(params=["single", "pair", "single-direct", "pair-direct"])
def gemm_mode(request, cuda_device, monkeypatch):
    from geo4d_b200 import ops
    lib = ops.lib()
    monkeypatch.setattr(ops, "_AUTOTUNE", False)     # the mode under test must not be overridden by the tuner
    lib.geo4d_debug_gemm_pair_mode(1 if request.param.startswith("pair") else 0)
    lib.geo4d_debug_gemm_direct_store(1 if request.param.endswith("direct") else 0)
    yield request.param
    lib.geo4d_debug_gemm_pair_mode(-1)
    lib.geo4d_debug_gemm_direct_store(0)


def _lin(M, K, n, bias=True, act=0, residual=False, out_dtype=torch.bfloat16, seed=0):
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(n, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(n, device="cuda", generator=g) if bias else None
    n_store = n // 2 if act == 2 else n
    res = torch.randn(M, n_store, device="cuda", generator=g).bfloat16() if residual else None
    ref = x.float() @ w.float().t()
    if bias:
        ref = ref + b
    if act == 1:
        ref = F.silu(ref)
    wk, bk = w, b
    if act == 2:   # GEGLU: value * gelu(gate); kernel layout interleaves [32 value | 32 gate] weight rows
        val, gate = ref[:, : n // 2], ref[:, n // 2:]
        ref = val * F.gelu(gate)
        h = n // 2
        wk = torch.stack([w[:h].reshape(h // 32, 32, K), w[h:].reshape(h // 32, 32, K)], 1).reshape(n, K).contiguous()
        if bias:
            bk = torch.stack([b[:h].reshape(h // 32, 32), b[h:].reshape(h // 32, 32)], 1).reshape(n).contiguous()
    if residual:
        ref = ref + res.float()
    out = ops.linear(x, wk, bk, act=act, residual=res, out_dtype=out_dtype)
    torch.cuda.synchronize()
    return _rel(out, ref)


@pytest.mark.parametrize("shape", [
    dict(M=128, K=64, n=32), dict(M=256, K=256, n=64), dict(M=384, K=128, n=128, residual=True),
    dict(M=1024, K=320, n=320, act=1), dict(M=300, K=192, n=200), dict(M=130, K=64, n=96, residual=True),
    dict(M=512, K=320, n=2560, act=2), dict(M=256, K=64, n=128, act=2),
    dict(M=128 * 40, K=640, n=1280, residual=True),          # more tiles than SMs: accumulator / ring phase wraps
    dict(M=128 * 37 + 5, K=128, n=320, residual=True),       # odd number of row boxes (pair tail), ragged rows
    dict(M=2560, K=1280, n=1280), dict(M=640, K=1280, n=1280, act=1),
])
def test_linear(gemm_mode, shape):
    assert _lin(**shape) < BF16_TOL


@pytest.mark.parametrize("shape", [dict(M=1000, K=512, n=512), dict(M=512, K=128, n=16), dict(M=512, K=128, n=3)])
def test_linear_fp32_out(gemm_mode, shape):
    assert _lin(out_dtype=torch.float32, **shape) < 1e-5


def _conv(N, H, W, Cin, Cout, seed=0, row_bias=False, residual=False):
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    b = torch.randn(Cout, device="cuda", generator=g)
    ref = F.conv2d(x.float(), w.float(), b, padding=1)
    rb = None
    if row_bias:
        rb = torch.randn(N, Cout, device="cuda", generator=g)
        ref = ref + rb[:, :, None, None]
    res = torch.randn(N * H * W, Cout, device="cuda", generator=g).bfloat16() if residual else None
    x2 = x.permute(0, 2, 3, 1).reshape(N * H * W, Cin).contiguous()
    w9 = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous()
    out = ops.conv3x3(x2, N, H, W, w9, b, row_bias=rb, rows_per_bias=H * W, residual=res)
    torch.cuda.synchronize()
    ref2 = ref.permute(0, 2, 3, 1).reshape(N * H * W, Cout)
    if residual:
        ref2 = ref2 + res.float()
    return _rel(out, ref2)


@pytest.mark.parametrize("shape", [
    dict(N=4, H=8, W=16, Cin=64, Cout=64), dict(N=2, H=40, W=64, Cin=128, Cout=320, row_bias=True, residual=True),
    dict(N=16, H=5, W=8, Cin=128, Cout=128), dict(N=16, H=10, W=16, Cin=192, Cout=256),
    dict(N=1, H=24, W=256, Cin=64, Cout=32), dict(N=2, H=7, W=200, Cin=64, Cout=40),
    dict(N=3, H=5, W=8, Cin=64, Cout=64, row_bias=True),     # 3 frames per 128-row box: a tile straddles emb rows
])
def test_conv3x3(gemm_mode, shape):
    assert _conv(**shape) < BF16_TOL


def test_temporal_conv(gemm_mode):
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    B, T, HW, Cc = 2, 4, 160, 128
    x = torch.randn(B, Cc, T, HW, 1, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cc, Cc, 3, 1, 1, device="cuda", generator=g) / (3 * Cc) ** 0.5).bfloat16()
    b = torch.randn(Cc, device="cuda", generator=g)
    ref = F.conv3d(x.float(), w.float(), b, padding=(1, 0, 0))
    x2 = x[..., 0].permute(0, 2, 3, 1).reshape(B * T * HW, Cc).contiguous()
    w3 = w[:, :, :, 0, 0].permute(2, 0, 1).contiguous()
    res = torch.randn(B * T * HW, Cc, device="cuda", generator=g).bfloat16()
    out = ops.temporal_conv3(x2, B, T, HW, w3, b, residual=res)
    torch.cuda.synchronize()
    ref2 = ref[..., 0].permute(0, 2, 3, 1).reshape(B * T * HW, Cc) + res.float()
    assert _rel(out, ref2) < BF16_TOL


def test_bmm(gemm_mode):
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    a = torch.randn(3, 320, 128, device="cuda", generator=g).bfloat16()
    b = torch.randn(3, 192, 128, device="cuda", generator=g).bfloat16()
    out = ops.bmm_nt(a, b, alpha=0.125, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert _rel(out, 0.125 * torch.einsum("bmk,bnk->bmn", a.float(), b.float())) < 1e-5


def test_every_tile_configuration_gives_identical_bits(cuda_device, monkeypatch):
    """tile_n / cta_pair only change the schedule: each output element sees the same k-step sequence, so all
    legal configurations -- and therefore whatever the autotuner pins -- produce bit-identical results."""
    from geo4d_b200 import ops
    monkeypatch.setattr(ops, "_AUTOTUNE", False)
    g = torch.Generator(device="cuda").manual_seed(11)
    M, K, n = 128 * 9 + 40, 320, 640
    x = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(n, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(n, device="cuda", generator=g)
    res = torch.randn(M, n, device="cuda", generator=g).bfloat16()
    base = ops.linear(x, w, b, residual=res)
    for pair in (1, 2):
        for tn in (32, 64, 128, 160, 256):
            monkeypatch.setattr(ops, "_FORCE_TILE", (tn, pair))
            out = ops.linear(x, w, b, residual=res)
            torch.cuda.synchronize()
            assert torch.equal(out, base), (tn, pair)
    monkeypatch.setattr(ops, "_FORCE_TILE", None)


def test_autotune_pins_a_configuration(cuda_device, monkeypatch):
    from geo4d_b200 import ops
    monkeypatch.setattr(ops, "_AUTOTUNE", True)
    g = torch.Generator(device="cuda").manual_seed(12)
    x = torch.randn(4096, 640, device="cuda", generator=g).bfloat16()
    w = (torch.randn(640, 640, device="cuda", generator=g) / 25).bfloat16()
    n0 = len(ops.tuned_configs())
    a = ops.linear(x, w)
    b = ops.linear(x, w)
    assert len(ops.tuned_configs()) == n0 + 1            # tuned once, reused afterwards
    key, cfg, timings = ops.tuned_configs()[-1]
    assert cfg in timings and timings[cfg] == min(timings.values())
    assert torch.equal(a, b)
    assert _rel(a, x.float() @ w.float().t()) < BF16_TOL


def test_bad_arguments_raise(cuda_device):
    from geo4d_b200 import ops
    from geo4d_b200._cabi import Geo4DError
    x = torch.randn(128, 48, device="cuda").bfloat16()       # K not a multiple of 64
    w = torch.randn(32, 48, device="cuda").bfloat16()
    with pytest.raises(Geo4DError):
        ops.linear(x, w)


@pytest.mark.parametrize("S,rows,Cc,silu,eps", [(4, 128, 64, True, 1e-5), (16, 2560, 320, True, 1e-5),
                                                  (1, 16 * 640, 640, False, 1e-6), (3, 40, 1920, True, 1e-5),
                                                  (2, 77, 192, True, 1e-6), (1, 40960, 320, True, 1e-5)])
def test_groupnorm(cuda_device, S, rows, Cc, silu, eps):
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(S)
    x = (torch.randn(S * rows, Cc, device="cuda", generator=g) * 1.5 + 0.3).bfloat16()
    gamma = torch.randn(Cc, device="cuda", generator=g)
    beta = torch.randn(Cc, device="cuda", generator=g)
    y = ops.groupnorm(x, S, rows, gamma, beta, eps, silu)
    y2 = ops.groupnorm(x, S, rows, gamma, beta, eps, silu)
    assert torch.equal(y, y2)                                   # fixed-order reduction: bit-reproducible
    ref = F.group_norm(x.float().reshape(S, rows, Cc).permute(0, 2, 1), 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    assert _rel(y, ref.permute(0, 2, 1).reshape(S * rows, Cc)) < BF16_TOL


@pytest.mark.parametrize("M,Cc", [(100, 64), (4096, 320), (1000, 640), (333, 1280), (64, 512), (7, 2048)])
def test_layernorm(cuda_device, M, Cc):
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M)
    x = (torch.randn(M, Cc, device="cuda", generator=g) * 2 + 0.5).bfloat16()
    gamma = torch.randn(Cc, device="cuda", generator=g)
    beta = torch.randn(Cc, device="cuda", generator=g)
    y = ops.layernorm(x, gamma, beta)
    assert _rel(y, F.layer_norm(x.float(), (Cc,), gamma, beta, 1e-5)) < BF16_TOL


def _attn_ref(q, k, v, H, scale):
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    qh = q.float().reshape(B, Lq, H, 64).permute(0, 2, 1, 3)
    kh = k.float().reshape(k.shape[0], Lk, H, 64).permute(0, 2, 1, 3)
    vh = v.float().reshape(v.shape[0], Lk, H, 64).permute(0, 2, 1, 3)
    p = (torch.einsum("bhid,bhjd->bhij", qh, kh) * scale).softmax(-1)
    return torch.einsum("bhij,bhjd->bhid", p, vh).permute(0, 2, 1, 3).reshape(B, Lq, H * 64)


@pytest.mark.parametrize("B,H,Lq,Lk,kv_shared,accumulate", [
    (1, 1, 128, 128, False, False), (2, 3, 640, 640, False, False), (2, 2, 160, 160, False, False),
    (3, 20, 40, 40, False, False), (4, 5, 256, 77, True, False), (4, 5, 256, 16, False, True),
    (2, 5, 2560, 2560, False, False)])
def test_attention(cuda_device, B, H, Lq, Lk, kv_shared, accumulate):
    """geo4d_attention vs softmax(QK^T/8)V; P is rounded to bf16 before PV, hence 4e-3."""
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(Lq + Lk)
    inner = H * 64
    q = torch.randn(B * Lq, inner, device="cuda", generator=g).bfloat16()
    Bk = 1 if kv_shared else B
    kv = torch.randn(Bk * Lk, 2 * inner, device="cuda", generator=g).bfloat16()
    out = (torch.randn(B * Lq, inner, device="cuda", generator=g).bfloat16() if accumulate
           else torch.empty(B * Lq, inner, device="cuda", dtype=torch.bfloat16))
    prev = out.float().clone()
    ops.attention(q, kv[:, :inner], kv[:, inner:], out, B, H, Lq, Lk, kv_batch_div=(B if kv_shared else 1),
                  accumulate=accumulate)
    torch.cuda.synchronize()
    k = kv[:, :inner].reshape(Bk, Lk, inner)
    v = kv[:, inner:].reshape(Bk, Lk, inner)
    if kv_shared:
        k, v = k.expand(B, Lk, inner), v.expand(B, Lk, inner)
    ref = _attn_ref(q.reshape(B, Lq, inner), k, v, H, 0.125).reshape(B * Lq, inner)
    if accumulate:
        ref = ref + prev
    assert _rel(out, ref) < 4e-3


@pytest.mark.parametrize("B,T,H,Lq,Lk1,Lk2", [(8, 4, 5, 300, 77, 16), (4, 4, 3, 128, 200, 16), (16, 16, 5, 2560, 77, 16),
                                             (2, 2, 10, 640, 77, 130)])
def test_cross_attention_two_key_sets_one_launch(cuda_device, B, T, H, Lq, Lk1, Lk2):
    """geo4d_cross_attention2: text keys shared by the T frames of a clip + per-frame image keys, two independent
    softmaxes, summed (attention.py:166-207) -- vs the two-launch form and vs fp32 torch."""
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(Lq + Lk1)
    inner = H * 64
    q = torch.randn(B * Lq, inner, device="cuda", generator=g).bfloat16()
    kv1 = torch.randn((B // T) * Lk1, 2 * inner, device="cuda", generator=g).bfloat16()
    kv2 = torch.randn(B * Lk2, 2 * inner, device="cuda", generator=g).bfloat16()
    out = torch.empty(B * Lq, inner, device="cuda", dtype=torch.bfloat16)
    ops.cross_attention2(q, kv1[:, :inner], kv1[:, inner:], Lk1, T, kv2[:, :inner], kv2[:, inner:], Lk2, 1, out, B, H, Lq)
    two = torch.empty_like(out)
    ops.attention(q, kv1[:, :inner], kv1[:, inner:], two, B, H, Lq, Lk1, kv_batch_div=T)
    ops.attention(q, kv2[:, :inner], kv2[:, inner:], two, B, H, Lq, Lk2, kv_batch_div=1, accumulate=True)
    torch.cuda.synchronize()
    k1 = kv1[:, :inner].reshape(B // T, Lk1, inner).repeat_interleave(T, dim=0)
    v1 = kv1[:, inner:].reshape(B // T, Lk1, inner).repeat_interleave(T, dim=0)
    k2, v2 = kv2[:, :inner].reshape(B, Lk2, inner), kv2[:, inner:].reshape(B, Lk2, inner)
    qq = q.reshape(B, Lq, inner)
    ref = (_attn_ref(qq, k1, v1, H, 0.125) + _attn_ref(qq, k2, v2, H, 0.125)).reshape(B * Lq, inner)
    assert _rel(out, ref) < 4e-3
    assert _rel(out, two) < 6e-3        # the two-launch form rounds the first branch to bf16 before adding the second


@pytest.mark.parametrize("B,T,HW,H", [(1, 16, 160, 5), (2, 4, 128, 2), (1, 16, 40, 20), (1, 7, 33, 3)])
def test_temporal_attention(cuda_device, B, T, HW, H):
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(T)
    inner = H * 64
    qkv = torch.randn(B * T * HW, 3 * inner, device="cuda", generator=g).bfloat16()
    out = torch.empty(B * T * HW, inner, device="cuda", dtype=torch.bfloat16)
    ops.temporal_attention(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], out, B, T, HW, H)

    def seq(x):
        return x.reshape(B, T, HW, inner).permute(0, 2, 1, 3).reshape(B * HW, T, inner)
    ref = _attn_ref(seq(qkv[:, :inner]), seq(qkv[:, inner:2 * inner]), seq(qkv[:, 2 * inner:]), H, 0.125)
    ref = ref.reshape(B, HW, T, inner).permute(0, 2, 1, 3).reshape(B * T * HW, inner)
    assert _rel(out, ref) < 4e-3


def test_data_movement_exact(cuda_device):
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    B, T, H, W = 2, 4, 8, 16
    a = torch.randn(B, 16, T, H, W, device="cuda", generator=g)
    b = torch.randn(B, 4, T, H, W, device="cuda", generator=g)
    rows = ops.bcthw_to_rows(a, b, 64)
    ref = torch.cat([a, b], 1).permute(0, 2, 3, 4, 1).reshape(-1, 20)
    assert torch.equal(rows[:, :20], ref.bfloat16()) and float(rows[:, 20:].abs().max()) == 0.0
    r32 = torch.randn(B * T * H * W, 16, device="cuda", generator=g)
    assert torch.equal(ops.rows_to_bcthw(r32, 16, B, T, H, W), r32.reshape(B, T, H, W, 16).permute(0, 4, 1, 2, 3))
    x = torch.randn(100, 64, device="cuda", generator=g).bfloat16()
    y = torch.randn(100, 128, device="cuda", generator=g).bfloat16()
    assert torch.equal(ops.concat_rows(x, y), torch.cat([x, y], 1))
    N, H2, W2, Cc = 3, 5, 8, 64
    x = torch.randn(N * H2 * W2, Cc, device="cuda", generator=g).bfloat16()
    ref = F.interpolate(x.float().reshape(N, H2, W2, Cc).permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    assert torch.equal(ops.upsample2x(x, N, H2, W2), ref.permute(0, 2, 3, 1).reshape(-1, Cc).bfloat16())


@pytest.mark.parametrize("pad_before", [1, 0])
def test_stride2_conv_via_im2col(gemm_mode, pad_before):
    """Downsample convs: U-Net pads 1 on every side (openaimodel3d.py Downsample), the VAE pads (0,1,0,1)
    (ae_modules.py Downsample)."""
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    N, Ci, Co, H3, W3 = 3, 64, 128, 8, 16
    xi = torch.randn(N, Ci, H3, W3, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / 24).bfloat16()
    bias = torch.randn(Co, device="cuda", generator=g)
    ref = (F.conv2d(xi.float(), w.float(), bias, stride=2, padding=1) if pad_before == 1
           else F.conv2d(F.pad(xi.float(), (0, 1, 0, 1)), w.float(), bias, stride=2))
    Ho, Wo = ref.shape[2], ref.shape[3]
    col = ops.im2col_s2(xi.permute(0, 2, 3, 1).reshape(-1, Ci).contiguous(), N, H3, W3, pad_before, Ho, Wo)
    out = ops.linear(col, w.permute(0, 2, 3, 1).reshape(Co, 9 * Ci).contiguous(), bias)
    assert _rel(out, ref.permute(0, 2, 3, 1).reshape(-1, Co)) < BF16_TOL


def test_ddim_step_counter_gather(cuda_device):
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    xx = torch.randn(1000, device="cuda", generator=g)
    vv = torch.randn(1000, device="cuda", generator=g)
    coef = torch.tensor([[0.9, 0.4, 1.0, 0.8, 0.6, 0.0], [0.5, 0.85, 0.98, 0.7, 0.7, 0.0]], device="cuda")
    idx = torch.zeros(1, dtype=torch.int32, device="cuda")
    x1, p0 = xx.clone(), torch.empty_like(xx)
    ops.advance_counter(idx, 1)
    ops.ddim_step(x1, vv, coef, idx, pred_x0=p0)
    sa, s1, rs, sap, dr, _ = coef[1].tolist()
    e_t = sa * vv + s1 * xx
    x0 = (sa * xx - s1 * vv) * rs
    assert _rel(x1, sap * x0 + dr * e_t) < 1e-6 and _rel(p0, x0) < 1e-6
    tab = torch.randn(5, 40, device="cuda", generator=g)
    o = torch.empty(40, device="cuda")
    ops.gather_row(tab, idx, o)
    assert torch.equal(o, tab[1])


@pytest.mark.parametrize("split", [2, 5, 16])
def test_split_k_matches_fused_path(cuda_device, monkeypatch, split):
    """Split-K tap-GEMM (the 5x8 level: 640 rows, K = 9 x 1280): partial fp32 planes + ordered reduce with the full
    epilogue (bias, per-frame embedding row, residual) vs the single-pass kernel and vs fp32 torch."""
    from geo4d_b200 import ops
    monkeypatch.setattr(ops, "_AUTOTUNE", False)
    g = torch.Generator(device="cuda").manual_seed(21)
    N, H, W, Cin, Cout = 16, 5, 8, 1280, 1280
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    b = torch.randn(Cout, device="cuda", generator=g)
    rb = torch.randn(1, Cout, device="cuda", generator=g)
    res = torch.randn(N * H * W, Cout, device="cuda", generator=g).bfloat16()
    ref = (F.conv2d(x.float(), w.float(), b, padding=1) + rb[:, :, None, None]).permute(0, 2, 3, 1).reshape(N * H * W, Cout) \
        + res.float()
    x2 = x.permute(0, 2, 3, 1).reshape(N * H * W, Cin).contiguous()
    w9 = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous()
    monkeypatch.setattr(ops, "_FORCE_TILE", (256, 1))
    monkeypatch.setattr(ops, "_FORCE_SPLIT", 1)
    base = ops.conv3x3(x2, N, H, W, w9, b, row_bias=rb, rows_per_bias=N * H * W, residual=res)
    monkeypatch.setattr(ops, "_FORCE_SPLIT", split)
    out = ops.conv3x3(x2, N, H, W, w9, b, row_bias=rb, rows_per_bias=N * H * W, residual=res)
    out2 = ops.conv3x3(x2, N, H, W, w9, b, row_bias=rb, rows_per_bias=N * H * W, residual=res)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)                      # ordered reduction: bit-reproducible
    assert _rel(out, ref) < BF16_TOL and _rel(base, ref) < BF16_TOL
    assert _rel(out, base) < 3e-3                      # same products, different fp32 summation order, bf16 output
    # fp32 output + SiLU through the reduce pass
    wl = (torch.randn(256, 2560, device="cuda", generator=g) / 50).bfloat16()
    xl = torch.randn(640, 2560, device="cuda", generator=g).bfloat16()
    bl = torch.randn(256, device="cuda", generator=g)
    o = ops.linear(xl, wl, bl, act=ops.ACT_SILU, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert _rel(o, F.silu(xl.float() @ wl.float().t() + bl)) < 1e-4
    monkeypatch.setattr(ops, "_FORCE_TILE", None)
    monkeypatch.setattr(ops, "_FORCE_SPLIT", None)
