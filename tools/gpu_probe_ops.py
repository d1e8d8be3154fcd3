#!/usr/bin/env python
"""GPU probe for the non-GEMM kernels (norms, data movement, temporal attention, tcgen05 attention).
Each case runs in its own subprocess.  Writes gpurun_out/probe_ops.json."""
import json, os, subprocess, sys, time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
OUT = os.path.join(REPO, "gpurun_out")
CASES = {}

def case(fn):
    CASES[fn.__name__] = fn
    return fn

def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))

def _time(fn, iters=20, warm=3):
    import torch
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

@case
def groupnorm():
    import torch, torch.nn.functional as F
    from geo4d_b200 import ops
    res = {}
    for (S, rows, Cc, silu, eps) in [(4, 128, 64, True, 1e-5), (16, 2560, 320, True, 1e-5), (1, 16 * 640, 640, False, 1e-6), (3, 40, 1920, True, 1e-5), (2, 77, 192, True, 1e-6)]:
        g = torch.Generator(device="cuda").manual_seed(S)
        x = (torch.randn(S * rows, Cc, device="cuda", generator=g) * 1.5 + 0.3).bfloat16()
        gamma = torch.randn(Cc, device="cuda", generator=g); beta = torch.randn(Cc, device="cuda", generator=g)
        y = ops.groupnorm(x, S, rows, gamma, beta, eps, silu)
        xr = x.float().reshape(S, rows, Cc).permute(0, 2, 1)
        ref = F.group_norm(xr, 32, gamma, beta, eps)
        if silu: ref = F.silu(ref)
        ref = ref.permute(0, 2, 1).reshape(S * rows, Cc)
        res[f"S{S}_r{rows}_C{Cc}"] = _rel(y, ref)
    x = torch.randn(16 * 2560, 320, device="cuda").bfloat16(); gamma = torch.ones(320, device="cuda"); beta = torch.zeros(320, device="cuda")
    out = torch.empty_like(x)
    ms = _time(lambda: ops.groupnorm(x, 16, 2560, gamma, beta, 1e-5, True, out=out))
    res["time_16x2560x320_ms"] = ms
    res["time_16x2560x320_GBs_algorithmic_2x"] = 2 * x.numel() * 2 / ms / 1e6
    return res

@case
def layernorm():
    import torch, torch.nn.functional as F
    from geo4d_b200 import ops
    res = {}
    for (M, Cc) in [(100, 64), (4096, 320), (1000, 640), (333, 1280), (64, 512)]:
        g = torch.Generator(device="cuda").manual_seed(M)
        x = (torch.randn(M, Cc, device="cuda", generator=g) * 2 + 0.5).bfloat16()
        gamma = torch.randn(Cc, device="cuda", generator=g); beta = torch.randn(Cc, device="cuda", generator=g)
        y = ops.layernorm(x, gamma, beta)
        ref = F.layer_norm(x.float(), (Cc,), gamma, beta, 1e-5)
        res[f"M{M}_C{Cc}"] = _rel(y, ref)
    x = torch.randn(40960, 320, device="cuda").bfloat16(); gamma = torch.ones(320, device="cuda"); beta = torch.zeros(320, device="cuda")
    out = torch.empty_like(x)
    ms = _time(lambda: ops.layernorm(x, gamma, beta, out=out))
    res["time_40960x320_ms"] = ms
    return res

@case
def movement():
    import torch, torch.nn.functional as F
    from geo4d_b200 import ops
    res = {}
    g = torch.Generator(device="cuda").manual_seed(0)
    B, T, H, W = 2, 4, 8, 16
    a = torch.randn(B, 16, T, H, W, device="cuda", generator=g); b = torch.randn(B, 4, T, H, W, device="cuda", generator=g)
    rows = ops.bcthw_to_rows(a, b, 64)
    ref = torch.cat([a, b], 1).permute(0, 2, 3, 4, 1).reshape(-1, 20)
    res["bcthw_to_rows"] = _rel(rows[:, :20], ref.bfloat16()); res["bcthw_pad_zero"] = float(rows[:, 20:].abs().max())
    r32 = torch.randn(B * T * H * W, 16, device="cuda", generator=g)
    back = ops.rows_to_bcthw(r32, 16, B, T, H, W)
    res["rows_to_bcthw"] = _rel(back, r32.reshape(B, T, H, W, 16).permute(0, 4, 1, 2, 3))
    x = torch.randn(100, 64, device="cuda", generator=g).bfloat16(); y = torch.randn(100, 128, device="cuda", generator=g).bfloat16()
    res["concat"] = _rel(ops.concat_rows(x, y), torch.cat([x, y], 1))
    N, H2, W2, Cc = 3, 5, 8, 64
    x = torch.randn(N * H2 * W2, Cc, device="cuda", generator=g).bfloat16()
    up = ops.upsample2x(x, N, H2, W2)
    ref = F.interpolate(x.float().reshape(N, H2, W2, Cc).permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1).reshape(-1, Cc)
    res["upsample"] = _rel(up, ref)
    # stride-2 conv via im2col + linear, both padding conventions
    for name, pad_before, H3, W3 in (("unet_pad1", 1, 8, 16), ("vae_pad0", 0, 8, 16)):
        Ci, Co = 64, 128
        xi = torch.randn(N, Ci, H3, W3, device="cuda", generator=g).bfloat16()
        w = (torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / 24).bfloat16()
        bias = torch.randn(Co, device="cuda", generator=g)
        if pad_before == 1:
            ref = F.conv2d(xi.float(), w.float(), bias, stride=2, padding=1)
        else:
            ref = F.conv2d(F.pad(xi.float(), (0, 1, 0, 1)), w.float(), bias, stride=2)
        Ho, Wo = ref.shape[2], ref.shape[3]
        xr = xi.permute(0, 2, 3, 1).reshape(-1, Ci).contiguous()
        col = ops.im2col_s2(xr, N, H3, W3, pad_before, Ho, Wo)
        wk = w.permute(0, 2, 3, 1).reshape(Co, 9 * Ci).contiguous()
        out = ops.linear(col, wk, bias)
        res["s2conv_" + name] = _rel(out, ref.permute(0, 2, 3, 1).reshape(-1, Co))
    # ddim step + counter + gather
    xx = torch.randn(1000, device="cuda", generator=g); vv = torch.randn(1000, device="cuda", generator=g)
    coef = torch.tensor([[0.9, 0.4, 1.0, 0.8, 0.6, 0.0], [0.5, 0.85, 0.98, 0.7, 0.7, 0.0]], device="cuda")
    idx = torch.zeros(1, dtype=torch.int32, device="cuda")
    x1 = xx.clone(); p0 = torch.empty_like(xx)
    ops.advance_counter(idx, 1)
    ops.ddim_step(x1, vv, coef, idx, pred_x0=p0)
    sa, s1, rs, sap, dr, _ = coef[1].tolist()
    e_t = sa * vv + s1 * xx; x0 = (sa * xx - s1 * vv) * rs
    res["ddim_step"] = _rel(x1, sap * x0 + dr * e_t); res["ddim_x0"] = _rel(p0, x0)
    tab = torch.randn(5, 40, device="cuda", generator=g); o = torch.empty(40, device="cuda")
    ops.gather_row(tab, idx, o)
    res["gather"] = float((o - tab[1]).abs().max())
    return res

def _attn_ref(q, k, v, H, scale):
    import torch
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    qh = q.float().reshape(B, Lq, H, 64).permute(0, 2, 1, 3)
    kh = k.float().reshape(k.shape[0], Lk, H, 64).permute(0, 2, 1, 3)
    vh = v.float().reshape(v.shape[0], Lk, H, 64).permute(0, 2, 1, 3)
    s = torch.einsum("bhid,bhjd->bhij", qh, kh) * scale
    p = s.softmax(-1)
    o = torch.einsum("bhij,bhjd->bhid", p, vh)
    return o.permute(0, 2, 1, 3).reshape(B, Lq, H * 64)

@case
def temporal_attention():
    import torch
    from geo4d_b200 import ops
    res = {}
    for (B, T, HW, H) in [(1, 16, 160, 5), (2, 4, 128, 2), (1, 16, 40, 20), (1, 7, 33, 3)]:
        g = torch.Generator(device="cuda").manual_seed(T)
        inner = H * 64
        qkv = torch.randn(B * T * HW, 3 * inner, device="cuda", generator=g).bfloat16()
        out = torch.empty(B * T * HW, inner, device="cuda", dtype=torch.bfloat16)
        ops.temporal_attention(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], out, B, T, HW, H)
        # reference: sequences over t for each (b, p)
        def seq(x):
            return x.reshape(B, T, HW, inner).permute(0, 2, 1, 3).reshape(B * HW, T, inner)
        ref = _attn_ref(seq(qkv[:, :inner]), seq(qkv[:, inner:2 * inner]), seq(qkv[:, 2 * inner:]), H, 0.125)
        ref = ref.reshape(B, HW, T, inner).permute(0, 2, 1, 3).reshape(B * T * HW, inner)
        res[f"B{B}_T{T}_HW{HW}_H{H}"] = _rel(out, ref)
    B, T, HW, H = 1, 16, 2560, 5
    qkv = torch.randn(B * T * HW, 3 * H * 64, device="cuda").bfloat16(); out = torch.empty(B * T * HW, H * 64, device="cuda", dtype=torch.bfloat16)
    ms = _time(lambda: ops.temporal_attention(qkv[:, :320], qkv[:, 320:640], qkv[:, 640:], out, B, T, HW, H))
    res["time_ms_level0"] = ms; res["GBs_algorithmic"] = 4 * B * T * HW * 320 * 2 / ms / 1e6
    return res

def _attn_case(B, H, Lq, Lk, kv_shared=False, accumulate=False, seed=0):
    import torch
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    inner = H * 64
    q = torch.randn(B * Lq, inner, device="cuda", generator=g).bfloat16()
    Bk = 1 if kv_shared else B
    kv = torch.randn(Bk * Lk, 2 * inner, device="cuda", generator=g).bfloat16()
    out = torch.randn(B * Lq, inner, device="cuda", generator=g).bfloat16() if accumulate else torch.empty(B * Lq, inner, device="cuda", dtype=torch.bfloat16)
    prev = out.float().clone()
    ops.attention(q, kv[:, :inner], kv[:, inner:], out, B, H, Lq, Lk, kv_batch_div=(B if kv_shared else 1), accumulate=accumulate)
    torch.cuda.synchronize()
    k = kv[:, :inner].reshape(Bk, Lk, inner); v = kv[:, inner:].reshape(Bk, Lk, inner)
    if kv_shared: k = k.expand(B, Lk, inner); v = v.expand(B, Lk, inner)
    ref = _attn_ref(q.reshape(B, Lq, inner), k, v, H, 0.125).reshape(B * Lq, inner)
    if accumulate: ref = ref + prev
    return _rel(out, ref)

@case
def attn_128x128():
    return {"rel": _attn_case(1, 1, 128, 128)}

@case
def attn_multi_tile():
    return {"rel": _attn_case(2, 3, 640, 640, seed=1)}

@case
def attn_tails():
    return {"rel_160": _attn_case(2, 2, 160, 160, seed=2), "rel_40": _attn_case(3, 20, 40, 40, seed=3)}

@case
def attn_cross():
    return {"text77_shared": _attn_case(4, 5, 256, 77, kv_shared=True, seed=4),
            "img16_accumulate": _attn_case(4, 5, 256, 16, accumulate=True, seed=5)}

@case
def attn_2560():
    return {"rel": _attn_case(2, 5, 2560, 2560, seed=6)}

@case
def attn_timing():
    import torch
    from geo4d_b200 import ops
    res = {}
    for (B, H, L) in [(16, 5, 2560), (16, 10, 640), (16, 20, 160)]:
        inner = H * 64
        qkv = torch.randn(B * L, 3 * inner, device="cuda").bfloat16(); out = torch.empty(B * L, inner, device="cuda", dtype=torch.bfloat16)
        ms = _time(lambda: ops.attention(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], out, B, H, L, L))
        fl = 4 * B * H * L * L * 64
        q4 = qkv[:, :inner].reshape(B, L, H, 64).permute(0, 2, 1, 3); k4 = qkv[:, inner:2 * inner].reshape(B, L, H, 64).permute(0, 2, 1, 3); v4 = qkv[:, 2 * inner:].reshape(B, L, H, 64).permute(0, 2, 1, 3)
        ms_t = _time(lambda: torch.nn.functional.scaled_dot_product_attention(q4, k4, v4))
        res[f"B{B}_H{H}_L{L}"] = {"ms": ms, "tflops": fl / ms / 1e9, "torch_sdpa_ms": ms_t, "torch_tflops": fl / ms_t / 1e9}
    return res

def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--case":
        r = CASES[sys.argv[2]]()
        print("RESULT " + json.dumps(r))
        return
    os.makedirs(OUT, exist_ok=True)
    results = {}
    names = [n for n in CASES if (len(sys.argv) < 2 or n in sys.argv[1:])]
    for name in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--case", name], capture_output=True, text=True, timeout=300)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if p.returncode == 0 and line:
                results[name] = {"ok": True, "result": json.loads(line[-1][7:])}
            else:
                results[name] = {"ok": False, "rc": p.returncode, "stdout": p.stdout[-1500:], "stderr": p.stderr[-2500:]}
        except subprocess.TimeoutExpired:
            results[name] = {"ok": False, "timeout": True}
        results[name]["sec"] = round(time.time() - t0, 1)
        print(name, json.dumps(results[name])[:700], flush=True)
        with open(os.path.join(OUT, "probe_ops.json"), "w") as f:
            json.dump(results, f, indent=1)

if __name__ == "__main__":
    main()
