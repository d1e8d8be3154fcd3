#!/usr/bin/env python
"""GPU probe for the tap-GEMM kernel: correctness vs torch (fp32 reference on bf16-rounded inputs)
and a few timings.  Each case runs in its own subprocess so a trap in one does not poison the rest.
Writes gpurun_out/probe_gemm.json."""
import json, os, subprocess, sys, time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
OUT = os.path.join(REPO, "gpurun_out")

CASES = {}

def case(fn):
    CASES[fn.__name__] = fn
    return fn

def _rel(a, b):
    import torch
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))

def _lin(M, K, n, bias=True, act=0, residual=False, out_dtype="bf16", seed=0):
    import torch
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(n, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(n, device="cuda", generator=g) if bias else None
    n_store = n // 2 if act == 2 else n
    res = torch.randn(M, n_store, device="cuda", generator=g).bfloat16() if residual else None
    ref = x.float() @ w.float().t()
    if bias: ref = ref + b
    if act == 1: ref = torch.nn.functional.silu(ref)
    wk = w
    bk = b
    if act == 2:
        val, gate = ref[:, : n // 2], ref[:, n // 2:]
        ref = val * torch.nn.functional.gelu(gate)
        # interleave weight rows in blocks of 32: [32 value | 32 gate]
        h = n // 2
        wv, wg = w[:h].reshape(h // 32, 32, K), w[h:].reshape(h // 32, 32, K)
        wk = torch.stack([wv, wg], 1).reshape(n, K).contiguous()
        if bias:
            bk = torch.stack([b[:h].reshape(h // 32, 32), b[h:].reshape(h // 32, 32)], 1).reshape(n).contiguous()
    if residual: ref = ref + res.float()
    out = ops.linear(x, wk, bk, act=act, residual=res,
                     out_dtype=torch.float32 if out_dtype == "fp32" else torch.bfloat16)
    torch.cuda.synchronize()
    return _rel(out, ref)

@case
def linear_small():
    return {"rel": _lin(128, 64, 32)}

@case
def linear_k256_n64():
    return {"rel": _lin(256, 256, 64)}

@case
def linear_n128():
    return {"rel": _lin(384, 128, 128, residual=True)}

@case
def linear_n160():
    return {"rel": _lin(1024, 320, 320, act=1)}

@case
def linear_n256():
    return {"rel": _lin(1000, 512, 512, out_dtype="fp32")}

@case
def linear_n_tail():
    return {"rel": _lin(300, 192, 200)}

@case
def linear_tiny_n():
    return {"rel16": _lin(512, 128, 16, out_dtype="fp32"), "rel3": _lin(512, 128, 3, out_dtype="fp32")}

@case
def linear_geglu():
    return {"rel": _lin(512, 320, 2560, act=2), "rel128": _lin(256, 64, 128, act=2)}

@case
def linear_multi_tile_persistent():
    # more tiles than SMs, exercises accumulator double buffering and phase wrap
    return {"rel": _lin(128 * 40, 640, 1280, residual=True)}

def _conv(N, H, W, Cin, Cout, seed=0, row_bias=False, residual=False):
    import torch
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    b = torch.randn(Cout, device="cuda", generator=g)
    ref = torch.nn.functional.conv2d(x.float(), w.float(), b, padding=1)
    rb = None
    if row_bias:
        rb = torch.randn(N, Cout, device="cuda", generator=g)
        ref = ref + rb[:, :, None, None]
    res = None
    if residual:
        res = torch.randn(N * H * W, Cout, device="cuda", generator=g).bfloat16()
    x2 = x.permute(0, 2, 3, 1).reshape(N * H * W, Cin).contiguous()
    w9 = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous()
    out = ops.conv3x3(x2, N, H, W, w9, b, row_bias=rb, rows_per_bias=H * W, residual=res)
    torch.cuda.synchronize()
    ref2 = ref.permute(0, 2, 3, 1).reshape(N * H * W, Cout)
    if residual: ref2 = ref2 + res.float()
    return _rel(out, ref2)

@case
def conv_8x16():
    return {"rel": _conv(4, 8, 16, 64, 64)}

@case
def conv_40x64():
    return {"rel": _conv(2, 40, 64, 128, 320, row_bias=True, residual=True)}

@case
def conv_5x8():
    return {"rel": _conv(16, 5, 8, 128, 128)}

@case
def conv_10x16():
    return {"rel": _conv(16, 10, 16, 192, 256)}

@case
def conv_wide():
    return {"rel": _conv(1, 24, 256, 64, 32), "rel_odd": _conv(2, 7, 200, 64, 3)}

@case
def temporal_conv():
    import torch
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    B, T, HW, Cc = 2, 4, 160, 128
    x = torch.randn(B, Cc, T, HW, 1, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cc, Cc, 3, 1, 1, device="cuda", generator=g) / (3 * Cc) ** 0.5).bfloat16()
    b = torch.randn(Cc, device="cuda", generator=g)
    ref = torch.nn.functional.conv3d(x.float(), w.float(), b, padding=(1, 0, 0))
    x2 = x[..., 0].permute(0, 2, 3, 1).reshape(B * T * HW, Cc).contiguous()
    w3 = w[:, :, :, 0, 0].permute(2, 0, 1).contiguous()
    res = torch.randn(B * T * HW, Cc, device="cuda", generator=g).bfloat16()
    out = ops.temporal_conv3(x2, B, T, HW, w3, b, residual=res)
    torch.cuda.synchronize()
    ref2 = ref[..., 0].permute(0, 2, 3, 1).reshape(B * T * HW, Cc) + res.float()
    return {"rel": _rel(out, ref2)}

@case
def bmm():
    import torch
    from geo4d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    a = torch.randn(3, 320, 128, device="cuda", generator=g).bfloat16()
    b = torch.randn(3, 192, 128, device="cuda", generator=g).bfloat16()
    out = ops.bmm_nt(a, b, alpha=0.125, out_dtype=torch.float32)
    torch.cuda.synchronize()
    ref = 0.125 * torch.einsum("bmk,bnk->bmn", a.float(), b.float())
    return {"rel": _rel(out, ref)}

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
def timing():
    import torch
    from geo4d_b200 import ops
    res = {}
    for (M, K, n) in [(40960, 320, 320), (40960, 320, 2560), (40960, 1280, 320), (10240, 640, 640), (8192, 8192, 8192), (2560, 1280, 5120)]:
        x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(n, K, device="cuda").bfloat16()
        out = torch.empty(M, n, device="cuda", dtype=torch.bfloat16)
        ms = _time(lambda: ops.linear(x, w, out=out))
        ms_t = _time(lambda: torch.matmul(x, w.t()))
        res[f"linear_{M}x{K}x{n}"] = {"ms": ms, "tflops": 2 * M * K * n / ms / 1e9, "torch_ms": ms_t, "torch_tflops": 2 * M * K * n / ms_t / 1e9}
    for (N, H, W, Ci, Co) in [(16, 40, 64, 320, 320), (16, 20, 32, 640, 640), (16, 10, 16, 1280, 1280), (16, 5, 8, 1280, 1280), (16, 40, 64, 960, 320), (4, 320, 512, 128, 128), (16, 80, 128, 512, 512)]:
        x = torch.randn(N * H * W, Ci, device="cuda").bfloat16(); w9 = torch.randn(9, Co, Ci, device="cuda").bfloat16()
        out = torch.empty(N * H * W, Co, device="cuda", dtype=torch.bfloat16)
        ms = _time(lambda: ops.conv3x3(x, N, H, W, w9, out=out))
        xt = x.reshape(N, H, W, Ci).permute(0, 3, 1, 2); wt = w9.reshape(3, 3, Co, Ci).permute(2, 3, 0, 1).contiguous(memory_format=torch.channels_last)
        xt = xt.contiguous(memory_format=torch.channels_last)
        ms_t = _time(lambda: torch.nn.functional.conv2d(xt, wt, padding=1))
        fl = 2 * N * H * W * Ci * Co * 9
        res[f"conv_{N}x{H}x{W}_{Ci}to{Co}"] = {"ms": ms, "tflops": fl / ms / 1e9, "torch_cudnn_ms": ms_t, "torch_tflops": fl / ms_t / 1e9}
    return res

def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--case":
        import torch
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
        print(name, json.dumps(results[name])[:600], flush=True)
        with open(os.path.join(OUT, "probe_gemm.json"), "w") as f:
            json.dump(results, f, indent=1)

if __name__ == "__main__":
    main()
