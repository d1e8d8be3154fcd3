#!/usr/bin/env python
"""Does any tap-GEMM tile configuration change the bits of a U-Net forward?  Records an exact checksum of every
op output in call order under several forced configurations and reports the first op that differs."""
import os, sys
os.environ["GEO4D_AUTOTUNE"] = "0"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from geo4d_b200 import ops, synthetic

REC = []
def bits(t):
    v = t.contiguous().view(torch.int16 if t.element_size() == 2 else torch.int32)
    return int(v.to(torch.int64).sum().item()) ^ int((v.to(torch.int64) * 31 % 1000003).sum().item())

def wrap(name, fn):
    def inner(*a, **k):
        r = fn(*a, **k)
        if isinstance(r, torch.Tensor):
            desc = name
            if name in ("linear", "conv3x3", "temporal_conv3"):
                x = a[0]
                desc = f"{name} in={tuple(x.shape)} act={k.get('act', 0)} res={k.get('residual') is not None} rb={k.get('row_bias') is not None} out={tuple(r.shape)} {r.dtype}"
            REC.append((desc, bits(r)))
        return r
    return inner

for n in ["linear", "conv3x3", "temporal_conv3", "bmm_nt", "groupnorm", "layernorm", "attention", "temporal_attention",
          "bcthw_to_rows", "rows_to_bcthw", "concat_rows", "upsample2x", "im2col_s2"]:
    setattr(ops, n, wrap(n, getattr(ops, n)))

def main():
    dev = torch.device("cuda")
    H, W = 320, 512
    model, pm_vae, cfg = synthetic.build_model(device=dev, seed=0)
    unet = model.model.diffusion_model
    x = torch.randn(1, 20, 16, H // 8, W // 8, device=dev)
    ctx = torch.cat([model.get_learned_conditioning([""]), model.get_image_conditioning(1)], 1)
    ts = torch.tensor([499], device=dev); fs = torch.tensor([24], device=dev)
    unet(x, ts, context=ctx, fs=fs)   # warm-up: prepare()/set_context ops are not part of the comparison
    runs = {}
    for label, force in [("auto-model", None), ("auto-model again", None), ("single-160", (160, 1)), ("pair-160", (160, 2)),
                         ("single-256", (256, 1)), ("pair-256", (256, 2)), ("single-128", (128, 1)), ("single-64", (64, 1))]:
        REC.clear()
        ops._FORCE_TILE = force
        try:
            y = unet(x, ts, context=ctx, fs=fs)
            torch.cuda.synchronize()
            runs[label] = list(REC)
            print(label, "ops", len(REC), "final", bits(y), flush=True)
        except Exception as e:
            print(label, "failed:", str(e)[:200], flush=True)
            ops._FORCE_TILE = None
    base = runs["auto-model"]
    for label, rec in runs.items():
        bad = [i for i, (a, b) in enumerate(zip(base, rec)) if a[1] != b[1]]
        if bad:
            i = bad[0]
            print(f"{label}: {len(bad)} of {len(rec)} ops differ; first at op #{i}: {rec[i][0]}  (previous op: {rec[i-1][0] if i else None})")
        else:
            print(f"{label}: bit-identical to auto-model")

if __name__ == "__main__":
    main()
