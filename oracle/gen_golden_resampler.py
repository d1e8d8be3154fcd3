#!/usr/bin/env python
"""Pin geo4d_b200.resampler.Resampler against the reference's own class (test infrastructure; BUILD container only):

    python oracle/gen_golden_resampler.py

Imports lvdm/modules/encoders/resampler.py UNMODIFIED, builds it with the `image_proj_stage_config` of
configs/inference_geo4d.yaml, fills it with seeded weights (oracle.unet.init_params over its own state-dict keys: the
keys and shapes are stored, so the test also pins the checkpoint layout), runs it on seeded CLIP-like tokens
[1, 257, 1280] (and on the per-frame form [1, 2, 257, 1280] of cross_attention=True) and stores the REFERENCE's outputs
(fp16) in tests/golden/resampler_ref.pt."""
import os
import sys
from collections import OrderedDict

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GEO4D_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

KW = dict(dim=1024, depth=4, dim_head=64, heads=12, num_queries=16, embedding_dim=1280, output_dim=1024, ff_mult=4,
          video_length=16)


def seeded_state(shapes, seed=5):
    from oracle import unet as ou
    sd = ou.init_params(OrderedDict(shapes), seed=seed)
    sd["latents"] = sd["latents"] * 30.0      # the reference draws latents ~ N(0, 1/dim): keep them O(1/sqrt(dim)) * a few
    return sd


def main():
    import yaml
    from lvdm.modules.encoders.resampler import Resampler as Ref
    cfg = yaml.safe_load(open(os.path.join(REPO, "configs", "inference_geo4d.yaml")))
    assert cfg["model"]["params"]["image_proj_stage_config"]["params"] == KW
    ref = Ref(**KW).eval()
    shapes = [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]
    sd = seeded_state(shapes)
    ref.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 257, 1280, generator=g)
    xf = torch.randn(1, 2, 257, 1280, generator=g)
    with torch.no_grad():
        y = ref(x)
        ref.video_length = 2   # per-frame form: latents are viewed as [B, T, num_queries, dim]; use the first 2 frames' queries
        lat_full = ref.latents.data
        ref.latents.data = lat_full[:, :2 * KW["num_queries"]]
        yf = ref(xf)
        ref.latents.data = lat_full
    # inputs are re-drawn from the seed by the test (torch.Generator().manual_seed(9): x then xf); a checksum pins them
    torch.save({"kw": KW, "shapes": shapes, "seed": 5, "input_seed": 9, "x_sum": float(x.double().sum()),
                "xf_sum": float(xf.double().sum()), "y": y.half(), "yf": yf.half()},
               os.path.join(REPO, "tests", "golden", "resampler_ref.pt"))
    print("reference output", tuple(y.shape), float(y.abs().mean()), "per-frame", tuple(yf.shape))


if __name__ == "__main__":
    main()
