"""GPU: DDIMSampler (CUDA-graph path and eager path) + tiny CUDA U-Net vs the CPU oracle sampler driving the
oracle U-Net with identical seeded weights, x_T and conditioning.
Tolerance: S-step latent rel-L2 <= 5e-2 (bf16 network inside an fp32 recurrence)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL_LATENT = 5e-2


def rel_l2(a, b):
    return float((a.float().cpu() - b.float().cpu()).norm() / (b.float().cpu().norm() + 1e-12))


class _TinyModel(torch.nn.Module):
    """Minimal stand-in for LatentVisualDiffusion exposing what DDIMSampler reads."""

    def __init__(self, unet, device):
        super().__init__()
        from geo4d_b200 import schedule as sched
        bufs = sched.register_schedule_buffers(1000, 0.00085, 0.012, "linear", True)
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                  "sqrt_one_minus_alphas_cumprod"):
            self.register_buffer(k, torch.tensor(bufs[k]).to(device))
        self.register_buffer("scale_arr", torch.tensor(sched.make_scale_arr()).to(device))
        self.num_timesteps = 1000
        self.parameterization = "v"
        self.use_dynamic_rescale = True

        class W(torch.nn.Module):
            conditioning_key = "hybrid"
        self.model = W()
        self.model.diffusion_model = unet

    def apply_model(self, x, t, cond, **kw):
        xc = torch.cat([x] + cond["c_concat"], 1)
        return self.model.diffusion_model(xc, t, context=cond["c_crossattn"][0], **kw)


@pytest.mark.parametrize("S", [3, 6])
def test_ddim_graph_and_eager_vs_oracle(cuda_device, S):
    from oracle import unet as ou
    from oracle import ddim as od
    from geo4d_b200.sampler import DDIMSampler
    from tests.test_unet_gpu import make_unet
    cfg = ou.UNetConfig.tiny()
    sd = ou.init_params(ou.param_shapes(cfg), seed=21)
    net = make_unet(dict(model_channels=64, context_dim=64, temporal_length=4), sd, cuda_device)
    g = torch.Generator().manual_seed(5)
    b, t, hh, ww = 1, 4, 8, 16
    x_T = torch.randn(b, 16, t, hh, ww, generator=g)
    zc = torch.randn(b, 4, t, hh, ww, generator=g)
    ctx = torch.randn(b, 77 + 16 * t, 64, generator=g)
    fs = torch.tensor([24])
    sch = od.Schedule.geo4d()

    def oracle_model(x, ts):
        return ou.forward(cfg, sd, torch.cat([x, zc], 1), ts, ctx, fs)

    ref, _ = od.ddim_sample(oracle_model, x_T, sch, S)
    model = _TinyModel(net, cuda_device)
    cond = {"c_crossattn": [ctx.to(cuda_device)], "c_concat": [zc.to(cuda_device)]}
    smp = DDIMSampler(model)
    out_g, inter = smp.sample(S=S, batch_size=b, shape=(16, t, hh, ww), conditioning=cond, eta=0.0,
                              verbose=False, x_T=x_T.to(cuda_device), fs=fs.to(cuda_device),
                              timestep_spacing="uniform_trailing")
    torch.cuda.synchronize()
    assert rel_l2(out_g, ref) < TOL_LATENT
    # replaying the cached graph (all S steps from the graph this time) must reproduce the result
    out_g2, _ = smp.sample(S=S, batch_size=b, shape=(16, t, hh, ww), conditioning=cond, eta=0.0, verbose=False,
                           x_T=x_T.to(cuda_device), fs=fs.to(cuda_device), timestep_spacing="uniform_trailing")
    # GroupNorm statistics use float atomics, so replays agree to rounding noise, not bitwise
    assert rel_l2(out_g2, out_g) < 1e-2
    smp_e = DDIMSampler(model, use_cuda_graph=False)
    out_e, _ = smp_e.sample(S=S, batch_size=b, shape=(16, t, hh, ww), conditioning=cond, eta=0.0, verbose=False,
                            x_T=x_T.to(cuda_device), fs=fs.to(cuda_device), timestep_spacing="uniform_trailing")
    assert rel_l2(out_e, ref) < TOL_LATENT
    assert rel_l2(out_e, out_g) < 1e-2
