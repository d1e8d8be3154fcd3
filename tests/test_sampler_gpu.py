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
    # every kernel of the step is deterministic (fixed-order GroupNorm statistics, no float atomics); the first call
    # ran step 0 eagerly before capturing, so the two calls may only differ through the autotuner's tile choice
    assert rel_l2(out_g2, out_g) < 1e-2
    smp_e = DDIMSampler(model, use_cuda_graph=False)
    out_e, _ = smp_e.sample(S=S, batch_size=b, shape=(16, t, hh, ww), conditioning=cond, eta=0.0, verbose=False,
                            x_T=x_T.to(cuda_device), fs=fs.to(cuda_device), timestep_spacing="uniform_trailing")
    assert rel_l2(out_e, ref) < TOL_LATENT
    assert rel_l2(out_e, out_g) < 1e-2


def _tiny_setup(cuda_device, seed=21):
    from oracle import unet as ou
    from tests.test_unet_gpu import make_unet
    cfg = ou.UNetConfig.tiny()
    sd = ou.init_params(ou.param_shapes(cfg), seed=seed)
    net = make_unet(dict(model_channels=64, context_dim=64, temporal_length=4), sd, cuda_device)
    g = torch.Generator().manual_seed(5)
    b, t, hh, ww = 1, 4, 8, 16
    x_T = torch.randn(b, 16, t, hh, ww, generator=g)
    zc = torch.randn(b, 4, t, hh, ww, generator=g)
    ctx = torch.randn(b, 77 + 16 * t, 64, generator=g)
    ctx_u = torch.randn(b, 77 + 16 * t, 64, generator=g)      # a different ("unconditional") context
    ctx_ui = torch.cat([ctx_u[:, :77], ctx[:, 77:]], 1)        # text of the unconditional one, image tokens of the conditional
    return cfg, sd, net, (b, t, hh, ww), x_T, zc, ctx, ctx_u, ctx_ui, torch.tensor([24])


def test_classifier_free_guidance_batched_in_the_graph(cuda_device):
    """SURVEY N4 / ddim.py:216-229: cond and uncond evaluated as ONE b = 2 U-Net pass inside the captured step,
    guidance mix + rescale_noise_cfg; vs the oracle sampler.  Identical conditionings collapse to one pass."""
    from oracle import unet as ou, ddim as od
    from geo4d_b200.sampler import DDIMSampler
    cfg, sd, net, (b, t, hh, ww), x_T, zc, ctx, ctx_u, _, fs = _tiny_setup(cuda_device)
    S, scale, resc = 4, 7.5, 0.7
    sch = od.Schedule.geo4d()
    f = lambda c: (lambda x, ts: ou.forward(cfg, sd, torch.cat([x, zc], 1), ts, c, fs))
    ref, _ = od.ddim_sample(f(ctx), x_T, sch, S, cfg_scale=scale, apply_model_uncond=f(ctx_u), guidance_rescale=resc)
    model = _TinyModel(net, cuda_device)
    d = lambda c: {"c_crossattn": [c.to(cuda_device)], "c_concat": [zc.to(cuda_device)]}
    smp = DDIMSampler(model)
    kw = dict(S=S, batch_size=b, shape=(16, t, hh, ww), eta=0.0, verbose=False, x_T=x_T.to(cuda_device),
              fs=fs.to(cuda_device), timestep_spacing="uniform_trailing")
    out, _ = smp.sample(conditioning=d(ctx), unconditional_guidance_scale=scale, unconditional_conditioning=d(ctx_u),
                        guidance_rescale=resc, **kw)
    e = rel_l2(out, ref)
    assert e < TOL_LATENT, e
    assert any(k[7] == 2 for k in smp._graphs), "cond/uncond were not batched into one captured step"
    # the shipped settings: the unconditional conditioning IS the conditional one -> one pass, same result as no guidance
    plain, _ = smp.sample(conditioning=d(ctx), **kw)
    same, _ = smp.sample(conditioning=d(ctx), unconditional_guidance_scale=scale,
                         unconditional_conditioning=d(ctx.clone()), guidance_rescale=resc, **kw)
    assert torch.equal(plain, same)
    assert all(k[7] == 1 for k in smp._graphs if k[9][0] == "plain" or torch.equal(torch.tensor(k[8]), torch.zeros(len(k[8]), dtype=torch.long)))
    assert any(k[9][0] == "cfg" and k[7] == 1 for k in smp._graphs)   # guidance with identical conditionings: one pass


def test_three_way_guidance_multicond(cuda_device):
    """ddim_multiplecond.py:226-236: v = v_u + s_img (v_ui - v_u) + s (v_c - v_ui), three conditionings in one pass"""
    from oracle import unet as ou, ddim as od
    from geo4d_b200.sampler import DDIMSampler_multicond
    cfg, sd, net, (b, t, hh, ww), x_T, zc, ctx, ctx_u, ctx_ui, fs = _tiny_setup(cuda_device)
    S, scale, s_img = 3, 5.0, 2.0
    f = lambda c: (lambda x, ts: ou.forward(cfg, sd, torch.cat([x, zc], 1), ts, c, fs))
    ref, _ = od.ddim_sample(f(ctx), x_T, od.Schedule.geo4d(), S, cfg_scale=scale, apply_model_uncond=f(ctx_u),
                            apply_model_uncond_img=f(ctx_ui), cfg_img=s_img)
    d = lambda c: {"c_crossattn": [c.to(cuda_device)], "c_concat": [zc.to(cuda_device)]}
    smp = DDIMSampler_multicond(_TinyModel(net, cuda_device))
    out, _ = smp.sample(S=S, batch_size=b, shape=(16, t, hh, ww), conditioning=d(ctx), eta=0.0, verbose=False,
                        x_T=x_T.to(cuda_device), fs=fs.to(cuda_device), timestep_spacing="uniform_trailing",
                        unconditional_guidance_scale=scale, unconditional_conditioning=d(ctx_u), cfg_img=s_img,
                        unconditional_conditioning_img_nonetext=d(ctx_ui))
    e = rel_l2(out, ref)
    assert e < TOL_LATENT, e
    assert any(k[7] == 3 for k in smp._graphs)


def test_eta_noise_in_the_graph(cuda_device):
    """eta = 1 (the README's 'better visual' setting, ddim.py:273-277): per-step sigma_t * randn drawn from the CUDA
    generator in the reference's order; the oracle is fed the very same draws."""
    from oracle import unet as ou, ddim as od
    from geo4d_b200.sampler import DDIMSampler
    cfg, sd, net, (b, t, hh, ww), x_T, zc, ctx, _, _, fs = _tiny_setup(cuda_device)
    S = 4
    torch.cuda.manual_seed(77)
    draws = [torch.randn((b, 16, t, hh, ww), device=cuda_device).cpu() for _ in range(S)]
    it = iter(draws)
    ref, _ = od.ddim_sample(lambda x, ts: ou.forward(cfg, sd, torch.cat([x, zc], 1), ts, ctx, fs), x_T,
                            od.Schedule.geo4d(), S, eta=1.0, noise_fn=lambda shape: next(it))
    smp = DDIMSampler(_TinyModel(net, cuda_device))
    torch.cuda.manual_seed(77)
    out, _ = smp.sample(S=S, batch_size=b, shape=(16, t, hh, ww), eta=1.0, verbose=False, x_T=x_T.to(cuda_device),
                        conditioning={"c_crossattn": [ctx.to(cuda_device)], "c_concat": [zc.to(cuda_device)]},
                        fs=fs.to(cuda_device), timestep_spacing="uniform_trailing")
    e = rel_l2(out, ref)
    assert e < TOL_LATENT, e
    assert any(k[10] for k in smp._graphs)      # the stochastic step was captured, not run eagerly
