"""GPU parity of the drop-in modules (forward only), the encoder, checkpoint round trips and error behaviour."""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("size,B", [(32, 8), (64, 4), (128, 2)])
def test_generator_and_critic_forward(cuda_dev, size, B):
    from rnagan_b200 import dcgan
    lrelu, tanh = torch.nn.LeakyReLU(0.2), torch.nn.Tanh()
    oG = O.OracleGenerator(2048, size, 3, 64, nonlinearity=lrelu, last_nonlinearity=tanh)
    oD = O.OracleCritic(size, 3, 64, nonlinearity=lrelu, last_nonlinearity=lrelu)
    O.reinit_(oG, 1); O.reinit_(oD, 2)
    G = dcgan.DCGANGenerator(2048, size, 3, 64, nonlinearity=torch.nn.LeakyReLU(0.2),
                             last_nonlinearity=torch.nn.Tanh()).to(cuda_dev)
    D = dcgan.DCGANDiscriminator(size, 3, 64, nonlinearity=torch.nn.LeakyReLU(0.2),
                                 last_nonlinearity=torch.nn.LeakyReLU(0.2)).to(cuda_dev)
    G.load_state_dict(oG.state_dict()); D.load_state_dict(oD.state_dict())
    g = torch.Generator().manual_seed(3)
    z = torch.randn(B, 2048, generator=g)
    x = torch.rand(B, 3, size, size, generator=g) * 2 - 1
    for mode in ("train", "eval"):
        getattr(oG, mode)(); getattr(oD, mode)(); getattr(G, mode)(); getattr(D, mode)()
        with torch.no_grad():
            ref_img, ref_out = oG(z), oD(x)
        img, out = G(z.to(cuda_dev)), D(x.to(cuda_dev))
        assert img.shape == ref_img.shape and img.dtype == torch.float32
        assert _rel(img, ref_img) <= 3e-2, mode
        assert (out.cpu() - ref_out).abs().max().item() <= 0.03 + 0.03 * ref_out.abs().max().item(), mode
    # train-mode forwards updated the running statistics like the reference's modules do
    for (n, bo), (_, bm) in zip(oG.named_buffers(), G.named_buffers()):
        if n.endswith("num_batches_tracked"):
            assert int(bo) == int(bm) == 1
        else:
            assert _rel(bm.float(), bo.float()) <= 2e-2, n
    feat = D(x.to(cuda_dev), feature_matching=True)
    with torch.no_grad():
        assert feat.shape == oD(x, feature_matching=True).shape


def test_encoder_matches_oracle(cuda_dev):
    from rnagan_b200.betaVAE import betaVAE
    feats, B = 300, 16
    oV = O.OracleVAE(feats, beta=0.005).eval()
    O.reinit_(oV, 23)
    vae = betaVAE(feats, 2048, [6000, 4000, 2048], [4000, 6000], beta=0.005)
    vae.load_state_dict(oV.state_dict())
    vae = vae.to(cuda_dev).eval()
    x = torch.randn(B, feats, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        zm, zl, h = oV.encode(x)
    gm, gl, gh = vae.encode(x.to(cuda_dev))
    assert _rel(gm, zm) <= 2e-2 and _rel(gl, zl) <= 2e-2 and _rel(gh, h) <= 2e-2
    vae.train()
    with pytest.raises(NotImplementedError):
        vae.encode(x.to(cuda_dev))


def test_decoder_forward_sample_match_oracle(cuda_dev):
    """Eval-mode betaVAE.decode / forward / sample (src/betaVAE.py:109-143) against the oracle's fp32 modules;
    the feature count (302) is not a multiple of 4 or 64 on purpose (ragged last Linear)."""
    from rnagan_b200.betaVAE import betaVAE
    feats, B = 302, 12
    oV = O.OracleVAE(feats, beta=0.005).eval()
    O.reinit_(oV, 29)
    with torch.no_grad():          # non-trivial running statistics for the decoder's BatchNorm1d layers
        for m in oV.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
    vae = betaVAE(feats, 2048, [6000, 4000, 2048], [4000, 6000], beta=0.005)
    vae.load_state_dict(oV.state_dict())
    vae = vae.to(cuda_dev).eval()
    z = torch.randn(B, 2048, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        ref = oV.decode(z)
    got = vae.decode(z.to(cuda_dev))
    # three bf16-operand GEMMs in a row (K up to 6000) ahead of the tanh: rel-L2 at the encoder test's level
    assert got.shape == (B, feats) and _rel(got, ref) <= 3e-2 and (got.cpu() - ref).abs().max().item() <= 8e-2
    # sample: CPU-drawn latents in the reference's order
    torch.manual_seed(5)
    with torch.no_grad():
        ref_s = oV.decode(torch.randn(7, 2048))
    torch.manual_seed(5)
    got_s = vae.sample(7, cuda_dev)
    assert _rel(got_s, ref_s) <= 3e-2
    # forward: same z_mean / z_log_var as the oracle; the reconstruction uses device noise, so check it through decode
    x = torch.randn(B, feats, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        zm, zl, _ = oV.encode(x)
    out, gm, gl = vae(x.to(cuda_dev))
    assert out.shape == (B, feats) and torch.isfinite(out).all()
    assert _rel(gm, zm) <= 2e-2 and _rel(gl, zl) <= 2e-2
    vae.train()
    with pytest.raises(NotImplementedError):
        vae.decode(z.to(cuda_dev))


def test_latent_prep_matches_reference_formula(cuda_dev):
    from rnagan_b200 import ops
    g = torch.Generator().manual_seed(9)
    noise = torch.rand(16, 2048, generator=g) * 0.6 - 0.3
    z = torch.randn(16, 2048, generator=g)
    ref = O.latent_prep(noise, z)
    lat = torch.empty(16, 2048, device=cuda_dev)
    ops.latent_prep(noise.to(cuda_dev), z.to(cuda_dev), lat_f32=lat)
    assert (lat.cpu() - ref).abs().max().item() <= 2e-5
    # single profile broadcast (generate_images): conditioning cancels, SURVEY.md section 3.3
    ref1 = O.latent_prep(noise, z[:1])
    ops.latent_prep(noise.to(cuda_dev), z[:1].contiguous().to(cuda_dev), lat_f32=lat)
    assert (lat.cpu() - ref1).abs().max().item() <= 2e-5
    # B == 1 gives NaN like the reference (unbiased std of one sample)
    one = torch.empty(1, 2048, device=cuda_dev)
    ops.latent_prep(noise[:1].contiguous().to(cuda_dev), z[:1].contiguous().to(cuda_dev), lat_f32=one)
    assert torch.isnan(one).all()


def test_adam_matches_torch(cuda_dev):
    from rnagan_b200.optim import adam_step
    g = torch.Generator().manual_seed(1)
    shapes = [(2048, 64, 4, 4), (1000,), (3,), (77, 5)]
    ps = [torch.nn.Parameter(torch.randn(s, generator=g).to(cuda_dev)) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    o1 = torch.optim.Adam(ps, lr=4e-4, betas=(0.5, 0.999))
    o2 = torch.optim.Adam(qs, lr=4e-4, betas=(0.5, 0.999))
    for _ in range(3):
        for p, q in zip(ps, qs):
            gr = torch.randn(p.shape, generator=g).to(cuda_dev)
            p.grad = gr.clone(); q.grad = gr.clone()
        adam_step(o1)
        o2.step()
    for p, q in zip(ps, qs):
        assert (p - q).abs().max().item() <= 1e-6
    sd = o1.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 3.0
    # bf16 shadows re-emitted by the step: a dense one and a PITCHED one (rows padded like a K = 19198 -> 19200 operand;
    # 70001 columns: odd, so rows start at every alignment and chunks of 65536 elements straddle rows); the pad columns
    # must stay untouched and the payload must equal a cast of the updated parameter bit for bit
    w = torch.nn.Parameter(torch.randn(37, 70001, generator=g).to(cuda_dev))
    d = torch.nn.Parameter(torch.randn(129, 64, generator=g).to(cuda_dev))
    w._rg_shadow = torch.full((37, 70008), -7.0, dtype=torch.bfloat16, device=cuda_dev)
    d._rg_shadow = torch.zeros(129, 64, dtype=torch.bfloat16, device=cuda_dev)
    o3 = torch.optim.Adam([w, d], lr=1e-2)
    for _ in range(2):
        w.grad = torch.randn(w.shape, generator=g).to(cuda_dev)
        d.grad = torch.randn(d.shape, generator=g).to(cuda_dev)
        adam_step(o3)
    assert torch.equal(w._rg_shadow[:, :70001], w.detach().to(torch.bfloat16))
    assert (w._rg_shadow[:, 70001:] == -7.0).all()
    assert torch.equal(d._rg_shadow, d.detach().to(torch.bfloat16))


def test_checkpoint_roundtrip_and_errors(cuda_dev):
    from rnagan_b200 import dcgan, wgan_loss
    from rnagan_b200.trainer import Trainer
    feats = 64
    oV = O.OracleVAE(feats).eval()
    d = tempfile.mkdtemp()
    ckpt = os.path.join(d, "vae.pt")
    torch.save(oV.state_dict(), ckpt)
    net = {
        "generator": {"name": dcgan.DCGANGenerator, "args": {"encoding_dims": 2048, "out_size": 32, "step_channels": 64},
                      "optimizer": {"name": torch.optim.Adam, "args": {"lr": 1e-4, "betas": (0.5, 0.999)}}},
        "discriminator": {"name": dcgan.DCGANDiscriminator, "args": {"in_size": 32, "step_channels": 64},
                          "optimizer": {"name": torch.optim.Adam, "args": {"lr": 4e-4, "betas": (0.5, 0.999)}}},
    }
    mk = lambda: [wgan_loss.WassersteinGeneratorLossVAE(ckpt, feats), wgan_loss.WassersteinDiscriminatorLossVAE(ckpt, feats),
                  wgan_loss.WassersteinGradientPenaltyVAE(ckpt, feats)]
    tr = Trainer(net, mk(), device=cuda_dev, checkpoints=os.path.join(d, "gan"), epochs=1, devices=[0])
    data = O.make_batch(8, feats, 32, 1)
    tr.real_inputs = data
    vals = tr.train_iter()
    assert list(vals) == ["WassersteinGeneratorLossVAE", "WassersteinDiscriminatorLossVAE", "WassersteinGradientPenaltyVAE"]
    assert all(np.isfinite(v) for v in vals.values())
    path = tr.save_model(0)
    saved = torch.load(path, weights_only=False)
    for k in ("epoch", "loss_information", "loss_objects", "generator", "discriminator", "optimizer_generator",
              "optimizer_discriminator"):
        assert k in saved
    tr2 = Trainer(net, mk(), device=cuda_dev, checkpoints=os.path.join(d, "gan"), epochs=1)
    tr2.load_model(load_path=path)
    for (n, a), (_, b) in zip(tr.generator.state_dict().items(), tr2.generator.state_dict().items()):
        assert torch.equal(a, b), n
    assert tr2.start_epoch == 1
    # same batch, same RNG -> the restored trainer continues identically (deterministic kernels)
    tr.real_inputs = tr2.real_inputs = data
    torch.manual_seed(3); v1 = tr.train_iter()
    torch.manual_seed(3); v2 = tr2.train_iter()
    assert list(v1.values()) == list(v2.values())
    # deferred read-back (Trainer.train's per-batch call) = the synchronous one, value for value, over fresh batches
    batches = [O.make_batch(8, feats, 32, 10 + i) for i in range(4)]
    torch.manual_seed(5)
    sync_vals = []
    for b in batches:
        tr.real_inputs = b
        sync_vals.append(list(tr.train_iter().values()))
    torch.manual_seed(5)
    n0 = len(tr2.loss_logs["WassersteinGeneratorLossVAE"])
    for b in batches:
        tr2.real_inputs = b
        assert tr2.train_iter(defer=True) is None
    logs = tr2.loss_logs
    lag_vals = [[logs[k][n0 + i] for k in logs] for i in range(4)]
    assert lag_vals == sync_vals
    assert tr2.loss_information["discriminator_iters"] == tr.loss_information["discriminator_iters"]
    # reference error behaviour
    tr.generator.label_type = "required"
    with pytest.raises(Exception, match="GAN model requires labels for training"):
        tr.train_iter()


def test_vae_train_step_matches_oracle(cuda_dev):
    """betaVAE training step (config 5) vs the oracle on identical Dropout mask / reparametrisation noise:
    losses within 2e-2 relative, per-parameter gradient cosine >= 0.98 (bf16 operands), BN1d running stats, and the
    frozen-encoder path sees the updated weights afterwards."""
    from rnagan_b200 import betaVAE as bv
    feats, B, beta = 300, 32, 0.0005
    oV = O.OracleVAE(feats, beta=beta)
    O.reinit_(oV, 23)
    oV.train()
    vae = bv.betaVAE(feats, 2048, [6000, 4000, 2048], [4000, 6000], beta=beta)
    vae.load_state_dict(oV.state_dict())
    vae = vae.to(cuda_dev).train()
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, feats, generator=g)
    keep = (torch.rand(B, feats, generator=g) >= 0.5).float()
    eps = torch.randn(B, 2048, generator=g)
    oo = torch.optim.Adam(oV.parameters(), lr=5e-5)
    om = torch.optim.Adam(vae.parameters(), lr=5e-5)
    ref = O.vae_train_step_explicit(oV, oo, x, beta, keep, eps)
    out3 = bv.train_step(vae, om, x.to(cuda_dev), beta, keep_mask=keep.to(cuda_dev), eps=eps.to(cuda_dev)).cpu()
    for got, key in zip(out3.tolist(), ("total_loss", "reconstruction_loss", "kl_loss")):
        assert abs(got - ref[key]) <= 2e-2 * abs(ref[key]) + 1e-4, key
    for (n, po), (_, pm) in zip(oV.named_parameters(), vae.named_parameters()):
        a, b = pm.grad.double().flatten().cpu(), po.grad.double().flatten()
        if n.endswith(".0.bias") and not n.startswith("decoder.2"):
            # bias of a Linear that feeds BatchNorm: its exact gradient is 0 (BN removes the mean); both sides only
            # hold rounding noise, so compare magnitudes instead of directions
            assert b.norm().item() <= 1e-4 and a.norm().item() <= 1e-3, n
            continue
        if b.norm() < 1e-12:
            continue
        cos = (a @ b / (a.norm() * b.norm())).item()
        assert cos >= 0.98, f"{n}: cosine {cos:.4f}"
        assert _rel(pm.data, po.data) <= 1e-2, n
    for (n, bo), (_, bm) in zip(oV.named_buffers(), vae.named_buffers()):
        if n.endswith("num_batches_tracked"):
            assert int(bo) == int(bm)
        else:
            assert _rel(bm.float(), bo.float()) <= 2e-2, n
    # a second step with internally drawn randomness runs and stays finite; eval-mode encode uses the new weights
    out = bv.train_step(vae, om, x.to(cuda_dev), beta)
    assert torch.isfinite(out).all()
    vae.eval(); oV.eval()
    zm = vae.encode(x.to(cuda_dev))[0]
    assert torch.isfinite(zm).all()


def test_up_generator_forward_matches_reference_golden(cuda_dev):
    """DCGANUpGenerator (src/dcgan.py:8-99) forward against the output of the reference's own class (golden) and the
    oracle, train-mode BN, same weights and latent."""
    from rnagan_b200 import dcgan
    from tests import _util as U
    gold = U.load_golden("modules.npz")
    oU = O.OracleUpGenerator(2048, 32, 3, 64, nonlinearity=torch.nn.LeakyReLU(0.2), last_nonlinearity=torch.nn.Tanh())
    O.reinit_(oU, 21)
    oU.train()
    G = dcgan.DCGANUpGenerator(2048, 32, 3, 64, nonlinearity=torch.nn.LeakyReLU(0.2),
                               last_nonlinearity=torch.nn.Tanh()).to(cuda_dev)
    G.load_state_dict(oU.state_dict())
    G.train()
    z = torch.randn(6, 2048, generator=torch.Generator().manual_seed(22))
    with torch.no_grad():
        ref = oU(z)
    out = G(z.to(cuda_dev))
    assert out.shape == ref.shape
    assert _rel(out, ref) <= 3e-2
    got = U.sample_of(out, 4096)
    assert U.rel_l2(got, gold["upgen/out_sample"]) <= 3e-2


@pytest.mark.parametrize("size,B", [(32, 8), (64, 4)])
def test_up_generator_g_step_matches_oracle(cuda_dev, size, B):
    """G step (src/wgan_loss.py:82-129) with the resize-conv generator: loss and generator gradients vs the oracle."""
    import copy
    import os
    import tempfile
    from torch.optim import Adam
    from rnagan_b200 import dcgan, wgan_loss
    feats = 128
    lrelu, tanh = torch.nn.LeakyReLU(0.2), torch.nn.Tanh()
    oG = O.OracleUpGenerator(2048, size, 3, 64, nonlinearity=lrelu, last_nonlinearity=tanh).train()
    oD = O.OracleCritic(size, 3, 64, nonlinearity=lrelu, last_nonlinearity=lrelu).train()
    oV = O.OracleVAE(feats, beta=0.005).eval()
    O.reinit_(oV, 13); O.reinit_(oG, 31); O.reinit_(oD, 12)
    ckpt = os.path.join(tempfile.mkdtemp(), "vae.pt")
    torch.save(oV.state_dict(), ckpt)
    G = dcgan.DCGANUpGenerator(2048, size, 3, 64, nonlinearity=torch.nn.LeakyReLU(0.2),
                               last_nonlinearity=torch.nn.Tanh()).to(cuda_dev).train()
    D = dcgan.DCGANDiscriminator(size, 3, 64, nonlinearity=torch.nn.LeakyReLU(0.2),
                                 last_nonlinearity=torch.nn.LeakyReLU(0.2)).to(cuda_dev).train()
    G.load_state_dict(oG.state_dict()); D.load_state_dict(oD.state_dict())
    loss = wgan_loss.WassersteinGeneratorLossVAE(ckpt, feats)
    data = O.make_batch(B, feats, size, 14)
    # torch's own bf16 deviation as the yardstick
    aG, aD = copy.deepcopy(oG), copy.deepcopy(oD)
    torch.manual_seed(99)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        O.g_step(aG, aD, Adam(aG.parameters(), lr=0.0), oV, data)
    torch.manual_seed(99)
    v_ref = O.g_step(oG, oD, Adam(oG.parameters(), lr=0.0), oV, data)
    torch.manual_seed(99)
    opt = Adam(G.parameters(), lr=0.0)
    v = loss.train_ops(G, D, opt, cuda_dev, B, data)
    assert abs(v - v_ref) <= 0.02 + 0.02 * abs(v_ref)
    c_bf16, pairs = 1.0, []
    for (n, po), (_, pm), (_, pa) in zip(oG.named_parameters(), G.named_parameters(), aG.named_parameters()):
        a, b, c = pm.grad.double().flatten().cpu(), po.grad.double().flatten(), pa.grad.double().flatten()
        cos = lambda u, w: (u @ w / (u.norm() * w.norm()).clamp_min(1e-30)).item()
        if b.norm() < 1e-7 * max(1.0, b.numel() ** 0.5):      # conv biases feeding BatchNorm: exact gradient is 0
            continue
        c_bf16 = min(c_bf16, cos(c, b))
        pairs.append((n, cos(a, b)))
    bound = 1.0 - 2.0 * (1.0 - c_bf16) - 0.01
    for n, cs in pairs:
        assert cs >= bound, f"{n}: cosine {cs:.4f} < {bound:.4f} (torch-bf16 worst {c_bf16:.4f})"


def test_tiles_u8_normalise_bit_exact_and_prefetcher(cuda_dev):
    """Data path: rg_tiles_u8_to_nchw equals the reference's CPU transforms bit for bit
    (cv2 BGR->RGB, permute(2,0,1), ConvertImageDtype(float) = x / 255, Normalize(0.5, 0.5) = (x - 0.5) / 0.5;
    src/read_data.py:341-343, src/histopathology_gan.py:106-109), and DevicePrefetcher yields device batches."""
    from rnagan_b200 import data as D
    g = torch.Generator().manual_seed(3)
    tiles = torch.randint(0, 256, (5, 64, 64, 3), dtype=torch.uint8, generator=g)
    tiles[0, 0, 0] = torch.tensor([0, 128, 255], dtype=torch.uint8)

    def cpu_transform(t, bgr):
        t = t.flip(-1) if bgr else t                       # BGR -> RGB swaps channels 0 and 2
        x = t.permute(0, 3, 1, 2).to(torch.float32) / 255  # ConvertImageDtype(torch.float)
        return (x - 0.5) / 0.5                             # Normalize(mean=0.5, std=0.5)

    for bgr in (True, False):
        got = D.normalise_tiles(tiles.to(cuda_dev), bgr=bgr)
        ref = cpu_transform(tiles, bgr)
        assert got.shape == (5, 3, 64, 64) and torch.equal(got.cpu(), ref)
    rna = torch.randn(5, 40, generator=g)
    loader = [{"image": tiles[:3], "rna_data": rna[:3], "labels": torch.zeros(3)},
              {"image": tiles[3:], "rna_data": rna[3:], "labels": torch.zeros(2)}]
    seen = list(D.DevicePrefetcher(loader, cuda_dev, bgr=True))
    assert len(seen) == 2 and seen[0]["image"].device.type == "cuda" and seen[0]["image"].dtype == torch.float32
    assert torch.equal(torch.cat([b["image"] for b in seen]).cpu(), cpu_transform(tiles, True))
    assert torch.equal(torch.cat([b["rna_data"] for b in seen]).cpu(), rna)


def test_autograd_through_modules(cuda_dev):
    """`forward` of the drop-in modules is differentiable once through torch.autograd (a caller's own train step, e.g.
    override_train_ops): loss.backward() through D(G(z)) and through two critic passes of one graph gives the oracle's
    parameter / input gradients within the bf16 yardstick, gradients accumulate across backward calls like autograd's,
    and a double backward (create_graph=True) is refused."""
    import copy
    from rnagan_b200 import dcgan
    size, B = 32, 8
    lrelu, tanh = torch.nn.LeakyReLU(0.2), torch.nn.Tanh()
    oG = O.OracleGenerator(2048, size, 3, 64, nonlinearity=lrelu, last_nonlinearity=tanh).train()
    oD = O.OracleCritic(size, 3, 64, nonlinearity=lrelu, last_nonlinearity=lrelu).train()
    O.reinit_(oG, 1); O.reinit_(oD, 2)
    G = dcgan.DCGANGenerator(2048, size, 3, 64, nonlinearity=torch.nn.LeakyReLU(0.2),
                             last_nonlinearity=torch.nn.Tanh()).to(cuda_dev).train()
    D = dcgan.DCGANDiscriminator(size, 3, 64, nonlinearity=torch.nn.LeakyReLU(0.2),
                                 last_nonlinearity=torch.nn.LeakyReLU(0.2)).to(cuda_dev).train()
    G.load_state_dict(oG.state_dict()); D.load_state_dict(oD.state_dict())
    g = torch.Generator().manual_seed(3)
    z = torch.randn(B, 2048, generator=g)
    x = torch.rand(B, 3, size, size, generator=g) * 2 - 1

    def cos(a, b):
        a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
        return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()

    # (1) generator loss through both networks, gradient w.r.t. the latent too
    aG, aD = copy.deepcopy(oG), copy.deepcopy(oD)
    zr = z.clone().requires_grad_()
    (-oD(oG(zr))).mean().backward()
    za = z.clone().requires_grad_()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        (-aD(aG(za))).mean().backward()
    zc = z.clone().to(cuda_dev).requires_grad_()
    loss = (-D(G(zc))).mean()
    assert loss.requires_grad
    loss.backward()
    worst_bf16 = min(cos(pa.grad, po.grad) for pa, po in zip(aG.parameters(), oG.parameters()) if po.grad.norm() > 0)
    bound = 1.0 - 2.0 * (1.0 - min(worst_bf16, cos(za.grad, zr.grad))) - 0.01
    for (n, po), (_, pm) in zip(oG.named_parameters(), G.named_parameters()):
        if po.grad.norm() > 0:
            assert cos(pm.grad, po.grad) >= bound, n
    assert cos(zc.grad, zr.grad) >= bound
    for (n, po), (_, pm) in zip(oD.named_parameters(), D.named_parameters()):
        if po.grad.norm() > 0:
            assert cos(pm.grad, po.grad) >= bound, n
    # (2) two critic passes in ONE graph, accumulated ON TOP of the gradients already there (no zero_grad)
    with torch.no_grad():
        fake_o = oG(z)
    fake_m = G(z.to(cuda_dev)).detach()
    (oD(fake_o).mean() - oD(x).mean()).backward()
    (D(fake_m).mean() - D(x.to(cuda_dev)).mean()).backward()
    for (n, po), (_, pm) in zip(oD.named_parameters(), D.named_parameters()):
        if po.grad.norm() > 0:
            assert cos(pm.grad, po.grad) >= bound, n
            assert abs(pm.grad.norm().item() / po.grad.norm().item() - 1.0) <= 0.1, n
    # (3) double backward is refused with a clear error
    xg = x.to(cuda_dev).requires_grad_()
    out = D(xg)
    with pytest.raises(RuntimeError):
        gx = torch.autograd.grad(out, xg, torch.ones_like(out), create_graph=True)[0]
        (gx.norm() - 1).pow(2).backward()
    # no_grad / eval paths still return plain tensors
    with torch.no_grad():
        assert not G(z.to(cuda_dev)).requires_grad


def test_train_betavae_loop_val_phase_and_best_checkpoint(cuda_dev, tmp_path):
    """train_betaVAE / evaluate_betaVAE with the reference's signatures (src/betaVAE.py:164-331): train + val phases,
    `model_dict_best.pt` at the best validation epoch, `model_last.pt`, best weights loaded back, warm-up + cosine
    schedule (src/betaVAE_training.py:165-166) driving the fused Adam through param_groups['lr']."""
    from rnagan_b200 import betaVAE as bv
    feats, beta = 300, 0.0005
    torch.manual_seed(3)
    vae = bv.betaVAE(feats, 2048, [6000, 4000, 2048], [4000, 6000], beta=beta).to(cuda_dev)
    g = torch.Generator().manual_seed(9)
    basis = torch.randn(8, feats, generator=g)

    def loader(n):
        return [{"rna_data": torch.tanh(torch.randn(32, 8, generator=g) @ basis * 0.3)} for _ in range(n)]

    loaders = {"train": loader(6), "val": loader(2)}
    opt = torch.optim.Adam(vae.parameters(), lr=2e-4)
    sched = bv.GradualWarmupScheduler(opt, multiplier=1, total_epoch=4,
                                      after_scheduler=torch.optim.lr_scheduler.CosineAnnealingLR(opt, 50))
    model, res = bv.train_betaVAE(vae, opt, loaders, save_dir=str(tmp_path), num_epochs=3, scheduler=sched, verbose=False)
    assert model is vae and set(res) == {"best_epoch", "best_loss"} and 0 <= res["best_epoch"] <= 2
    assert os.path.exists(tmp_path / "model_dict_best.pt") and os.path.exists(tmp_path / "model_last.pt")
    best = torch.load(tmp_path / "model_dict_best.pt")
    for k, v in vae.state_dict().items():
        assert torch.equal(v.cpu(), best[k].cpu()), k                   # the best weights were loaded back
    assert opt.param_groups[0]["lr"] < 2e-4                              # the schedule ran (18 steps: past the warm-up)
    test_loss, preds, real = bv.evaluate_betaVAE(vae, loaders["val"], verbose=False)
    # (the eval forward still draws reparametrisation noise, like the reference's: equal only statistically)
    assert abs(test_loss["total_loss"] - res["best_loss"]["total_loss"]) <= 0.15 * res["best_loss"]["total_loss"] + 1e-3
    assert test_loss["total_loss"] == test_loss["reconstruction_loss"]    # eval: total = reconstruction
    assert len(preds) == 2 and len(preds[0]) == 32 and len(preds[0][0]) == feats and len(real[1]) == 32
    assert not vae.training


def test_gradient_penalty_forward_value(cuda_dev):
    """WassersteinGradientPenalty[VAE].forward(interpolate, d_interpolate) (src/wgan_loss.py:296-312, 32-44): the penalty
    value through first-order autograd of the critic module equals the oracle's; differentiating it again is refused."""
    from rnagan_b200 import dcgan, wgan_loss
    size, B = 32, 8
    lrelu = torch.nn.LeakyReLU(0.2)
    oD = O.OracleCritic(size, 3, 64, nonlinearity=lrelu, last_nonlinearity=lrelu).train()
    O.reinit_(oD, 2)
    D = dcgan.DCGANDiscriminator(size, 3, 64, nonlinearity=torch.nn.LeakyReLU(0.2),
                                 last_nonlinearity=torch.nn.LeakyReLU(0.2)).to(cuda_dev).train()
    D.load_state_dict(oD.state_dict())
    # a head pre-activation next to zero flips the LeakyReLU' of the whole sample between bf16 and fp32 (slope 0.2 vs 1):
    # a legitimate kink, not an error -- take the first seeded batch whose critic outputs stay clear of it
    for seed in range(5, 40):
        x = torch.rand(B, 3, size, size, generator=torch.Generator().manual_seed(seed)) * 2 - 1
        with torch.no_grad():
            if oD(x).abs().min().item() > 0.05:
                break
    xo = x.clone().requires_grad_()
    ref = O.gradient_penalty(xo, oD(xo)).item()
    xm = x.to(cuda_dev).requires_grad_()
    pen = wgan_loss.WassersteinGradientPenalty().forward(xm, D(xm))
    assert abs(pen.item() - ref) <= 0.02 + 0.03 * abs(ref), (pen.item(), ref)
    with pytest.raises(RuntimeError):
        pen.backward()
