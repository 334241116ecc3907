"""GPU parity of the three train_ops (G step, critic step, gradient-penalty step), of the post-step weights /
BatchNorm running statistics and of tile synthesis against the CPU oracle on identical seeds and inputs, through the
reference-facing API (rnagan_b200.trainer.Trainer -> wgan_loss.*.train_ops -> C ABI).

Tolerances (bf16 operands, fp32 accumulation; stated per BASELINE.json north_star):
  * every step is started from EXACTLY the oracle's state (weights, BN buffers, Adam moments), so steps are compared
    in isolation and bf16 noise is not compounded by trajectory divergence;
  * losses: |cuda - oracle| <= 0.02 + 0.02*|oracle|, and within the same bound of the reference's own golden values;
  * per-parameter gradients: cosine >= 1 - 2*(1 - c_bf16) - 0.01 where c_bf16 is the worst cosine torch's OWN bf16
    autocast reaches against fp32 on the same step (computed live) -- i.e. at most twice torch-bf16's deviation;
  * post-step weights: rel-L2 <= 2e-2 (Adam's first steps are sign-like: lr-sized flips where a gradient is ~0);
  * BatchNorm running statistics: rel-L2 <= 5e-2; num_batches_tracked exact;
  * synthesized tiles (in [0,1]): rel-L2 <= 3e-2, max abs <= 0.12.
"""
import copy
import os
import tempfile

import numpy as np
import pytest
import torch
from torch.optim import Adam

from oracle import ref_oracle as O
from tests import _util as U

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()


def _build(size, feats, dev):
    from rnagan_b200 import dcgan, wgan_loss
    from rnagan_b200.trainer import Trainer
    lrelu, tanh = torch.nn.LeakyReLU(0.2), torch.nn.Tanh()
    oG = O.OracleGenerator(2048, size, 3, 64, nonlinearity=lrelu, last_nonlinearity=tanh).train()
    oD = O.OracleCritic(size, 3, 64, nonlinearity=lrelu, last_nonlinearity=lrelu).train()
    oV = O.OracleVAE(feats, beta=0.005).eval()
    O.reinit_(oV, U.SEED_V); O.reinit_(oG, U.SEED_G); O.reinit_(oD, U.SEED_D)
    ckpt = os.path.join(tempfile.mkdtemp(), "vae.pt")
    torch.save(oV.state_dict(), ckpt)
    net = {
        "generator": {"name": dcgan.DCGANGenerator,
                      "args": {"encoding_dims": 2048, "out_channels": 3, "step_channels": 64, "out_size": size,
                               "nonlinearity": torch.nn.LeakyReLU(0.2), "last_nonlinearity": torch.nn.Tanh()},
                      "optimizer": {"name": Adam, "args": {"lr": 0.0001, "betas": (0.5, 0.999)}}},
        "discriminator": {"name": dcgan.DCGANDiscriminator,
                          "args": {"in_size": size, "in_channels": 3, "step_channels": 64,
                                   "nonlinearity": torch.nn.LeakyReLU(0.2),
                                   "last_nonlinearity": torch.nn.LeakyReLU(0.2)},
                          "optimizer": {"name": Adam, "args": {"lr": 0.0004, "betas": (0.5, 0.999)}}},
    }
    losses = [wgan_loss.WassersteinGeneratorLossVAE(ckpt, feats), wgan_loss.WassersteinDiscriminatorLossVAE(ckpt, feats),
              wgan_loss.WassersteinGradientPenaltyVAE(ckpt, feats)]
    tr = Trainer(net, losses, device=dev, sample_size=64, epochs=1, devices=[0])
    tr.generator.train(); tr.discriminator.train()
    return oG, oD, oV, tr


def _bf16_yardstick(which, oG, oD, oV, data):
    st = torch.get_rng_state()
    cG, cD = copy.deepcopy(oG), copy.deepcopy(oD)
    cg = Adam(cG.parameters(), lr=1e-4, betas=(0.5, 0.999))
    cd = Adam(cD.parameters(), lr=4e-4, betas=(0.5, 0.999))
    with torch.autocast("cpu", dtype=torch.bfloat16):
        if which == 0:
            O.g_step(cG, cD, cg, oV, data)
        elif which == 1:
            O.critic_step(cG, cD, cd, oV, data)
        else:
            O.gp_step(cG, cD, cd, oV, data)
    torch.set_rng_state(st)
    return cG if which == 0 else cD


def _check_grads(tag, onet, mnet, anet):
    c_bf16 = 1.0
    pairs = []
    for (n, po), (_, pm), (_, pa) in zip(onet.named_parameters(), mnet.named_parameters(), anet.named_parameters()):
        if po.grad is None or po.grad.norm() == 0:
            assert pm.grad is None or pm.grad.abs().max().item() == 0.0, f"{tag} {n}: expected a zero gradient"
            continue
        c_bf16 = min(c_bf16, _cos(pa.grad, po.grad))
        pairs.append((n, _cos(pm.grad, po.grad)))
    bound = 1.0 - 2.0 * (1.0 - c_bf16) - 0.01
    for n, c in pairs:
        assert c >= bound, f"{tag} {n}: cosine {c:.4f} < {bound:.4f} (torch-bf16 worst {c_bf16:.4f})"


def _sync(tr, oG, oD, og, od):
    tr.generator.load_state_dict(oG.state_dict())
    tr.discriminator.load_state_dict(oD.state_dict())
    if len(og.state_dict()["state"]):
        tr.optimizer_generator.load_state_dict(og.state_dict())
    if len(od.state_dict()["state"]):
        tr.optimizer_discriminator.load_state_dict(od.state_dict())


def _check_state(tag, onet, mnet):
    for (n, po), (_, pm) in zip(onet.named_parameters(), mnet.named_parameters()):
        assert _rel(pm.data, po.data) <= 2e-2, f"{tag} weight {n}"
    for (n, bo), (_, bm) in zip(onet.named_buffers(), mnet.named_buffers()):
        if n.endswith("num_batches_tracked"):
            assert int(bo) == int(bm), f"{tag} {n}"
        else:
            assert _rel(bm.float(), bo.float()) <= 5e-2, f"{tag} buffer {n}"


@pytest.mark.parametrize("cfg", ["mini32", "mini64", "full256"])
def test_train_steps_match_oracle(cuda_dev, cfg):
    size, batch, feats, iters = U.CONFIGS[cfg]
    gold = U.load_golden(f"gan_{cfg}.npz")
    torch.set_num_threads(os.cpu_count() or 8)
    oG, oD, oV, tr = _build(size, feats, cuda_dev)
    data = O.make_batch(batch, feats, size, U.SEED_BATCH)
    tr.real_inputs, tr.batch_size = data, batch
    og = Adam(oG.parameters(), lr=1e-4, betas=(0.5, 0.999))
    od = Adam(oD.parameters(), lr=4e-4, betas=(0.5, 0.999))
    names = list(tr.losses.keys())
    steps = [(0, O.g_step, og, oG, "generator"), (1, O.critic_step, od, oD, "discriminator"),
             (2, O.gp_step, od, oD, "discriminator")]
    torch.manual_seed(U.SEED_RUN)
    for it in range(iters):
        for which, fn, opt, onet, mname in steps:
            ac = _bf16_yardstick(which, oG, oD, oV, data)
            _sync(tr, oG, oD, og, od)
            st = torch.get_rng_state()
            v_ref = fn(oG, oD, opt, oV, data)
            st_after = torch.get_rng_state()
            torch.set_rng_state(st)
            v = tr._call(names[which])
            # the product consumes the CPU RNG stream exactly like the reference (noise, then eps)
            assert torch.equal(torch.get_rng_state(), st_after)
            assert abs(v - v_ref) <= 0.02 + 0.02 * abs(v_ref), f"it{it} step{which}: {v} vs oracle {v_ref}"
            g_ref = float(gold["losses"][it][which])
            assert abs(v - g_ref) <= 0.02 + 0.02 * abs(g_ref), f"it{it} step{which}: {v} vs golden {g_ref}"
            _check_grads(f"it{it} step{which}", onet, getattr(tr, mname), ac)
            _check_state(f"it{it} step{which}", onet, getattr(tr, mname))
    _check_state("final G", oG, tr.generator)
    _check_state("final D", oD, tr.discriminator)


def test_synthesis_matches_oracle_and_golden(cuda_dev):
    from rnagan_b200 import gan_utils
    size, batch, feats, _ = U.CONFIGS["mini32"]
    gold = U.load_golden("gan_mini32.npz")
    oG, oD, oV, tr = _build(size, feats, cuda_dev)
    tr.generator.load_state_dict(oG.state_dict())
    vae = tr.losses["WassersteinGeneratorLossVAE"].betavae
    data = O.make_batch(batch, feats, size, U.SEED_BATCH)
    n = 64
    # (a) one profile, chunks of 10, train-mode BN -- the reference's generate_images semantics
    torch.manual_seed(5)
    ref = O.synth_tiles(oG, oV, data["rna_data"][:1], n)
    torch.manual_seed(5)
    got = gan_utils.generate_images(tr, gene_exp=data["rna_data"][:1], sample_size=n, betavae=vae)
    assert got.shape == (n, size, size, 3) and got.dtype == np.float32
    assert U.rel_l2(got, ref) <= 3e-2 and np.abs(got - ref).max() <= 0.12
    assert got.min() >= 0.0 and got.max() <= 1.0
    # (b) one profile per row (conditioning does not cancel)
    profiles = torch.randn(n, feats, generator=torch.Generator().manual_seed(U.SEED_BATCH + 1))
    oG2 = copy.deepcopy(oG)
    torch.manual_seed(6)
    ref2 = O.synth_tiles(oG2, oV, profiles, n)
    tr.generator.load_state_dict(oG.state_dict())
    torch.manual_seed(6)
    got2 = gan_utils.generate_images(tr, gene_exp=profiles, sample_size=n, betavae=vae)
    assert U.rel_l2(got2, ref2) <= 3e-2 and np.abs(got2 - ref2).max() <= 0.12


def test_synthesis_after_golden_training_state(cuda_dev):
    """Tiles from the reference's own generate_images (golden) after its 2 training iterations: replay the oracle to
    that state, load it into the CUDA generator and synthesize with the same seed."""
    from rnagan_b200 import gan_utils
    size, batch, feats, iters = U.CONFIGS["mini32"]
    gold = U.load_golden("gan_mini32.npz")
    oG, oD, oV, tr = _build(size, feats, cuda_dev)
    data = O.make_batch(batch, feats, size, U.SEED_BATCH)
    og = Adam(oG.parameters(), lr=1e-4, betas=(0.5, 0.999))
    od = Adam(oD.parameters(), lr=4e-4, betas=(0.5, 0.999))
    torch.manual_seed(U.SEED_RUN)
    for _ in range(iters):
        O.train_iter(oG, oD, og, od, oV, data)
    tr.generator.load_state_dict(oG.state_dict())
    vae = tr.losses["WassersteinGeneratorLossVAE"].betavae
    n_syn = int(gold["tiles/n"][0])
    torch.manual_seed(U.SEED_RUN + 1)
    got = gan_utils.generate_images(tr, gene_exp=data["rna_data"][:1], sample_size=n_syn, betavae=vae)
    flat = np.ascontiguousarray(got.transpose(0, 3, 1, 2)).astype(np.float64).reshape(-1)
    sample = flat[U.sample_idx(flat.size, 4096)]
    assert U.rel_l2(sample, gold["tiles/sample"]) <= 3e-2
    assert np.abs(sample - gold["tiles/sample"]).max() <= 0.12


def test_plain_wgan_steps_match_oracle(cuda_dev):
    """The un-conditioned `wgan` losses (torchgan WassersteinGeneratorLoss / WassersteinDiscriminatorLoss(clip) /
    WassersteinGradientPenalty, src/histopathology_gan.py:267-272) through Trainer.train_iter vs the oracle's
    restatement; the device-RNG noise of each step is reproduced by re-seeding the CUDA generator."""
    from rnagan_b200 import dcgan, wgan_loss
    from rnagan_b200.trainer import Trainer
    size, batch = 32, 8
    lrelu, tanh = torch.nn.LeakyReLU(0.2), torch.nn.Tanh()
    oG = O.OracleGenerator(2048, size, 3, 64, nonlinearity=lrelu, last_nonlinearity=tanh).train()
    oD = O.OracleCritic(size, 3, 64, nonlinearity=lrelu, last_nonlinearity=lrelu).train()
    O.reinit_(oG, U.SEED_G); O.reinit_(oD, U.SEED_D)
    net = {
        "generator": {"name": dcgan.DCGANGenerator,
                      "args": {"encoding_dims": 2048, "out_channels": 3, "step_channels": 64, "out_size": size,
                               "nonlinearity": torch.nn.LeakyReLU(0.2), "last_nonlinearity": torch.nn.Tanh()},
                      "optimizer": {"name": Adam, "args": {"lr": 0.0001, "betas": (0.5, 0.999)}}},
        "discriminator": {"name": dcgan.DCGANDiscriminator,
                          "args": {"in_size": size, "in_channels": 3, "step_channels": 64,
                                   "nonlinearity": torch.nn.LeakyReLU(0.2),
                                   "last_nonlinearity": torch.nn.LeakyReLU(0.2)},
                          "optimizer": {"name": Adam, "args": {"lr": 0.0004, "betas": (0.5, 0.999)}}},
    }
    clip = (-0.05, 0.05)
    losses = [wgan_loss.WassersteinGeneratorLoss(), wgan_loss.WassersteinDiscriminatorLoss(clip=clip),
              wgan_loss.WassersteinGradientPenalty()]
    tr = Trainer(net, losses, device=cuda_dev, sample_size=64, epochs=1, devices=[0])
    tr.generator.train(); tr.discriminator.train()
    real = O.make_batch(batch, 64, size, U.SEED_BATCH)["image"]
    tr.real_inputs, tr.batch_size = real, batch
    og = Adam(oG.parameters(), lr=1e-4, betas=(0.5, 0.999))
    od = Adam(oD.parameters(), lr=4e-4, betas=(0.5, 0.999))
    names = list(tr.losses.keys())
    torch.manual_seed(U.SEED_RUN)
    for which in range(3):
        _sync(tr, oG, oD, og, od)
        torch.cuda.manual_seed(100 + which)
        noise = torch.randn(batch, 2048, device=cuda_dev).cpu()      # what the product will draw
        st = torch.get_rng_state()
        if which == 0:
            v_ref = O.plain_g_step(oG, oD, og, noise)
        elif which == 1:
            v_ref = O.plain_critic_step(oG, oD, od, noise, real, clip=clip)
        else:
            v_ref = O.plain_gp_step(oG, oD, od, noise, real)
        st_after = torch.get_rng_state()
        torch.set_rng_state(st)
        torch.cuda.manual_seed(100 + which)
        v = tr._call(names[which])
        v = float(v.item()) if torch.is_tensor(v) else v
        assert torch.equal(torch.get_rng_state(), st_after)          # CPU stream: only the GP step's eps
        assert abs(v - v_ref) <= 0.02 + 0.02 * abs(v_ref), f"plain step{which}: {v} vs oracle {v_ref}"
        onet, mnet = (oG, tr.generator) if which == 0 else (oD, tr.discriminator)
        _check_state(f"plain step{which}", onet, mnet)
    # the clamp ran before the critic step: every critic weight of the oracle was inside the clip range at that point
    assert all(float(p.abs().max()) <= 0.06 for p in oD.parameters())


def test_training_is_bitwise_deterministic(cuda_dev):
    """cudnn.deterministic=True in the reference (src/histopathology_gan.py:289): every reduction on the sm_100a path
    (split-K, fused BatchNorm statistics, column sums, gradient norm) has a fixed order, so two runs of the same two
    iterations from the same state and seeds give bit-identical losses, weights, BatchNorm buffers and Adam moments."""
    size, batch, feats, _ = U.CONFIGS["mini32"]
    data = O.make_batch(batch, feats, size, U.SEED_BATCH)
    results = []
    for _ in range(2):
        oG, oD, oV, tr = _build(size, feats, cuda_dev)
        tr.generator.load_state_dict(oG.state_dict())
        tr.discriminator.load_state_dict(oD.state_dict())
        tr.real_inputs, tr.batch_size = data, batch
        torch.manual_seed(U.SEED_RUN)
        losses = [tuple(tr.train_iter().values()) for _ in range(2)]
        torch.cuda.synchronize()
        state = {f"G.{k}": v.detach().clone().cpu() for k, v in tr.generator.state_dict().items()}
        state.update({f"D.{k}": v.detach().clone().cpu() for k, v in tr.discriminator.state_dict().items()})
        for name, opt in (("og", tr.optimizer_generator), ("od", tr.optimizer_discriminator)):
            for i, st in enumerate(opt.state.values()):
                state[f"{name}.{i}.m"] = st["exp_avg"].detach().clone().cpu()
                state[f"{name}.{i}.v"] = st["exp_avg_sq"].detach().clone().cpu()
        results.append((losses, state))
    (l0, s0), (l1, s1) = results
    assert l0 == l1
    assert s0.keys() == s1.keys()
    for k in s0:
        assert torch.equal(s0[k], s1[k]), f"{k} differs between two identical runs"


def test_uint8_tiles_and_synthesis_job(cuda_dev):
    """uint8 tile output (what src/generate_tissue_images.py:127-129 computes on the host: `img *= 255;
    img.astype(np.uint8)`, then RGB->BGR for cv2.imwrite) written by the generator's last kernel is BIT-EXACT against
    that host arithmetic applied to the fp32 tiles of the same call, and the streaming job (gan_utils.synthesize_job,
    BASELINE config 4) hands every tile of its range to the sink exactly once, in order, ragged last batch included."""
    from rnagan_b200 import gan_utils
    size, batch, feats, _ = U.CONFIGS["mini32"]
    oG, oD, oV, tr = _build(size, feats, cuda_dev)
    tr.generator.load_state_dict(oG.state_dict())
    vae = tr.losses["WassersteinGeneratorLossVAE"]._encoder(cuda_dev)
    profiles = torch.randn(7, feats, generator=torch.Generator().manual_seed(3))
    n = 48
    rows = profiles[torch.arange(n) % 7]
    torch.manual_seed(21)
    f32 = gan_utils.generate_tiles(tr.generator, vae, rows, n, chunk=n, device=cuda_dev).cpu().numpy()
    for bgr in (False, True):
        torch.manual_seed(21)
        u8 = gan_utils.generate_tiles(tr.generator, vae, rows, n, chunk=n, device=cuda_dev, u8=True, bgr=bgr)
        assert u8.dtype == torch.uint8 and u8.shape == (n, size, size, 3)
        want = (f32 * np.float32(255)).astype(np.uint8)
        if bgr:
            want = want[..., ::-1]
        assert np.array_equal(u8.cpu().numpy(), want)
    # streaming job: tiles [5, 5 + 100) in batches of 32 (ragged tail of 4), profile row t % 7
    got = {}

    def sink(t0, tiles):
        assert tiles.dtype == np.uint8 and tiles.shape[1:] == (size, size, 3)
        got[t0] = tiles.copy()

    torch.manual_seed(33)
    done = gan_utils.synthesize_job(tr.generator, vae, profiles, 5, 105, batch=32, sink=sink)
    assert done == 100 and sorted(got) == [5, 37, 69, 101] and [got[k].shape[0] for k in sorted(got)] == [32, 32, 32, 4]
    # same RNG stream, same batches through generate_tiles
    torch.manual_seed(33)
    for t0 in sorted(got):
        nb = got[t0].shape[0]
        idx = torch.arange(t0, t0 + nb) % 7
        ref = gan_utils.generate_tiles(tr.generator, vae, profiles[idx], nb, chunk=nb, device=cuda_dev, u8=True)
        assert np.array_equal(ref.cpu().numpy(), got[t0]), t0


def test_cuda_graph_replay_is_bit_identical_to_eager(cuda_dev):
    """Single-process steps are captured into CUDA graphs after two eager calls (steps._run_step).  A replay launches the
    same kernels in the same order on the same buffers, Adam reads its step count from device memory: six iterations with
    graphs must be BIT-identical to six eager iterations (losses, weights, BatchNorm buffers, Adam moments and step
    counters), the graphs must actually have been replayed, and an optimizer.load_state_dict in between (new moment
    buffers) must trigger a fresh capture instead of replaying into stale memory."""
    from rnagan_b200 import steps
    size, batch, feats, _ = U.CONFIGS["mini32"]
    data = O.make_batch(batch, feats, size, U.SEED_BATCH)
    results = []
    for use in (False, True):
        steps.use_graphs(use)
        steps._GRAPHS.clear()
        replays0 = steps.GRAPH_STATS["replays"]
        try:
            oG, oD, oV, tr = _build(size, feats, cuda_dev)
            tr.generator.load_state_dict(oG.state_dict())
            tr.discriminator.load_state_dict(oD.state_dict())
            tr.real_inputs, tr.batch_size = data, batch
            torch.manual_seed(U.SEED_RUN)
            losses = [tuple(tr.train_iter().values()) for _ in range(4)]
            # swap the optimizer state for an equal copy in new buffers, then keep going
            for opt in (tr.optimizer_generator, tr.optimizer_discriminator):
                opt.load_state_dict(copy.deepcopy(opt.state_dict()))
            losses += [tuple(tr.train_iter().values()) for _ in range(4)]
            torch.cuda.synchronize()
            state = {f"G.{k}": v.detach().clone().cpu() for k, v in tr.generator.state_dict().items()}
            state.update({f"D.{k}": v.detach().clone().cpu() for k, v in tr.discriminator.state_dict().items()})
            for name, opt in (("og", tr.optimizer_generator), ("od", tr.optimizer_discriminator)):
                for i, st in enumerate(opt.state.values()):
                    state[f"{name}.{i}.m"] = st["exp_avg"].detach().clone().cpu()
                    state[f"{name}.{i}.v"] = st["exp_avg_sq"].detach().clone().cpu()
                    state[f"{name}.{i}.step"] = torch.tensor(float(st["step"]))
            results.append((losses, state, steps.GRAPH_STATS["replays"] - replays0))
        finally:
            steps.use_graphs(True)
    (l0, s0, r0), (l1, s1, r1) = results
    assert r0 == 0 and r1 >= 6, (r0, r1)          # 3 step kinds x (2 + 2 replays after each capture) at least
    assert steps.GRAPH_STATS["failed"] == 0
    assert l0 == l1
    for k in s0:
        assert torch.equal(s0[k], s1[k]), f"{k} differs between eager and graph replay"
