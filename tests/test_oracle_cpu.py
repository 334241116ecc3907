"""The oracle (oracle/ref_oracle.py, a CPU restatement) must reproduce the outputs of the reference's OWN modules
stored in tests/golden/ by oracle/make_golden.py: losses of every train_ops call, per-parameter gradients after
each of the three steps, post-step weights / BatchNorm running statistics, and synthesized tiles.
fp32 CPU arithmetic, same op order -> tolerance 1e-5 relative (bit-identical on the generating machine)."""
import numpy as np
import pytest
import torch
from torch.optim import Adam

from oracle import ref_oracle as O
from tests import _util as U

TOL = 1e-5


def _build(size, feats):
    G = O.OracleGenerator(encoding_dims=2048, out_size=size, out_channels=3, step_channels=64,
                          nonlinearity=torch.nn.LeakyReLU(0.2), last_nonlinearity=torch.nn.Tanh())
    D = O.OracleCritic(in_size=size, in_channels=3, step_channels=64, nonlinearity=torch.nn.LeakyReLU(0.2),
                       last_nonlinearity=torch.nn.LeakyReLU(0.2))
    vae = O.OracleVAE(feats, beta=0.005)
    O.reinit_(vae, U.SEED_V)
    vae.eval()
    O.reinit_(G, U.SEED_G)
    O.reinit_(D, U.SEED_D)
    G.train()
    D.train()
    return G, D, vae


def _cmp_named(gold, prefix, named, grads=False):
    n = 0
    for name, p in named:
        t = p.grad if grads else p
        if t is None:
            continue
        key = f"{prefix}/{name}"
        ref = gold[key + "/sample"]
        got = U.sample_of(t)
        scale = max(np.abs(ref).max(), 1e-12)
        assert np.abs(got - ref).max() <= TOL * scale + 1e-9, key
        assert abs(U.stats_of(t)[1] - gold[key + "/l2"]) <= 1e-4 * max(gold[key + "/l2"], 1e-9), key
        n += 1
    assert n > 0


@pytest.mark.parametrize("cfg", ["mini32", "mini64", "full256"])
def test_oracle_matches_reference_training(cfg):
    size, batch, feats, iters = U.CONFIGS[cfg]
    gold = U.load_golden(f"gan_{cfg}.npz")
    torch.set_num_threads(8)
    G, D, vae = _build(size, feats)
    opt_g = Adam(G.parameters(), lr=1e-4, betas=(0.5, 0.999))
    opt_d = Adam(D.parameters(), lr=4e-4, betas=(0.5, 0.999))
    data = O.make_batch(batch, feats, size, U.SEED_BATCH)
    torch.manual_seed(U.SEED_RUN)
    for it in range(iters):
        vg = O.g_step(G, D, opt_g, vae, data)
        _cmp_named(gold, f"it{it}/g_step/grad_G", G.named_parameters(), grads=True)
        vd = O.critic_step(G, D, opt_d, vae, data)
        _cmp_named(gold, f"it{it}/critic_step/grad_D", D.named_parameters(), grads=True)
        vp = O.gp_step(G, D, opt_d, vae, data)
        _cmp_named(gold, f"it{it}/gp_step/grad_D", D.named_parameters(), grads=True)
        np.testing.assert_allclose([vg, vd, vp], gold["losses"][it], rtol=TOL, atol=1e-6)
        _cmp_named(gold, f"it{it}/end/G", G.state_dict().items())
        _cmp_named(gold, f"it{it}/end/D", D.state_dict().items())

    # synthesis (reference generate_images, src/gan_utils.py:197-244) continues from the trained state
    n_syn = int(gold["tiles/n"][0])
    torch.manual_seed(U.SEED_RUN + 1)
    tiles = O.synth_tiles(G, vae, data["rna_data"][:1], n_syn)
    assert tiles.shape == (n_syn, size, size, 3) and tiles.dtype == np.float32
    flat = np.ascontiguousarray(tiles.transpose(0, 3, 1, 2)).astype(np.float64).reshape(-1)
    np.testing.assert_allclose(flat[U.sample_idx(flat.size, 4096)], gold["tiles/sample"], rtol=0, atol=2e-6)
    torch.manual_seed(U.SEED_RUN + 2)
    profiles = torch.randn(n_syn, feats, generator=torch.Generator().manual_seed(U.SEED_BATCH + 1))
    tiles2 = O.synth_tiles(G, vae, profiles, n_syn)
    flat2 = np.ascontiguousarray(tiles2.transpose(0, 3, 1, 2)).astype(np.float64).reshape(-1)
    np.testing.assert_allclose(flat2[U.sample_idx(flat2.size, 4096)], gold["tiles_multi/sample"], rtol=0, atol=2e-6)


def test_oracle_matches_reference_modules():
    gold = U.load_golden("modules.npz")
    up = O.OracleUpGenerator(encoding_dims=2048, out_size=32, out_channels=3, step_channels=64,
                             nonlinearity=torch.nn.LeakyReLU(0.2), last_nonlinearity=torch.nn.Tanh())
    O.reinit_(up, 21)
    up.train()
    g = torch.Generator().manual_seed(22)
    z = torch.randn(6, 2048, generator=g)
    y = up(z)
    np.testing.assert_allclose(U.sample_of(y, 4096), gold["upgen/out_sample"], rtol=0, atol=1e-5)
    _cmp_named(gold, "upgen/end", up.state_dict().items())

    feats = 300
    vae = O.OracleVAE(feats, beta=0.0005)
    O.reinit_(vae, 23)
    x = torch.randn(16, feats, generator=g)
    vae.eval()
    zm, zl, _ = vae.encode(x)
    np.testing.assert_allclose(U.sample_of(zm, 2048), gold["vae/z_mean_sample"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(U.sample_of(zl, 2048), gold["vae/z_logvar_sample"], rtol=0, atol=1e-5)
    vae.train()
    torch.manual_seed(24)
    opt = Adam(vae.parameters(), lr=5e-5, weight_decay=0)
    logs = []
    for _ in range(2):
        r = O.vae_train_step(vae, opt, x, 0.0005)
        logs.append([r["total_loss"], r["reconstruction_loss"], r["kl_loss"]])
    np.testing.assert_allclose(logs, gold["vae/train_losses"], rtol=1e-5)
    _cmp_named(gold, "vae/end", vae.state_dict().items())


def test_reference_error_behaviour():
    with pytest.raises(Exception, match="at least 16\\*16 and an exact power of 2"):
        O.OracleGenerator(out_size=24)
    G, D, vae = _build(32, 64)
    G.label_type = "required"
    data = O.make_batch(4, 64, 32, 1)
    with pytest.raises(Exception, match="GAN model requires labels for training"):
        O.g_step(G, D, None, vae, data)
