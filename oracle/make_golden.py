"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN MODULES (build container only).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/

The reference's src/dcgan.py, src/wgan_loss.py, src/betaVAE.py are imported unmodified (torchgan names come from
oracle/torchgan_shim); src/gan_utils.py cannot be imported (lmdb/lz4framed/matplotlib missing), so its
`generate_images` function is extracted from the source text with `ast` and executed as is.
Nothing here is used at run time on the GPU box; only the .npz fixtures travel.
"""
import ast
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/src"
sys.path.insert(0, os.path.join(HERE, "torchgan_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from oracle import ref_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CONFIGS = {
    # name: (image size, batch, rna_features, iterations)
    "mini32": (32, 8, 256, 2),
    "mini64": (64, 4, 192, 1),
    # BASELINE config-2 shapes (gan_run_lung.json: 256x256 tiles, 19198 protein-coding genes), one iteration
    "full256": (256, 8, 19198, 1),
}
SEED_G, SEED_D, SEED_V, SEED_BATCH, SEED_RUN = 11, 12, 13, 14, 99


def sample_idx(n, k=256):
    return np.unique(np.linspace(0, n - 1, num=min(k, n)).astype(np.int64))


def snap(prefix, named, out, grads=False):
    for name, p in named:
        t = p.grad if grads else p
        if t is None:
            continue
        v = t.detach().double().flatten()
        out[f"{prefix}/{name}/sum"] = np.float64(v.sum().item())
        out[f"{prefix}/{name}/l2"] = np.float64(v.norm().item())
        out[f"{prefix}/{name}/sample"] = v[torch.from_numpy(sample_idx(v.numel()))].numpy()


def load_generate_images():
    import torchvision.transforms as transforms
    src = open(os.path.join(REF, "gan_utils.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "generate_images"][0]
    ns = {"torch": torch, "np": np, "transforms": transforms}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "gan_utils.generate_images", "exec"), ns)
    return ns["generate_images"]


def run_config(name, size, batch, feats, iters):
    import betaVAE as ref_vae
    import wgan_loss as ref_loss
    from torch.optim import Adam
    from torchgan.models import DCGANDiscriminator, DCGANGenerator
    from torchgan.trainer import Trainer

    out = {}
    vae = ref_vae.betaVAE(feats, 2048, [6000, 4000, 2048], [4000, 6000], beta=0.005)
    O.reinit_(vae, SEED_V)
    ckpt = os.path.join(tempfile.mkdtemp(), "vae.pt")
    torch.save(vae.state_dict(), ckpt)

    net = {
        "generator": {"name": DCGANGenerator,
                      "args": {"encoding_dims": 2048, "out_channels": 3, "step_channels": 64, "out_size": size,
                               "nonlinearity": nn.LeakyReLU(0.2), "last_nonlinearity": nn.Tanh()},
                      "optimizer": {"name": Adam, "args": {"lr": 0.0001, "betas": (0.5, 0.999)}}},
        "discriminator": {"name": DCGANDiscriminator,
                          "args": {"in_size": size, "in_channels": 3, "step_channels": 64,
                                   "nonlinearity": nn.LeakyReLU(0.2), "last_nonlinearity": nn.LeakyReLU(0.2)},
                          "optimizer": {"name": Adam, "args": {"lr": 0.0004, "betas": (0.5, 0.999)}}},
    }
    losses = [ref_loss.WassersteinGeneratorLossVAE(ckpt, feats),
              ref_loss.WassersteinDiscriminatorLossVAE(ckpt, feats),
              ref_loss.WassersteinGradientPenaltyVAE(ckpt, feats)]
    tr = Trainer(net, losses, device=torch.device("cpu"), sample_size=64, epochs=1, devices=[0])
    O.reinit_(tr.generator, SEED_G)
    O.reinit_(tr.discriminator, SEED_D)
    tr.generator.train()
    tr.discriminator.train()
    batch_data = O.make_batch(batch, feats, size, SEED_BATCH)
    tr.real_inputs = batch_data
    tr.batch_size = batch

    torch.manual_seed(SEED_RUN)
    lg, ld, lp = list(tr.losses.values())
    loss_log = []
    for it in range(iters):
        v_g = tr._call_train_ops(lg)
        snap(f"it{it}/g_step/grad_G", tr.generator.named_parameters(), out, grads=True)
        v_d = tr._call_train_ops(ld)
        snap(f"it{it}/critic_step/grad_D", tr.discriminator.named_parameters(), out, grads=True)
        v_p = tr._call_train_ops(lp)
        snap(f"it{it}/gp_step/grad_D", tr.discriminator.named_parameters(), out, grads=True)
        loss_log.append([v_g, v_d, v_p])
        snap(f"it{it}/end/G", tr.generator.state_dict().items(), out)
        snap(f"it{it}/end/D", tr.discriminator.state_dict().items(), out)
    out["losses"] = np.asarray(loss_log, dtype=np.float64)

    # synthesis through the reference's own generate_images (train-mode G, chunks of 10)
    gen_images = load_generate_images()
    vae.eval()
    torch.manual_seed(SEED_RUN + 1)
    tr.device = torch.device("cpu")
    # generate_images hard-codes view(-1, 3, 256, 256) (src/gan_utils.py:224): pick sample sizes whose element
    # count is a multiple of 3*256*256 and store values in flat NCHW order, which is independent of that view.
    n_syn = (3 * 256 * 256) // (3 * size * size)
    if size == 256:
        n_syn = 12            # one full chunk of 10 + a ragged chunk of 2 (a single sample would make std() NaN)
    profile = batch_data["rna_data"][:1]
    tiles = gen_images(tr, gene_exp=profile, sample_size=n_syn, betavae=vae)
    flat = np.ascontiguousarray(tiles.transpose(0, 3, 1, 2)).astype(np.float64).reshape(-1)
    out["tiles/n"] = np.asarray([n_syn])
    out["tiles/sample"] = flat[sample_idx(flat.size, 4096)]
    out["tiles/mean"] = np.float64(flat.mean())
    out["tiles/std"] = np.float64(flat.std())
    # multi-profile conditioning (one profile per row), same function
    torch.manual_seed(SEED_RUN + 2)
    profiles = torch.randn(n_syn, feats, generator=torch.Generator().manual_seed(SEED_BATCH + 1))
    tiles2 = gen_images(tr, gene_exp=profiles, sample_size=n_syn, betavae=vae)
    flat2 = np.ascontiguousarray(tiles2.transpose(0, 3, 1, 2)).astype(np.float64).reshape(-1)
    out["tiles_multi/sample"] = flat2[sample_idx(flat2.size, 4096)]

    out["meta"] = np.asarray([size, batch, feats, iters, SEED_G, SEED_D, SEED_V, SEED_BATCH, SEED_RUN])
    np.savez_compressed(os.path.join(GOLD, f"gan_{name}.npz"), **out)
    print(name, "losses", loss_log)


def run_modules():
    """Module-level goldens: DCGANUpGenerator forward (src/dcgan.py), betaVAE forward + loss (src/betaVAE.py)."""
    import betaVAE as ref_vae
    import dcgan as ref_dcgan

    out = {}
    up = ref_dcgan.DCGANUpGenerator(encoding_dims=2048, out_size=32, out_channels=3, step_channels=64,
                                    nonlinearity=nn.LeakyReLU(0.2), last_nonlinearity=nn.Tanh())
    O.reinit_(up, 21)
    up.train()
    g = torch.Generator().manual_seed(22)
    z = torch.randn(6, 2048, generator=g)
    y = up(z)
    out["upgen/out_sample"] = y.detach().double().flatten()[torch.from_numpy(sample_idx(y.numel(), 4096))].numpy()
    out["upgen/out_sum"] = np.float64(y.double().sum().item())
    snap("upgen/end", up.state_dict().items(), out)

    feats = 300
    vae = ref_vae.betaVAE(feats, 2048, [6000, 4000, 2048], [4000, 6000], beta=0.0005)
    O.reinit_(vae, 23)
    x = torch.randn(16, feats, generator=g)
    vae.eval()
    zm, zl, h = vae.encode(x)
    out["vae/z_mean_sample"] = zm.detach().double().flatten()[torch.from_numpy(sample_idx(zm.numel(), 2048))].numpy()
    out["vae/z_logvar_sample"] = zl.detach().double().flatten()[torch.from_numpy(sample_idx(zl.numel(), 2048))].numpy()
    vae.train()
    torch.manual_seed(24)
    opt = torch.optim.Adam(vae.parameters(), lr=5e-5, weight_decay=0)
    logs = []
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        rec, m, lv = vae(x)
        ls = ref_vae.betaVAEloss(x, rec, m, lv, 0.0005, training=True)
        ls["total_loss"].backward()
        opt.step()
        logs.append([ls["total_loss"].item(), ls["reconstruction_loss"].item(), ls["kl_loss"].item()])
    out["vae/train_losses"] = np.asarray(logs, dtype=np.float64)
    snap("vae/end", vae.state_dict().items(), out)
    out["meta"] = np.asarray([feats, 16, 21, 22, 23, 24])
    np.savez_compressed(os.path.join(GOLD, "modules.npz"), **out)
    print("modules: vae losses", logs)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for cfg_name, cfg in CONFIGS.items():
        if not only or cfg_name in only:
            run_config(cfg_name, *cfg)
    if not only or "modules" in only:
        run_modules()
