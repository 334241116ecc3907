"""Diagnostic: layer-by-layer forward deviation of the CUDA engines from the fp32 oracle, next to torch's own bf16."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_oracle as O  # noqa: E402
from rnagan_b200 import dcgan, ops  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def block_outputs(net_blocks, x, autocast):
    outs = []
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast), torch.no_grad():
        for blk in net_blocks:
            x = blk(x)
            outs.append(x.float())
    return outs


def main(size=64, B=4):
    dev = torch.device("cuda:0")
    lrelu, tanh = torch.nn.LeakyReLU(0.2), torch.nn.Tanh()
    oG = O.OracleGenerator(2048, size, 3, 64, nonlinearity=lrelu, last_nonlinearity=tanh)
    oD = O.OracleCritic(size, 3, 64, nonlinearity=lrelu, last_nonlinearity=lrelu)
    O.reinit_(oG, 11); O.reinit_(oD, 12)
    oG.train(); oD.train()
    G = dcgan.DCGANGenerator(2048, size, 3, 64, nonlinearity=torch.nn.LeakyReLU(0.2), last_nonlinearity=torch.nn.Tanh()).to(dev)
    D = dcgan.DCGANDiscriminator(size, 3, 64, nonlinearity=torch.nn.LeakyReLU(0.2), last_nonlinearity=torch.nn.LeakyReLU(0.2)).to(dev)
    G.load_state_dict(oG.state_dict()); D.load_state_dict(oD.state_dict())
    G.train(); D.train()
    g = torch.Generator().manual_seed(5)
    z = torch.randn(B, 2048, generator=g)
    z = (z - z.mean(0)) / z.std(0)
    ref = block_outputs(list(oG.model), z.view(B, 2048, 1, 1), False)
    ac = block_outputs(list(oG.model), z.view(B, 2048, 1, 1), True)
    img = G(z.to(dev))
    eng = G._engine()
    H = 4
    for l in range(eng.n + 1):
        C = ref[l].shape[1]
        h = eng.bufs.get(f"module.h{l}", (B, H, H, C)).float().permute(0, 3, 1, 2)
        print(f"G block {l}: cuda rel {rel(h, ref[l]):.5f} | torch-bf16 rel {rel(ac[l], ref[l]):.5f}")
        H *= 2
    print(f"G image  : cuda rel {rel(img, ref[-1]):.5f} | torch-bf16 rel {rel(ac[-1], ref[-1]):.5f}")
    x = torch.rand(B, 3, size, size, generator=g) * 2 - 1
    ref = block_outputs(list(oD.model) + [oD.disc], x, False)
    ac = block_outputs(list(oD.model) + [oD.disc], x, True)
    out = D(x.to(dev))
    eng = D._engine()
    H = size // 2
    for l in range(eng.n + 1):
        C = ref[l].shape[1]
        h = eng.bufs.get(f"module.h{l}", (B, H, H, C)).float().permute(0, 3, 1, 2)
        print(f"D block {l}: cuda rel {rel(h, ref[l]):.5f} | torch-bf16 rel {rel(ac[l], ref[l]):.5f}")
        H //= 2
    print(f"D out    : cuda rel {rel(out, ref[-1].view(B)):.5f} | torch-bf16 rel {rel(ac[-1].view(B), ref[-1].view(B)):.5f}")
    print("D out oracle", ref[-1].view(B).tolist(), "cuda", out.tolist())


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 4)
