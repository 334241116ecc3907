"""Host-logic check (CPU, float64): the hand-scheduled gradient-penalty second-order pass that
rnagan_b200/engine.py drives on the GPU (in the per-channel-sum form implemented by
rg_bn_bwd_reduce / rg_bn_bwd_apply / rg_bn_gp_reduce / rg_bn_gp_apply) equals torch.autograd's double backward of
wasserstein_gradient_penalty_vae (src/wgan_loss.py:32-44) on a small critic with the same layer types."""
import torch
import torch.nn.functional as F

SLOPE = 0.2
EPS = 1e-5


def lrelu_mask(u):
    return torch.where(u > 0, torch.ones_like(u), torch.full_like(u, SLOPE))


def S(t):   # per-channel sum over (B, H, W)
    return t.sum(dim=(0, 2, 3), keepdim=True)


def manual_gp_grads(x, W0, b0, Ws, gammas, betas, w_head, lambd):
    """Returns (penalty, grads dict) using only first-order building blocks: conv fprop/dgrad/wgrad + BN formulas."""
    L = len(Ws)
    conv = lambda t, W: F.conv2d(t, W, stride=2, padding=1)
    dgrad = lambda t, W: F.conv_transpose2d(t, W, stride=2, padding=1)
    wgrad = lambda xin, dy, W: torch.nn.grad.conv2d_weight(xin, W.shape, dy, stride=2, padding=1)

    # ---- step 1: forward
    a0 = conv(x, W0) + b0.view(1, -1, 1, 1)
    m = [lrelu_mask(a0)]
    h = [F.leaky_relu(a0, SLOPE)]
    a, mean, rstd, xhat = [None], [None], [None], [None]
    for l in range(1, L + 1):
        al = conv(h[l - 1], Ws[l - 1])
        M = al.numel() // al.shape[1]
        mu = S(al) / M
        var = S((al - mu) ** 2) / M
        r = (var + EPS).rsqrt()
        xh = (al - mu) * r
        u = gammas[l - 1].view(1, -1, 1, 1) * xh + betas[l - 1].view(1, -1, 1, 1)
        a.append(al); mean.append(mu); rstd.append(r); xhat.append(xh)
        m.append(lrelu_mask(u))
        h.append(F.leaky_relu(u, SLOPE))
    a_head = F.conv2d(h[L], w_head)                      # [B,1,1,1]
    m_head = lrelu_mask(a_head)

    # ---- step 2: first backward to the input (d_out = ones)
    d_a_head = m_head
    dh = F.conv_transpose2d(d_a_head, w_head)            # delta h_L
    du, da, s1, s2 = [None] * (L + 1), [None] * (L + 1), [None] * (L + 1), [None] * (L + 1)
    for l in range(L, 0, -1):
        du[l] = dh * m[l]
        M = du[l].numel() // du[l].shape[1]
        s1[l] = S(du[l]); s2[l] = S(du[l] * xhat[l])
        g_r = gammas[l - 1].view(1, -1, 1, 1) * rstd[l]
        da[l] = g_r * (du[l] - s1[l] / M - xhat[l] * s2[l] / M)
        dh = dgrad(da[l], Ws[l - 1])
    da0 = dh * m[0]
    g = dgrad(da0, W0)
    norm = g.norm(2)
    penalty = (norm - 1) ** 2
    seed = lambd * 2 * (norm - 1) / norm

    # ---- step 3: adjoint sweep bottom -> top with A_g = seed * g
    grads = {"W0": None, "b0": None, "W": [None] * L, "gamma": [None] * L, "beta": [None] * L, "w_head": None}
    A_g = seed * g
    grads["W0"] = wgrad(A_g, da0, W0)                    # dW0[p,s] = sum lo=da0 (x) hi=A_g
    A_da = conv(A_g, W0)
    A_dh = A_da * m[0]
    A_a = [None] * (L + 1)
    for l in range(1, L + 1):
        grads["W"][l - 1] = wgrad(A_dh, da[l], Ws[l - 1])
        ggI = conv(A_dh, Ws[l - 1])
        gO = du[l]
        M = ggI.numel() // ggI.shape[1]
        q1, q2, q3 = S(ggI), S(ggI * xhat[l]), S(ggI * gO)
        gam = gammas[l - 1].view(1, -1, 1, 1)
        r = rstd[l]
        A_du = gam * r * (ggI - q1 / M - xhat[l] * q2 / M)
        grads["gamma"][l - 1] = (r * (q3 - (q1 * s1[l] + q2 * s2[l]) / M)).flatten()
        A_a[l] = gam * r * r / M * (xhat[l] * (q1 * s1[l] / M - q3 + 3 * s2[l] * q2 / M)
                                    + q2 * (s1[l] / M - gO) + s2[l] * (q1 / M - ggI))
        A_dh = A_du * m[l]
    grads["w_head"] = (d_a_head.view(-1, 1, 1, 1) * A_dh).sum(0, keepdim=True)

    # ---- step 4: ordinary backward over the forward graph, seeded by A_a
    T = A_a[L]
    for l in range(L, 0, -1):
        grads["W"][l - 1] = grads["W"][l - 1] + wgrad(h[l - 1], T, Ws[l - 1])
        A_h = dgrad(T, Ws[l - 1])
        if l - 1 >= 1:
            k = l - 1
            A_u = A_h * m[k]
            M = A_u.numel() // A_u.shape[1]
            t1, t2 = S(A_u), S(A_u * xhat[k])
            grads["gamma"][k - 1] = grads["gamma"][k - 1] + t2.flatten()
            grads["beta"][k - 1] = t1.flatten()
            g_r = gammas[k - 1].view(1, -1, 1, 1) * rstd[k]
            T = A_a[k] + g_r * (A_u - t1 / M - xhat[k] * t2 / M)
        else:
            A_a0 = A_h * m[0]
            grads["W0"] = grads["W0"] + wgrad(x, A_a0, W0)
            grads["b0"] = A_a0.sum(dim=(0, 2, 3))
    grads["beta"][L - 1] = torch.zeros_like(betas[L - 1])
    return penalty, grads


def test_manual_gp_matches_autograd_double_backward():
    torch.manual_seed(0)
    dt = torch.float64
    B, C0, chans = 3, 4, [6, 8]
    x = torch.randn(B, 3, 32, 32, dtype=dt, requires_grad=True)
    W0 = (torch.randn(C0, 3, 4, 4, dtype=dt) * 0.2).requires_grad_()
    b0 = (torch.randn(C0, dtype=dt) * 0.1).requires_grad_()
    Ws, gammas, betas = [], [], []
    cin = C0
    for c in chans:
        Ws.append((torch.randn(c, cin, 4, 4, dtype=dt) * 0.2).requires_grad_())
        gammas.append((1 + 0.2 * torch.randn(c, dtype=dt)).requires_grad_())
        betas.append((0.1 * torch.randn(c, dtype=dt)).requires_grad_())
        cin = c
    w_head = (torch.randn(1, cin, 4, 4, dtype=dt) * 0.2).requires_grad_()
    lambd = 10.0

    # autograd reference (the reference's own formulation)
    hcur = F.leaky_relu(F.conv2d(x, W0, b0, stride=2, padding=1), SLOPE)
    for W, ga, be in zip(Ws, gammas, betas):
        hcur = F.leaky_relu(F.batch_norm(F.conv2d(hcur, W, stride=2, padding=1), None, None, ga, be, True, 0.1, EPS),
                            SLOPE)
    out = F.leaky_relu(F.conv2d(hcur, w_head), SLOPE).view(B)
    gref = torch.autograd.grad(out, x, torch.ones_like(out), create_graph=True, retain_graph=True)[0]
    pen = (gref.norm(2) - 1) ** 2
    params = [W0, b0] + Ws + gammas + betas + [w_head]
    ref = torch.autograd.grad(lambd * pen, params, allow_unused=True)

    with torch.no_grad():
        p, gm = manual_gp_grads(x.detach(), W0, b0, Ws, gammas, betas, w_head, lambd)
    assert abs(p.item() - pen.item()) < 1e-10 * max(1.0, abs(pen.item()))
    got = [gm["W0"], gm["b0"]] + gm["W"] + gm["gamma"] + gm["beta"] + [gm["w_head"]]
    for r, g_ in zip(ref, got):
        r = torch.zeros_like(g_) if r is None else r
        assert (r - g_).abs().max().item() <= 1e-9 * max(1.0, r.abs().max().item())
