"""The three optimiser steps of one RNA-GAN iteration as device-side schedules (no host synchronisation).

``wgan_loss.*.train_ops`` (the reference-facing API) and ``bench.py`` (device-resident timing) both call these; the
only difference is where the noise / eps / batch come from and whether the loss is read back with ``.item()``.
"""
import os

import torch

from . import ops
from .optim import adam_step

F32 = torch.float32
BF16 = torch.bfloat16


def latent(ge, noise_d, z):
    """standardise_0(noise + z) -> bf16 [B, E] (src/wgan_loss.py:105-106); z=None (the un-conditioned `wgan` losses,
    src/histopathology_gan.py:267-272): the N(0, I) noise itself, cast to bf16."""
    B, E = noise_d.shape
    lat = ge.bufs.get("lat", (B, E), BF16)
    if z is None:
        ops.cast_pad_bf16(noise_d, E, out=lat)
    else:
        ops.latent_prep(noise_d, z, lat_bf16=lat)
    return lat


def _loss_buf(ge, name):
    return ge.bufs.get(name, (1,), F32)


# ------------------------------------------------------------------------------------------------ optimiser updates
# Single process: Adam runs right after backward, like the reference.  Data-parallel: the all-reduce of the LAST layer
# a backward produces cannot overlap that backward (the generator's layer 0 is 60 % of its gradient bytes and is produced
# last), so the update is deferred to the first consumer of the new weights -- G's update to just before the next
# G forward (after D(real) of the critic step), D's to just before the next D forward (after the G forward of the GP /
# next G step) -- and the tail of the reduction hides behind that work.  Every path that reads the weights from outside
# these schedules (module.forward, state_dict, checkpoints, synthesis) goes through `module._engine()`, which applies
# what is pending first; observable values are those of the reference's order of operations.
DEFER_UPDATES = os.environ.get("RG_DP_DEFER", "1") != "0"


def _queue_update(eng, opt):
    if eng.sync.world() == 1 or not DEFER_UPDATES:
        adam_step(opt, grad_scale=eng.sync.finish())        # gradients averaged over ranks (no-op single-process)
        eng.pack(full=False)                                # Adam re-emitted the bf16 GEMM operands itself
        return
    eng.sync.launch_rest()
    eng.pending_update = opt


def apply_update(eng):
    """Apply the engine's deferred optimiser step, if any (waits for its gradient reductions on the stream)."""
    opt = getattr(eng, "pending_update", None)
    if opt is not None:
        eng.pending_update = None
        adam_step(opt, grad_scale=eng.sync.wait())
        eng.pack(full=False)


def _need_train_mode(what, *modules):
    """The hand-scheduled backward passes differentiate train-mode BatchNorm (batch statistics), which is what every
    reference driver runs (Trainer.train puts the models in train mode each epoch).  With a module in eval mode the
    forward would use running statistics while the backward subtracted batch-mean terms: refuse instead of returning
    silently wrong gradients."""
    for m in modules:
        if not m.training:
            raise NotImplementedError(f"{what}: {type(m).__name__} is in eval mode; the sm_100a backward of BatchNorm is "
                                      "implemented for train mode (batch statistics) only -- call .train() first")


# ------------------------------------------------------------------------------------------------ CUDA graphs
# One optimiser step is ~130 launches, a third of them a few microseconds long; at 11 ms per iteration the host spends
# ~5 ms per iteration in ctypes launch calls.  In a single-process job every step is therefore captured into a CUDA graph
# after two eager calls and replayed from then on: the inputs (noise, z, real tiles, eps) are copied into static buffers
# first, Adam reads its learning rate and step count from device memory (rg_adam_step_dyn), and everything else a step
# touches already lives in static buffers (engine._Bufs).  Data-parallel jobs stay eager (their gradient exchange runs
# on side streams with host-driven bucketing).  RG_GRAPHS=0 keeps everything eager.
USE_GRAPHS = os.environ.get("RG_GRAPHS", "1") != "0"
GRAPH_WARMUP = 2
_GRAPHS = {}
GRAPH_STATS = {"replays": 0, "kernel_launches": 0, "captured": 0, "failed": 0}


def use_graphs(flag):
    """Switch graph replay on / off at run time (bench.py's per-kernel event timing needs eager launches)."""
    global USE_GRAPHS
    USE_GRAPHS = bool(flag)


def _opt_sig(opt):
    """Identity of everything a captured Adam launch has baked in: the addresses of parameters, gradients, moments and
    bf16 shadows (optimizer.load_state_dict or a re-bound gradient buffer gives a new signature, hence a new capture)."""
    sig = []
    for group in opt.param_groups:
        for p in group["params"]:
            if p.grad is None:
                continue
            st = opt.state.get(p)
            if not st or "exp_avg" not in st:
                return None
            sh = getattr(p, "_rg_shadow", None)
            sig.append((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                        0 if sh is None else sh.data_ptr()))
    return tuple(sig)


class _Graphed:
    def __init__(self):
        self.calls, self.graph, self.static, self.out, self.dyn, self.failed = 0, None, {}, None, None, False
        self.launches, self.step_mirror, self.lr_mirror = 0, None, None


def _run_step(kind, generator, discriminator, opt, inputs, body, extra_key=()):
    """inputs: {name: tensor or None}; body(inputs, dyn) launches the step and returns its output tensor."""
    ge = generator._engine(flush=False)
    graphable = (USE_GRAPHS and ge.sync.world() == 1 and isinstance(opt, torch.optim.Adam)
                 and not torch.cuda.is_current_stream_capturing() and ops.PROFILE is None)
    if not graphable:
        return body(inputs, None)
    sig = _opt_sig(opt)
    if sig is None:                          # optimizer state not materialised yet (first step): eager
        return body(inputs, None)
    key = (kind, id(ge), id(discriminator._engine(flush=False)), id(opt), sig, generator.training,
           discriminator.training, extra_key,
           tuple((k, None if v is None else (tuple(v.shape), v.dtype)) for k, v in inputs.items()),
           tuple((g["betas"], g["eps"]) for g in opt.param_groups))
    ent = _GRAPHS.get(key)
    if ent is None:
        if len(_GRAPHS) >= 32:               # engines / optimizers were rebuilt many times: drop the stale captures
            _GRAPHS.clear()
        ent = _GRAPHS[key] = _Graphed()
    if ent.failed or ent.calls < GRAPH_WARMUP:
        ent.calls += 1
        return body(inputs, None)
    from . import _lib
    from .optim import advance_host_steps, group_step
    for k, v in inputs.items():
        if v is None:
            continue
        buf = ent.static.get(k)
        if buf is None:
            buf = ent.static[k] = torch.empty_like(v, memory_format=torch.contiguous_format)
        if buf.data_ptr() != v.data_ptr():
            buf.copy_(v, non_blocking=True)
    steps_now = [group_step(opt, gi) for gi in range(len(opt.param_groups))]
    lrs_now = [float(g["lr"]) for g in opt.param_groups]
    if ent.graph is None:
        try:
            dev = ge.device
            ent.dyn = {gi: torch.tensor([lrs_now[gi], float(steps_now[gi])], dtype=F32, device=dev)
                       for gi in range(len(opt.param_groups))}
            st_in = {k: (None if v is None else ent.static[k]) for k, v in inputs.items()}
            torch.cuda.synchronize()
            n0 = _lib.lib().rg_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                ent.out = body(st_in, ent.dyn)
            ent.launches = int(_lib.lib().rg_launch_count() - n0)
            ent.graph = g
            ent.step_mirror, ent.lr_mirror = steps_now, lrs_now
            GRAPH_STATS["captured"] += 1
        except Exception as ex:          # capture is an optimisation: fall back to eager launches, loudly, once
            import warnings
            ent.failed, ent.graph = True, None
            GRAPH_STATS["failed"] += 1
            warnings.warn(f"rnagan_b200: CUDA-graph capture of the {kind} step failed ({type(ex).__name__}: {ex}); "
                          "running it eagerly")
            torch.cuda.synchronize()
            return body(inputs, None)
    # the device-side (lr, step) pair follows the optimizer object (schedulers, steps taken outside the graph)
    for gi in range(len(opt.param_groups)):
        if steps_now[gi] != ent.step_mirror[gi] or lrs_now[gi] != ent.lr_mirror[gi]:
            ent.dyn[gi].copy_(torch.tensor([lrs_now[gi], float(steps_now[gi])], dtype=F32), non_blocking=False)
    ent.graph.replay()
    advance_host_steps(opt)
    ent.step_mirror = [s_ + 1 for s_ in steps_now]
    ent.lr_mirror = lrs_now
    GRAPH_STATS["replays"] += 1
    GRAPH_STATS["kernel_launches"] += ent.launches
    return ent.out


def _update(eng, opt, dyn):
    if dyn is None:
        _queue_update(eng, opt)
    else:                                   # captured: single process, no deferral, device-side step count
        adam_step(opt, grad_scale=1.0, dyn=dyn)
        eng.pack(full=False)


def g_step(generator, discriminator, opt_g, noise_d, z):
    """WassersteinGeneratorLossVAE.train_ops body (src/wgan_loss.py:100-128). Returns the device loss tensor [1]."""
    _need_train_mode("generator step", generator, discriminator)      # both BatchNorm stacks are differentiated

    def body(t, dyn):
        ge, de = generator._engine(flush=False), discriminator._engine(flush=False)
        B = t["noise"].shape[0]
        lat = latent(ge, t["noise"], t["z"])
        apply_update(ge)
        fake = ge.forward(lat, tag="g", training=generator.training)
        apply_update(de)
        out = de.forward(fake, tag="gstep", training=discriminator.training)
        loss = _loss_buf(ge, "loss_g")
        ops.wgan_loss(out, -1.0, loss)                                   # mean(-D(G(z)))
        d_img = de.backward(B, -1.0 / B, tag="gstep", params=False, want_dimg=True)
        ge.backward(lat, d_img, fake, tag="g")
        _update(ge, opt_g, dyn)
        return loss

    return _run_step("g", generator, discriminator, opt_g, {"noise": noise_d, "z": z}, body)


def critic_step(generator, discriminator, opt_d, noise_d, z, real, clip=None):
    """WassersteinDiscriminatorLossVAE.train_ops body (src/wgan_loss.py:213-262)."""
    _need_train_mode("critic step", discriminator)                    # G is forward-only here: any mode

    def body(t, dyn):
        ge, de = generator._engine(flush=False), discriminator._engine(flush=False)
        apply_update(de)
        if clip is not None:                                             # src/wgan_loss.py:213-215
            for p in discriminator.parameters():
                ops.clamp_(p.data, clip[0], clip[1])
            de.pack()
        B = t["noise"].shape[0]
        lat = latent(ge, t["noise"], t["z"])
        out_real = de.forward(t["real"], tag="real", training=discriminator.training)
        apply_update(ge)                                    # G's reduction tail overlapped D(real)
        fake = ge.forward(lat, tag="g", training=generator.training, keep=False)      # G is not differentiated here
        out_fake = de.forward(fake, tag="fake", training=discriminator.training)
        loss = _loss_buf(ge, "loss_d")
        ops.wgan_loss(out_fake, 1.0, loss, b=out_real, sign_b=-1.0)      # mean(D(G(z)) - D(x))
        de.backward_pair(B, (("real", -1.0), ("fake", 1.0)))     # one wgrad / dgrad launch per layer over both passes
        _update(de, opt_d, dyn)
        return loss

    return _run_step("critic", generator, discriminator, opt_d, {"noise": noise_d, "z": z, "real": real}, body,
                     extra_key=(None if clip is None else tuple(clip),))


def gp_step(generator, discriminator, opt_d, noise_d, z, real, eps_d, lambd=10.0):
    """WassersteinGradientPenaltyVAE.train_ops body (src/wgan_loss.py:357-388). Returns device [P, seed, ||g||]."""
    _need_train_mode("gradient-penalty step", discriminator)

    def body(t, dyn):
        ge, de = generator._engine(flush=False), discriminator._engine(flush=False)
        lat = latent(ge, t["noise"], t["z"])
        apply_update(ge)
        fake = ge.forward(lat, tag="g", training=generator.training, keep=False)      # G is not differentiated here
        apply_update(de)                                    # D's critic-step reduction overlapped this G forward
        out3 = de.gradient_penalty(t["real"], fake, t["eps"], lambd=lambd)
        _update(de, opt_d, dyn)
        return out3

    return _run_step("gp", generator, discriminator, opt_d, {"noise": noise_d, "z": z, "real": real, "eps": eps_d},
                     body, extra_key=(float(lambd),))


def flush_updates(*modules):
    """Apply every deferred optimiser step of the given generator / critic modules (data-parallel runs)."""
    for m in modules:
        m._engine()
