"""Optimizer step on the fused multi-tensor Adam kernel (rg_adam_step), driven by the caller's torch.optim.Adam.

The reference builds ``torch.optim.Adam(lr=1e-4 / 4e-4, betas=(0.5, 0.999))`` objects (src/histopathology_gan.py:252,257)
and passes them to ``train_ops``; to stay a drop-in, the optimizer object remains the owner of the hyper-parameters and
of the state (``state[p] = {step, exp_avg, exp_avg_sq}`` in torch's own format, so ``optimizer.state_dict()`` and the
torchgan checkpoint layout keep working) while the arithmetic runs in one kernel launch over all tensors.
"""
import torch

from . import ops


def _ensure_state(opt, p):
    st = opt.state[p]
    if len(st) == 0:
        st["step"] = torch.tensor(0.0, dtype=torch.float32)
        st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
    for k in ("exp_avg", "exp_avg_sq"):
        # the kernel walks physical memory: moments must share the parameter's layout (a state loaded from an upstream
        # checkpoint is contiguous while the engine keeps conv weights channels_last) -- converted once
        if st[k].stride() != p.stride() or st[k].device != p.device or st[k].dtype != p.dtype:
            new = torch.empty_like(p, memory_format=torch.preserve_format)
            new.copy_(st[k])
            st[k] = new
    return st


def _dense(t):
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


def group_step(opt, gi):
    """Host-side step count of param group gi (0 before the first step)."""
    for p in opt.param_groups[gi]["params"]:
        st = opt.state.get(p)
        if st and "step" in st:
            return int(st["step"])
    return 0


def advance_host_steps(opt):
    """Advance the optimizer's per-parameter `step` counters by one without launching anything (a replayed CUDA graph
    has done the update; `optimizer.state_dict()` must keep telling the truth)."""
    for group in opt.param_groups:
        for p in group["params"]:
            st = opt.state.get(p)
            if st and "step" in st:
                st["step"] += 1


def adam_step(opt, clamp=None, grad_scale=1.0, dyn=None):
    """One Adam step over every parameter of ``opt`` that has a gradient (gradients are multiplied by grad_scale
    first: 1/world_size after a SUM all-reduce). Returns nothing; asynchronous.
    dyn: {group index: device fp32 [2] = (lr, step count)} -- the launch reads the learning rate and the step count from
    the device and advances the count itself (CUDA-graph capture, steps.py); the host counters are then advanced by the
    caller with advance_host_steps, once per replay."""
    if not isinstance(opt, torch.optim.Adam):
        raise NotImplementedError(f"only torch.optim.Adam is implemented on the sm_100a path (got {type(opt).__name__})")
    tables = opt.__dict__.setdefault("_rg_tables", {})
    for gi, group in enumerate(opt.param_groups):
        if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False) or group.get("maximize", False):
            raise NotImplementedError("Adam with weight_decay / amsgrad / maximize is not implemented")
        params = [p for p in group["params"] if p.grad is not None]
        if not params:
            continue
        states = [_ensure_state(opt, p) for p in params]
        for p in params:
            if p.dtype != torch.float32 or not _dense(p) or p.grad.stride() != p.stride():
                raise NotImplementedError("fused Adam needs dense fp32 parameters with gradients in the same layout")
        shadows = [getattr(p, "_rg_shadow", None) for p in params]
        key = tuple(t.data_ptr() for p, s in zip(params, states) for t in (p, p.grad, s["exp_avg"], s["exp_avg_sq"])) + \
            tuple(0 if t is None else t.data_ptr() for t in shadows)
        ent = tables.get(gi)
        if ent is None or ent[0] != key:
            tab = ops.AdamTable([p.detach() for p in params], [p.grad for p in params],
                                [s["exp_avg"] for s in states], [s["exp_avg_sq"] for s in states], shadows)
            ent = (key, tab)
            tables[gi] = ent
        b1, b2 = group["betas"]
        if dyn is not None:
            ent[1].step_dyn(dyn[gi], b1, b2, group["eps"], clamp=clamp, grad_scale=grad_scale)
            continue
        for s in states:
            s["step"] += 1
        step = int(states[0]["step"])
        if len(states) > 1 and any(int(s["step"]) != step for s in states[1:]):
            # one bias correction per launch: a partially populated / merged optimizer state would be stepped wrongly
            raise NotImplementedError("fused Adam needs every parameter of a group at the same step count")
        ent[1].step(group["lr"], b1, b2, group["eps"], step, clamp=clamp, grad_scale=grad_scale)
