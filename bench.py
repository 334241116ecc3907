#!/usr/bin/env python
"""bench.py -- RNA-GAN hot-path benchmark (BASELINE.json metric: WGAN G+D train steps/s, plus synthesized tiles/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

Workload (config 2 of BASELINE.json, "RNA-GAN lung training"): gan_run_lung.json shapes -- 19198 protein-coding
genes -> betaVAE latent 2048 -> DCGANGenerator -> 3x256x256 tile -> DCGANDiscriminator; per-GPU batch 64; one STEP =
one full reference iteration on one 64-sample batch = G step + critic step + gradient-penalty step
(src/wgan_loss.py:82-129, 181-263, 314-389).  Synthetic data of the configured shapes, random-init networks.
N>1: one process per GPU (torchrun), batch-sharded (64 per GPU, weak scaling), NCCL gradient all-reduce; `value`
is summed over ranks.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definition of every key.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GENES, LATENT, SIZE, STEP_CH = 19198, 2048, 256, 64
FLOP_PER_SAMPLE_STEP = 105.06e9        # algorithmic FLOPs of G+critic+GP steps per sample
FLOP_PER_TILE = 5.6036e9


def env_int(name, default):
    return int(os.environ.get(name, default))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(family):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json, written by tools/ncu_summary.py); None when absent."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        key = "gemm_fwd_kernel" if family == "fwd_dgrad" else "gemm_wgrad_kernel"
        for name, v in d.get("kernels", {}).items():
            if key in name:
                return v.get("dram_bytes_per_launch")
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons = [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8:
                    continue
                sm.append(float(f[1]))
                out["sm_max_mhz"] = float(f[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ====================================================================================================== our arm
def build_job(device, batch, seed=99):
    import torch.nn as nn
    from torch.optim import Adam

    from rnagan_b200 import dcgan, wgan_loss
    from rnagan_b200.betaVAE import betaVAE
    from rnagan_b200.trainer import Trainer

    torch.manual_seed(seed)
    vae = betaVAE(GENES, LATENT, [6000, 4000, 2048], [4000, 6000], beta=0.005)
    ckpt_dir = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    ckpt = os.path.join(ckpt_dir, f"rnagan_bench_vae_{os.getpid()}.pt")
    torch.save(vae.state_dict(), ckpt)
    del vae
    net = {
        "generator": {"name": dcgan.DCGANGenerator,
                      "args": {"encoding_dims": LATENT, "out_channels": 3, "step_channels": STEP_CH, "out_size": SIZE,
                               "nonlinearity": nn.LeakyReLU(0.2), "last_nonlinearity": nn.Tanh()},
                      "optimizer": {"name": Adam, "args": {"lr": 0.0001, "betas": (0.5, 0.999)}}},
        "discriminator": {"name": dcgan.DCGANDiscriminator,
                          "args": {"in_size": SIZE, "in_channels": 3, "step_channels": STEP_CH,
                                   "nonlinearity": nn.LeakyReLU(0.2), "last_nonlinearity": nn.LeakyReLU(0.2)},
                          "optimizer": {"name": Adam, "args": {"lr": 0.0004, "betas": (0.5, 0.999)}}},
    }
    try:
        losses = [wgan_loss.WassersteinGeneratorLossVAE(ckpt, GENES), wgan_loss.WassersteinDiscriminatorLossVAE(ckpt, GENES),
                  wgan_loss.WassersteinGradientPenaltyVAE(ckpt, GENES)]
    finally:
        os.unlink(ckpt)
    tr = Trainer(net, losses, device=device, sample_size=64, epochs=1, devices=[0])
    tr.generator.train()
    tr.discriminator.train()
    tr.batch_size = batch
    return tr


def host_batch(batch, seed):
    g = torch.Generator().manual_seed(seed)
    return {"image": (torch.rand(batch, 3, SIZE, SIZE, generator=g) * 2 - 1).pin_memory(),
            "rna_data": torch.randn(batch, GENES, generator=g).pin_memory(),
            "labels": torch.zeros(batch)}


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist

    from rnagan_b200 import _lib, ops, steps

    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)
    _lib.check(_lib.lib().rg_check_device(), "rg_check_device")
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    B, K, W = args.batch, args.steps, args.warmup
    tr = build_job(device, B)
    G, D = tr.generator, tr.discriminator
    vae = tr.losses["WassersteinGeneratorLossVAE"]._encoder(device)
    names = list(tr.losses.keys())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return ms

    # ------------------------------------------------------------------ device-resident loop (`value`)
    hb = [host_batch(B, 1000 + rank), host_batch(B, 2000 + rank)]
    real_d = [b["image"].to(device) for b in hb]
    rna_d = [b["rna_data"].to(device) for b in hb]
    n_it = K + W
    g = torch.Generator().manual_seed(77 + rank)
    noise_all = (torch.rand(n_it, 3, B, LATENT, generator=g) * 0.6 - 0.3).to(device)
    eps_all = torch.rand(n_it, 1, generator=g).to(device)
    loss_log = torch.zeros(n_it, 3, device=device)

    def resident_step(i):
        i %= n_it
        j = i & 1
        z = vae.encode_mean(rna_d[j])                      # encoder runs once per iteration (same batch for 3 steps)
        l1 = steps.g_step(G, D, tr.optimizer_generator, noise_all[i, 0], z)
        l2 = steps.critic_step(G, D, tr.optimizer_discriminator, noise_all[i, 1], z, real_d[j])
        l3 = steps.gp_step(G, D, tr.optimizer_discriminator, noise_all[i, 2], z, real_d[j], eps_all[i])
        loss_log[i, 0:1].copy_(l1)
        loss_log[i, 1:2].copy_(l2)
        loss_log[i, 2:3].copy_(l3[0:1])

    for i in range(W):
        resident_step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.lib().rg_launch_count() + steps.GRAPH_STATS["kernel_launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(W, W + K):
        resident_step(i)
    e1.record()
    barrier()
    # kernels launched in the timed region: eager launches counted by the library + kernels inside replayed CUDA graphs
    launches = _lib.lib().rg_launch_count() + steps.GRAPH_STATS["kernel_launches"] - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / K
    value = world * 1000.0 / ms_step
    losses_host = loss_log.cpu()
    finite = bool(torch.isfinite(losses_host).all())

    # ------------------------------------------------------------------ end-to-end loop through train_ops (`e2e`)
    Ke = max(3, min(K, args.e2e_steps))
    tr.real_inputs = hb[0]
    for i in range(2):
        tr.real_inputs = hb[i & 1]
        tr.train_iter()
    barrier()
    e0.record()
    n_log0 = len(tr.loss_logs[names[0]])
    for i in range(Ke):
        tr.real_inputs = hb[i & 1]                          # alternate batches: the per-batch H2D/encoder cache misses
        tr.train_iter(defer=True)                           # what Trainer.train runs per batch: the host reads the
    tr.flush()                                              # three losses of step i once step i+1 is queued
    e1.record()
    assert len(tr.loss_logs[names[0]]) == n_log0 + Ke       # every step's losses reached the host inside the region
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / Ke
    h2d = B * 3 * SIZE * SIZE * 4 + B * GENES * 4 + 3 * B * LATENT * 4 + 4
    e2e = {"value": world * 1000.0 / ms_e2e, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
           "ms_per_step": ms_e2e, "steps": Ke, "api": "rnagan_b200.trainer.Trainer.train_iter(defer=True) [the per-batch call of Trainer.train] -> wgan_loss.*.device_ops; losses read on the host one step late, all inside the timed region"}

    # ------------------------------------------------------------------ steady state: hundreds of steps, clocks sampled
    sustained = None
    if args.sustained_steps > 0:
        Ks = args.sustained_steps
        barrier()
        s2 = ClockSampler(local_rank)
        if rank == 0:
            s2.start()
        e0.record()
        for i in range(Ks):
            resident_step(i)
        e1.record()
        barrier()
        c2 = s2.stop() if rank == 0 else None
        ms_s = max_over_ranks(e0.elapsed_time(e1)) / Ks
        sustained = {"value": world * 1000.0 / ms_s, "unit": "steps/s", "steps": Ks, "ms_per_step": ms_s,
                     "seconds": ms_s * Ks / 1000.0, "clocks": c2,
                     "note": "same device-resident loop as `value`, run for hundreds of steps so the power-capped "
                             "steady-state clock applies"}

    # ------------------------------------------------------------------ data-parallel correctness (after timing)
    dp = None
    if world > 1 and not args.no_dp_check:
        from rnagan_b200 import steps as _steps
        _steps.flush_updates(G, D)
        torch.cuda.synchronize()
        same = True
        for m in (G, D):
            flat = torch.cat([p.detach().flatten() for p in m.parameters()] +
                             [b.detach().flatten().float() for n_, b in m.named_buffers() if "num_batches" in n_])
            lo_, hi_ = flat.clone(), flat.clone()
            dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
            same = same and bool(torch.equal(lo_, hi_))
        from oracle import dp_check                      # the oracle as the checker, outside every timed region
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0)) // world))
        dp = dp_check.run(rank, world, device)
        dp["bench_weights_identical_after_timed_steps"] = same
        dp["note"] = ("after the timed regions: all ranks hold bit-identical G/D weights (all-reduce MIN == MAX over the "
                      "flattened parameters); then three train_ops on a 32x32 / batch-8-per-rank job vs the CPU oracle "
                      "run per shard with averaged gradients (oracle/dp_check.py); ok = within twice torch-bf16's own "
                      "deviation")

    # ------------------------------------------------------------------ live per-kernel roofline (outside timing)
    peaks = measured_peaks()
    ops.PROFILE = []
    for i in range(2):
        resident_step(W + (i % max(K, 1)))
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    agg = {}
    for kind, flops, a, b in prof:
        d = agg.setdefault(kind, [0.0, 0.0, 0])
        d[0] += flops
        d[1] += a.elapsed_time(b)
        d[2] += 1
    fam = {"fwd_dgrad": ("conv_down", "conv_up"), "wgrad": ("conv_wgrad",),
           "edge_hbm_bound": ("conv_up_img", "gemm_nt", "gemm_nn", "gemm_tn", "proj_wgrad")}
    fam_stats = {}
    for name, kinds in fam.items():
        fl = sum(agg[k][0] for k in kinds if k in agg)
        ms = sum(agg[k][1] for k in kinds if k in agg)
        n = sum(agg[k][2] for k in kinds if k in agg)
        if n:
            fam_stats[name] = {"tflops": fl / ms / 1e9, "ms_per_step": ms / 2, "launches_per_step": n // 2,
                               "gflop_per_launch": fl / n / 1e9}
    dom = max(("fwd_dgrad", "wgrad"), key=lambda k: fam_stats.get(k, {"ms_per_step": 0})["ms_per_step"])
    ds = fam_stats[dom]
    roofline = {"bound": "tensor",
                "kernel": "rg::gemm_fwd_kernel (conv fprop/dgrad)" if dom == "fwd_dgrad" else "rg::gemm_wgrad_kernel",
                "achieved": ds["tflops"], "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": ds["tflops"] / peaks["bf16_sustained"], "traffic": ncu_traffic(dom), "peak_source": peaks["source"],
                "traffic_source": "committed ncu --set full capture (profiles/ncu_traffic.json), not this run",
                "avg_launch_ms": ds["ms_per_step"] / ds["launches_per_step"],
                "algorithmic_gflop_per_launch": ds["gflop_per_launch"], "share_of_step": ds["ms_per_step"] / ms_step,
                "families": fam_stats,
                "whole_step": {"tflops": FLOP_PER_SAMPLE_STEP * B / (ms_step * 1e-3) / 1e12,
                               "frac": FLOP_PER_SAMPLE_STEP * B / (ms_step * 1e-3) / 1e12 / peaks["bf16_sustained"]}}

    # ------------------------------------------------------------------ synthesis throughput (config 4 shape)
    synth = None
    if args.synth_chunk > 0:
        from rnagan_b200 import gan_utils
        S = args.synth_chunk
        prof_rows = torch.randn(S, GENES, generator=torch.Generator().manual_seed(5)).to(device)
        out = torch.empty(S, SIZE, SIZE, 3, dtype=torch.float32, device=device)
        for _ in range(2):
            gan_utils.generate_tiles(G, vae, prof_rows, S, chunk=S, device=device, out=out)
        barrier()
        reps = 5
        e0.record()
        for _ in range(reps):
            gan_utils.generate_tiles(G, vae, prof_rows, S, chunk=S, device=device, out=out)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / reps
        tps = world * S * 1000.0 / ms
        synth = {"tiles_per_s": tps, "chunk": S, "ms_per_chunk": ms, "device_resident": True,
                 "tflops": FLOP_PER_TILE * S / (ms * 1e-3) / 1e12,
                 "frac_of_bf16_sustained": FLOP_PER_TILE * S / (ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
                 "note": "one synthetic RNA profile per row, train-mode BN over the chunk, latent prep included"}
        # (b) BASELINE config 4 end to end: a 100k-tile job sharded over the ranks (parallel.shard_range, no collective),
        # batch 1024, uint8 NHWC tiles written by the generator's last kernel and copied to pinned HOST memory inside the
        # timed region (gan_utils.synthesize_job); wall clock on the host because the result is a host buffer
        from rnagan_b200.parallel import shard_range
        job = args.synth_job
        if job > 0:
            first, last = shard_range(job, rank, world)
            seen = [0, 0]

            def sink(t0, tiles):
                seen[0] += tiles.shape[0]
                seen[1] += int(tiles[0, 0, 0, 0]) + int(tiles[-1, -1, -1, -1])     # touch the host data

            gan_utils.synthesize_job(G, vae, prof_rows, first, min(last, first + 2 * S), batch=S, sink=sink)   # warm
            seen[0] = 0
            barrier()
            t0 = time.perf_counter()
            gan_utils.synthesize_job(G, vae, prof_rows, first, last, batch=S, sink=sink)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert seen[0] == last - first
            dt = max_over_ranks(dt * 1000.0) / 1000.0
            synth["job"] = {"tiles": job, "batch": S, "tiles_per_s": job / dt, "seconds": dt,
                            "d2h_bytes_per_tile": SIZE * SIZE * 3, "output": "uint8 NHWC on the host (pinned ring)",
                            "sharding": f"shard_range over {world} rank(s), no collective",
                            "tiles_per_rank": last - first}
        # (c) the reference's own semantics (src/gan_utils.py:197-244): ONE profile, generator in chunks of 10 with
        # train-mode BN per chunk, fp32 NHWC numpy result on the host -- through generate_images
        if rank == 0:
            n_ref = 640
            one = prof_rows[:1].cpu()
            gan_utils.generate_images(tr, gene_exp=one, sample_size=n_ref, betavae=vae)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tiles = gan_utils.generate_images(tr, gene_exp=one, sample_size=n_ref, betavae=vae)
            dt = time.perf_counter() - t0
            synth["reference_semantics"] = {"tiles_per_s": n_ref / dt, "tiles": n_ref, "chunk": 10, "profiles": 1,
                                            "output": "fp32 NHWC numpy on the host", "shape": list(tiles.shape),
                                            "api": "rnagan_b200.gan_utils.generate_images"}
        barrier()

    # ------------------------------------------------------------------ betaVAE training (config 5 shape)
    vae_train = None
    if args.vae_steps > 0:
        from rnagan_b200 import betaVAE as bv
        torch.manual_seed(7)
        vmod = bv.betaVAE(GENES, LATENT, [6000, 4000, 2048], [4000, 6000], beta=0.0005)
        for m in vmod.modules():                       # init_weights_xavier (src/utils.py:12-15)
            if isinstance(m, torch.nn.Linear):
                torch.nn.init.xavier_uniform_(m.weight)
                m.bias.data.fill_(0.01)
        vmod = vmod.to(device).train()
        vopt = torch.optim.Adam(vmod.parameters(), lr=5e-5, weight_decay=0)
        VB = 128
        xs = torch.randn(VB, GENES, generator=torch.Generator().manual_seed(11 + rank)).to(device)
        for _ in range(3):
            bv.train_step(vmod, vopt, xs, 0.0005)
        barrier()
        e0.record()
        for _ in range(args.vae_steps):
            out3 = bv.train_step(vmod, vopt, xs, 0.0005)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / args.vae_steps
        nparam = sum(p.numel() for p in vmod.parameters())
        # algorithmic HBM bytes per step: Adam 28 B/param + bf16 weight read x2 (fwd, dgrad) + fp32 grad write + repack 6 B
        algo = nparam * (28 + 4 + 4 + 6)
        vae_train = {"steps_per_s": world * 1000.0 / ms, "ms_per_step": ms, "batch": VB, "params": nparam,
                     "hbm_gbs_algorithmic": algo / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": algo / (ms * 1e-3) / 1e9 / peaks["hbm"],
                     "losses": [float(v) for v in out3.cpu()], "device_resident": True}
        del vmod, vopt

    # ------------------------------------------------------------------ CPU baseline (oracle port) on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_port(batch=16, iters=2, scale_to=B)
    stock = None
    if rank == 0 and world == 1 and not args.no_stock:
        stock = stock_torch_gpu(device, B, args.synth_chunk if args.synth_chunk > 0 else 1024)

    if rank == 0:
        line = {
            "metric": "wgan_gd_train_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "RNA-GAN lung training (gan_run_lung.json shapes): betaVAE(19198->2048) + DCGAN "
                                   "G/D 256x256, G step + critic step + GP step per batch",
                       "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "step_unit": "one full iteration (3 optimiser steps) on a 64-sample batch; value sums ranks",
                       "l2_policy": "working set per step (>1 GB of activations) exceeds the 126 MB L2",
                       "precision": "bf16 operands / fp32 accumulate, fp32 master weights, stats, losses"},
            "e2e": e2e, "gpu_launches": int(launches), "cuda_graphs": dict(steps.GRAPH_STATS), "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu, "stock_torch_gpu": stock, "sustained": sustained, "dp_check": dp,
            "synthesis": synth, "vae_train": vae_train, "losses_finite": finite,
            "last_losses": [float(x) for x in losses_host[-1]],
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# ====================================================================================================== CPU arms
def cpu_baseline_port(batch, iters, scale_to, size=SIZE, genes=GENES):
    """The oracle (CPU restatement of the reference path) timed on this box's host cores; bounded sample."""
    from torch.optim import Adam

    from oracle import ref_oracle as O

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    lrelu, tanh = torch.nn.LeakyReLU(0.2), torch.nn.Tanh()
    torch.manual_seed(99)
    G = O.OracleGenerator(LATENT, size, 3, STEP_CH, nonlinearity=lrelu, last_nonlinearity=tanh).train()
    D = O.OracleCritic(size, 3, STEP_CH, nonlinearity=lrelu, last_nonlinearity=lrelu).train()
    vae = O.OracleVAE(genes, beta=0.005).eval()
    og = Adam(G.parameters(), lr=1e-4, betas=(0.5, 0.999))
    od = Adam(D.parameters(), lr=4e-4, betas=(0.5, 0.999))
    data = O.make_batch(batch, genes, size, 14)
    O.train_iter(G, D, og, od, vae, data)                   # warm-up
    t0 = time.perf_counter()
    for _ in range(iters):
        O.train_iter(G, D, og, od, vae, data)
    dt = (time.perf_counter() - t0) / iters
    t1 = time.perf_counter()
    tiles = O.synth_tiles(G, vae, data["rna_data"][:1], 60)
    dt_s = time.perf_counter() - t1
    return {"value": (batch / dt) / scale_to, "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": f"{iters} timed iterations (after 1 warm-up) of oracle.train_iter at batch {batch}, fp32, "
                      f"scaled to {scale_to}-sample steps (samples/s / {scale_to})",
            "sec_per_iter_at_sample_batch": dt, "samples_per_s": batch / dt,
            "synthesis_tiles_per_s": tiles.shape[0] / dt_s}


def stock_torch_gpu(device, batch, synth_chunk):
    """what the reference itself would do on this B200 -- the same modules (the oracle restates
    them in plain torch: nn.ConvTranspose2d / nn.Conv2d / nn.BatchNorm2d / nn.Linear, autograd incl. the double
    backward of the penalty, torch.optim.Adam) executed eagerly by stock PyTorch + cuDNN / cuBLASLt on the same GPU,
    same batch, same step definition.  Baseline leg (outside every timed region of our arm; may use oracle/)."""
    from torch.optim import Adam

    from oracle import ref_oracle as O

    lrelu, tanh = torch.nn.LeakyReLU(0.2), torch.nn.Tanh()
    torch.manual_seed(99)
    orig_noise = O.draw_noise
    O.draw_noise = lambda b, d: orig_noise(b, d).to(device, non_blocking=True)
    out = {"batch": batch, "modes": {}, "note": "oracle modules (= the reference's layers) on cuda, eager; losses read "
           "with .item() per train_ops like the reference; CUDA events, data resident on the device"}
    old_tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    old_bench = torch.backends.cudnn.benchmark
    try:
        data = O.make_batch(batch, GENES, SIZE, 14)
        data = {k: v.to(device) for k, v in data.items()}
        vae = O.OracleVAE(GENES, beta=0.005).eval().to(device)
        rows = torch.randn(synth_chunk, GENES, generator=torch.Generator().manual_seed(5)).to(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.backends.cudnn.benchmark = True
        for mode, (tf32, amp, cl, warm, iters) in {"fp32": (False, False, False, 3, 8),
                                                    "tf32": (True, False, False, 5, 20),
                                                    "bf16_autocast_channels_last": (True, True, True, 10, 50),
                                                    "bf16_autocast_nchw": (True, True, False, 10, 30)}.items():
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            G = O.OracleGenerator(LATENT, SIZE, 3, STEP_CH, nonlinearity=lrelu, last_nonlinearity=tanh).train().to(device)
            D = O.OracleCritic(SIZE, 3, STEP_CH, nonlinearity=lrelu, last_nonlinearity=lrelu).train().to(device)
            d = dict(data)
            if cl:
                G, D = G.to(memory_format=torch.channels_last), D.to(memory_format=torch.channels_last)
                d["image"] = d["image"].contiguous(memory_format=torch.channels_last)
            og = Adam(G.parameters(), lr=1e-4, betas=(0.5, 0.999))
            od = Adam(D.parameters(), lr=4e-4, betas=(0.5, 0.999))
            try:
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                    for _ in range(warm):
                        O.train_iter(G, D, og, od, vae, d)
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(iters):
                        losses = O.train_iter(G, D, og, od, vae, d)
                    e1.record()
                    torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                ent = {"steps_per_s": 1000.0 / ms, "ms_per_step": ms, "iters": iters, "losses": [float(v) for v in losses]}
                # synthesis: one profile per row, chunk = synth_chunk, train-mode BN, (x+1)/2 NHWC on the device
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                    def synth():
                        lat = O.latent_prep(O.draw_noise(synth_chunk, LATENT), vae.encode(rows)[0])
                        return ((G(lat).float() + 1.0) / 2.0).permute(0, 2, 3, 1).contiguous()
                    for _ in range(2):
                        synth()
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(5):
                        synth()
                    e1.record()
                    torch.cuda.synchronize()
                ent["tiles_per_s"] = synth_chunk * 5 * 1000.0 / e0.elapsed_time(e1)
            except Exception as ex:                    # e.g. an op without a bf16 double-backward kernel
                ent = {"error": f"{type(ex).__name__}: {ex}"[:300]}
            out["modes"][mode] = ent
            del G, D, og, od
            torch.cuda.empty_cache()
        ok = {k: v for k, v in out["modes"].items() if "steps_per_s" in v}
        if ok:
            best = max(ok, key=lambda k: ok[k]["steps_per_s"])
            out.update({"best_mode": best, "steps_per_s": ok[best]["steps_per_s"], "tiles_per_s": ok[best]["tiles_per_s"],
                        "dtype": "bf16" if "bf16" in best else best})
    finally:
        O.draw_noise = orig_noise
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old_tf32
        torch.backends.cudnn.benchmark = old_bench
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference itself is
    Python importing torchgan, which cannot travel to the GPU box) on all host threads."""
    if rank != 0:
        return
    from torch.optim import Adam

    from oracle import ref_oracle as O

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    lrelu, tanh = torch.nn.LeakyReLU(0.2), torch.nn.Tanh()
    torch.manual_seed(99)
    G = O.OracleGenerator(LATENT, SIZE, 3, STEP_CH, nonlinearity=lrelu, last_nonlinearity=tanh).train()
    D = O.OracleCritic(SIZE, 3, STEP_CH, nonlinearity=lrelu, last_nonlinearity=lrelu).train()
    vae = O.OracleVAE(GENES, beta=0.005).eval()
    og = Adam(G.parameters(), lr=1e-4, betas=(0.5, 0.999))
    od = Adam(D.parameters(), lr=4e-4, betas=(0.5, 0.999))
    K, W = args.steps, args.warmup
    # every timed "step" is a BOUNDED SAMPLE of the 64-sample step: one full oracle iteration (G + critic + GP) on a fixed
    # 16-sample batch (the same sample as `cpu_baseline`), so the configuration does not depend on the box; `value`
    # converts the measured samples/s to 64-sample steps/s, `ms_per_step` is the MEASURED time of one sample step
    sb = 16
    data = O.make_batch(sb, GENES, SIZE, 14)
    for _ in range(W):
        O.train_iter(G, D, og, od, vae, data)
    t0 = time.perf_counter()
    for _ in range(K):
        O.train_iter(G, D, og, od, vae, data)
    dt = (time.perf_counter() - t0) / K
    value = (sb / dt) / args.batch              # 64-sample steps per second on the host cores (rank 0 only)
    line = {
        "impl": "reference", "metric": "wgan_gd_train_steps_per_s", "value": value, "unit": "steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1000.0 * dt, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "RNA-GAN lung training (gan_run_lung.json shapes): betaVAE(19198->2048) + DCGAN "
                               "G/D 256x256, G step + critic step + GP step per batch",
                   "per_gpu_batch": args.batch, "step_unit": "one full iteration on a 64-sample batch",
                   "sample_batch": sb, "sample_fraction_of_step": sb / args.batch,
                   "ms_per_step_is": "measured seconds of one 16-sample oracle iteration (the bounded sample), not "
                                     "extrapolated; value = 16 / seconds / 64"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": f"each timed step = one oracle.train_iter at batch {sb} (fp32, {cores} threads); "
                                   f"value = samples/s / {args.batch}"},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_STDOUT_FD = None


def emit(line):
    """The ONE JSON line of this process, written to the real stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_STDOUT_FD, data)


def main():
    # Rank 0 prints exactly ONE line on stdout.  Native libraries write there too (NCCL prints its version banner on
    # file descriptor 1 when a communicator is created): point fd 1 at stderr for the whole run and keep a private
    # duplicate of the real stdout for the JSON line.
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (config 2: 64)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--synth-chunk", type=int, default=1024, help="0 disables the synthesis measurement")
    ap.add_argument("--vae-steps", type=int, default=20, help="betaVAE (config 5) training steps to time; 0 disables")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stock", action="store_true", help="skip the stock-PyTorch-on-this-GPU baseline (B0)")
    ap.add_argument("--no-dp-check", action="store_true", help="skip the data-parallel correctness check (world > 1)")
    ap.add_argument("--sustained-steps", type=int, default=300, help="extra steady-state steps after the timed region")
    ap.add_argument("--synth-job", type=int, default=100000, help="tiles of the end-to-end synthesis job (config 4)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the sm_100a path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
