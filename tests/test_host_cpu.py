"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/rnagan_b200.h declares, the
ctypes table matches the header, the product refuses to run without CUDA (no fallback), host logic of the trainer /
data-parallel plumbing (gloo, world_size 2), state_dict layout equals the reference layout."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "rnagan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rnagan_b200 import _lib
    if _lib._needs_build():
        _lib.build()
    lib = _lib.lib()
    names = _header_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rnagan_b200.h but not exported by the .so"
        assert n in _lib.SIGNATURES, f"{n} missing from the ctypes signature table"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.rg_version() >= 100
    assert lib.rg_launch_count() == 0          # loading launches nothing; no compute without a GPU


def test_dependent_launch_kernels_wait_before_memory():
    """Every kernel launched through `launch_pdl` (programmatic dependent launch, csrc/rg_host.cuh) may become resident
    while its predecessor still runs: each of them must execute griddepcontrol.wait (SASS `ACQBULK`) -- and the sources
    must place it before the kernel's first statement that can touch global memory (checked structurally: it is the
    first statement of the body, or directly follows the tile engine's prologue).  Static check, no GPU."""
    import shutil
    from rnagan_b200 import _lib
    if shutil.which("cuobjdump") is None or shutil.which("c++filt") is None:
        pytest.skip("cuobjdump / c++filt not on PATH")
    if _lib._needs_build():
        _lib.build()
    csrc = os.path.join(ROOT, "rnagan_b200", "csrc")
    text = {f: open(os.path.join(csrc, f)).read() for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))}
    kernels = set()
    for t in text.values():
        for head in re.findall(r"__global__([^;{]*)[;{]", t):
            head = re.sub(r"__launch_bounds__\([^)]*\)", "", head)
            m = re.search(r"([A-Za-z_0-9]+)\s*\(", head)
            if m:
                kernels.add(m.group(1))
    launched = sorted({m for t in text.values() for m in re.findall(r"launch_pdl\(\s*([A-Za-z_0-9]+)", t)} & kernels)
    assert "gemm_fwd_kernel" in launched and "ew_split_kernel" in launched and len(launched) >= 6
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    mangled = [f.split("\n", 1)[0].strip() for f in funcs]
    pretty = subprocess.run(["c++filt"] + mangled, capture_output=True, text=True, check=True).stdout.split("\n")
    seen = set()
    for name, body in zip(pretty, funcs):
        for k in launched:
            if re.search(r"\b" + k + r"\b", name.split("(")[0]):
                seen.add(k)
                assert "ACQBULK" in body, f"{name}: launched with the PDL attribute but has no griddepcontrol.wait"
    assert seen == set(launched), f"PDL-launched kernels not found in the library: {set(launched) - seen}"
    for k in launched:
        defs = [m for t in text.values()
                for m in re.finditer(r"__global__[^;{]*?\b" + k + r"\s*\([^{;]*\)\s*\{(.*?)griddep_wait\(\);", t, flags=re.S)]
        assert defs, f"{k}: no griddep_wait() in its definition"
        for m in defs:
            before = m.group(1)
            # nothing but declarations / the barrier+TMEM prologue may precede the wait: no global loads or stores
            touches = r"\bld8\(|\bst8\(|__ldg|\[[^\]]*\]\s*=[^=]|tma_ld_\w*\(|tma_load_\w*\(|tma_store_\w*\(|tma_prefetch_4d\("
            assert not re.search(touches, before), f"{k}: memory access before the wait"


def test_shipped_kernels_use_the_blackwell_paths():
    """Static check of the shipped library's SASS (no GPU): the tile engine issues tcgen05.mma (`UTC*MMA`), loads its
    operands with TMA (`UTMALDG`), stores through TMA (`UTMASTG`) and reads accumulators from TMEM (`LDTM`); the fused
    image-side kernels take their activation tile through a TMA box load and their image patch through cp.async
    (`LDGSTS`), contract with mma.sync (`HMMA`) and -- conv_up -- issue 18 ldmatrix per channel chunk (6 tile rows x 3
    column shifts: the fragment reuse that halves its shared-memory traffic); the NVSwitch exchange kernel reduces with
    multimem.ld_reduce (`LDGMC...ADD.F32x4`)."""
    import shutil
    from rnagan_b200 import _lib
    if shutil.which("cuobjdump") is None or shutil.which("c++filt") is None:
        pytest.skip("cuobjdump / c++filt not on PATH")
    if _lib._needs_build():
        _lib.build()
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    mangled = [f.split("\n", 1)[0].strip() for f in funcs]
    pretty = subprocess.run(["c++filt"] + mangled, capture_output=True, text=True, check=True).stdout.split("\n")

    def bodies(kernel):
        out = [b for n, b in zip(pretty, funcs) if re.search(r"\b" + kernel + r"\b", n.split("(")[0])]
        assert out, f"{kernel}: not in the library"
        return out

    def count(body, pat):
        return len(re.findall(r"\b" + pat, body))

    for b in bodies("gemm_fwd_kernel") + bodies("gemm_wgrad_kernel"):
        assert count(b, r"UTC\w*MMA") > 0 and count(b, "UTMALDG") > 0 and count(b, "LDTM") > 0
        assert count(b, r"HMMA\.") == 0, "legacy mma.sync inside the tile engine"
    assert any(count(b, "UTMASTG") > 0 for b in bodies("gemm_fwd_kernel"))
    for b in bodies("img_conv_up_kernel"):
        assert count(b, "UTMALDG") == 1 and count(b, r"HMMA\.") == 48 and count(b, "LDSM") == 18
    for b in bodies("img_conv_wgrad_kernel"):
        assert count(b, "UTMALDG") == 2 and count(b, "LDGSTS") > 0 and count(b, r"HMMA\.") > 0
    for b in bodies("img_conv_down_kernel"):
        assert count(b, "LDGSTS") > 0 and count(b, r"HMMA\.") > 0
    for b in bodies("nvls_allreduce_kernel"):      # multimem.ld_reduce: the sum happens inside the NVSwitch
        assert count(b, r"LDGMC\.E\.ADD\.F32x4") > 0


def test_size_queries_work_without_gpu():
    from rnagan_b200 import _lib
    lib = _lib.lib()
    assert lib.rg_conv_wgrad_ws_bytes(64, 64, 64, 128, 64) > 0
    assert lib.rg_gemm_tn_ws_bytes(1 << 20, 64, 64) > 0
    assert lib.rg_reduce_ws_bytes(1 << 20, 64) > 0
    assert lib.rg_adam_table_bytes(10) == 10 * 56   # 5 pointers + 4 ints (count, pitched-shadow geometry)


def test_no_cpu_fallback():
    from rnagan_b200 import dcgan, ops
    from rnagan_b200.betaVAE import betaVAE
    G = dcgan.DCGANGenerator(2048, 32, 3, 64)
    D = dcgan.DCGANDiscriminator(32, 3, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G(torch.randn(2, 2048))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        D(torch.randn(2, 3, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        betaVAE(32, 2048, [6000, 4000, 2048], [4000, 6000]).eval().encode(torch.randn(2, 32))
    with pytest.raises(ValueError, match="no CPU fallback"):
        ops.gemm_nt(torch.zeros(64, 64, dtype=torch.bfloat16), torch.zeros(64, 64, dtype=torch.bfloat16))


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under rnagan_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "rnagan_b200")
    pat = re.compile(r"^\s*(from|import)\s+\.*oracle\b|importlib.*oracle|ref_oracle", re.M)
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert not pat.search(open(os.path.join(pkg, fn)).read()), fn


def test_state_dict_layout_matches_reference_layout():
    from oracle import ref_oracle as O
    from rnagan_b200 import dcgan
    from rnagan_b200.betaVAE import betaVAE
    pairs = [(dcgan.DCGANGenerator(2048, 64, 3, 64), O.OracleGenerator(2048, 64, 3, 64)),
             (dcgan.DCGANDiscriminator(64, 3, 64), O.OracleCritic(64, 3, 64)),
             (dcgan.DCGANUpGenerator(2048, 32, 3, 64), O.OracleUpGenerator(2048, 32, 3, 64)),
             (betaVAE(100, 2048, [6000, 4000, 2048], [4000, 6000]), O.OracleVAE(100))]
    for mine, ref in pairs:
        a, b = mine.state_dict(), ref.state_dict()
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
    # SURVEY.md Appendix D spot checks
    sd = dcgan.DCGANGenerator(2048, 256, 3, 64).state_dict()
    assert tuple(sd["model.0.0.weight"].shape) == (2048, 2048, 4, 4)
    assert tuple(sd["model.6.0.weight"].shape) == (64, 3, 4, 4) and tuple(sd["model.6.0.bias"].shape) == (3,)
    sd = dcgan.DCGANDiscriminator(256, 3, 64).state_dict()
    assert tuple(sd["model.5.0.weight"].shape) == (2048, 1024, 4, 4) and tuple(sd["disc.0.weight"].shape) == (1, 2048, 4, 4)


def test_reference_constructor_errors():
    from rnagan_b200 import dcgan
    for cls in (dcgan.DCGANGenerator, dcgan.DCGANUpGenerator):
        with pytest.raises(Exception, match="Target Image Size must be at least 16\\*16 and an exact power of 2"):
            cls(out_size=24)
    with pytest.raises(Exception, match="at least 16\\*16 and an exact power of 2"):
        dcgan.DCGANDiscriminator(in_size=8)


def test_trainer_binds_train_ops_by_parameter_name():
    from rnagan_b200 import wgan_loss
    from rnagan_b200.trainer import Trainer
    import inspect
    g = list(inspect.signature(wgan_loss.WassersteinGeneratorLossVAE.train_ops).parameters)
    d = list(inspect.signature(wgan_loss.WassersteinDiscriminatorLossVAE.train_ops).parameters)
    p = list(inspect.signature(wgan_loss.WassersteinGradientPenaltyVAE.train_ops).parameters)
    assert g == ["self", "generator", "discriminator", "optimizer_generator", "device", "batch_size", "real_inputs", "labels"]
    assert d == p == ["self", "generator", "discriminator", "optimizer_discriminator", "real_inputs", "device", "labels"]

    calls = []

    class FakeG(wgan_loss.GeneratorLoss):
        def train_ops(self, generator, optimizer_generator, real_inputs, labels=None):
            calls.append(("g", generator, optimizer_generator, real_inputs))
            return 1.0

    class FakeD(wgan_loss.DiscriminatorLoss):
        def train_ops(self, discriminator, device, batch_size):
            calls.append(("d", discriminator, str(device), batch_size))
            return 2.0

    net = {"generator": {"name": torch.nn.Linear, "args": {"in_features": 2, "out_features": 2},
                         "optimizer": {"name": torch.optim.Adam, "args": {"lr": 1e-3}}},
           "discriminator": {"name": torch.nn.Linear, "args": {"in_features": 2, "out_features": 1},
                             "optimizer": {"name": torch.optim.Adam, "args": {"lr": 1e-3}}}}
    tr = Trainer(net, [FakeG(), FakeD()], device="cpu", devices=[0])
    tr.real_inputs, tr.batch_size = {"image": torch.zeros(3, 1)}, 3
    out = tr.train_iter()
    assert out == {"FakeG": 1.0, "FakeD": 2.0}
    assert calls[0][0] == "g" and calls[0][1] is tr.generator and calls[1] == ("d", tr.discriminator, "cpu", 3)
    assert tr.devices == [0]                      # unknown kwargs become attributes, like torchgan
    assert tr.loss_information["generator_iters"] == 1 and tr.loss_information["discriminator_iters"] == 1
    # deferred read-back (what Trainer.train uses): at most one iteration stays pending, reading the logs flushes it,
    # and nothing is lost or reordered
    for _ in range(3):
        assert tr.train_iter(defer=True) is None
        assert len(tr._pending) == 1
    assert tr.loss_logs == {"FakeG": [1.0] * 4, "FakeD": [2.0] * 4} and tr._pending == []
    info = tr.loss_information
    assert info == {"generator_losses": 4.0, "discriminator_losses": 8.0, "generator_iters": 4, "discriminator_iters": 4}


def test_shard_range_partitions_units():
    from rnagan_b200.parallel import shard_range
    for n, w in [(100000, 8), (10, 3), (5, 8), (0, 2)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from rnagan_b200.parallel import allreduce_mean_, init_from_env
rank, world = init_from_env(backend="gloo")
torch.manual_seed(rank)
ts = [torch.full((5,), float(rank + 1)), torch.full((3, 2), float(10 * (rank + 1))), torch.arange(4.0) * (rank + 1)]
allreduce_mean_(ts, bucket_bytes=40)       # tiny buckets: exercises the multi-bucket path
exp = [torch.full((5,), 1.5), torch.full((3, 2), 15.0), torch.arange(4.0) * 1.5]
ok = all(torch.allclose(a, b) for a, b in zip(ts, exp))
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_gradient_allreduce_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2", OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)))
             for r in range(2)]
    codes = [p.wait(timeout=180) for p in procs]
    assert codes == [0, 0]


def test_native_weight_layout_plumbing_cpu():
    """The engines keep 4x4 conv weights channels_last (physically [Cp][kh][kw][Cs]): values, shapes and state_dict
    round trips are unchanged, the flat-buffer gradient view and the Adam moments share the parameter's strides, and a
    contiguous optimiser state (an upstream checkpoint) is converted on first use."""
    import torch
    import torch.nn as nn

    from rnagan_b200 import ops
    from rnagan_b200.engine import _grad_of, _to_native
    from rnagan_b200.optim import _ensure_state
    from rnagan_b200.parallel import GradSync

    torch.manual_seed(0)
    m = nn.Sequential(nn.Conv2d(8, 16, 4, 2, 1, bias=False), nn.BatchNorm2d(16))
    w = m[0].weight
    ref = w.detach().clone()
    assert len(_to_native([w])) == 1 and ops.is_native4(w) and torch.equal(w, ref)
    assert len(_to_native([w])) == 0                      # idempotent
    assert ops.phys2d(w.detach()).shape == (16, 16 * 8) and ops.phys2d(w.detach()).is_contiguous()
    # physical order = [p][kh][kw][s]
    assert torch.equal(ops.phys2d(w.detach()).view(16, 4, 4, 8), ref.permute(0, 2, 3, 1))
    gs = GradSync(m)
    assert w.grad.stride() == w.stride() and _grad_of(w).data_ptr() == w.grad.data_ptr()
    w.grad.copy_(torch.arange(w.numel(), dtype=torch.float32).view_as(w))
    o, n = gs.offs[id(w)]
    assert torch.equal(gs.flat[o:o + n].view(16, 4, 4, 8).permute(0, 3, 1, 2), w.grad)
    # state_dict round trip through a contiguous module and back keeps the native layout
    m2 = nn.Sequential(nn.Conv2d(8, 16, 4, 2, 1, bias=False), nn.BatchNorm2d(16))
    m2.load_state_dict(m.state_dict())
    assert torch.equal(m2[0].weight, w) and m2[0].weight.is_contiguous()
    m.load_state_dict(m2.state_dict())
    assert ops.is_native4(m[0].weight) and torch.equal(m[0].weight, ref)
    # Adam moments: created with the parameter's strides; a contiguous loaded state is converted, values kept
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    st = _ensure_state(opt, w)
    assert st["exp_avg"].stride() == w.stride()
    vals = torch.randn(16, 8, 4, 4)
    opt.state[w]["exp_avg"] = vals.clone()                # contiguous, as torch.load of an upstream checkpoint gives
    st = _ensure_state(opt, w)
    assert st["exp_avg"].stride() == w.stride() and torch.equal(st["exp_avg"], vals)


def test_vae_checkpoint_with_module_prefix_loads(tmp_path):
    """Checkpoint interop: a betaVAE state_dict saved from an nn.DataParallel wrapper (`module.` keys)
    loads into the loss objects exactly like the bare layout the reference writes (src/betaVAE.py:265-278)."""
    import torch

    from rnagan_b200 import wgan_loss
    from rnagan_b200.betaVAE import betaVAE

    torch.manual_seed(1)
    vae = betaVAE(24, 2048, [6000, 4000, 2048], [4000, 6000], beta=0.005)
    bare = tmp_path / "bare.pt"
    wrapped = tmp_path / "wrapped.pt"
    torch.save(vae.state_dict(), bare)
    torch.save({"module." + k: v for k, v in vae.state_dict().items()}, wrapped)
    a = wgan_loss.WassersteinGeneratorLossVAE(str(bare), 24).betavae
    b = wgan_loss.WassersteinGeneratorLossVAE(str(wrapped), 24).betavae
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    assert not a.training and not b.training


def test_rna_scaler_matches_reference_preprocessing():
    """data.RNAScaler vs the reference's own preprocessing expressions (pandas `_get_log` + scikit-learn StandardScaler,
    src/histopathology_gan.py:133-149; transform of a held-out table, src/read_data.py:495-496; inverse_transform,
    src/betaVAE_sample.py:132) on a table with zero counts, an all-zero gene and a constant gene."""
    pd = pytest.importorskip("pandas")
    skp = pytest.importorskip("sklearn.preprocessing")
    from rnagan_b200.data import RNAScaler, log_expression
    rng = np.random.default_rng(0)
    x = rng.gamma(2.0, 50.0, size=(37, 211))
    x[rng.random(x.shape) < 0.2] = 0.0
    x[:, 5] = 0.0
    x[:, 7] = 3.0
    held = rng.gamma(2.0, 50.0, size=(9, 211))
    held[rng.random(held.shape) < 0.2] = 0.0

    def _get_log(col):
        col = np.log(col.replace(0, np.nan))
        return col.replace(np.nan, 0)

    cols = [f"rna_{i}" for i in range(x.shape[1])]
    ref_log = pd.DataFrame(x, columns=cols).apply(_get_log).values
    ref_held_log = pd.DataFrame(held, columns=cols).apply(_get_log).values
    sk = skp.StandardScaler()
    ref = sk.fit_transform(ref_log)
    mine = RNAScaler()
    got = mine.fit_transform(x)
    assert np.array_equal(log_expression(x), ref_log)                     # the log step is exact
    assert np.abs(got - ref).max() <= 1e-12 and np.abs(mine.scale_ - sk.scale_).max() <= 1e-12
    assert np.array_equal(got.astype(np.float32)[:, 5], np.zeros(37, np.float32)) and mine.scale_[5] == 1.0
    assert mine.scale_[7] == 1.0 and np.abs(got[:, 7]).max() <= 1e-12     # constant gene: scale 1, like scikit-learn
    assert np.abs(mine.transform(held) - sk.transform(ref_held_log)).max() <= 1e-12
    assert np.abs(mine.inverse_transform(got) - sk.inverse_transform(ref)).max() <= 1e-12
    with pytest.raises(RuntimeError):
        RNAScaler().transform(x)
    with pytest.raises(ValueError):
        mine.transform(x[:, :10])


def test_peer_exchange_schedule_simulated():
    """The copy-engine gradient exchange (parallel.PeerExchange, RG_DP_EXCHANGE=ce) as data movement only: W virtual
    ranks share a list of flat buffers (what symmetric memory gives each rank), phase 1 (pull my slice from every peer,
    sum in rank order) runs on every rank, then phase 2 (pull every peer's reduced slice).  Afterwards every rank holds
    the SUM on the bucket -- bit-identical across ranks -- and nothing outside the bucket moved.  Ragged buckets, buckets
    smaller than the world and world sizes 1..8."""
    from rnagan_b200.parallel import exchange_gather, exchange_pull_reduce, slice_bounds

    def copy(dst, src):
        dst.copy_(src)

    def reduce(stage, world, n, out):                   # stand-in for ops.slices_sum: ascending rank order
        acc = stage[0, :n].clone()
        for r in range(1, world):
            acc += stage[r, :n]
        out.copy_(acc)

    g = torch.Generator().manual_seed(0)
    for world in (1, 2, 3, 4, 8):
        for numel, lo, n in ((4096, 0, 4096), (4096, 128, 1000), (4096, 4000, 96), (64, 8, 8), (64, 0, 4)):
            b = slice_bounds(n, world)
            assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b, b[1:]))
            assert all((hi - lo_) % 4 == 0 for lo_, hi in b[:-1]) and all(lo_ % 4 == 0 for lo_, hi in b if hi > lo_)
            flats = [torch.randn(numel, generator=g) for _ in range(world)]
            before = [f.clone() for f in flats]
            want = torch.zeros(n)
            for r in range(world):                      # the order rg_slices_sum uses
                want = want + before[r][lo:lo + n] if r else before[r][lo:lo + n].clone()
            per = max(b[0][1] - b[0][0], 4)
            stages = [torch.full((world, per), float("nan")) for _ in range(world)]
            for r in range(world):
                exchange_pull_reduce(r, world, flats, stages[r], lo, n, copy, reduce)
            for r in range(world):
                exchange_gather(r, world, flats, lo, n, copy)
            for r in range(world):
                assert torch.equal(flats[r][lo:lo + n], want), (world, numel, lo, n, r)
                assert torch.equal(flats[r][:lo], before[r][:lo]) and torch.equal(flats[r][lo + n:], before[r][lo + n:])


_GRID_SEEN = []


class _GridTinyG(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(4, 3 * 8 * 8)

    def sampler(self, sample_size, device):
        return [torch.randn(sample_size, 4, device=device)]

    def forward(self, z):
        _GRID_SEEN.append((self.training, z.clone()))
        return torch.tanh(self.lin(z)).view(-1, 3, 8, 8)


def _grid_losses():
    """Picklable stand-in loss classes (save_model pickles the loss objects), created once at module level."""
    from rnagan_b200 import wgan_loss
    g = globals()
    if "_GridFakeG" not in g:
        class _GridFakeG(wgan_loss.GeneratorLoss):
            def train_ops(self, generator, optimizer_generator):
                return 1.0

        class _GridFakeD(wgan_loss.DiscriminatorLoss):
            def train_ops(self, discriminator, optimizer_discriminator):
                return 2.0

        for c in (_GridFakeG, _GridFakeD):
            c.__qualname__ = c.__name__
            g[c.__name__] = c
    return g["_GridFakeG"], g["_GridFakeD"]


def test_sample_grid_matches_torchvision_and_trainer_writes_it(tmp_path):
    """image_grid (the per-epoch sample grid torchgan's Trainer writes to `recon` [tg]) reproduces
    torchvision.utils.save_image(..., normalize=True) pixel for pixel, and Trainer.train writes
    `{recon}/epoch{N}_{model}.png` from an eval-mode generator on a fixed test noise after the checkpoint."""
    from rnagan_b200 import wgan_loss
    from rnagan_b200.image_grid import save_image_grid
    from rnagan_b200.trainer import Trainer
    Image = pytest.importorskip("PIL.Image")
    tv = pytest.importorskip("torchvision.utils")
    g = torch.Generator().manual_seed(0)
    for shape, nrow in (((64, 3, 32, 32), 8), ((10, 3, 16, 24), 8), ((5, 1, 8, 8), 3), ((1, 3, 4, 4), 8)):
        x = torch.randn(shape, generator=g)
        a, b = str(tmp_path / "a.png"), str(tmp_path / "b.png")
        save_image_grid(x, a, nrow=nrow)
        tv.save_image(x, b, nrow=nrow, normalize=True)
        assert np.array_equal(np.array(Image.open(a)), np.array(Image.open(b))), shape

    seen = _GRID_SEEN
    seen.clear()
    TinyG = _GridTinyG
    FakeG, FakeD = _grid_losses()
    net = {"generator": {"name": TinyG, "args": {}, "optimizer": {"name": torch.optim.Adam, "args": {"lr": 1e-3}}},
           "discriminator": {"name": torch.nn.Linear, "args": {"in_features": 2, "out_features": 1},
                             "optimizer": {"name": torch.optim.Adam, "args": {"lr": 1e-3}}}}
    tr = Trainer(net, [FakeG(), FakeD()], device="cpu", epochs=2, sample_size=6, nrow=4,
                 checkpoints=str(tmp_path / "ckpt" / "gan"), recon=str(tmp_path / "images"))
    tr([{"image": torch.zeros(2, 3, 8, 8)}] * 3)
    for e in (1, 2):
        img = np.array(Image.open(tmp_path / "images" / f"epoch{e}_generator.png"))
        assert img.shape == (2 * 10 + 2, 4 * 10 + 2, 3)              # 6 tiles, 4 per row, 2 px padding
    assert os.path.exists(tmp_path / "ckpt" / "gan0.model") and os.path.exists(tmp_path / "ckpt" / "gan1.model")
    assert len(seen) == 2 and not seen[0][0] and torch.equal(seen[0][1], seen[1][1])   # eval mode, same noise
    assert tr.loss_logs["_GridFakeG"] == [1.0] * 6


def test_reference_written_checkpoint_loads(tmp_path):
    """a `{dir}{k}.model` in torchgan's save_model layout whose `loss_objects` are instances of the
    REFERENCE's classes pickled under its top-level module names (`wgan_loss`, `betaVAE`; fixture written by
    oracle/make_upstream_ckpt.py from the reference's own modules) loads through compat.load_checkpoint / Trainer.load_model
    without the reference on the path: classes resolve to this package's, the instance state survives, the state_dicts
    load strictly into this package's modules and into torch.optim.Adam."""
    from rnagan_b200 import betaVAE as bv
    from rnagan_b200 import compat, dcgan, wgan_loss
    from rnagan_b200.trainer import Trainer
    path = os.path.join(ROOT, "tests", "golden", "upstream_style_ckpt.model")
    with pytest.raises(Exception):                 # plain torch.load cannot resolve the reference's module names
        import importlib
        if importlib.util.find_spec("wgan_loss") is not None:
            raise ModuleNotFoundError("reference importable here: the negative check does not apply")
        torch.load(path, weights_only=False)
    ck = compat.load_checkpoint(path, map_location="cpu")
    exp = ck["_expected"]
    assert set(exp["loss_classes"].values()) == {"wgan_loss.WassersteinGeneratorLossVAE",
                                                 "wgan_loss.WassersteinDiscriminatorLossVAE",
                                                 "wgan_loss.WassersteinGradientPenaltyVAE"}
    assert list(ck["loss_objects"]) == list(exp["loss_classes"])
    for name, obj in ck["loss_objects"].items():
        assert type(obj) is getattr(wgan_loss, name)
        assert type(obj.betavae) is bv.betaVAE and type(obj.betavae.encoder) is bv.RNAEncoder
        assert not obj.betavae.training and obj._ckpt_key[0] == "unpickled"
        got = float(sum(p.double().sum() for p in obj.betavae.state_dict().values()))
        assert abs(got - exp["vae_param_sum"]) <= 1e-9 * max(1.0, abs(exp["vae_param_sum"]))
    g0 = ck["loss_objects"]["WassersteinGeneratorLossVAE"]
    assert g0.reduction == exp["reduction_attr"] and g0.override_train_ops == exp["override_attr"] == exp["dims"]["feats"]
    d = exp["dims"]
    net = {
        "generator": {"name": dcgan.DCGANGenerator,
                      "args": {"encoding_dims": d["z"], "out_channels": 3, "step_channels": d["step"], "out_size": d["size"]},
                      "optimizer": {"name": torch.optim.Adam, "args": {"lr": 1e-4, "betas": (0.5, 0.999)}}},
        "discriminator": {"name": dcgan.DCGANDiscriminator,
                          "args": {"in_size": d["size"], "in_channels": 3, "step_channels": d["step"]},
                          "optimizer": {"name": torch.optim.Adam, "args": {"lr": 4e-4, "betas": (0.5, 0.999)}}},
    }
    tr = Trainer(net, list(ck["loss_objects"].values()), device="cpu", checkpoints=str(tmp_path / "gan"))
    tr.load_model(load_path=path)
    assert tr.start_epoch == 1 and tr.loss_information["generator_iters"] == 1
    got = float(sum(p.double().sum() for p in tr.generator.state_dict().values()))
    assert abs(got - exp["generator_param_sum"]) <= 1e-9 * max(1.0, abs(exp["generator_param_sum"]))
    st = tr.optimizer_generator.state_dict()["state"]
    assert len(st) == len(list(tr.generator.parameters())) and all(float(s["step"]) == 1.0 for s in st.values())
    # and the round trip: what this package writes, it reads (loss objects included)
    again = compat.load_checkpoint(tr.save_model(0), map_location="cpu")
    assert type(again["loss_objects"]["WassersteinGradientPenaltyVAE"]) is wgan_loss.WassersteinGradientPenaltyVAE


def test_partition_properties_hypothesis():
    """Property tests of the two partitioners the multi-GPU paths rest on: parallel.shard_range (tile synthesis, whole
    units) and parallel.slice_bounds (gradient exchange, 4-float aligned slices)."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    from rnagan_b200.parallel import shard_range, slice_bounds

    @hyp.settings(max_examples=300, deadline=None)
    @hyp.given(st.integers(0, 10 ** 7), st.integers(1, 64))
    def shards(n, world):
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)

    @hyp.settings(max_examples=300, deadline=None)
    @hyp.given(st.integers(1, 10 ** 7).map(lambda k: 4 * k), st.integers(1, 64))
    def slices(n, world):
        b = slice_bounds(n, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
        assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
        lens = [hi - lo for lo, hi in b]
        assert all(v % 4 == 0 for v in lens) and all(lo % 4 == 0 for lo, hi in b)
        assert max(lens) == lens[0] and max(lens) - 4 * ((n // 4 + world - 1) // world) == 0
        assert lens == sorted(lens, reverse=True)            # full slices first, then one short one, then empties

    shards()
    slices()


def test_gradsync_bucketing_merges_adjacent_layers():
    """GradSync with bucketing on (the copy-engine exchange mode): layers announced in backward order are merged into
    contiguous buckets of at least `bucket_floats`, every float of every gradient is reduced exactly once per step, and
    nothing is reduced before its layer was announced.  The reductions are recorded instead of executed."""
    from rnagan_b200.parallel import GradSync

    class Rec(GradSync):
        world = staticmethod(lambda: 2)

        def _reduce_range(self, lo, hi):
            self.calls.append((lo, hi))

            class H:
                def wait(self_inner):
                    return None
            return H()

    sizes = [3, 130, 520, 2100, 8400, 33500, 31]             # layer sizes in the critic's proportions, first -> last
    net = torch.nn.ModuleList([torch.nn.Linear(n, 1, bias=False) for n in sizes])
    gs = Rec(net)
    gs.calls = []
    gs.bucket_floats = 4000
    params = list(net.parameters())
    for p in reversed(params):                                # backward order: last layer first
        gs.layer_done(p)
    assert gs.finish() == 0.5
    total = gs.flat.numel()
    covered = torch.zeros(total, dtype=torch.int32)
    for lo, hi in gs.calls:
        covered[lo:hi] += 1
    for p in params:
        o, n = gs.offs[id(p)]
        assert bool((covered[o:o + n] == 1).all())
    assert int(covered.max()) == 1
    # [31 + 33500] reaches the threshold, 8400 alone does, the four small layers leave together at finish()
    assert [hi - lo >= 4000 for lo, hi in gs.calls] == [True, True, False] and len(gs.calls) == 3
    # second step, announcements out of order: non-adjacent ranges are not merged, still exactly-once coverage
    gs.calls = []
    for i in (6, 0, 5, 2, 1, 4, 3):
        gs.layer_done(params[i])
    gs.finish()
    covered.zero_()
    for lo, hi in gs.calls:
        covered[lo:hi] += 1
    for p in params:
        o, n = gs.offs[id(p)]
        assert bool((covered[o:o + n] == 1).all())
    assert int(covered.max()) == 1
    # bucketing off (the NCCL default): one reduction per announcement
    gs.calls, gs.bucket_floats = [], 0
    for p in reversed(params):
        gs.layer_done(p)
    gs.finish()
    assert len(gs.calls) == len(params)


def test_frechet_distance_matches_reference_formula():
    """fid.frechet_distance (symmetric eigen formulation) vs the reference's formula evaluated with scipy's general
    matrix square root (src/fid.py:147-163), on full-rank, rank-deficient (N < D) and identical feature sets; and
    activation_statistics vs np.mean / np.cov(rowvar=False) (src/fid.py:110-111)."""
    linalg = pytest.importorskip("scipy.linalg")
    from rnagan_b200.fid import activation_statistics, fid_from_features, frechet_distance
    rng = np.random.default_rng(0)
    for n1, n2, d in ((500, 400, 32), (40, 50, 64), (300, 300, 8)):
        a = rng.normal(size=(n1, d)) @ rng.normal(size=(d, d)) * 0.3 + rng.normal(size=d)
        b = rng.normal(size=(n2, d)) @ rng.normal(size=(d, d)) * 0.3
        mu1, s1 = activation_statistics(a)
        mu2, s2 = activation_statistics(b)
        assert np.allclose(mu1, a.mean(0)) and np.allclose(s1, np.cov(a, rowvar=False))
        covmean = linalg.sqrtm(s1.dot(s2))
        covmean = covmean.real if np.iscomplexobj(covmean) else covmean
        ref = (mu1 - mu2).dot(mu1 - mu2) + np.trace(s1) + np.trace(s2) - 2 * np.trace(covmean)
        got = frechet_distance(mu1, s1, mu2, s2)
        assert abs(got - ref) <= 1e-6 * max(1.0, abs(ref)), (n1, n2, d, got, ref)
        assert abs(fid_from_features(b, a) - got) <= 1e-9 * max(1.0, abs(got))
    same = frechet_distance(mu1, s1, mu1, s1)
    assert abs(same) <= 1e-8 * np.trace(s1)
    with pytest.raises(ValueError):
        frechet_distance(mu1, s1, mu2[:-1], s2)


def test_fid_pipeline_with_inception_architecture():
    """src/fid.py:33-235 end to end on the CPU with RANDOM Inception weights (the pretrained file cannot be downloaded
    here): the extractor refuses to run unweighted unless told to, loads a state_dict with torchvision's own key layout,
    produces [N, 2048] activations of Mixed_7c, and calculate_fid is ~0 for identical stacks and > 0 for different ones."""
    import numpy as np
    import pytest
    import torch
    from rnagan_b200 import fid
    with pytest.raises(ValueError):
        fid.PartialInceptionNetwork()
    torch.manual_seed(0)
    net = fid.PartialInceptionNetwork(allow_random_weights=True)
    for m in net.inception_network.modules():                  # init_weights=False leaves default init: make it tame
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
    state = net.inception_network.state_dict()
    assert "Mixed_7c.branch_pool.conv.weight" in state and "Conv2d_1a_3x3.conv.weight" in state
    net2 = fid.PartialInceptionNetwork(weights=state)           # same key layout as torchvision's checkpoint
    rng = np.random.default_rng(0)
    a = rng.random((6, 40, 40, 3), dtype=np.float32)
    b = (rng.random((6, 40, 40, 3)) * 255).astype(np.uint8)
    x = fid.preprocess_images(a)
    assert x.shape == (6, 3, 299, 299) and x.dtype == torch.float32 and 0.0 <= float(x.min()) and float(x.max()) <= 1.0
    act = fid.get_activations(x, 4, device="cpu", network=net2)
    assert act.shape == (6, 2048) and np.isfinite(act).all()
    same = fid.calculate_fid(a, a, False, 4, device="cpu", network=net2)
    diff = fid.calculate_fid(a, b, True, 4, device="cpu", network=net2)
    assert abs(same) <= 1e-6 * max(1.0, abs(diff)) and diff > 0
