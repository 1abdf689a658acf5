"""The one-process-per-GPU driver of a sharded factorisation (`ShardedLDLRank`: torch.distributed all-gathers of the
packed contributions between the phases) with world_size 2 and 3 on the CPU: gloo backend, every rank running the
CUDA-on-CPU emulated build (tests/emu), i.e. the product's kernels and phase logic with host memory as device memory.
On GPUs the same class runs over NCCL."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
os.environ["CLARABEL_EMU"] = "1"
import clarabel_rs_b200 as cb
cb.pkg._LIBPATH = ROOT + "/tests/emu/libclarabel_emu_full.so"
from helpers import small_kkt
from oracle import QDLDL
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N, cp, rv, nz, ds = small_kkt(700, 1100, seed=4, window=30, k=3)
s = cb.ShardedLDLRank(N, cp, rv, nz, ds, ordering=cb.ORDER_ND, nd_leaf=50)
assert s.refactor()
f = QDLDL((N, N), cp, rv, nz, s.solver.perm(), dsigns=ds, regularize_eps=1e-13, regularize_delta=2e-7)
rng = np.random.default_rng(0)
for _ in range(2):
    b = rng.standard_normal(N)
    x, xo = s.solve(b), f.solve(b)
    err = float(np.max(np.abs(x - xo)))
    assert err <= 1e-9 * max(1.0, float(np.max(np.abs(xo)))), err
owned = int(s._L.cldl_shard_count(s.solver._h, 2, rank))
assert 0 < owned < N
print("RANK_OK %d/%d owned_x=%d err=%.2e" % (rank, world, owned, err), flush=True)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ldl_over_torch_distributed_gloo(world, tmp_path):
    lib = os.path.join(ROOT, "tests", "emu", "libclarabel_emu_full.so")
    assert os.path.exists(lib), "tests/emu/libclarabel_emu_full.so missing: run `make`"
    script = tmp_path / "worker.py"
    script.write_text("ROOT = %r\n" % ROOT + WORKER)
    port = 29500 + (os.getpid() % 500) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    assert out.stdout.count("RANK_OK") == world, out.stdout[-2000:]


IPM_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
os.environ["CLARABEL_EMU"] = "1"
import clarabel_rs_b200 as cb
cb.pkg._LIBPATH = ROOT + "/tests/emu/libclarabel_emu_full.so"
import oracle
from helpers import workloads
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
pr = workloads.random_sparse_qp(n=400, m=800, nnz_per_row=4, seed=2, window=30)
dev = cb.CudaSolver(pr["P"], pr["q"], pr["A"], pr["b"], pr["cones"], ordering=cb.ORDER_ND, nd_leaf=60, shard=(world, rank))
r = dev.solve()
ora = oracle.IPM(pr["P"], pr["q"], pr["A"], pr["b"], pr["cones"])
ora.set_perm(dev.kkt_perm()); ro = ora.solve()
assert r["status"] == ro["status"] == "Solved", (r["status"], ro["status"])
assert r["iterations"] == ro["iterations"], (r["iterations"], ro["iterations"])
assert np.max(np.abs(r["x"] - ro["x"])) <= 1e-6 * max(1.0, np.max(np.abs(ro["x"])))
# every rank ran the same iterations on identical data: gather x and compare bit for bit
import torch
xs = [torch.zeros(r["x"].size, dtype=torch.float64) for _ in range(world)]
dist.all_gather(xs, torch.from_numpy(r["x"].copy()))
assert all(torch.equal(xs[0], t) for t in xs)
print("IPM_OK %d/%d it=%d exchanges=%d" % (rank, world, r["iterations"], dev._transport.calls), flush=True)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_interior_point_solve_over_gloo(world, tmp_path):
    """the whole cipm_* driver with its LDL split over `world` ranks: identical iterations on every rank, the oracle's
    status / iteration count / solution"""
    script = tmp_path / "worker.py"
    script.write_text("ROOT = %r\n" % ROOT + IPM_WORKER)
    port = 29600 + (os.getpid() % 300) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    assert out.stdout.count("IPM_OK") == world, out.stdout[-2000:]
