"""The switch-side (NVLS) sums -- cngi_b200_multimem_reduce_f32 / cngi_b200_multimem_allreduce_f64 through
distributed.SymmetricCollectives -- against NCCL on two GPUs.  Skipped on boxes with one GPU (the driver's test box);
tools/probe_collectives.py runs the same comparison inside its multi-GPU measurements."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from cngi_prototype_b200 import distributed as D
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
D.init_nccl(dev)
if not D.SymmetricCollectives.supported(dev):
    print("no multicast"); dist.destroy_process_group(); sys.exit(0)
sym = D.SymmetricCollectives(dev)
n = 1024
g = torch.Generator(device=dev).manual_seed(3 + rank)
a = sym.empty((1, 2, n, n), torch.complex64)
torch.view_as_real(a).normal_(generator=g)
ref = a.clone()
dist.reduce(torch.view_as_real(ref), 1)
sym.reduce_grid(a, 1).wait()
torch.cuda.synchronize()
if rank == 1:
    assert float((a - ref).abs().max() / ref.abs().max()) < 1e-6
d = sym.empty((1, 2, n, n), torch.float64)
d.normal_(generator=g)
keep = d[:, 1:].clone()
dref = d[:, :1].clone()
dist.all_reduce(dref)
sym.allreduce_density(d, n * n).wait()
torch.cuda.synchronize()
assert float((d[:, :1] - dref).abs().max() / dref.abs().max()) < 1e-14 and torch.equal(d[:, 1:], keep)
dist.barrier()
if rank == 0:
    print("multimem ok")
dist.destroy_process_group()
''' % ROOT


def test_multimem_reduce_and_allreduce_match_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one NVSwitch")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29583", str(script)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=300)
    assert r.returncode == 0 and ("multimem ok" in r.stdout or "no multicast" in r.stdout), r.stdout[-3000:]
