"""2-rank data-parallel equivalence worker (launched by tests/test_dp_gpu.py with torchrun):
with -bnSync the two-GPU step must equal the single-GPU step on the concatenated batch -- same
log-probabilities per shard and the same (all-reduced) parameter gradients."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multigrid-neural-architectures_b200"))
import numpy as np, torch, torch.distributed as dist
from mgconv import builders as B, multigpu

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
rng = np.random.default_rng(11)
Bg = 16
x = torch.from_numpy(rng.standard_normal((Bg, 3, 32, 32)).astype(np.float32))
t = torch.from_numpy(rng.integers(1, 101, Bg))

def run(nGPU, xs, ts, bnSync):
    torch.manual_seed(5)
    model = B.cifar_rnmg.createModel(B.Opt(nLayer=1, nGPU=nGPU, bnSync=bnSync))
    (model.model if nGPU > 1 else model).precision = precision
    model.cuda()
    params, grads = model.getParameters()
    crit = B.cifar_rnmg.createCriterion()
    model.zeroGradParameters()
    out, err = B.cifar_rnmg.ftrain(xs.cuda(), ts.cuda(), model, crit)
    torch.cuda.synchronize()
    return out.clone(), float(err), grads.clone()

lo, hi = multigpu.shard_range(Bg, world, rank)
out2, err2, g2 = run(world, x[lo:hi], t[lo:hi], True)          # data parallel, sync-BN
out1, err1, g1 = run(1, x, t, False)                            # single device, whole batch
tol = 1e-4 if precision == "fp32" else 2e-2
rel = lambda a, b: float((a - b).norm() / b.norm())
e_out, e_g = rel(out2, out1[lo:hi]), rel(g2, g1)
assert e_out <= tol, ("log-probabilities", e_out)
assert e_g <= 10 * tol, ("gradients", e_g)
# local-BN mode (the reference's DataParallelTable): statistics per shard -> differs from the whole-batch run
out3, _, _ = run(world, x[lo:hi], t[lo:hi], False)
e_local = rel(out3, out1[lo:hi])
assert e_local > max(1e-3, 5 * e_out), ("per-replica BN must differ from whole-batch BN", e_local, e_out)
dist.barrier()
print(f"dp-equivalence ok rank {rank} [{precision}]: sync-BN vs single GPU: out {e_out:.2e} grad {e_g:.2e}; per-replica BN differs by {e_local:.2e}", flush=True)
dist.destroy_process_group()
