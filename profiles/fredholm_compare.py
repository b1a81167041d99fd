"""torchrun -n P: time MPIFredholm1 (config 5 slice shapes) in its two multi-GPU modes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import numpy as np, torch
import pylops_mpi_b200 as pm
comm = pm.get_comm_world(); rank, size = comm.Get_rank(), comm.Get_size()
nsl, ns, nr, nv = 64, 256, 256, 64
G = torch.randn(nsl, ns, nr, device="cuda", dtype=torch.complex64)
xm = pm.DistributedArray(global_shape=nsl * size * nr * nv, partition=pm.Partition.BROADCAST, dtype=np.complex64)
xm.local_array.normal_()
res = {}
for name, kw in (("chunked_nccl", {"fused": False}), ("fused_peer", {"fused": True})):
    Fr = pm.MPIFredholm1(G, nz=nv, dtype=np.complex64, **kw)
    for _ in range(5): y = Fr.matvec(xm)
    torch.cuda.synchronize(); comm.Barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): y = Fr.matvec(xm)
    e1.record(); torch.cuda.synchronize()
    res[name] = comm.allreduce(e0.elapsed_time(e1) / 20 * 1e3, "max")
    ref = y.local_array.clone() if name == "chunked_nccl" else ref
    if name == "fused_peer":
        assert torch.allclose(y.local_array, ref, rtol=1e-4, atol=1e-3), "fused != chunked"
# scalar all-reduce latency: NCCL vs peer-memory mailbox
from pylops_mpi_b200.Distributed import allreduce_
buf = torch.ones(2, dtype=torch.float64, device="cuda")
lat = {}
for mode in ("peer", "nccl"):
    os.environ["B2_PEER_ALLREDUCE"] = "1" if mode == "peer" else "0"
    for _ in range(10): allreduce_(comm, buf)
    torch.cuda.synchronize(); comm.Barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): allreduce_(comm, buf)
    e1.record(); torch.cuda.synchronize()
    lat[mode] = comm.allreduce(e0.elapsed_time(e1) / 200 * 1e3, "max")
vec = {}
for nel in (10 ** 3, 10 ** 4, 65536):
    vb = torch.ones(nel, dtype=torch.float32, device="cuda")
    for mode in ("peer", "nccl"):
        os.environ["B2_PEER_ALLREDUCE"] = "1" if mode == "peer" else "0"
        for _ in range(10): allreduce_(comm, vb)
        torch.cuda.synchronize(); comm.Barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100): allreduce_(comm, vb)
        e1.record(); torch.cuda.synchronize()
        vec[f"{nel}_{mode}"] = round(comm.allreduce(e0.elapsed_time(e1) / 100 * 1e3, "max"), 2)
chk = torch.full((3,), float(rank + 1), dtype=torch.float64, device="cuda")
os.environ["B2_PEER_ALLREDUCE"] = "1"
allreduce_(comm, chk)
assert abs(chk[0].item() - size * (size + 1) / 2) < 1e-12
if rank == 0:
    print("FREDHOLM_US", size, res, "SCALAR_ALLREDUCE_US", lat, "VECTOR_ALLREDUCE_F32_US", vec)
