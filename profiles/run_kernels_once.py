"""Launch each non-stencil hot kernel a few times at its BASELINE size (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import pylops_mpi_b200 as pm
from pylops_mpi_b200 import _lib as L

torch.manual_seed(0)
m = n = k = 8192
A = (torch.randn(m, k, device="cuda") / 90).to(torch.bfloat16)
B = (torch.randn(k, n, device="cuda") / 90).to(torch.bfloat16)
C = torch.empty(m, n, device="cuda")
for _ in range(2):
    L.check(L.lib.b2_gemm_bf16(L.ctx(), A.data_ptr(), k, B.data_ptr(), n, C.data_ptr(), n, m, n, k, 0, 0, L.stream()))
torch.cuda.synchronize()
del A, B, C
a = pm.DistributedArray(global_shape=1 << 28, dtype=np.float32); a.local_array.normal_()
b = pm.DistributedArray(global_shape=1 << 28, dtype=np.float32); b.local_array.normal_()
for _ in range(2):
    a._dot_device(b); a._norm_device(2)
A2 = torch.randn(32768, 8192, device="cuda")
op = pm.MatrixMult(A2)
x1, x2 = torch.randn(8192, device="cuda"), torch.randn(32768, device="cuda")
for _ in range(2):
    op.matvec(x1); op.rmatvec(x2)
G = torch.randn(64, 256, 256, device="cuda", dtype=torch.complex64)
Fr = pm.MPIFredholm1(G, nz=64, dtype=np.complex64)
xm = pm.DistributedArray(global_shape=64 * 256 * 64, partition=pm.Partition.BROADCAST, dtype=np.complex64)
xm.local_array.normal_()
for _ in range(2):
    Fr.matvec(xm)
torch.cuda.synchronize()
print("done")
