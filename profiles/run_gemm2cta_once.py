import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pylops_mpi_b200 import _lib as L
m = n = k = 8192
A = (torch.randn(m, k, device="cuda") / 90).to(torch.bfloat16)
B = (torch.randn(k, n, device="cuda") / 90).to(torch.bfloat16)
C = torch.empty(m, n, device="cuda")
for _ in range(3):
    L.check(L.lib.b2_gemm_bf16(L.ctx(), A.data_ptr(), k, B.data_ptr(), n, C.data_ptr(), n, m, n, k, 0, 0, L.stream()))
torch.cuda.synchronize()
