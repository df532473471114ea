import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from after_b200.engine import Engine
eng = Engine()
M, N, K = 6144, 1536, 512
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
for _ in range(3):
    eng.debug_gemm(A, W, b, sys.argv[1] if len(sys.argv) > 1 else "fp32")
torch.cuda.synchronize()
eng.close()
