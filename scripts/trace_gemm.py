import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["AFTER_DEBUG_TRACE"] = "1"
import torch
from after_b200.engine import Engine
eng = Engine()
for (M, N, K) in [(6144, 1536, 512), (6144, 512, 1536)]:
    for prec in ["fp32", "bf16"]:
        A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
        eng.debug_gemm(A, W, b, prec)
eng.close()
