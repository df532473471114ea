"""In-kernel %globaltimer trace of the first tiles of the CTA-pair tap-GEMM (debug build only:
    python -m after_b200.build --debug --force && AFTER_DEBUG_TRACE=1 python scripts/trace_gemm.py [M N K ...])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["AFTER_DEBUG_TRACE"] = "1"
os.environ["AFTER_B200_DEBUG_BUILD"] = "1"
import torch
from after_b200.engine import Engine
eng = Engine()
args = [int(a) for a in sys.argv[1:]]
shapes = [tuple(args[i:i + 3]) for i in range(0, len(args), 3)] or [(6144, 1536, 512), (6144, 512, 1536)]
for (M, N, K) in shapes:
    for prec in ["fp32", "bf16"]:
        A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
        eng.debug_gemm(A, W, b, prec)
        torch.cuda.synchronize()
eng.close()
