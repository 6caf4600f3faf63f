import sys, torch, numpy as np
sys.path.insert(0, '.')
from tests.test_gpu_kernels import _fc_case
for (R,K,M,T,t0,T_live,mode,cg) in [(50,192,256,8,0,6,1,1),(50,192,256,8,0,6,1,2),(45,128,256,12,0,3,1,2),(45,128,256,12,0,3,1,1),(45,128,256,12,0,7,1,1)]:
    z, w, trains, dump = _fc_case(R,K,M,T,t0,T_live,mode,cg)
    from tests._util_cpu import split_reconstruct
    ref = torch.einsum("trk,mk->trm", z.double(), split_reconstruct(w,1))
    err = (dump.double()-ref).abs()
    bad = err > 1e-3
    print("case", (R,K,M,T,t0,T_live,mode,cg), "max err", err.max().item(), "bad frac", bad.float().mean().item())
    print("  bad by t:", bad.any(dim=2).any(dim=1).tolist())
    print("  bad rows:", bad.any(dim=2).any(dim=0).nonzero().flatten().tolist())
    print("  bad cols count:", bad.any(dim=0).any(dim=0).sum().item(), "of", M)
