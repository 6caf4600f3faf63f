"""Does tcgen05.mma's 128-byte swizzle follow ABSOLUTE shared-memory address bits?  The fc kernel's producers
write the spike tile with chunk' = chunk ^ ((addr >> 7) & 7) at a row shift / 8-row-group stride given by
SNN_DBG_SWIZZLE=shift,sbo,base_offset and the MMA descriptor uses the same start / SBO / base-offset field."""
import os, sys, subprocess
if len(sys.argv) > 1:
    sys.path.insert(0, '.')
    import torch
    from tests.test_gpu_kernels import _fc_case
    from tests._util_cpu import split_reconstruct
    for cg in (1, 2):
        z, w, trains, dump = _fc_case(50, 192, 256, 8, 0, 6, 1, cg)
        ref = torch.einsum("trk,mk->trm", z.double(), split_reconstruct(w, 1))
        print("   cg", cg, "max err %.3g" % (dump.double() - ref).abs().max().item(), flush=True)
else:
    for cfg in ["0,1024,0", "3,1024,0", "3,1024,3", "0,1280,0", "2,1280,0", "2,1280,2", "5,1152,5", "5,1152,0", "1,2048,1", "1,2048,0"]:
        print("SNN_DBG_SWIZZLE =", cfg, flush=True)
        env = dict(os.environ, SNN_DBG_SWIZZLE=cfg)
        subprocess.run([sys.executable, __file__, "child"], env=env, timeout=120)
