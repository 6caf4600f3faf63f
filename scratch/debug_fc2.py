import sys, ctypes, torch, numpy as np
sys.path.insert(0, '.')
from snn_automotive_object_detection_b200 import _lib
from tests.test_gpu_kernels import pack_words
from tests._util import vp, stream, prepared_fc, _TRAIN_DTYPE
lib = _lib.load()
lib.snn_debug_set_btile_dump.argtypes = [ctypes.c_void_p]
R,K,M,T,t0,T_live,mode,cg = 45,128,256,12,0,3,1,1
J, T_box = 32, 3
g = torch.Generator().manual_seed(0)
z = (torch.rand(T_live, R, K, generator=g) < 0.15).float()
w = torch.randn(M, K, generator=g) * 0.3
words = pack_words(z, 0, 1).cuda()
wp = prepared_fc(w.cuda(), mode)
trains = torch.zeros(R, M, dtype=torch.int16, device="cuda")
dump = torch.full((T_live, R, M), float("nan"), device="cuda")
tiles, kbs, nh = 4, K // 64, T_box * J
dbg = torch.full((tiles, kbs, 1, nh, 64), -1, dtype=torch.int16, device="cuda")
lib.snn_debug_set_btile_dump(vp(dbg))
rc = lib.snn_fc_lif_layer(vp(words), 1, 0, R, K, M, T, t0, T_live, mode, vp(wp), vp(trains), vp(dump), cg, stream())
_lib.check(rc, "fc")
torch.cuda.synchronize()
lib.snn_debug_set_btile_dump(None)
d = dbg.cpu().view(torch.bfloat16).float()      # [tiles][kb][1][t*J + j][64]
for tile in range(tiles):
    ut = tile // 2
    for kb in range(kbs):
        got = d[tile, kb, 0].view(T_box, J, 64)
        exp = torch.zeros(T_box, J, 64)
        for j in range(J):
            r = ut * J + j
            if r < R:
                exp[:, j] = z[:, r, kb*64:(kb+1)*64]
        bad = (got != exp)
        print("tile", tile, "kb", kb, "bad elems", bad.sum().item(), "bad j:", bad.any(dim=2).any(dim=0).nonzero().flatten().tolist(), "got sum", got.sum().item(), "exp sum", exp.sum().item())
ref = torch.einsum("trk,mk->trm", z.double(), w.to(torch.bfloat16).double())
err = (dump.cpu().double() - ref).abs()
print("dump bad rows", (err > 1e-3).any(dim=2).any(dim=0).nonzero().flatten().tolist())
