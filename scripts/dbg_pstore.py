import ctypes, sys, os, torch
sys.path.insert(0, '/root/repo')
import mi_b200
from mi_b200 import ops, _lib
lib = ctypes.CDLL(_lib.LIB_PATH)
dev = torch.device('cuda:0')
B, D = 65536, 1024
g = torch.Generator().manual_seed(0)
X = torch.relu(torch.randn(B, D, generator=g)).to(dev).bfloat16()
Y = torch.tanh(torch.randn(B, D, generator=g)).to(dev).bfloat16()
sid = torch.arange(B, dtype=torch.int32, device=dev)
ref = torch.full((B,), 20.0, device=dev)
L = _lib.load()
import ctypes as C
for dbg in (0, 1, 2, 0):
    lib.mi_set_debug(dbg)
    L.mi_set_profiling(1)
    ms = (C.c_double * 3)(); cnt = (C.c_int64 * 3)()
    for it in range(3):
        L.mi_profile_read(ms, cnt)
        ops.score_grad(X, Y, sid, sid, 0, 1.0 / 32, ref, 1.0, None, 0.0, False, 'fast', 1.0, 1.0 / B, want_k=True)
        torch.cuda.synchronize()
        L.mi_profile_read(ms, cnt)
    print('dbg', dbg, 'ds_panel ms', round(ms[1], 3), 'gemm ms', round(ms[2], 3), flush=True)
