import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodn_b200 import _lib
lib = _lib.get_lib()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for n_mma in (0, 1, 12, 48):
    for flags, name in ((0, "bare"), (1, "+fence.proxy.async"), (3, "+fence +tcgen05.ld"), (7, "1 worker warp, +fence +ld")):
        lib.check(lib.dll.mmn_selftest_protocol(2000, n_mma, flags, out.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        print(f"n_mma {n_mma:2d}  {name:28s} {int(out[0])} cycles/round")
