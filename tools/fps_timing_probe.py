import ctypes, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
from b200pc import synth
lib = ctypes.CDLL(os.path.join(ROOT, "point-cloud-interpolation-_b200/build/timing/libb200pc_fpstiming.so"))
dev = torch.device("cuda:0")
a, _ = synth.batch_pairs(0, 1, 16384)
for N, npt in ((16384, 1024), (1024, 256), (256, 64)):
    x = torch.from_numpy(a[:, :N].copy()).to(dev); st = torch.zeros(1, dtype=torch.long, device=dev); out = torch.empty(1, npt, dtype=torch.long, device=dev)
    for _ in range(2):
        rc = lib.b200pc_fps(ctypes.c_void_p(x.data_ptr()), 1, N, npt, ctypes.c_void_p(st.data_ptr()), ctypes.c_void_p(out.data_ptr()), None, ctypes.c_size_t(0), None)
        torch.cuda.synchronize()
    print("N", N, "rc", rc, flush=True)
