import ctypes, glob, os, sys, torch
from ctypes import c_void_p, c_int, c_int64
F, S = 3836960, 58
dev = torch.device("cuda", 0)
feat = torch.randn(F, 39, device=dev)
n_tiles = (S + 5) // 6
cst = torch.zeros(n_tiles * 6, device=dev)
out = torch.empty(F, S, device=dev)
for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "libs", "*.so"))):
    lib = ctypes.CDLL(path)
    lib.loe_emission_h16_tile_bytes.restype = c_int
    b = (torch.randn(n_tiles * lib.loe_emission_h16_tile_bytes() // 2, device=dev) * 0.1).half()
    fn = lib.loe_emission_h16_dev
    fn.restype = c_int
    fn.argtypes = [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        assert fn(feat.data_ptr(), F, 39, b.data_ptr(), cst.data_ptr(), S, out.data_ptr(), S, st) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn(feat.data_ptr(), F, 39, b.data_ptr(), cst.data_ptr(), S, out.data_ptr(), S, st)
    e1.record(); torch.cuda.synchronize()
    print(os.path.basename(path), round(e0.elapsed_time(e1) / 10, 3), "ms")
