"""Time loe_emission_h16_img_dev of every library build under scratch/libs on the config-2 shape (random image / model:
timing only) and compare the scores bit for bit with the first library's."""
import ctypes, glob, os, sys, torch
from ctypes import c_void_p, c_int, c_int64
F, S = 3836960, 58
dev = torch.device("cuda", 0)
n_tiles = (S + 5) // 6
cst = torch.zeros(n_tiles * 6, device=dev)
out = torch.empty(F, S, device=dev)
torch.manual_seed(0)
ref = None
for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "libs", "*.so"))):
    lib = ctypes.CDLL(path)
    lib.loe_emission_h16_tile_bytes.restype = c_int
    lib.loe_emission_h16_img_bytes.restype = c_int64
    lib.loe_emission_h16_img_bytes.argtypes = [c_int64]
    nb = lib.loe_emission_h16_img_bytes(F)
    torch.manual_seed(1)
    img = (torch.randn(nb // 2, device=dev) * 0.5).half()
    inv2 = torch.ones(((F + 127) // 128) * 128, device=dev)
    b = (torch.randn(n_tiles * lib.loe_emission_h16_tile_bytes() // 2, device=dev) * 0.1).half()
    fn = lib.loe_emission_h16_img_dev
    fn.restype = c_int
    fn.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        assert fn(img.data_ptr(), inv2.data_ptr(), F, b.data_ptr(), cst.data_ptr(), S, out.data_ptr(), S, st) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn(img.data_ptr(), inv2.data_ptr(), F, b.data_ptr(), cst.data_ptr(), S, out.data_ptr(), S, st)
    e1.record(); torch.cuda.synchronize()
    o = out.clone()
    same = "" if ref is None else f" identical={bool(torch.equal(o, ref))}"
    if ref is None: ref = o
    print(os.path.basename(path), round(e0.elapsed_time(e1) / 10, 3), "ms", same)
