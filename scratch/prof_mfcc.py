import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "cs-304-speech-recognition-code_b200"))
from loe_speech_recognition.mfcc import mel_lane_tables
from ctypes import c_void_p, c_int, c_int64
dev = torch.device("cuda", 0)
bins, w, na, nb = mel_lane_tables(16000)
bins_d, w_d = torch.from_numpy(bins).to(dev), torch.from_numpy(w).to(dev)
rng = np.random.default_rng(0)
n = 2000
lens = rng.integers(46400, 73600, n)
pcm_off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
frames = 1 + lens // 160
frm_off = np.concatenate(([0], np.cumsum(frames))).astype(np.int64)
pcm = (torch.randn(int(pcm_off[-1]), device=dev) * 1000)
F = int(frm_off[-1])
mel = torch.zeros(F, 40, device=dev); um = torch.zeros(n, device=dev); feat = torch.zeros(F, 39, device=dev)
po, fo = torch.from_numpy(pcm_off).to(dev), torch.from_numpy(frm_off).to(dev)
lib = ctypes.CDLL(sys.argv[1])
fn = lib.loe_mfcc_phase_dev
fn.restype = c_int
fn.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
               c_void_p, c_void_p, c_void_p, c_void_p, c_int]
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    assert fn(pcm.data_ptr(), 0, po.data_ptr(), fo.data_ptr(), n, F, int(frames.max()), int(frames.min()),
              bins_d.data_ptr(), w_d.data_ptr(), na, nb, mel.data_ptr(), um.data_ptr(), feat.data_ptr(), st, 3) == 0
torch.cuda.synchronize()
