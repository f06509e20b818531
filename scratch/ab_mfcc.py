"""A/B the MFCC kernels of the library builds under scratch/libs (first one = baseline): max difference of the mel
energies and features on the same random PCM (float32 and int16, even and odd utterance starts), then timings."""
import ctypes, glob, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "cs-304-speech-recognition-code_b200"))
from loe_speech_recognition.mfcc import mel_lane_tables
from ctypes import c_void_p, c_int, c_int64
dev = torch.device("cuda", 0)
bins, w, na, nb = mel_lane_tables(16000)
bins_d, w_d = torch.from_numpy(bins).to(dev), torch.from_numpy(w).to(dev)

def corpus(n, lo, hi, seed, odd=False):
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi, n)
    if odd:
        lens |= 1
    t = np.arange(hi)
    sig = []
    for L in lens:
        f = rng.uniform(200, 4000, 3)
        s = sum(3000 * np.sin(2 * np.pi * fi * t[:L] / 16000 + rng.uniform(0, 6)) for fi in f) + rng.normal(0, 30, L)
        sig.append(np.round(s).astype(np.int16))
    pcm_off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    frames = 1 + lens // 160
    frm_off = np.concatenate(([0], np.cumsum(frames))).astype(np.int64)
    return np.concatenate(sig), pcm_off, frm_off, frames

def run(lib, pcm, fmt, pcm_off, frm_off, frames, phases=3, reps=0):
    n = len(frames); F = int(frm_off[-1])
    mel = torch.zeros(F, 40, device=dev); um = torch.zeros(n, device=dev); feat = torch.zeros(F, 39, device=dev)
    po, fo = torch.from_numpy(pcm_off).to(dev), torch.from_numpy(frm_off).to(dev)
    fn = lib.loe_mfcc_phase_dev
    fn.restype = c_int
    fn.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                   c_void_p, c_void_p, c_void_p, c_void_p, c_int]
    st = torch.cuda.current_stream().cuda_stream
    call = lambda ph: fn(pcm.data_ptr(), fmt, po.data_ptr(), fo.data_ptr(), n, F, int(frames.max()), int(frames.min()),
                         bins_d.data_ptr(), w_d.data_ptr(), na, nb, mel.data_ptr(), um.data_ptr(), feat.data_ptr(), st, ph)
    assert call(3) == 0
    torch.cuda.synchronize()
    times = {}
    for ph in ((1, 2) if reps else ()):
        for _ in range(3): call(ph)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): call(ph)
        e1.record(); torch.cuda.synchronize()
        times[ph] = round(e0.elapsed_time(e1) / reps, 3)
    return mel.cpu().numpy(), feat.cpu().numpy(), times

libs = [(os.path.basename(p), ctypes.CDLL(p)) for p in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "libs", "*.so")))]
for odd in (False, True):
    s16, pcm_off, frm_off, frames = corpus(64, 1500, 70000, 1, odd)
    for fmt, arr in ((0, s16.astype(np.float32)), (1, s16)):
        pcm = torch.from_numpy(arr).to(dev)
        ref = None
        for name, lib in libs:
            mel, feat, _ = run(lib, pcm, fmt, pcm_off, frm_off, frames)
            if ref is None:
                ref = (mel, feat); continue
            dm = np.abs(mel - ref[0]) / (np.abs(ref[0]) + 1e-30)
            scale = np.abs(ref[1][:, :13]).max()
            df = np.abs(feat - ref[1]) / np.maximum(np.abs(ref[1]), 1e-4 * scale)
            print(f"odd={odd} fmt={fmt} {name}: mel rel max {dm.max():.3e} (p99.9 {np.quantile(dm, 0.999):.3e}), feat rel max {df.max():.3e}, finite {np.isfinite(feat).all()}")
# timing on the bench shape
s16, pcm_off, frm_off, frames = corpus(500, 46400, 73600, 2)
reps = 20
s16 = np.tile(s16, 20); n0 = len(frames)
pcm_off = np.concatenate(([0], np.cumsum(np.tile(np.diff(pcm_off), 20)))).astype(np.int64)
frames = np.tile(frames, 20); frm_off = np.concatenate(([0], np.cumsum(frames))).astype(np.int64)
for fmt, arr in ((0, s16.astype(np.float32)), (1, s16)):
    pcm = torch.from_numpy(arr).to(dev)
    for name, lib in libs:
        _, _, tm = run(lib, pcm, fmt, pcm_off, frm_off, frames, reps=10)
        print(f"fmt={fmt} {name}: frames {frm_off[-1]} mel {tm[1]} ms ceps {tm[2]} ms")
