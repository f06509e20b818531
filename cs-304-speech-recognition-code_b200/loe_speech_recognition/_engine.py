"""Per-process device engine: owns the CUDA device buffers (torch tensors) and drives the C ABI.

PyTorch is plumbing here (device memory, streams, torch.distributed); every computation is a
hand-written kernel behind include/loe_b200.h.  The engine is created lazily on first use and
never crosses a fork: callers that use ``ProcessPoolExecutor`` (as the reference's scripts do)
get a fresh engine per worker process.  There is no CPU fallback -- without a CUDA device or
without the built library every entry point raises.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from ._trellis import HostTrellis, stack

_ENGINE = None
_ENGINE_PID = None

PRECISIONS = {"fp32": 0, "fp64": 1, "tc": 2, "h16": 3}


def default_precision() -> str:
    """"auto" = for 39-dimensional models the tcgen05 3xFP16 kernel ("h16"; the 3xTF32 kernel "tc" when a
    model's whitening matrix leaves the binary16 range), the float32 SIMT kernel otherwise.
    LOE_B200_EMISSION=fp32|fp64|tc|h16 overrides (fp64 reproduces scipy bit for bit almost
    everywhere and is the mode to use when chasing a path difference)."""
    return os.environ.get("LOE_B200_EMISSION", "auto")


class NoCudaDevice(RuntimeError):
    pass


@dataclass
class GaussPack:
    """Flat device copy of a list of MultivariateNormal (scipy frozen) objects."""
    n_states: int
    dim: int
    mean32: "torch.Tensor"
    u32: "torch.Tensor"
    cst32: "torch.Tensor"
    mean64: "torch.Tensor"
    u64: "torch.Tensor"
    cst64: "torch.Tensor"
    b_packed: "torch.Tensor" = None     # tensor-core image, 3xTF32 (dim == 39 only)
    cst_pad: "torch.Tensor" = None
    b_h16: "torch.Tensor" = None        # tensor-core image, 3xFP16 (dim == 39 and |W| < 32768 only)


@dataclass
class GmmPack:
    """Flat device copy of a DiagGMM (gmm.py): per-component arrays for the SIMT kernels, the tensor-core image if its
    entries fit the binary16 pair."""
    n_states: int
    n_mix: int
    mean32: "torch.Tensor"
    iv32: "torch.Tensor"
    cst32: "torch.Tensor"
    mean64: "torch.Tensor"
    iv64: "torch.Tensor"
    cst64: "torch.Tensor"
    b_img: "torch.Tensor" = None
    shift_scale: "torch.Tensor" = None


@dataclass
class TrellisPack:
    tr_off: "torch.Tensor"
    col: "torch.Tensor"
    band: "torch.Tensor"
    flags: "torch.Tensor"
    word: "torch.Tensor"
    word_lo: "torch.Tensor"
    max_pos: int
    max_ends: int
    n_trellis: int


@dataclass
class Batch:
    """A batch of utterances resident on the device."""
    feat: "torch.Tensor"          # [F, D] float32
    frm_off: "torch.Tensor"       # [n+1] int64 (device)
    frm_off_host: np.ndarray      # [n+1] int64
    n_utt: int
    max_frames: int

    @property
    def total_frames(self) -> int:
        return int(self.frm_off_host[-1])


def host_gauss_arrays(normals) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(mean [S,D], U [S,D,D], cst [S]) in float64 from MultivariateNormal objects
    (hidden_markov_model.py:20-48; scipy frozen: mean, cov_object._LP, ._log_pdet, ._rank)."""
    means = np.stack([np.asarray(mn._core.mean, dtype=np.float64) for mn in normals])
    us = np.stack([np.asarray(mn._core.cov_object._LP, dtype=np.float64) for mn in normals])
    cst = np.array([-0.5 * (mn._core.cov_object._rank * np.log(2 * np.pi) + mn._core.cov_object._log_pdet)
                    for mn in normals], dtype=np.float64)
    return means, us, cst


def _tf32_round(a: np.ndarray) -> np.ndarray:
    bits = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    return ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def pack_tc_image(means: np.ndarray, us: np.ndarray, cst: np.ndarray):
    """Host pre-pack of the tensor-core operand (layout documented at loe_emission_tc_dev):
    W_s = [U_s ; -mean_s.U_s] in float64, split into a TF32-rounded part and the float32 residual."""
    S, D = means.shape
    spt, cols, K = 6, 40, 40
    n_tiles = (S + spt - 1) // spt
    W = np.zeros((n_tiles * spt, K, cols), dtype=np.float64)
    W[:S, :D, :D] = us
    W[:S, D, :D] = -np.einsum("si,sij->sj", means, us)
    hi = _tf32_round(W.astype(np.float32))
    lo = (W - hi.astype(np.float64)).astype(np.float32)
    out = np.empty((n_tiles, 2, K // 4, spt * cols, 4), dtype=np.float32)
    for h, part in enumerate((hi, lo)):
        # part [tile, state_local, k, j] -> [tile, kc, n = state_local*40 + j, q]
        p = part.reshape(n_tiles, spt, K // 4, 4, cols)            # [t, sl, kc, q, j]
        out[:, h] = p.transpose(0, 2, 1, 4, 3).reshape(n_tiles, K // 4, spt * cols, 4)
    cst_pad = np.zeros(n_tiles * spt, dtype=np.float32)
    cst_pad[:S] = cst.astype(np.float32)
    return np.ascontiguousarray(out.reshape(-1)), cst_pad


H16_MAX = 32768.0


H16_BLOCK_WIDTHS = (8, 8, 8, 8, 8)    # accumulator column blocks per state (csrc/emission_h16.cu)


def pack_h16_image(means: np.ndarray, us: np.ndarray, cst: np.ndarray):
    """Host pre-pack of the 3xFP16 tensor-core operand (layout documented at loe_emission_h16_dev), or
    None when an entry of the operand leaves the binary16 range.

    The kernel only needs |U_s^T (x - mean_s)|, so the whitening matrix may be replaced by any W with
    W W^T = U_s U_s^T.  U_s^T = Q R gives W = R^T, LOWER TRIANGULAR: feature k only reaches the columns
    j <= k.  With the accumulator columns ordered [column block j // 8][state][j % 8], the K chunk c of 8
    features only reaches the first 48 (c + 1) columns and its MMAs are issued with that N instead of 240 --
    65 % of the dense MMA work.  Row 39 is the bias -mean_s . W (dense), column 39 an exact zero."""
    S, D = means.shape
    spt, cols, K = 6, 40, 40
    n_tiles = (S + spt - 1) // spt
    W = np.zeros((n_tiles * spt, K, cols), dtype=np.float64)
    for s in range(S):
        r = np.linalg.qr(np.asarray(us[s], dtype=np.float64).T, mode="r")      # upper triangular, |R d| = |U^T d|
        W[s, :D, :D] = np.tril(r.T)
        W[s, D, :D] = -means[s] @ W[s, :D, :D]
    if not np.all(np.abs(W) < H16_MAX):          # also rejects NaN / inf
        return None
    hi = W.astype(np.float16)
    lo = (W - hi.astype(np.float64)).astype(np.float16)
    out = np.empty((n_tiles, 15, spt * cols, 8), dtype=np.float16)
    for base, part in ((0, hi), (5, lo), (10, hi)):
        p = part.reshape(n_tiles, spt, K // 8, 8, cols)                 # [t, sl, kc, q, j]
        blocks, j0 = [], 0
        for w in H16_BLOCK_WIDTHS:                                      # block b: n = start_b + sl * w + (j - j0)
            blocks.append(p[..., j0:j0 + w].transpose(0, 2, 1, 4, 3).reshape(n_tiles, K // 8, spt * w, 8))
            j0 += w
        out[:, base:base + 5] = np.concatenate(blocks, axis=2)
    return np.ascontiguousarray(out.reshape(-1))


class Engine:
    def __init__(self, device: Optional[int] = None):
        import torch

        self.torch = torch
        self.lib = _native.load()
        if not torch.cuda.is_available():
            raise NoCudaDevice("no CUDA device visible: loe_speech_recognition (B200 build) has no CPU fallback")
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if torch.cuda.device_count() > 1 else 0
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self._mel_cache = {}
        self._copy_stream = None
        self.launches = 0           # kernels launched through the C ABI (bench.py reports it)

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> int:
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def copy_stream(self):
        """Side stream for host->device copies that overlap with compute (decode_pcm_flat)."""
        if self._copy_stream is None:
            self._copy_stream = self.torch.cuda.Stream(device=self.device)
        return self._copy_stream

    def _to_dev(self, arr: np.ndarray):
        t = self.torch.from_numpy(np.ascontiguousarray(arr))
        return t.to(self.device, non_blocking=False)

    def empty(self, shape, dtype):
        return self.torch.empty(shape, dtype=dtype, device=self.device)

    @staticmethod
    def _p(t) -> int:
        return 0 if t is None else t.data_ptr()

    # ------------------------------------------------------------------ packing
    def pack_gaussians(self, normals) -> GaussPack:
        means, us, cst = host_gauss_arrays(normals)
        return self.pack_gauss_arrays(means, us, cst)

    def pack_gauss_arrays(self, means, us, cst) -> GaussPack:
        gp = GaussPack(
            n_states=means.shape[0], dim=means.shape[1],
            mean32=self._to_dev(means.astype(np.float32)), u32=self._to_dev(us.astype(np.float32)),
            cst32=self._to_dev(cst.astype(np.float32)),
            mean64=self._to_dev(means), u64=self._to_dev(us), cst64=self._to_dev(cst))
        if gp.dim == 39:
            b, c = pack_tc_image(means, us, cst)
            gp.b_packed, gp.cst_pad = self._to_dev(b), self._to_dev(c)
            h = pack_h16_image(means, us, cst)
            if h is not None:
                gp.b_h16 = self._to_dev(h)
        return gp

    def pack_tc_words(self, words):
        """Tensor-core images of several word models in ONE host pass and ONE upload: ``words`` =
        [(means [S,D], U [S,D,D], cst [S])]; every word starts on a 6-state tile boundary so that a launch
        can address it by pointer offset.  Returns (b_packed, cst_pad, first tile of each word)."""
        D = words[0][0].shape[1]
        first, slots, t = [], [], 0
        for means, _, _ in words:
            n_t = (means.shape[0] + 5) // 6
            first.append(t)
            slots.append(np.arange(means.shape[0]) + 6 * t)
            t += n_t
        slots = np.concatenate(slots)
        means_p = np.zeros((6 * t, D)); us_p = np.zeros((6 * t, D, D)); cst_p = np.zeros(6 * t)
        means_p[slots] = np.concatenate([w[0] for w in words])
        us_p[slots] = np.concatenate([w[1] for w in words])
        cst_p[slots] = np.concatenate([w[2] for w in words])
        b, c = pack_tc_image(means_p, us_p, cst_p)        # zero states give zero columns, like the padding
        return self._to_dev(b), self._to_dev(c), first

    def emission_tc_into(self, feat, b_packed, cst_pad, first_tile: int, n_states: int, out, col0: int):
        """Tensor-core emission of one word (tiles from ``first_tile``) into columns [col0, col0+n_states) of ``out``."""
        n_frames = int(feat.shape[0])
        if n_frames == 0:
            return
        ld = int(out.shape[1])
        _native.check(self.lib.loe_emission_tc_dev(feat.data_ptr(), n_frames, int(feat.shape[1]),
                                                   b_packed.data_ptr() + first_tile * 19200 * 4, cst_pad.data_ptr() + first_tile * 6 * 4,
                                                   n_states, out.data_ptr() + col0 * 4, ld, self._stream()))
        self.launches += 1

    def emission_h16_into(self, feat, b_h16, cst_pad, first_tile: int, n_states: int, out, col0: int):
        """3xFP16 tensor-core emission of one word (image tiles from ``first_tile``) into columns [col0, col0+n_states) of ``out``."""
        n_frames = int(feat.shape[0])
        if n_frames == 0:
            return
        ld = int(out.shape[1])
        tile_bytes = int(self.lib.loe_emission_h16_tile_bytes())
        _native.check(self.lib.loe_emission_h16_dev(feat.data_ptr(), n_frames, int(feat.shape[1]),
                                                    b_h16.data_ptr() + first_tile * tile_bytes, cst_pad.data_ptr() + first_tile * 6 * 4,
                                                    n_states, out.data_ptr() + col0 * 4, ld, self._stream()))
        self.launches += 1

    def pack_trellises(self, trellises: List[HostTrellis]) -> TrellisPack:
        off, col, band, flags, word, word_lo, max_pos, max_ends = stack(trellises)
        if max_pos > _native.LOE_MAX_POS:
            raise OverflowError(f"{max_pos} trellis positions: the int8 path of the reference holds at most "
                                f"{_native.LOE_MAX_POS}")
        return TrellisPack(self._to_dev(off), self._to_dev(col), self._to_dev(band), self._to_dev(flags),
                           self._to_dev(word), self._to_dev(word_lo), max_pos, max_ends, len(trellises))

    # ------------------------------------------------------------------ batches
    def upload_features(self, feats: Sequence[np.ndarray], dim: Optional[int] = None) -> Batch:
        lens = np.array([f.shape[0] for f in feats], dtype=np.int64)
        if dim is not None:
            for f in feats:
                assert f.shape[1] == dim
        off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
        flat = np.concatenate([np.asarray(f, dtype=np.float32) for f in feats], axis=0) if len(feats) > 1 \
            else np.ascontiguousarray(feats[0], dtype=np.float32)
        return Batch(self._to_dev(flat), self._to_dev(off), off, len(feats), int(lens.max()))

    # ------------------------------------------------------------------ MFCC
    def _mel_tables(self, sample_rate):
        key = float(sample_rate)
        if key not in self._mel_cache:
            from .mfcc import mel_lane_tables
            bins, w, na, nb = mel_lane_tables(sample_rate)
            self._mel_cache[key] = (self._to_dev(bins), self._to_dev(w), na, nb)
        return self._mel_cache[key]

    def image_buffers(self, total_frames: int):
        """(a_img uint8, inv2 float32) sized for ``total_frames`` frames: the pre-split A operand of the 3xFP16 emission kernel."""
        torch = self.torch
        n_bytes = int(self.lib.loe_emission_h16_img_bytes(total_frames))
        return self.empty((n_bytes,), torch.uint8), self.empty((((total_frames + 127) // 128) * 128,), torch.float32)

    def mfcc_device(self, pcm, pcm_off, frm_off, n_utt, total_frames, max_frames, min_frames, sample_rate=16000,
                    out=None, mel_ws=None, utt_max=None, phases=3, image=None, want_feat=True):
        """PCM (device) -> features [total_frames, 39] (device).  All tensors on this device.
        ``image=(a_img, inv2)`` (see image_buffers) additionally writes every feature row as the pre-split operand of
        emission_image(); with ``want_feat=False`` the float32 matrix is not written at all (decode path) and None returned."""
        torch = self.torch
        bins, w, na, nb = self._mel_tables(sample_rate)
        if out is None and want_feat:
            out = self.empty((total_frames, 39), torch.float32)
        if mel_ws is None:
            mel_ws = self.empty((total_frames, 40), torch.float32)
        if utt_max is None:
            utt_max = self.empty((n_utt,), torch.float32)
        if pcm.dtype == torch.int16:
            fmt = 1
        elif pcm.dtype == torch.float32:
            fmt = 0
        else:
            raise TypeError(f"PCM must be float32 or int16 on the device (got {pcm.dtype})")
        if image is not None:
            _native.check(self.lib.loe_mfcc_img_dev(pcm.data_ptr(), fmt, pcm_off.data_ptr(), frm_off.data_ptr(), n_utt, total_frames,
                                                    max_frames, min_frames, bins.data_ptr(), w.data_ptr(), na, nb,
                                                    mel_ws.data_ptr(), utt_max.data_ptr(), self._p(out if want_feat else None),
                                                    image[0].data_ptr(), image[1].data_ptr(), self._stream(), phases))
            self.launches += (phases & 1) + ((phases >> 1) & 1)
            return out if want_feat else None
        _native.check(self.lib.loe_mfcc_phase_dev(pcm.data_ptr(), fmt, pcm_off.data_ptr(), frm_off.data_ptr(), n_utt, total_frames,
                                                  max_frames, min_frames, bins.data_ptr(), w.data_ptr(), na, nb,
                                                  mel_ws.data_ptr(), utt_max.data_ptr(), out.data_ptr(), self._stream(), phases))
        self.launches += (phases & 1) + ((phases >> 1) & 1)
        return out

    def _ex_tables(self, sample_rate, config):
        key = (float(sample_rate), config)
        if key not in self._mel_cache:
            start, length, w = config.mel_tables(sample_rate)
            self._mel_cache[key] = (self._to_dev(config.window_table()), self._to_dev(start), self._to_dev(length), self._to_dev(w),
                                    int(w.shape[1]), self._to_dev(config.dct_table()))
        return self._mel_cache[key]

    def mfcc_ex_device(self, pcm, pcm_off, frm_off, n_utt, total_frames, max_frames, min_frames, sample_rate, config,
                       out=None, mel_ws=None, ceps_ws=None, utt_stat=None):
        """PCM (device) -> features [total_frames, 3 n_mfcc] (device) under an MFCCConfig (loe_mfcc_ex_dev)."""
        import ctypes
        torch = self.torch
        win, start, length, w, pitch, dct = self._ex_tables(sample_rate, config)
        nc = int(config.n_mfcc)
        out = self.empty((total_frames, 3 * nc), torch.float32) if out is None else out
        mel_ws = self.empty((total_frames, config.n_mels), torch.float32) if mel_ws is None else mel_ws
        ceps_ws = self.empty((total_frames, nc), torch.float32) if ceps_ws is None else ceps_ws
        utt_stat = self.empty((n_utt, 2 * nc), torch.float32) if utt_stat is None else utt_stat
        if pcm.dtype == torch.int16:
            fmt = 1
        elif pcm.dtype == torch.float32:
            fmt = 0
        else:
            raise TypeError(f"PCM must be float32 or int16 on the device (got {pcm.dtype})")
        cfg = _native.MfccConfigStruct(int(config.n_fft), int(config.hop_length), int(config.n_mels), nc,
                                       _native.LOG_MODES[config.log], _native.NORM_MODES[config.norm], float(config.preemphasis), 0.0)
        _native.check(self.lib.loe_mfcc_ex_dev(pcm.data_ptr(), fmt, pcm_off.data_ptr(), frm_off.data_ptr(), n_utt, total_frames,
                                               max_frames, min_frames, ctypes.addressof(cfg), win.data_ptr(), start.data_ptr(),
                                               length.data_ptr(), w.data_ptr(), pitch, dct.data_ptr(), mel_ws.data_ptr(),
                                               ceps_ws.data_ptr(), utt_stat.data_ptr(), out.data_ptr(), self._stream()))
        self.launches += 3 + (1 if config.norm in ("cmn", "cmvn") else 0)
        return out

    def upload_pcm(self, signals: Sequence[np.ndarray], hop: int = 160):
        lens = np.array([s.shape[0] for s in signals], dtype=np.int64)
        pcm_off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
        frames = 1 + lens // hop
        frm_off = np.concatenate(([0], np.cumsum(frames))).astype(np.int64)
        # int16 signals (raw WAV samples) stay int16 on the wire; anything else is shipped as float32
        dt = np.int16 if all(s.dtype == np.int16 for s in signals) else np.float32
        flat = np.concatenate([np.asarray(s, dtype=dt) for s in signals]) if len(signals) > 1 \
            else np.ascontiguousarray(signals[0], dtype=dt)
        return self._to_dev(flat), self._to_dev(pcm_off), self._to_dev(frm_off), frm_off, frames

    def mfcc(self, signals: Sequence[np.ndarray], sample_rate=16000, config=None) -> Batch:
        if config is not None and not config.is_reference:
            pcm, pcm_off, frm_off_dev, frm_off, frames = self.upload_pcm(signals, int(config.hop_length))
            feat = self.mfcc_ex_device(pcm, pcm_off, frm_off_dev, len(signals), int(frm_off[-1]), int(frames.max()),
                                       int(frames.min()), sample_rate, config)
            return Batch(feat, frm_off_dev, frm_off, len(signals), int(frames.max()))
        pcm, pcm_off, frm_off_dev, frm_off, frames = self.upload_pcm(signals)
        feat = self.mfcc_device(pcm, pcm_off, frm_off_dev, len(signals), int(frm_off[-1]), int(frames.max()),
                                int(frames.min()), sample_rate)
        return Batch(feat, frm_off_dev, frm_off, len(signals), int(frames.max()))

    # ------------------------------------------------------------------ emission
    def emission(self, feat, gp: GaussPack, precision: Optional[str] = None, out=None, ld: Optional[int] = None):
        torch = self.torch
        precision = precision or default_precision()
        if precision == "auto":
            precision = "h16" if gp.b_h16 is not None else "tc" if gp.b_packed is not None else "fp32"
        if precision == "h16" and gp.b_h16 is None and gp.b_packed is not None:
            precision = "tc"                     # whitening matrix outside the binary16 range
        n_frames, dim = int(feat.shape[0]), int(feat.shape[1])
        if dim != gp.dim:
            raise AssertionError(f"feature dimension {dim} != model dimension {gp.dim}")
        if ld is None:
            ld = gp.n_states
        if out is None:
            out = self.empty((n_frames, ld), torch.float32)
        code = PRECISIONS[precision]
        if code == 3:
            if gp.b_h16 is None:
                raise NotImplementedError("the tensor-core emission kernel is built for 39-dimensional features")
            _native.check(self.lib.loe_emission_h16_dev(feat.data_ptr(), n_frames, dim, gp.b_h16.data_ptr(),
                                                        gp.cst_pad.data_ptr(), gp.n_states, out.data_ptr(), ld, self._stream()))
            self.launches += 1
            return out
        if code == 2:
            if gp.b_packed is None:
                raise NotImplementedError("the tensor-core emission kernel is built for 39-dimensional features")
            _native.check(self.lib.loe_emission_tc_dev(feat.data_ptr(), n_frames, dim, gp.b_packed.data_ptr(),
                                                       gp.cst_pad.data_ptr(), gp.n_states, out.data_ptr(), ld, self._stream()))
            self.launches += 1
            return out
        mean, u, cst = (gp.mean64, gp.u64, gp.cst64) if code == 1 else (gp.mean32, gp.u32, gp.cst32)
        _native.check(self.lib.loe_emission_dev(feat.data_ptr(), n_frames, dim, mean.data_ptr(), u.data_ptr(), cst.data_ptr(),
                                                gp.n_states, out.data_ptr(), ld, code, self._stream()))
        self.launches += 1
        return out

    def emission_image(self, image, n_frames: int, gp: GaussPack, out=None):
        """[n_frames, S] scores from the pre-split operand mfcc_device(image=...) wrote: the 3xFP16 kernel without producer
        work.  Bit-identical to emission(feat, gp, "h16")."""
        if gp.b_h16 is None:
            raise NotImplementedError("model has no 3xFP16 image (whitening matrix outside the binary16 range)")
        if out is None:
            out = self.empty((n_frames, gp.n_states), self.torch.float32)
        _native.check(self.lib.loe_emission_h16_img_dev(image[0].data_ptr(), image[1].data_ptr(), n_frames, gp.b_h16.data_ptr(),
                                                        gp.cst_pad.data_ptr(), gp.n_states, out.data_ptr(), int(out.shape[1]), self._stream()))
        self.launches += 1
        return out

    # ------------------------------------------------------------------ diagonal GMM emission
    def pack_gmm(self, weights, means, variances) -> GmmPack:
        from .gmm import LOG_2PI, pack_gmm_image
        S, M, D = means.shape
        with np.errstate(divide="ignore"):
            cst = np.log(weights) - 0.5 * (D * LOG_2PI + np.sum(np.log(variances), axis=-1))
        m = np.ascontiguousarray(means.reshape(S * M, D)); iv = np.ascontiguousarray((1.0 / variances).reshape(S * M, D))
        c = np.ascontiguousarray(cst.reshape(S * M))
        gp = GmmPack(S, M, self._to_dev(m.astype(np.float32)), self._to_dev(iv.astype(np.float32)), self._to_dev(c.astype(np.float32)),
                     self._to_dev(m), self._to_dev(iv), self._to_dev(c))
        img = pack_gmm_image(weights, means, variances) if D == 39 else None
        if img is not None:
            gp.b_img, gp.shift_scale = self._to_dev(img[0]), self._to_dev(img[1])
        return gp

    def emission_gmm(self, feat, gp: GmmPack, precision: Optional[str] = None, out=None):
        """[F, S] state log-likelihoods of a diagonal GMM.  precision "auto" / None: the tcgen05 kernel when the model has
        a tensor-core image, else float32 SIMT; "tc", "fp32", "fp64" force a path."""
        torch = self.torch
        precision = precision or "auto"
        if precision == "auto":
            precision = "tc" if gp.b_img is not None else "fp32"
        n_frames, dim = int(feat.shape[0]), int(feat.shape[1])
        if dim != 39:
            raise AssertionError(f"feature dimension {dim} != model dimension 39")
        if out is None:
            out = self.empty((n_frames, gp.n_states), torch.float32)
        ld = int(out.shape[1])
        if precision == "tc":
            if gp.b_img is None:
                raise NotImplementedError("this mixture model has no tensor-core image (entries outside the binary16 pair's range)")
            _native.check(self.lib.loe_emission_gmm_tc_dev(feat.data_ptr(), n_frames, dim, gp.b_img.data_ptr(), gp.shift_scale.data_ptr(),
                                                           gp.mean32.data_ptr(), gp.iv32.data_ptr(), gp.cst32.data_ptr(), gp.n_states,
                                                           gp.n_mix, out.data_ptr(), ld, self._stream()))
        elif precision in ("fp32", "fp64"):
            mean, iv, cst = (gp.mean64, gp.iv64, gp.cst64) if precision == "fp64" else (gp.mean32, gp.iv32, gp.cst32)
            _native.check(self.lib.loe_emission_gmm_dev(feat.data_ptr(), n_frames, dim, mean.data_ptr(), iv.data_ptr(), cst.data_ptr(),
                                                        gp.n_states, gp.n_mix, out.data_ptr(), ld, 1 if precision == "fp64" else 0,
                                                        self._stream()))
        else:
            raise ValueError(f"unknown precision {precision!r}")
        self.launches += 1
        return out

    # ------------------------------------------------------------------ Viterbi
    def viterbi(self, scores, batch_frm_off, n_utt, max_frames, total_frames, tp: TrellisPack, utt_tr=None,
                loop=False, penalty=0.0, penalty_f64=False, want_end_scores=True, labels=None):
        """``labels=(skip_label, max_words)`` additionally decodes the word sequence of every path in the
        same launch; the return value then has two more entries (words int8 [n, max_words], count int32 [n])."""
        torch = self.torch
        words = count = None
        skip_label, max_words = -1, 0
        if labels is not None:
            skip_label, max_words = labels
            words = self.empty((n_utt, max_words), torch.int8)
            count = self.empty((n_utt,), torch.int32)
        path = self.empty((total_frames,), torch.int8)
        best = self.empty((n_utt,), torch.int32)
        best_score = self.empty((n_utt,), torch.float32)
        end_scores = self.empty((n_utt, tp.max_ends), torch.float32) if want_end_scores else None
        bp_ws = None
        if not self.lib.loe_viterbi_bp_fits(max_frames, tp.max_pos):
            bp_ws = self.empty((total_frames * _native.LOE_MAX_POS,), torch.uint8)
        _native.check(self.lib.loe_viterbi_dev(
            scores.data_ptr(), int(scores.shape[1]), batch_frm_off.data_ptr(), n_utt, max_frames,
            tp.tr_off.data_ptr(), tp.col.data_ptr(), tp.band.data_ptr(), tp.flags.data_ptr(), tp.max_pos,
            self._p(utt_tr), 1 if loop else 0, float(penalty), 1 if penalty_f64 else 0,
            path.data_ptr(), self._p(end_scores), tp.max_ends, best.data_ptr(), best_score.data_ptr(),
            self._p(bp_ws), tp.word.data_ptr(), tp.word_lo.data_ptr(), skip_label, self._p(words), max_words, self._p(count),
            self._stream()))
        self.launches += 1
        if labels is not None:
            return path, end_scores, best, best_score, words, count
        return path, end_scores, best, best_score

    def labels(self, path, frm_off, n_utt, tp: TrellisPack, utt_tr=None, skip_label=-1, max_words=32):
        """(words int8 [n_utt, max_words], count int32 [n_utt]) on the device."""
        torch = self.torch
        words = self.empty((n_utt, max_words), torch.int8)
        count = self.empty((n_utt,), torch.int32)
        _native.check(self.lib.loe_labels_dev(path.data_ptr(), frm_off.data_ptr(), n_utt, tp.tr_off.data_ptr(),
                                              tp.word.data_ptr(), tp.word_lo.data_ptr(), self._p(utt_tr), skip_label,
                                              words.data_ptr(), max_words, count.data_ptr(), self._stream()))
        self.launches += 1
        return words, count

    # ------------------------------------------------------------------ silence stripper
    def silence(self, signals: Sequence[np.ndarray], frame_size: int, high: float, low: float, max_silence_frames: int):
        """Host arrays (energies, noise mask, seg [n,4], max [n], energy offsets) for a batch of signals."""
        torch = self.torch
        pcm, pcm_off, _, _, _ = self.upload_pcm(signals)
        lens = np.array([s.shape[0] for s in signals], dtype=np.int64)
        efr = lens // frame_size + 1
        eoff = np.concatenate(([0], np.cumsum(efr))).astype(np.int64)
        n, tot = len(signals), int(eoff[-1])
        energy = self.empty((tot,), torch.float32)
        noise = self.empty((tot,), torch.uint8)
        seg = self.empty((n, 4), torch.int32)
        mx = self.empty((n,), torch.float32)
        fmt = 1 if pcm.dtype == torch.int16 else 0
        eoff_dev = self._to_dev(eoff)                 # must outlive the launch call
        _native.check(self.lib.loe_silence_dev(pcm.data_ptr(), fmt, pcm_off.data_ptr(), n, frame_size, float(high), float(low),
                                               int(max_silence_frames), eoff_dev.data_ptr(), energy.data_ptr(),
                                               noise.data_ptr(), seg.data_ptr(), mx.data_ptr(), self._stream()))
        self.launches += 1
        return energy.cpu().numpy(), noise.cpu().numpy().astype(bool), seg.cpu().numpy(), mx.cpu().numpy(), eoff

    # ------------------------------------------------------------------ template DTW
    def dtw(self, seq_feats: Sequence[np.ndarray], sample_feats: Sequence[np.ndarray], pruning: bool, pruning_factor: float,
            want_matrices: bool = False):
        """(best_idx int32 [n], best_dist float64 [n], dist float64 [n, W], cost, path) for samples vs templates;
        cost / path (float64 / int8, sample 0 only) are returned when ``want_matrices``."""
        torch = self.torch
        lens = np.array([f.shape[0] for f in seq_feats], dtype=np.int32)
        starts = np.concatenate(([0], np.cumsum(lens)[:-1])).astype(np.int32)
        H, W = int(lens.sum()), len(lens)
        D = int(seq_feats[0].shape[1])
        row_start = np.zeros(H + 1, dtype=np.int32)
        row_start[1:] = np.repeat(starts, lens)
        boundary = np.zeros(H + 1, dtype=np.int32)
        boundary[starts[1:]] = 1
        seq = self._to_dev(np.concatenate([np.asarray(f, dtype=np.float32) for f in seq_feats]))
        sl = np.array([f.shape[0] for f in sample_feats], dtype=np.int64)
        soff = np.concatenate(([0], np.cumsum(sl))).astype(np.int64)
        samp = self._to_dev(np.concatenate([np.asarray(f, dtype=np.float32) for f in sample_feats]))
        n = len(sample_feats)
        dist = self.empty((n, W), torch.float64)
        bi = self.empty((n,), torch.int32)
        bd = self.empty((n,), torch.float64)
        cost = path = None
        if want_matrices:
            L0 = int(sl[0])
            cost = self.empty((H + 1, L0 + 1), torch.float64)
            path = self.empty((H + 1, L0 + 1), torch.int8)
        # keep every table alive until the launch is enqueued (a temporary would be freed, and its block reused, first)
        t_rs, t_b, t_st, t_ln, t_so = (self._to_dev(row_start), self._to_dev(boundary), self._to_dev(starts), self._to_dev(lens),
                                       self._to_dev(soff))
        _native.check(self.lib.loe_dtw_dev(seq.data_ptr(), H, D, t_rs.data_ptr(), t_b.data_ptr(),
                                           t_st.data_ptr(), t_ln.data_ptr(), W, samp.data_ptr(),
                                           t_so.data_ptr(), n, 1 if pruning else 0, float(pruning_factor),
                                           dist.data_ptr(), bi.data_ptr(), bd.data_ptr(), self._p(cost), self._p(path), self._stream()))
        self.launches += 1
        return (bi.cpu().numpy(), bd.cpu().numpy(), dist.cpu().numpy(),
                None if cost is None else cost.cpu().numpy(), None if path is None else path.cpu().numpy())

    # ------------------------------------------------------------------ K-means statistics
    def kmeans_stats(self, feat, path, frm_off, n_utt, total_frames, tp: TrellisPack, utt_tr, remux: bool,
                     n_glob: int, shift):
        """Returns (stats [n_glob, 1+D+D(D+1)/2] float64, counts [n_glob, n_glob] int32) on the device."""
        torch = self.torch
        dim = int(feat.shape[1])
        bucket = self.empty((total_frames,), torch.int16)
        counts = torch.zeros((n_glob, n_glob), dtype=torch.int32, device=self.device)
        _native.check(self.lib.loe_align_dev(path.data_ptr(), frm_off.data_ptr(), n_utt, tp.tr_off.data_ptr(), tp.col.data_ptr(),
                                             tp.word.data_ptr(), tp.word_lo.data_ptr(), self._p(utt_tr), 1 if remux else 0,
                                             n_glob, bucket.data_ptr(), counts.data_ptr(), self._stream()))
        stride = 1 + dim + dim * (dim + 1) // 2
        ws = self.empty((int(self.lib.loe_kmeans_ws_doubles(total_frames, n_glob, dim)),), torch.float64)
        stats = self.empty((n_glob, stride), torch.float64)
        _native.check(self.lib.loe_kmeans_dev(feat.data_ptr(), bucket.data_ptr(), total_frames, dim, n_glob, shift.data_ptr(),
                                              ws.data_ptr(), stats.data_ptr(), self._stream()))
        self.launches += 3
        return stats, counts, bucket


def _prefer_spawn() -> None:
    """The reference's scripts map ``predict`` over ``concurrent.futures.ProcessPoolExecutor()``
    (scripts/project3_predict_simple.py:23-27, project5_test_*.py:33-41) after computing MFCCs in the
    parent.  A CUDA context does not survive fork(), so from the moment THIS process owns one, pools
    created with the default context must not fork it (workers create their own engine lazily).  The
    default becomes ``forkserver`` with this package and torch preloaded into the server: the server never
    touches CUDA, so its children are clean, and they start in milliseconds instead of re-importing torch
    for every pool (the scripts open one pool per label / per penalty value: 22 pools in
    project3_predict_simple.py, 100 in project5_find_trans_ndigits_with_sil.py).  Done when the engine is
    created -- never at import -- and only if the application has not chosen a (non-fork) start method itself;
    LOE_B200_START_METHOD=spawn selects plain spawn, LOE_B200_KEEP_START_METHOD=1 opts out."""
    import multiprocessing
    if os.environ.get("LOE_B200_KEEP_START_METHOD"):
        return
    method = os.environ.get("LOE_B200_START_METHOD", "forkserver")
    if method not in ("forkserver", "spawn"):
        raise ValueError(f"LOE_B200_START_METHOD={method!r}: 'forkserver' or 'spawn' (a CUDA context does not survive fork)")
    try:
        # "fork" counts as not chosen: it is what an unset default turns into as soon as anything asks for a context
        # (tqdm creates a multiprocessing lock for its first progress bar: scripts/project3_predict_simple.py:15 does so
        # before the first MFCC), and a forked worker could never use this process's CUDA context anyway
        current = multiprocessing.get_start_method(allow_none=True)
        if current is None or current == "fork":
            if method == "forkserver" and "forkserver" not in multiprocessing.get_all_start_methods():
                method = "spawn"
            multiprocessing.set_start_method(method, force=True)
            if method == "forkserver":
                multiprocessing.set_forkserver_preload(["__main__", "torch", __name__.split(".")[0]])
    except RuntimeError:
        pass


def get_engine() -> Engine:
    """The engine of this process (created on first use; never inherited across fork)."""
    global _ENGINE, _ENGINE_PID
    if _ENGINE is None or _ENGINE_PID != os.getpid():
        if _ENGINE is not None:
            raise RuntimeError("loe_speech_recognition: this process was forked from one that already owns a CUDA "
                               "context; use the 'forkserver' or 'spawn' start method (multiprocessing.get_context('spawn'))")
        _ENGINE = Engine()
        _ENGINE_PID = os.getpid()
        _prefer_spawn()
    return _ENGINE
