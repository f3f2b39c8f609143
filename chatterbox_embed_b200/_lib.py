"""ctypes binding of libcbx.so (include/cbx.h).  No CPU fallback: if the library is missing it is built with nvcc,
and if that is impossible the import of the compute classes fails loudly."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import numpy as np

from . import build as _build

OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_WORKSPACE = 0, -1, -2, -3, -4
CLIP_VE_TOO_SHORT, CLIP_XV_TOO_SHORT, CLIP_VE_NAN = 1, 2, 4
DO_VE, DO_XV, NO_TRIM, PCM_PINNED = 1, 2, 4, 8

_lib: Optional[C.CDLL] = None


class ClipPlan(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ("n_samples", "ve_frames", "ve_partials", "ve_target", "xv_frames", "xv_tdnn", "xv_segments")]


class CbxError(RuntimeError):
    pass


_P = C.POINTER
_SIGS = {
    "cbx_ve_frame_step": (C.c_int, [C.c_double, C.c_double]),
    "cbx_ve_num_wins": (C.c_int, [C.c_int64, C.c_int, C.c_double, _P(C.c_int64), _P(C.c_int64)]),
    "cbx_plan_clip": (C.c_int, [C.c_int64, C.c_int, C.c_double, _P(ClipPlan)]),
    "cbx_trim_num_frames": (C.c_int64, [C.c_int64]),
    "cbx_clip_cost": (C.c_double, [C.c_int64]),
    "cbx_partition": (C.c_int, [_P(C.c_int64), C.c_int64, C.c_int, _P(C.c_int32), _P(C.c_int64), _P(C.c_double)]),
    "cbx_create": (C.c_int, [C.c_int, _P(C.c_void_p)]),
    "cbx_destroy": (None, [C.c_void_p]),
    "cbx_last_error": (C.c_char_p, [C.c_void_p]),
    "cbx_version": (C.c_char_p, []),
    "cbx_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "cbx_get_option": (C.c_int64, [C.c_void_p, C.c_char_p]),
    "cbx_load_weights": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _P(C.c_char_p), _P(C.c_void_p), _P(C.c_int64)]),
    "cbx_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_int, _P(C.c_int64), C.c_int, C.c_double, C.c_int]),
    "cbx_embed": (C.c_int, [C.c_void_p, C.c_void_p, _P(C.c_int64), C.c_int, C.c_float, C.c_int, C.c_double,
                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "cbx_embed_host": (C.c_int, [C.c_void_p, C.c_void_p, _P(C.c_int64), C.c_int, C.c_float, C.c_int, C.c_double,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "cbx_embed_host_submit": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, _P(C.c_int64), C.c_int, C.c_float, C.c_int, C.c_double, C.c_int]),
    "cbx_embed_host_wait": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cbx_resample_out_len": (C.c_int64, [C.c_int, C.c_int, C.c_int64]),
    "cbx_prompt_mel_frames": (C.c_int64, [C.c_int64]),
    "cbx_prompt_mel": (C.c_int, [C.c_void_p, C.c_void_p, _P(C.c_int64), C.c_int, C.c_void_p, C.c_void_p]),
    "cbx_s3_log_mel_frames": (C.c_int64, [C.c_int64]),
    "cbx_s3_log_mel": (C.c_int, [C.c_void_p, C.c_void_p, _P(C.c_int64), C.c_int, C.c_void_p, C.c_void_p]),
    "cbx_project": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cbx_resample": (C.c_int, [C.c_void_p, C.c_void_p, _P(C.c_int64), C.c_int, C.c_int, C.c_int, C.c_void_p, _P(C.c_int64), C.c_void_p]),
    "cbx_ve_forward_partials": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "cbx_ve_forward_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_int]),
    "cbx_campplus_forward_feats": (C.c_int, [C.c_void_p, C.c_void_p, _P(C.c_int64), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "cbx_campplus_forward_workspace_bytes": (C.c_int64, [C.c_void_p, _P(C.c_int64), C.c_int]),
    "cbx_locate": (C.c_int, [C.c_void_p, C.c_char_p, _P(C.c_int64), _P(C.c_int64), _P(C.c_int64), _P(C.c_int64)]),
    "cbx_clip_rows": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_int64), _P(C.c_int64), _P(C.c_int64), _P(C.c_int64)]),
    "cbx_launch_count": (C.c_int64, [C.c_void_p]),
    "cbx_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "cbx_profile_report": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_int64]),
}
EXPORTS = tuple(_SIGS)


def lib() -> C.CDLL:
    """Load (building first if stale/missing) libcbx.so.  Raises if it cannot be had."""
    global _lib
    if _lib is None:
        path = os.environ.get("CBX_LIB") or _build.LIB            # CBX_LIB: A/B runs of two builds of the library (tools/ab_bench.sh)
        if os.environ.get("CBX_NO_BUILD") != "1" and not os.environ.get("CBX_LIB"):
            path = _build.build()
        if not os.path.exists(path):
            raise CbxError(f"{path} is missing and could not be built; there is no CPU fallback")
        L = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def frame_step(overlap: float = 0.5, rate: Optional[float] = None) -> int:
    s = lib().cbx_ve_frame_step(float(overlap), float(rate) if rate else 0.0)
    assert s > 0, "0 < frame_step <= ve_partial_frames (voice_encoder.py:80)"
    return s


def num_wins(n_frames: int, step: int, min_coverage: float):
    assert n_frames > 0
    a, b = C.c_int64(), C.c_int64()
    rc = lib().cbx_ve_num_wins(int(n_frames), int(step), float(min_coverage), C.byref(a), C.byref(b))
    if rc:
        raise CbxError("cbx_ve_num_wins: bad argument")
    return a.value, b.value


def plan_clip(n_samples: int, step: int = 77, min_coverage: float = 0.8) -> ClipPlan:
    p = ClipPlan()
    if lib().cbx_plan_clip(int(n_samples), int(step), float(min_coverage), C.byref(p)):
        raise CbxError("cbx_plan_clip: bad argument")
    return p


def resample_out_len(src_sr: int, dst_sr: int, n_samples: int) -> int:
    return int(lib().cbx_resample_out_len(int(src_sr), int(dst_sr), int(n_samples)))


def prompt_mel_frames(n_samples: int) -> int:
    return int(lib().cbx_prompt_mel_frames(int(n_samples)))


def s3_log_mel_frames(n_samples: int) -> int:
    return int(lib().cbx_s3_log_mel_frames(int(n_samples)))


def clip_cost(n_samples: int) -> float:
    return lib().cbx_clip_cost(int(n_samples))


def _i64(values):
    """Sequence / array of integers -> (keep-alive int64 array, const int64_t* for the C call); no per-element Python."""
    arr = np.ascontiguousarray(values, dtype=np.int64)
    return arr, arr.ctypes.data_as(_P(C.c_int64))


def partition(lengths: Sequence[int], world: int):
    """cbx_partition: (rank_of int32 [n], row_of int64 [n], rank_cost float64 [world])."""
    lens = np.ascontiguousarray(lengths, dtype=np.int64)
    rank_of, row_of, cost = np.empty(len(lens), np.int32), np.empty(len(lens), np.int64), np.empty(world, np.float64)
    rc = lib().cbx_partition(lens.ctypes.data_as(_P(C.c_int64)), len(lens), int(world), rank_of.ctypes.data_as(_P(C.c_int32)),
                             row_of.ctypes.data_as(_P(C.c_int64)), cost.ctypes.data_as(_P(C.c_double)))
    if rc:
        raise CbxError("cbx_partition: bad argument (negative length or world <= 0)")
    return rank_of, row_of, cost


class Context:
    """One libcbx context per GPU."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = lib().cbx_create(int(device), C.byref(self._h))
        if rc:
            raise CbxError(f"cbx_create failed ({rc}): {lib().cbx_last_error(None).decode()} -- "
                           "this package has no CPU fallback; a B200 (sm_100) is required")
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().cbx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc < 0:
            raise CbxError(f"{what} failed ({rc}): {lib().cbx_last_error(self._h).decode()}")
        return rc

    def set_option(self, key: str, value: int):
        self._check(lib().cbx_set_option(self._h, key.encode(), int(value)), "cbx_set_option")

    def get_option(self, key: str) -> int:
        return lib().cbx_get_option(self._h, key.encode())

    def load_weights(self, which: int, tensors: Dict[str, np.ndarray]):
        names = list(tensors)
        arrs = [np.ascontiguousarray(tensors[k], dtype=np.float32) for k in names]
        n = len(names)
        c_names = (C.c_char_p * n)(*[k.encode() for k in names])
        c_ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        c_num = (C.c_int64 * n)(*[a.size for a in arrs])
        self._check(lib().cbx_load_weights(self._h, which, n, c_names, c_ptrs, c_num), "cbx_load_weights")

    def workspace_bytes(self, lengths: Sequence[int], step: int, min_cov: float, flags: int) -> int:
        keep, arr = _i64(lengths)
        return self._check(lib().cbx_workspace_bytes(self._h, len(keep), arr, step, min_cov, flags), "cbx_workspace_bytes")

    def embed(self, pcm_ptr: int, offsets: Sequence[int], trim_top_db: float, step: int, min_cov: float,
              ve_ptr: int, xv_ptr: int, status_ptr: int, ws_ptr: int, ws_bytes: int, stream: int, flags: int):
        n = len(offsets) - 1
        keep, off = _i64(offsets)
        self._check(lib().cbx_embed(self._h, pcm_ptr, off, n, float(trim_top_db), step, min_cov,
                                    ve_ptr, xv_ptr, status_ptr, ws_ptr, ws_bytes, stream, flags), "cbx_embed")

    def embed_host(self, pcm: np.ndarray, offsets: np.ndarray, trim_top_db: float, step: int, min_cov: float, flags: int):
        """pcm: flat float32 numpy array (may be a view of page-locked memory: pass PCM_PINNED in flags)."""
        n = len(offsets) - 1
        assert pcm.dtype == np.float32 and pcm.flags.c_contiguous
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        ve = np.empty((n, 256), np.float32) if flags & DO_VE else None
        xv = np.empty((n, 192), np.float32) if flags & DO_XV else None
        status = np.zeros(n, np.int32)
        self._check(lib().cbx_embed_host(self._h, pcm.ctypes.data, off.ctypes.data_as(_P(C.c_int64)), n, float(trim_top_db),
                                         step, min_cov, ve.ctypes.data if ve is not None else None,
                                         xv.ctypes.data if xv is not None else None, status.ctypes.data, flags),
                    "cbx_embed_host")
        return ve, xv, status

    def embed_host_submit(self, slot: int, pcm: np.ndarray, offsets: np.ndarray, trim_top_db: float, step: int, min_cov: float, flags: int):
        """Streaming form: enqueue copies + kernels for one batch in slot 0/1 and return at once.  ``pcm`` (and, with
        PCM_PINNED, its contents) must stay alive until ``embed_host_wait(slot)``."""
        n = len(offsets) - 1
        assert pcm.dtype == np.float32 and pcm.flags.c_contiguous
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        self._check(lib().cbx_embed_host_submit(self._h, int(slot), pcm.ctypes.data, off.ctypes.data_as(_P(C.c_int64)), n,
                                                float(trim_top_db), step, min_cov, flags), "cbx_embed_host_submit")
        self.__dict__.setdefault("_pending", {})[int(slot)] = (n, flags, pcm)

    def embed_host_wait(self, slot: int):
        n, flags, _keepalive = self.__dict__.get("_pending", {}).pop(int(slot))
        ve = np.empty((n, 256), np.float32) if flags & DO_VE else None
        xv = np.empty((n, 192), np.float32) if flags & DO_XV else None
        status = np.zeros(n, np.int32)
        self._check(lib().cbx_embed_host_wait(self._h, int(slot), ve.ctypes.data if ve is not None else None,
                                              xv.ctypes.data if xv is not None else None, status.ctypes.data), "cbx_embed_host_wait")
        return ve, xv, status

    def campplus_forward_workspace_bytes(self, frame_offsets: Sequence[int]) -> int:
        keep, a = _i64(frame_offsets)
        n = lib().cbx_campplus_forward_workspace_bytes(self._h, a, len(frame_offsets) - 1)
        if n < 0:
            raise CbxError("cbx_campplus_forward_workspace_bytes: bad argument")
        return int(n)

    def campplus_forward_feats(self, feats_ptr: int, frame_offsets: Sequence[int], xv_ptr: int, status_ptr: int, ws_ptr: int, ws_bytes: int, stream: int):
        keep, a = _i64(frame_offsets)
        self._check(lib().cbx_campplus_forward_feats(self._h, feats_ptr, a, len(frame_offsets) - 1, xv_ptr, status_ptr, ws_ptr, ws_bytes, stream),
                    "cbx_campplus_forward_feats")

    def resample(self, x_ptr: int, in_offsets: Sequence[int], src_sr: int, dst_sr: int, y_ptr: int, out_offsets: Sequence[int], stream: int):
        n = len(in_offsets) - 1
        keep_a, a = _i64(in_offsets)
        keep_b, b = _i64(out_offsets)
        self._check(lib().cbx_resample(self._h, x_ptr, a, n, int(src_sr), int(dst_sr), y_ptr, b, stream), "cbx_resample")

    def prompt_mel(self, pcm_ptr: int, offsets: Sequence[int], out_ptr: int, stream: int):
        n = len(offsets) - 1
        keep, a = _i64(offsets)
        self._check(lib().cbx_prompt_mel(self._h, pcm_ptr, a, n, out_ptr, stream), "cbx_prompt_mel")

    def s3_log_mel(self, pcm_ptr: int, offsets: Sequence[int], out_ptr: int, stream: int):
        n = len(offsets) - 1
        keep, a = _i64(offsets)
        self._check(lib().cbx_s3_log_mel(self._h, pcm_ptr, a, n, out_ptr, stream), "cbx_s3_log_mel")

    def project(self, x_ptr: int, n: int, in_dim: int, w_ptr: int, b_ptr: Optional[int], out_dim: int, normalize: bool, y_ptr: int, stream: int):
        self._check(lib().cbx_project(self._h, x_ptr, int(n), int(in_dim), w_ptr, b_ptr, int(out_dim), int(bool(normalize)), y_ptr, stream), "cbx_project")

    def ve_forward_workspace_bytes(self, n: int) -> int:
        return self._check(lib().cbx_ve_forward_workspace_bytes(self._h, n), "cbx_ve_forward_workspace_bytes")

    def ve_forward_partials(self, mels_ptr: int, n: int, out_ptr: int, ws_ptr: int, ws_bytes: int, stream: int):
        self._check(lib().cbx_ve_forward_partials(self._h, mels_ptr, n, out_ptr, ws_ptr, ws_bytes, stream), "cbx_ve_forward_partials")

    def locate(self, name: str):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self._check(lib().cbx_locate(self._h, name.encode(), C.byref(a), C.byref(b), C.byref(c), C.byref(d)), "cbx_locate")
        return a.value, b.value, c.value, d.value

    def clip_rows(self, clip: int):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self._check(lib().cbx_clip_rows(self._h, clip, C.byref(a), C.byref(b), C.byref(c), C.byref(d)), "cbx_clip_rows")
        return {"mel_row": a.value, "slot": b.value, "fb_row": c.value, "td_row": d.value}

    def launch_count(self) -> int:
        return lib().cbx_launch_count(self._h)

    def profile_enable(self, on: bool):
        self._check(lib().cbx_profile_enable(self._h, 1 if on else 0), "cbx_profile_enable")

    def profile_report(self):
        """{tag: dict(launches, ms, flops, bytes)} for everything launched since profile_enable(True)."""
        n = lib().cbx_profile_report(self._h, None, 0)
        buf = C.create_string_buffer(int(n) + 16)
        lib().cbx_profile_report(self._h, buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            tag, cnt, ms, fl, by = line.split()
            out[tag] = dict(launches=int(cnt), ms=float(ms), flops=float(fl), bytes=float(by))
        return out


_contexts: Dict[int, Context] = {}


def context(device: int = 0) -> Context:
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]
