"""Minimal RIFF/WAVE reader for the ingest side of the path (replaces ``torchaudio.load`` / ``torchaudio.info`` in
``load_audio`` ref:53-59 and ``discover_two_files`` ref:129-137, which need TorchCodec and do not work in this image).

Only the container is parsed on the host.  PCM16 samples are handed to the GPU as they are (``(frames, channels)``
int16): the 2^-15 scaling and the channel mean are fused into the resampler (``zk_resample_pcm16``), which also halves
the H2D bytes of a 48 kHz recording.  Other encodings (8/24/32-bit PCM, 32/64-bit float) are converted to
``(channels, frames)`` float32 on the host with torchaudio's normalisation (full scale = 1.0).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import BinaryIO, Callable, Optional, Tuple

import numpy as np

WAVE_FORMAT_PCM, WAVE_FORMAT_IEEE_FLOAT, WAVE_FORMAT_EXTENSIBLE = 0x0001, 0x0003, 0xFFFE


class WavError(ValueError):
    pass


@dataclass
class WavInfo:
    sample_rate: int
    num_frames: int
    num_channels: int
    bits_per_sample: int
    encoding: str          # "PCM_S", "PCM_U" (8 bit) or "PCM_F"
    data_offset: int
    data_bytes: int


def _read_header(f: BinaryIO) -> WavInfo:
    head = f.read(12)
    if len(head) < 12 or head[:4] not in (b"RIFF", b"RF64") or head[8:12] != b"WAVE":
        raise WavError("not a RIFF/WAVE file")
    fmt = None
    ds64_data = None
    while True:
        ck = f.read(8)
        if len(ck) < 8:
            raise WavError("no data chunk")
        cid, size = ck[:4], struct.unpack("<I", ck[4:])[0]
        if cid == b"ds64":
            body = f.read(size)
            ds64_data = struct.unpack("<Q", body[8:16])[0]
        elif cid == b"fmt ":
            body = f.read(size)
            if size < 16:
                raise WavError("short fmt chunk")
            tag, ch, sr, _, align, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == WAVE_FORMAT_EXTENSIBLE and size >= 26:
                tag = struct.unpack("<H", body[24:26])[0]
            fmt = (tag, ch, sr, align, bits)
        elif cid == b"data":
            if fmt is None:
                raise WavError("data chunk before fmt chunk")
            if size == 0xFFFFFFFF and ds64_data is not None:
                size = ds64_data
            tag, ch, sr, align, bits = fmt
            if ch < 1 or bits % 8 or align != ch * bits // 8:
                raise WavError(f"unsupported layout: {ch} channels, {bits} bits, block align {align}")
            if tag == WAVE_FORMAT_PCM and bits in (8, 16, 24, 32):
                enc = "PCM_U" if bits == 8 else "PCM_S"
            elif tag == WAVE_FORMAT_IEEE_FLOAT and bits in (32, 64):
                enc = "PCM_F"
            else:
                raise WavError(f"unsupported encoding: format tag {tag:#x}, {bits} bits")
            off = f.tell()
            f.seek(0, 2)
            size = min(size, f.tell() - off)  # tolerate a truncated file / a streaming writer's placeholder size
            return WavInfo(sr, size // align, ch, bits, enc, off, (size // align) * align)
        else:
            f.seek(size, 1)
        if size & 1:
            f.seek(1, 1)  # chunks are word aligned


def info(path: str) -> WavInfo:
    """What ``torchaudio.info`` gives ``discover_two_files`` (ref:132-133): ``num_frames`` and the format."""
    with open(path, "rb") as f:
        return _read_header(f)


def read(path: str, alloc: Optional[Callable[[Tuple[int, int], str], np.ndarray]] = None) -> Tuple[np.ndarray, int]:
    """-> ``(samples, sample_rate)``.  PCM16: int16 ``(frames, channels)`` (interleaved, as stored); everything else:
    float32 ``(channels, frames)`` in [-1, 1) like ``torchaudio.load`` (ref:54).

    ``alloc(shape, dtype)`` (``dtype`` "<i2" or "<f4") supplies the array the samples are delivered in -- the batch
    runner hands out page-locked memory, so the file is read straight into the buffer the H2D copy starts from."""
    alloc = alloc or (lambda shape, dtype: np.empty(shape, dtype=dtype))
    with open(path, "rb") as f:
        wi = _read_header(f)
        f.seek(wi.data_offset)
        n, c = wi.num_frames, wi.num_channels
        if wi.encoding == "PCM_S" and wi.bits_per_sample == 16:
            pcm = alloc((n, c), "<i2")  # read straight into the (writable) array handed to the H2D copy
            if pcm.shape != (n, c) or pcm.dtype != np.dtype("<i2") or not pcm.flags.c_contiguous:
                raise WavError("alloc() must return a C-contiguous array of the requested shape and dtype")
            got = f.readinto(memoryview(pcm).cast("B"))
            if got != wi.data_bytes:
                raise WavError(f"{path}: short read ({got} of {wi.data_bytes} bytes)")
            return pcm, wi.sample_rate
        raw = f.read(wi.data_bytes)
    if wi.encoding == "PCM_U":
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif wi.encoding == "PCM_S" and wi.bits_per_sample == 24:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        x = ((v ^ 0x800000) - 0x800000).astype(np.float32) / 8388608.0
    elif wi.encoding == "PCM_S":
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif wi.bits_per_sample == 32:
        x = np.frombuffer(raw, dtype="<f4").astype(np.float32)
    else:
        x = np.frombuffer(raw, dtype="<f8").astype(np.float32)
    out = alloc((c, n), "<f4")
    if out.shape != (c, n) or out.dtype != np.dtype("<f4"):
        raise WavError("alloc() must return an array of the requested shape and dtype")
    out[...] = x.reshape(n, c).T
    return out, wi.sample_rate


def write_pcm16(path: str, samples: np.ndarray, sample_rate: int) -> None:
    """``samples``: float ``(channels, frames)`` / ``(frames,)`` in [-1, 1] or int16 ``(frames, channels)`` (tests, tools)."""
    a = np.asarray(samples)
    if a.dtype != np.int16:
        a = np.atleast_2d(a)
        a = np.clip(np.round(a.T * 32768.0), -32768, 32767).astype("<i2")
    elif a.ndim == 1:
        a = a[:, None]
    n, c = a.shape
    data = np.ascontiguousarray(a, dtype="<i2").tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, WAVE_FORMAT_PCM, c, sample_rate, sample_rate * c * 2, c * 2, 16))
        f.write(b"data" + struct.pack("<I", len(data)) + data)
