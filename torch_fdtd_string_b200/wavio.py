"""Minimal PCM wav writer (the image has no `soundfile`): PCM_16 / PCM_24 like the subtypes the
reference passes to ``sf.write`` (reference src/task/simulate.py:103-106, 416-425)."""
import wave

import numpy as np


def write_wav(path, data, sr, subtype="PCM_16"):
    x = np.asarray(data, dtype=np.float64).reshape(-1)
    x = np.nan_to_num(x)
    if subtype == "PCM_16":
        q = np.clip(np.round(x * 32768.0), -32768, 32767).astype("<i2").tobytes()
        width = 2
    elif subtype == "PCM_24":
        v = np.clip(np.round(x * 8388608.0), -8388608, 8388607).astype("<i4")
        b = v.view(np.uint8).reshape(-1, 4)[:, :3]
        q = b.tobytes()
        width = 3
    else:
        raise ValueError(subtype)
    with wave.open(path, "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(width)
        f.setframerate(int(sr))
        f.writeframes(q)


def write_wav_pcm(path, raw, sr, bits):
    """mono wav from already quantised little-endian PCM bytes (what ``sfdtd_postprocess`` produces on the device):
    ``raw`` = uint8 array of n_samples * bits/8 bytes"""
    with wave.open(path, "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(bits // 8)
        f.setframerate(int(sr))
        f.writeframes(np.ascontiguousarray(raw, dtype=np.uint8).tobytes())
