"""CPU: host-side mirror of the reference's chunk driver, driven by the oracle's forward_fn."""
import functools
import os
import wave

import numpy as np
import torch

import golden_util as gu


def test_process_mirror_chunked_matches_reference(oracle, tmp_path):
    """torch_fdtd_string_b200.process (mirror of reference src/task/simulate.py:16-119) with the
    oracle's forward_fn plugged in reproduces the reference's chunked run, incl. the in-place u_H carry."""
    from torch_fdtd_string_b200.simulate import process
    g = gu.load_golden("hammer_b2_chunked")
    assert int(g["chunk_size"]) < int(g["Nt"])
    inp = gu.build_inputs(g)
    save_path = str(tmp_path / "x" / str(int(g["sr"])))
    os.makedirs(tmp_path / "x", exist_ok=True)
    out = process("unused", inp["state_u"], inp["state_z"], inp["string_params"], inp["bow_params"],
                  inp["hammer_params"], inp["bow_mask"], inp["hammer_mask"], inp["consts"], inp["Nt"],
                  inp["chunk_size"], save_path, True, inp["relative_order"], inp["surface_integral"], False,
                  forward=oracle.forward_fn)
    uout, zout, su, sz, v_r, F_H, u_H, sig0, sig1 = out
    assert uout.shape == (int(g["B"]), int(g["Nt"]) - 2)
    assert gu.rel_l2(uout.numpy(), g["uout"]) < 1e-10
    assert gu.rel_l2(F_H.numpy(), g["F_H_out"]) < 1e-10
    assert gu.rel_l2(u_H.numpy(), g["u_H_out"]) < 1e-10
    # partial wav files like the reference's write_during_process
    w = wave.open(str(tmp_path / "x-0" / "output-u.wav"))
    assert w.getframerate() == int(g["sr"]) and w.getnframes() == int(g["Nt"]) - 2 and w.getsampwidth() == 2


def test_chunked_equals_unchunked(oracle):
    g = gu.load_golden("hammer_b2_chunked")
    a = gu.build_inputs(g)
    b = gu.build_inputs(g); b["chunk_size"] = b["Nt"]
    oa = gu.run_process(oracle.forward_fn, a); ob = gu.run_process(oracle.forward_fn, b)
    assert torch.equal(oa["uout"], ob["uout"]) and torch.equal(oa["u_H_out"], ob["u_H_out"])


def test_wav_writer_roundtrip(tmp_path):
    from torch_fdtd_string_b200.wavio import write_wav
    x = np.sin(np.arange(480) * 0.1) * 0.5
    for sub, width, scale in (("PCM_16", 2, 32768.0), ("PCM_24", 3, 8388608.0)):
        p = str(tmp_path / f"{sub}.wav")
        write_wav(p, x, 48000, sub)
        w = wave.open(p)
        assert (w.getsampwidth(), w.getnframes(), w.getframerate()) == (width, 480, 48000)
        raw = np.frombuffer(w.readframes(480), dtype=np.uint8).reshape(-1, width)
        v = np.zeros(480, dtype=np.int64)
        for i in range(width):
            v |= raw[:, i].astype(np.int64) << (8 * i)
        v = np.where(v >= 1 << (8 * width - 1), v - (1 << (8 * width)), v)
        assert np.abs(v / scale - x).max() <= 1.0 / scale
