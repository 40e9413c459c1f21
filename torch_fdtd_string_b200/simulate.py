"""Chunk driver: mirror of the reference's ``process()`` (reference src/task/simulate.py:16-119)
with the JIT-built extension replaced by this package's ``forward_fn``.  Same arguments, same
return tuple, same in-place behaviour and the same per-chunk partial wav writes."""
import os

import torch

from .forward_fn import forward_fn
from .wavio import write_wav


def _chunk(x, n, size, axis=1):
    # reference simulate.py:38-45
    if not isinstance(x, torch.Tensor):
        return x
    if x.dim() > 1 and x.size(axis) > 2:
        x = x.narrow(axis, n, size)
    return x


def _chunk_params(params, n, chunk_size):
    # reference simulate.py:46-55
    params = list(params)
    params[-1] = chunk_size
    for i, p in enumerate(params):
        if isinstance(p, torch.Tensor):
            params[i] = _chunk(p, n, chunk_size)
        elif isinstance(p, (tuple, list)):
            params[i] = tuple(_chunk(pp, n, chunk_size) for pp in p)
    return params


def process(root_dir, state_u, state_z, string_params, bow_params, hammer_params,
            bow_mask, hammer_mask, consts, Nt, chunk_size, save_path=None, skip_nan=True,
            relative_order=4, surface_integral=False, manufactured=False, forward=forward_fn):
    """``root_dir`` is accepted for signature compatibility (the reference uses it to find and
    JIT-build its C++ sources, simulate.py:28-36); nothing is compiled here."""
    cn = 0
    tot = [[] for _ in range(5)]
    sig0 = sig1 = None
    while cn < Nt - 2:
        output_size = min(chunk_size, state_u.size(1) - cn)
        outputs = forward(*_chunk_params(
            (state_u, state_z, string_params, bow_params, hammer_params, bow_mask, hammer_mask,
             consts, relative_order, surface_integral, manufactured, cn, Nt), cn, output_size))
        uout, zout, c_state_u, c_state_z, v_r_out, F_H_out, u_H_out, sig0, sig1 = outputs
        # the chunk views alias state_u/state_z and were updated in place; keep the reference's
        # explicit copy so that a forward() returning fresh tensors also works
        if c_state_u.data_ptr() != state_u.narrow(1, cn, output_size).data_ptr():
            state_u[:, cn + 2:cn + output_size, :] = c_state_u[:, 2:2 + output_size, :]
            state_z[:, cn + 2:cn + output_size, :] = c_state_z[:, 2:2 + output_size, :]
        for lst, t in zip(tot, (uout, zout, v_r_out, F_H_out, u_H_out)):
            lst.append(t.narrow(1, 2, output_size - 2))
        cn += chunk_size - 2

        state_is_nan = torch.isnan(c_state_u.flatten(1).sum(-1))
        if not skip_nan:
            assert not state_is_nan.any(), state_is_nan.nonzero()
        if save_path is not None:
            _u = torch.cat(tot[0], dim=1); _z = torch.cat(tot[1], dim=1)
            p = save_path.split('/')
            sr = int(p.pop(-1)); sp = '/'.join(p)
            for b in range(_u.size(0)):
                if not state_is_nan[b]:
                    os.makedirs(f"{sp}-{b}", exist_ok=True)
                    write_wav(f'{sp}-{b}/output-u.wav', _u[b].cpu(), sr, 'PCM_16')
                    write_wav(f'{sp}-{b}/output-z.wav', _z[b].cpu(), sr, 'PCM_16')
                    write_wav(f'{sp}-{b}/output.wav', _u[b].cpu() + _z[b].cpu(), sr, 'PCM_16')
    total = [torch.cat(x, dim=1) for x in tot]
    return (total[0], total[1], state_u, state_z, total[2], total[3], total[4], sig0, sig1)
