"""GPU: convergence-order harness for the manufactured-solution mode.

The reference's only known-answer check is the method of manufactured solutions: forcing `vnv.cpp:11-37` (applied at
`string.cpp:227-232`), initial condition `src/model/simulator.py:175-180`, target `src/model/analytic.py:21-27`
    u(x, t) = p_a cos^2(pi x) cos(gamma t) exp(-sigma t),   x in [-1/2, 1/2],
preset `experiment/linear-string.yaml` (B = 1, double, relative_order 8, alpha = 1).  The fixtures `manufactured_sr*`
hold the reference's inputs and its full state history for the SAME physical time (10 ms) at sr = 12 / 24 / 48 / 96 kHz
(grids N_t = 33 / 47 / 67 / 95: the stiffness term makes h ~ sqrt(k)).  The harness runs the CUDA stepper on each grid,
measures the max-norm error against the analytic solution at the final step and the observed order of accuracy in h between
successive grids -- and checks both against what the reference itself achieves (its error at 48 kHz is 1.2e-4 = 2.7 % of
the amplitude, observed order ~0.9 in h: the scheme's truncation error at these grids, not a property of this port).
"""
import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu

RATES = ("12k", "24k", "48k", "96k")


def grid_size(g):
    """N_t of the run (string.cpp:23-40 with the float32 constants of simulator.cpp:22)"""
    k = float(np.float32(g["consts"][0])); th = np.float32(g["consts"][1]); lam = float(np.float32(g["consts"][2]))
    tt1 = float(np.float32(2 * th - 1)); tt2 = float(np.float32(2 * np.float32(tt1)))
    gamma = 2 * float(g["f0"].ravel()[0]); K = gamma * float(g["kappa"].ravel()[0])
    h1 = lam * np.sqrt((gamma ** 2 * k ** 2 + np.sqrt(gamma ** 4 * k ** 4 + 16 * K ** 2 * k ** 2 * tt1)) / tt2)
    return int(np.floor(1 / h1))


def analytic(g, N, n):
    """src/model/analytic.py:21-27 at time sample n on the N+1 grid points"""
    x = np.arange(N + 1) / N - 0.5
    t = n / float(g["sr"])
    gamma = 2 * float(g["f0"].ravel()[0])
    return float(g["p_a"].ravel()[0]) * np.cos(np.pi * x) ** 2 * np.cos(gamma * t) * np.exp(-float(g["sig0"].ravel()[0]) * t)


def run(g):
    from torch_fdtd_string_b200 import process
    inp = gu.build_inputs(g, device="cuda")
    out = process("unused", inp["state_u"], inp["state_z"], inp["string_params"], inp["bow_params"], inp["hammer_params"],
                  inp["bow_mask"], inp["hammer_mask"], inp["consts"], inp["Nt"], inp["chunk_size"], None, True,
                  inp["relative_order"], inp["surface_integral"], True)
    return out[2][0].cpu().numpy()                      # state_u history (Nt, Nx_t1)


def test_manufactured_convergence_order():
    rows = []
    for r in RATES:
        name = f"manufactured_sr{r}"
        try:
            g = gu.load_golden(name)
        except FileNotFoundError:
            continue
        su = run(g)
        N, n = grid_size(g), int(g["Nt"]) - 1
        ua = analytic(g, N, n)
        e_cuda = np.abs(su[n, :N + 1] - ua).max()
        e_ref = np.abs(g["state_u_full"][0][n, :N + 1] - ua).max()
        d = np.abs(su - g["state_u_full"][0]).max() / np.abs(g["state_u_full"][0]).max()
        rows.append((r, N, e_cuda, e_ref, d))
        assert d < 1e-9, (name, d)                                             # the whole history is the reference's
        assert abs(e_cuda - e_ref) <= 1e-6 * e_ref, (name, e_cuda, e_ref)      # and so is its distance to the analytic solution
    assert len(rows) >= 3, rows
    orders = []
    for (r0, N0, e0, f0, _), (r1, N1, e1, f1, _) in zip(rows[:-1], rows[1:]):
        p_cuda = np.log(e0 / e1) / np.log(N1 / N0); p_ref = np.log(f0 / f1) / np.log(N1 / N0)
        orders.append((r0, r1, p_cuda, p_ref))
        assert abs(p_cuda - p_ref) < 1e-5
        assert 0.6 < p_cuda < 1.5, orders          # observed order in h of the reference scheme at these grids (measured 0.83 / 0.92)
    for row in rows:
        print("sr %s: N_t %d  |u - analytic|_max CUDA %.6e  reference %.6e  CUDA-vs-reference history %.1e" % row)
    for o in orders:
        print("order in h between %s and %s: CUDA %.4f  reference %.4f" % o)
    assert all(a[2] > b[2] for a, b in zip(rows[:-1], rows[1:]))               # the error decreases under refinement
