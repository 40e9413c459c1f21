"""Reference-side sampler arguments of the golden fixtures (shared by make_golden.py and the sampler tests)."""
SR = 48000

NSYNTH = dict(
    sr=SR, f0_inf=98.0, alpha_inf=1, lambda_c=1, relative_order=4, theta=("auto", 0.03, 98.0),
    string_kwargs=dict(
        sampling_f0='random', sampling_kappa='random', sampling_alpha='random',
        sampling_pickup='random', sampling_T60='random', precorrect=True,
        f0_min=98.0, f0_max=440.0, f0_diff_max=30, f0_mod_max=0.08,
        kappa_min=0.01, kappa_max=0.03, alpha_min=1., alpha_max=25.,
        t60_min_1=10., t60_max_1=25., t60_min_2=10., t60_max_2=30.,
        sampling_p_a='random', p_a_max=0.02, sampling_p_x='random', p_x_max=0.5),
    hammer_kwargs=dict(M_r_min=1.0, M_r_max=10., alpha_fixed=3),
    bow_kwargs=dict(),
)

ALLFIXED = dict(
    sr=SR, f0_inf=55.0, alpha_inf=20, lambda_c=1, relative_order=8, theta=("auto", 0.08, 55.0),
    string_kwargs=dict(
        sampling_f0='fix', sampling_kappa='fix', sampling_alpha='fix',
        sampling_pickup='fix', sampling_T60='fix', precorrect=True,
        f0_fixed=55.0, kappa_fixed=0.08, alpha_fixed=20., lossless=False,
        sampling_p_a='fix', p_a_fixed=0.02, sampling_p_x='fix', p_x_fixed=0.2),
    hammer_kwargs=dict(x_H_min=0.1, x_H_max=0.1, v_H_min=4.0, v_H_max=4.0, M_r_min=1.5, M_r_max=1.5,
                       w_H_min=2000, w_H_max=2000),
    bow_kwargs=dict(x_b_min=0.2, x_b_max=0.2, v_b_min=0.35, v_b_max=0.35, F_b_min=90, F_b_max=90.,
                    phi_0_max=9., phi_0_min=9., phi_1_max=0.01, phi_1_min=0.01, wid_min=4, wid_max=4),
)

LINEAR = dict(
    sr=SR, f0_inf=55.0, alpha_inf=1, lambda_c=1, relative_order=8, theta=("auto", 0.03, 55.0),
    string_kwargs=dict(
        sampling_f0='fix', sampling_kappa='fix', sampling_alpha='fix',
        sampling_pickup='random', sampling_T60='fix', precorrect=False,
        f0_fixed=55.0, f0_mod_max=0, lossless=False, t60_fixed=20., kappa_min=0.03, kappa_max=0.03,
        kappa_fixed=0.03, alpha_fixed=1., alpha_min=1., alpha_max=1.,
        sampling_p_a='fix', p_a_fixed=0.01, sampling_p_x='fix', p_x_fixed=0.3, pluck_profile='smooth'),
    hammer_kwargs=dict(x_H_min=0.5, x_H_max=0.5, v_H_min=2.5, v_H_max=2.5, M_r_min=10., M_r_max=10.,
                       w_H_min=3000, w_H_max=3000, alpha_fixed=3),
    bow_kwargs=dict(),
)

# hammered string with tension modulation on a finer grid (BASELINE config 4, scaled down)
FINEHAMMER = dict(
    sr=96000, f0_inf=55.0, alpha_inf=3, lambda_c=1, relative_order=8, theta=("auto", 0.01, 55.0),
    string_kwargs=dict(
        sampling_f0='fix', sampling_kappa='fix', sampling_alpha='fix',
        sampling_pickup='random', sampling_T60='fix', precorrect=False,
        f0_fixed=55.0, f0_mod_max=0, lossless=False, t60_fixed=20., kappa_fixed=0.01, alpha_fixed=3.,
        sampling_p_a='fix', p_a_fixed=0.01, sampling_p_x='fix', p_x_fixed=0.25, pluck_profile='smooth'),
    hammer_kwargs=dict(x_H_min=0.3, x_H_max=0.3, v_H_min=2.5, v_H_max=2.5, M_r_min=10., M_r_max=10.,
                       w_H_min=3000, w_H_max=3000, alpha_fixed=3),
    bow_kwargs=dict(),
)

# BASELINE config 4 at its stated rate: 192 kHz (N_t = 237, N_l = 593 with pre-correction off)
FINEHAMMER192 = dict(FINEHAMMER, sr=192000)

# the manufactured-solution preset on coarser / finer grids (convergence-order harness, tests/test_gpu_convergence.py)
LINEAR12 = dict(LINEAR, sr=12000); LINEAR24 = dict(LINEAR, sr=24000); LINEAR96 = dict(LINEAR, sr=96000)

# the reference's own default ranges reach f0_min = 27.5 Hz and kappa_min = 0 (src/model/simulator.py:123): N_t up to ~556 rows
# at 48 kHz (the stiffness term shortens the grid: kappa_rel = 0.03 gives N_t ~ 100 at 25 Hz, so kappa is kept small here)
LOWF0 = dict(NSYNTH, theta=("auto", 0.001, 25.0), f0_inf=25.0, alpha_inf=1.5,
             string_kwargs=dict(NSYNTH['string_kwargs'], f0_min=27.5, f0_max=36.0, f0_diff_max=3, kappa_min=0.0002, kappa_max=0.002,
                                alpha_min=1.5, alpha_max=3., p_a_max=0.004))

# strings that blow up to NaN within tens of milliseconds in the reference (large alpha, large pluck amplitude, fine grid):
# pins the NaN mask and onset (src/task/simulate.py:91-93,333-334 drops such strings)
HOT = dict(NSYNTH, string_kwargs=dict(NSYNTH['string_kwargs'], f0_min=98.0, f0_max=140.0, alpha_min=18., alpha_max=25.,
                                      sampling_p_a='fix', p_a_fixed=0.02, kappa_min=0.01, kappa_max=0.015))

PRESETS = dict(nsynth=NSYNTH, lowf0=LOWF0, hot=HOT, allfixed=ALLFIXED, linear=LINEAR, finehammer=FINEHAMMER, finehammer192=FINEHAMMER192,
               linear12=LINEAR12, linear24=LINEAR24, linear96=LINEAR96)

