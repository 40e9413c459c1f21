/*
 * sfdtd.h -- C ABI of the B200-native StringFDTD time stepper (libsfdtd.so).
 *
 * This is the drop-in boundary for the reference's only native entry point,
 *
 *   vector<Tensor> forward_fn(state_u, state_z, string_params, bow_params,
 *                             hammer_params, bow_mask, hammer_mask, constant,
 *                             relative_error, surface_integral, manufactured,
 *                             n_0, Nt)
 *   (reference src/model/cpp/simulator.cpp:14-27, bound with pybind11 at :61-63,
 *    loaded by torch.utils.cpp_extension.load at src/task/simulate.py:28-36 and
 *    called at src/task/simulate.py:65-76).
 *
 * Plain pointers, sizes and strides only -- no torch types.  All data pointers
 * are DEVICE pointers (CUDA, same device as the current context).  Every array
 * has its own element strides so that the reference's `narrow()`ed chunk views
 * (src/task/simulate.py:38-55) and compact inputs with a constant time axis
 * (time stride 0) are both expressible without copies.
 *
 * Semantics follow reference src/model/cpp/string.cpp:43-306 step by step (see
 * DESIGN.md); "group" = the strings of one reference batch: they share the
 * batch-max operator widths (misc.cpp:119-127) and the any-over-batch
 * convergence votes (string.cpp:252-253, hammer.cpp:51).
 */
#ifndef SFDTD_H_
#define SFDTD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFDTD_ABI_VERSION 2

/* dtype of all floating-point arrays */
enum { SFDTD_F64 = 0, SFDTD_F32 = 1 };

/* flags */
enum {
    SFDTD_SURFACE_INTEGRAL = 1u << 0, /* pickup = surface integral of velocities (string.cpp:274-291), else interpolated pickup at `pos` (:293-298) */
    SFDTD_MANUFACTURED     = 1u << 1, /* add the manufactured-solution forcing (vnv.cpp:11-37, string.cpp:227-232) */
    SFDTD_SAVE_STATE       = 1u << 2, /* state_u/state_z hold the full (B,Nt,Nx) history: row n is read-modified-written in place
                                         every step like the reference (string.cpp:264-265).  Without it they hold 2 time rows
                                         [n-2, n-1] per string: read at the start, overwritten with the last two rows at the end. */
    SFDTD_SKIP_AUX         = 1u << 3  /* do not evaluate v_r for un-bowed strings nor F_H/u_H for un-hammered strings (they are
                                         written as 0); audio outputs are unaffected.  Off = reference-faithful. */
};

/* per-string status bits written to `status` (may be NULL) */
enum {
    SFDTD_ST_SOLVER_CAP = 1u << 0, /* inner linear iteration hit its cap (result of that step not converged) */
    SFDTD_ST_OUTER_CAP  = 1u << 1, /* fixed-point loop (string.cpp:200) hit max_iter -- the reference would spin forever */
    SFDTD_ST_HAMMER_CAP = 1u << 2, /* hammer loop (hammer.cpp:33) hit max_iter */
    SFDTD_ST_BOW_WINDOW = 1u << 3, /* raised-cosine support wider than the kernel supports */
    SFDTD_ST_RANGE      = 1u << 4  /* a grid size left the range the launch was sized for */
};

/* return codes of sfdtd_forward */
enum {
    SFDTD_OK = 0,
    SFDTD_ERR_ARG = -1,        /* bad argument (null pointer, size, dtype, abi) */
    SFDTD_ERR_UNSUPPORTED = -2,/* configuration outside the built kernel set (grid too wide, group too large) */
    SFDTD_ERR_CUDA = -3        /* CUDA runtime error; see sfdtd_last_error() */
};

typedef struct sfdtd_array {
    void   *ptr;   /* device pointer, may be NULL only where stated */
    int64_t bs;    /* element stride between strings (batch axis) */
    int64_t ts;    /* element stride between time samples (0 = constant in time); ignored for per-string scalars */
} sfdtd_array;

typedef struct sfdtd_args {
    int32_t abi_version;   /* SFDTD_ABI_VERSION */
    int32_t dtype;         /* SFDTD_F64 */
    uint32_t flags;
    int32_t B;             /* number of strings */
    int32_t group_size;    /* strings per group (reference batch); groups are consecutive; the last may be short */
    int32_t Nt;            /* time samples in this call; steps n = 2 .. Nt-1 are computed (simulator.cpp:40) */
    int32_t Nx_t1, Nx_l1;  /* padded state widths = state_u.size(-1), state_z.size(-1) (string.cpp:123-124) */
    int32_t n_0;           /* global index of local sample 0 (simulator.cpp:26; only used by MANUFACTURED) */
    int32_t max_iter;      /* cap for the reference's uncapped loops (string.cpp:200, hammer.cpp:33); <=0 -> 100 */
    /* `constant` and `relative_error` arrive as float32 in the reference (simulator.cpp:22-23) */
    float k, theta_t, lambda_c, relative_order;

    /* states (space stride 1).  SAVE_STATE: (B,Nt,Nx) in/out.  else (B,2,Nx) in/out. */
    sfdtd_array state_u, state_z;
    /* string_params (simulator.cpp:17; string.cpp:68-70): kappa(B) alpha(B) p_a(B) f0(B,Nt) pos(B) T60(B,2,2 contiguous, bs = string stride) */
    sfdtd_array kappa, alpha, p_a, f0, pos, T60;
    /* bow_params (string.cpp:73-74): x_b v_b F_b wid (B,Nt); phi_0 phi_1 (B) */
    sfdtd_array x_b, v_b, F_b, wid, phi_0, phi_1;
    /* hammer_params (string.cpp:77-78): x_H w_H M_r alpha_H (B); u_H (B,Nt) updated IN PLACE: u_H[:,n] += u_H (string.cpp:303) */
    sfdtd_array x_H, w_H, M_r, alpha_H, u_H;
    /* excitation masks, uint8 (B) */
    const uint8_t *bow_mask, *hammer_mask;
    /* float32 table torch.linspace(1/Nx_t1, 1, Nx_t1) built by the host exactly as the reference does (misc.cpp:26-27) */
    const float *xax;

    /* outputs (B,Nt), columns 0 and 1 are left untouched (the caller discards them, src/task/simulate.py:82-86).
       uout and zout are required; v_r, F_H, u_H_out may have ptr == NULL (audio-only: not written). */
    sfdtd_array uout, zout, v_r, F_H, u_H_out;   /* u_H_out = u_H / k (simulator.cpp:57) */
    /* outputs (B): loss parameters of the last step (string.cpp:119-120) */
    void *sig0, *sig1;
    uint32_t *status;      /* (B) status bits, OR-ed into the array (zero it before the call), may be NULL */
    /* optional (may be NULL): per-string int64[4] counters {outer iterations, linear sweeps, hammer iterations, steps},
       added to the array (zero it before the call) */
    int64_t *counters;
    /* optional (may be NULL): control curves synthesised inside the stepper from per-string scalars instead of being read
       from (B,Nt) arrays.  When set, f0 / x_b / v_b / F_b / wid above are ignored (their ptr may be NULL) and u_H.ptr may
       be NULL (the hammer displacement is then carried inside the library and not written back). */
    const struct sfdtd_synth *synth;
} sfdtd_args;

/* Compact description of the (B,Nt) control curves the reference's samplers materialise on the host
 * (reference src/model/simulator.py:210-235 f0 glissando + vibrato, :419-484 bow ramps with shapers, :573-578 hammer
 * displacement; src/utils/control.py:5-45; src/utils/misc.py:74-82).  All pointers are DEVICE arrays of B doubles.
 * With t = 1..Nt_full the (1-based) global sample index, ramp = (t-1)/(Nt_full-1):
 *   f0(t)  = g(t) + v(t) g(t),  g = f0_a + (f0_b - f0_a) ramp,
 *            v = (t > vib_t0) ? mod_amp (1 - cos(2 pi mod_frq (t - vib_t0) k)) / 2 : 0
 *   x_b(t) = x_b1 + (x_b2 - x_b1) ramp
 *   v_b(t) = (v_b1 + (v_b2 - v_b1) ramp) tanh(10 t / sr)
 *   F_b(t) = (F_b1 + (F_b2 - F_b1) ramp) * (pulloff > 0 ? tanh(100 max(Nt_full - (t-1) - off, 0) / sr) : 1),
 *            off = Nt_full - floor(sr pulloff)
 *   wid(t) = wid
 *   u_H(t) = -1e-3 (t = 1), -1e-3 + k v_H (t = 2), 0 afterwards        (pre-loaded content of hammer_params[2])
 * sfdtd_synth_controls() writes exactly the values the stepper uses, for callers (and tests) that need the curves. */
typedef struct sfdtd_synth {
    int32_t Nt_full;       /* samples of the full-length curves */
    int32_t t_0;           /* global 0-based index of local sample 0 of this call */
    double sr;             /* sample rate */
    const double *f0_a, *f0_b, *mod_frq, *mod_amp, *vib_t0;
    const double *x_b1, *x_b2, *v_b1, *v_b2, *F_b1, *F_b2, *pulloff, *wid;
    const double *v_H;
} sfdtd_synth;

/* A plan holds everything sfdtd_forward derives from the parameters before it can launch: per-string grid maxima,
 * the kernel bucket of every string / group, launch geometry and the library-owned scratch.  Creating one costs a small
 * prepass kernel and ONE device->host read (it synchronises `cuda_stream`); running it is fully asynchronous. */
typedef struct sfdtd_plan sfdtd_plan;

/* Runs steps 2..Nt-1 for all B strings on `cuda_stream` (a cudaStream_t, NULL = default stream).
 * = sfdtd_plan_create + sfdtd_forward_plan + sfdtd_plan_destroy: the host blocks once, for the small device->host read
 * that sizes the launch (before any stepper kernel is queued); it does NOT wait for the stepper -- synchronise the
 * stream before reading results.  All device pointers must belong to the current device of the calling thread. */
int sfdtd_forward(const sfdtd_args *args, void *cuda_stream);

/* Plan for `args` (reads f0 / the synth scalars, kappa, alpha, the masks and state row n-1).  Synchronises `cuda_stream`
 * once.  The plan stays valid for any later args with the same B, group_size, Nt, Nx_t1, Nx_l1, flags and masks whose
 * per-string grid sizes do not exceed the ones seen here (otherwise SFDTD_ST_RANGE is raised per string). */
int sfdtd_plan_create(const sfdtd_args *args, void *cuda_stream, sfdtd_plan **plan);
/* Queues the whole call on `cuda_stream` and returns immediately (no host synchronisation).  A plan may be in flight on
 * one stream at a time; calls with the same plan on the same stream serialise correctly. */
int sfdtd_forward_plan(sfdtd_plan *plan, const sfdtd_args *args, void *cuda_stream);
/* Releases the plan; its device scratch is freed in stream order on `cuda_stream` (no host synchronisation). */
int sfdtd_plan_destroy(sfdtd_plan *plan, void *cuda_stream);

/* Writes the (B,Nt) control curves of args->synth exactly as the stepper evaluates them.  Any output ptr may be NULL. */
int sfdtd_synth_controls(const sfdtd_args *args, const sfdtd_array *f0, const sfdtd_array *x_b, const sfdtd_array *v_b,
                         const sfdtd_array *F_b, const sfdtd_array *u_H, void *cuda_stream);

/* Device-side post-processing of the audio, downstream of the stepper in the reference's run()
 * (reference src/task/simulate.py:333-335 NaN mask, :336-337 / src/utils/audio.py:72-76 silence test on 20 log10 RMS,
 * src/utils/audio.py:42-48 l-infinity gain, src/task/simulate.py:416-425 wav subtypes): per string b over the samples
 * n0 <= n < n0 + n_samples of uout / zout (B,Nt) fp64
 *   is_nan[b]    = any NaN in uout[b];  is_silent[b] = 20 log10 rms(uout[b]) <= silence_db
 *   gain[b]      = normalize ? 1 / max|uout[b]| (1 when that is 0 or NaN) : 1
 *   pcm_u/pcm_z/pcm_w[b, :] = PCM samples of gain*uout, gain*zout, gain*(uout+zout): bits = 16 -> int16, 24 -> 3 packed
 *   little-endian bytes per sample (round to nearest even, clipped, NaN -> 0); string b starts at byte b * pcm_pitch
 *   (pcm_pitch = 0: tightly packed rows of n_samples * bits/8 bytes; a multiple of 4 enables word stores).
 *   is_nan / is_silent / gain and the pcm pointers may each be NULL. */
int sfdtd_postprocess(const sfdtd_array *uout, const sfdtd_array *zout, int32_t B, int32_t n0, int32_t n_samples,
                      double silence_db, int32_t normalize, int32_t bits, int64_t pcm_pitch, uint8_t *is_nan,
                      uint8_t *is_silent, double *gain, void *pcm_u, void *pcm_z, void *pcm_w, void *cuda_stream);

/* The same for float32 uout / zout (outputs of an SFDTD_F32 call); gain stays double. */
int sfdtd_postprocess_f32(const sfdtd_array *uout, const sfdtd_array *zout, int32_t B, int32_t n0, int32_t n_samples,
                          double silence_db, int32_t normalize, int32_t bits, int64_t pcm_pitch, uint8_t *is_nan,
                          uint8_t *is_silent, double *gain, void *pcm_u, void *pcm_z, void *pcm_w, void *cuda_stream);

/* Human-readable description of the last error on this thread. */
const char *sfdtd_last_error(void);

/* FMA-pipe peak of the current device in TFLOP/s (2 flops per FMA), measured with a register-resident
 * FMA-chain kernel: which = 0 -> fp64, 1 -> fp32.  Roofline denominator for bench.py. */
int sfdtd_measure_fma_peak(int which, double *tflops);

/* ABI / build info */
int sfdtd_abi_version(void);
/* number of kernel launches issued by sfdtd_forward calls since load (for bench accounting) */
int64_t sfdtd_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SFDTD_H_ */
