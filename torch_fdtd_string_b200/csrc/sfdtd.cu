// sfdtd.cu -- B200 (sm_100a) time-loop-fused StringFDTD stepper behind the C ABI of include/sfdtd.h.
//
// Layout of the computation (DESIGN.md has the derivations; reference citations are to
// /root/reference/src/model/cpp/*.cpp):
//
//   * Every string is owned by L lanes of one warp for the whole call.  Transverse block: blocked
//     layout, ET consecutive grid rows per lane; the state rows u^{n-1}, u^{n-2} live in shared memory
//     (bank-conflict-free padded rows with guard cells, so stencil halos are plain loads), the per-step
//     working set (coefficients, tridiagonal factors, right-hand side, iterate) in registers.
//   * Only the rows that can differ from zero are solved: R = min(W_t, N_t+3) transverse rows.  The
//     reference solves all W_t = batch-max rows (misc.cpp:119-127); rows >= N_t+3 form a homogeneous
//     constant-coefficient tail whose exact effect is a continued-fraction correction of the pivot of
//     row R-1 (computed once per step).  Strings are therefore sized by their OWN grid, and a launch is
//     bucketed by (L, ET) so that small strings use few lanes.
//   * The implicit system  [A11 K_tl; K_lt A22] w = -RHS  (string.cpp:162-181,238) is solved matrix-free:
//     block Gauss-Seidel over the transverse/longitudinal blocks; A11 (tridiagonal, depends on
//     Lambda(u^{n-1})) by a register-resident partitioned Thomas factorisation (local LU of the ET-1
//     interior rows per lane + parallel cyclic reduction over the L interface rows via shuffles);
//     A22 (constant coefficients, off/diag ~ 1e-5) by Jacobi sweeps folded into the same iteration.
//     The sweeps stop on the predicted error (measured contraction rate), not on the last change.
//   * Longitudinal block: shared memory; the linear interpolation operators Int_tl / Int_lt
//     (misc.cpp:78-105) are gathers whose indices/weights are cached and rebuilt only when a grid
//     size changes (a few times per second of audio).
//   * Per-step scalars (grid sizes, loss, operator coefficients; string.cpp:16-41,96-120) are computed
//     TB steps at a time (16 in independent mode, 8 in grouped mode), one step per lane, into a shared-memory table;
//     outputs are staged there too and flushed as coalesced rows.
//   * Groups (reference batches) couple their strings only through the batch-max operator widths and
//     the any-over-batch convergence votes (string.cpp:252-253, hammer.cpp:51).  The widths come from a
//     prepass table.  A group without bowed/hammered strings needs no votes (the second fixed-point
//     pass of an unforced string reproduces the first), so its strings run fully independently
//     ("independent mode", any warp, no CTA barrier); a group with forced strings runs as one thread-block
//     cluster of 128-thread CTAs whose votes travel as 32-bit OR words through distributed shared memory
//     ("grouped mode").
//   * The stepper is templated on its arithmetic type: double (SFDTD_F64, the parity mode) and float (SFDTD_F32, the
//     reference's `precision: single`): state rows, working set, solves and outputs are of that type, the per-step
//     scalar table / bow window / hammer loop are evaluated in double in both builds, and an fp32 call takes its grid
//     sizes from the reference's float32 evaluation of get_derived_vars.
//   * Scheduling (independent mode): the grid of a bucket is what is resident at once, and every warp pulls its next
//     (time slice, set of 32/L strings) item from a per-bucket counter -- the hardest sets first, each for the whole call,
//     the sets of the last round in time slices whose state rows pass from warp to warp through global memory with a
//     release/acquire word per set.  Loop conditions inside the time loop are kept provably warp-uniform (kernel
//     parameters, __shfl_sync broadcasts, vote results): otherwise every shuffle is compiled with a reconvergence path.
// No tensor cores: no step is a dense contraction.  HBM traffic: controls in, audio out.

#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <map>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <chrono>

#include "sfdtd.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define FULLMASK 0xffffffffu

namespace cg = cooperative_groups;

namespace {

constexpr int WL_MARGIN = 2;        // ghost rows of the longitudinal block kept beyond N_l (decay (e/d)^m, e/d ~ 1e-5)
constexpr int NV_I = 22;            // doubles per time step in the per-string scalar table, independent mode
constexpr int NV_G = 26;            // grouped mode: + tolerances, bow force
constexpr int NI = 8;               // ints per time step
constexpr int NOUT = 5;             // staged outputs per step
constexpr int NCONST = 8;           // per-string constants kept in shared memory
constexpr int GS_CAP = 60;          // cap on block Gauss-Seidel sweeps per solve
constexpr int NLA_I = 5;            // longitudinal arrays per string, independent mode: Z1 Z2 ZA ZB RL
constexpr int NLA_G = 6;            // grouped mode: + ZP (previous fixed-point iterate)
// time steps per scalar-table block.  Independent mode: 16 = one step per lane of a 16-lane string, so that the table code
// (divisions, square roots) runs with every lane busy and the control prefetch / output flush move 128-byte rows (+6...8 %
// over 8); grouped mode: 8, its CTAs are at the shared-memory limit.
constexpr int TBS_I = 16, TBS_G = 8;
constexpr int TBS_MAX = TBS_I > TBS_G ? TBS_I : TBS_G;
#ifndef SFDTD_F32_TOL
#define SFDTD_F32_TOL 1e-7f     // fp32 build: predicted relative error that ends the sweeps
#endif
#ifndef SFDTD_PREDICT_SWEEPS
#define SFDTD_PREDICT_SWEEPS 1
#endif
// fewest sweeps of a step's first solve: with / without a longitudinal right-hand side of its own (keep_l > 0)
#ifndef SFDTD_MIN_SW_L
#define SFDTD_MIN_SW_L 4
#endif
#ifndef SFDTD_MIN_SW
#define SFDTD_MIN_SW 3
#endif
#ifndef SFDTD_DEFAULT_WLMIN
#define SFDTD_DEFAULT_WLMIN 16      // smallest longitudinal allocation class (rows incl. guards)
#endif
#ifndef SFDTD_DEFAULT_LANE_DIV
#define SFDTD_DEFAULT_LANE_DIV 4
#endif
#ifndef SFDTD_DEFAULT_QUEUE
#define SFDTD_DEFAULT_QUEUE 1
#endif

// table slots (doubles)
enum { T_IHT = 0, T_OFFA, T_DIAGA, T_CORR, T_OFFC, T_DIAGC, T_DIAGB, T_OFF1B, T_KH4, T_PH2, T_PHL, T_IDA, T_EIDA,
       T_RDW, T_CTR, T_VB, T_WID, T_UHPRE, T_IHL, T_S0K, T_S1K, T_GA2,      // <- independent mode uses the slots up to here
       T_TOLT, T_TOLL, T_FB, T_SPARE };
static_assert(T_GA2 < NV_I && T_SPARE < NV_G, "table too small");
// table slots (ints)
enum { I_NT = 0, I_NL, I_R, I_WLS, I_RK, I_KEEPL, I_IDXH, I_IC };
// per-string constants
enum { C_WPOW = 0, C_MR, C_AHM1, C_PHI0, C_PHI1, C_RP };

// grouped mode: what one CTA of a group's cluster runs.  A group (reference batch) is spread over the CTAs of one
// thread-block cluster; `cls` selects the lane/row shape of all string slots of this CTA.
constexpr int CTA_SLOTS = 8;
constexpr int GB_MAX = 64;          // strings per group in grouped mode (a cluster holds <= 8 CTAs x 8 slots)
struct CtaDesc {
    int32_t group;                  // group id (row of Wtab)
    int32_t cls;                    // 0: 16 lanes x 4 rows per string, 1: 32 lanes x 4 rows
    int32_t n;                      // string slots in use
    int32_t G;                      // strings of the group
    int32_t str[CTA_SLOTS];         // global string id per slot
    int32_t gidx[CTA_SLOTS];        // index of the string inside its group (its entry of the group board)
};

struct KArgs {
    sfdtd_args a;
    sfdtd_synth sy;                 // copy of *a.synth (valid when has_synth)
    int32_t has_synth;
    int32_t f32;                    // SFDTD_F32 call: grid sizes from the reference's float32 evaluation of get_derived_vars
    double k, ik, k2, k4, th, omth, tt1, tt2, lamc, order, mhd;
    const int32_t *Wtab;            // [n_groups][Nt]  W_t | W_l << 16  (batch-max operator widths, misc.cpp:119-127)
    const int32_t *ids;             // independent mode: string ids
    const CtaDesc *ctas;            // grouped mode: one descriptor per CTA
    double *uH_carry;               // (B,2) hammer displacement rows [n-2, n-1] handed over between time slices when a.u_H.ptr is NULL
    int32_t n_items;                // entries of ids
    const int32_t *maxNl;           // per string: largest N_l of this call (prepass); sizes the longitudinal block in grouped mode
    int32_t WLp;                    // independent mode: longitudinal rows allocated per string (incl. guards)
    int32_t need_xax;               // copy the bow axis to shared memory
    int32_t max_iter;
    int32_t n_lo, n_hi;             // time slice of this launch: steps n_lo <= n < n_hi (2 <= n_lo)
    int32_t *queue;                 // independent mode, optional: work counter; every warp pulls its next (time slice, set of 32/L strings) from it
    int32_t *done;                  // queue mode: per set, the number of time slices completed (release/acquire hand-over of the state rows)
    int32_t q_slice, q_nslices;     // queue mode: steps per time slice, number of slices (of the sets beyond q_full)
    int32_t q_full;                 // queue mode: the first q_full sets run the whole call as one item
};

__device__ __forceinline__ double ldx(const sfdtd_array &A, int b, int n) {
    return ((const double *)A.ptr)[(int64_t)b * A.bs + (int64_t)n * A.ts];
}
__device__ __forceinline__ double lds(const sfdtd_array &A, int b) {
    return ((const double *)A.ptr)[(int64_t)b * A.bs];
}

// ---- control curves: read from the caller's (B,Nt) arrays, or synthesised from per-string scalars (sfdtd_synth) -------
// Every kernel evaluates the same functions with explicitly rounded operations (no FMA contraction), so the prepass,
// the width table, the stepper and sfdtd_synth_controls see bit-identical values (floor(1/h) of f0 decides grid sizes).
__device__ __forceinline__ double sy_ramp(const sfdtd_synth &y, int g) {      // g: global 0-based sample index
    const int den = y.Nt_full > 1 ? y.Nt_full - 1 : 1;
    return __ddiv_rn((double)g, (double)den);
}
__device__ __forceinline__ double sy_lin(double a, double b, double ramp) { return __dadd_rn(a, __dmul_rn(__dsub_rn(b, a), ramp)); }
__device__ __forceinline__ double sy_f0(const sfdtd_synth &y, int b, int g) {
    const double f = sy_lin(y.f0_a[b], y.f0_b[b], sy_ramp(y, g));
    const double dt = __dsub_rn((double)(g + 1), y.vib_t0[b]);
    double vib = 0.0;
    if (dt > 0) {
        const double ph = __dmul_rn(__dmul_rn(__dmul_rn(2 * M_PI, y.mod_frq[b]), dt), __ddiv_rn(1.0, y.sr));
        vib = __ddiv_rn(__dmul_rn(y.mod_amp[b], __dsub_rn(1.0, cos(ph))), 2.0);
    }
    return __dadd_rn(f, __dmul_rn(vib, f));
}
__device__ __forceinline__ double sy_xb(const sfdtd_synth &y, int b, int g) { return sy_lin(y.x_b1[b], y.x_b2[b], sy_ramp(y, g)); }
__device__ __forceinline__ double sy_vb(const sfdtd_synth &y, int b, int g) {
    return __dmul_rn(sy_lin(y.v_b1[b], y.v_b2[b], sy_ramp(y, g)), tanh(__dmul_rn(__ddiv_rn((double)(g + 1), y.sr), 10.0)));
}
__device__ __forceinline__ double sy_Fb(const sfdtd_synth &y, int b, int g) {
    double F = sy_lin(y.F_b1[b], y.F_b2[b], sy_ramp(y, g));
    const double po = y.pulloff[b];
    if (po > 0) {
        const double off = __dsub_rn((double)y.Nt_full, floor(__dmul_rn(y.sr, po)));
        const double rem = fmax(__dsub_rn(__dsub_rn((double)y.Nt_full, (double)g), off), 0.0);
        F = __dmul_rn(F, tanh(__dmul_rn(__ddiv_rn(rem, y.sr), 100.0)));
    }
    return F;
}
__device__ __forceinline__ double sy_uH(const sfdtd_synth &y, int b, int g) {
    if (g == 0) return -1e-3;
    if (g == 1) return __dadd_rn(-1e-3, __dmul_rn(__ddiv_rn(1.0, y.sr), y.v_H[b]));
    return 0.0;
}
__device__ __forceinline__ double ctl_f0(const KArgs &A, int b, int n) { return A.has_synth ? sy_f0(A.sy, b, n + A.sy.t_0) : ldx(A.a.f0, b, n); }
__device__ __forceinline__ double ctl_xb(const KArgs &A, int b, int n) { return A.has_synth ? sy_xb(A.sy, b, n + A.sy.t_0) : ldx(A.a.x_b, b, n); }
__device__ __forceinline__ double ctl_vb(const KArgs &A, int b, int n) { return A.has_synth ? sy_vb(A.sy, b, n + A.sy.t_0) : ldx(A.a.v_b, b, n); }
__device__ __forceinline__ double ctl_Fb(const KArgs &A, int b, int n) { return A.has_synth ? sy_Fb(A.sy, b, n + A.sy.t_0) : ldx(A.a.F_b, b, n); }
__device__ __forceinline__ double ctl_wid(const KArgs &A, int b, int n) { return A.has_synth ? A.sy.wid[b] : ldx(A.a.wid, b, n); }
// pre-loaded content of hammer_params[2] at sample n (string.cpp:303 adds the new displacement onto it)
template <typename T> __device__ __forceinline__ double ctl_uH(const KArgs &A, int b, int n) {
    if (A.a.u_H.ptr) return (double)((const T *)A.a.u_H.ptr)[(int64_t)b * A.a.u_H.bs + (int64_t)n * A.a.u_H.ts];
    return A.has_synth ? sy_uH(A.sy, b, n + A.sy.t_0) : 0.0;
}

// 1/x to full double precision without the slow-path branch of the IEEE division (MUFU.RCP64H + 2 Newton steps)
__device__ __forceinline__ double frcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// ---- get_derived_vars (string.cpp:16-41), reference operation order, no FMA contraction ----------
// floor(1/h) decides the grid sizes: one ulp flips them, so this part is evaluated exactly like the reference.
struct Derived { double gamma, K, Nt, ht, Nl, hl; };
// the same in float32, as the reference's `precision: single` run evaluates it (tensors float32, C++ scalars rounded to
// float32 by ATen): its floor(1/h) can differ from the double one, and the grid size decides the whole waveform
__device__ __forceinline__ Derived derive_f32(float f0, float kappa_rel, float alpha, const KArgs &A) {
    Derived d;
    const float pi = (float)M_PI, k = (float)A.k, k2 = (float)A.k2, k4 = (float)A.k4;
    const float gamma = __fmul_rn(2.0f, f0);
    const float kappa = __fmul_rn(gamma, kappa_rel);
    const float t0 = __fdiv_rn(__fmul_rn(pi, kappa), gamma);
    const float IHP = __fmul_rn(t0, t0);
    const float K = __fmul_rn(__fsqrt_rn(IHP), __fdiv_rn(gamma, pi));
    const float g2 = __fmul_rn(gamma, gamma), g4 = __fmul_rn(g2, g2), K2 = __fmul_rn(K, K);
    const float in = __fadd_rn(__fmul_rn(g4, k4), __fmul_rn(__fmul_rn(__fmul_rn(16.0f, K2), k2), (float)A.tt1));
    const float num = __fadd_rn(__fmul_rn(g2, k2), __fsqrt_rn(in));
    const float h1 = __fmul_rn((float)A.lamc, __fsqrt_rn(__fdiv_rn(num, (float)A.tt2)));
    const float Nt = floorf(__frcp_rn(h1));
    const float h2 = __fmul_rn(__fmul_rn(__fmul_rn((float)A.lamc, gamma), alpha), k);
    const float Nl = floorf(__frcp_rn(h2));
    d.Nt = (double)Nt; d.ht = (double)__frcp_rn(Nt); d.Nl = (double)Nl; d.hl = (double)__frcp_rn(Nl);
    d.gamma = (double)gamma; d.K = (double)K;
    return d;
}
__device__ __forceinline__ Derived derive(double f0, double kappa_rel, double alpha, const KArgs &A) {
    if (A.f32) return derive_f32((float)f0, (float)kappa_rel, (float)alpha, A);
    Derived d;
    const double gamma = __dmul_rn(2.0, f0);
    const double kappa = __dmul_rn(gamma, kappa_rel);
    const double t0 = __ddiv_rn(__dmul_rn(M_PI, kappa), gamma);
    const double IHP = __dmul_rn(t0, t0);
    const double K = __dmul_rn(__dsqrt_rn(IHP), __ddiv_rn(gamma, M_PI));
    const double g2 = __dmul_rn(gamma, gamma);
    const double g4 = __dmul_rn(g2, g2);
    const double K2 = __dmul_rn(K, K);
    const double in = __dadd_rn(__dmul_rn(g4, A.k4), __dmul_rn(__dmul_rn(__dmul_rn(16.0, K2), A.k2), A.tt1));
    const double num = __dadd_rn(__dmul_rn(g2, A.k2), __dsqrt_rn(in));
    const double h1 = __dmul_rn(A.lamc, __dsqrt_rn(__ddiv_rn(num, A.tt2)));
    d.Nt = floor(__ddiv_rn(1.0, h1));
    d.ht = __ddiv_rn(1.0, d.Nt);
    const double h2 = __dmul_rn(__dmul_rn(__dmul_rn(A.lamc, gamma), alpha), A.k);
    d.Nl = floor(__ddiv_rn(1.0, h2));
    d.hl = __ddiv_rn(1.0, d.Nl);
    d.gamma = gamma; d.K = K;
    return d;
}
__device__ __forceinline__ int clampN(double v) { return (int)fmin(fmax(v, 0.0), 60000.0); }   // NaN -> 0

// ---- prepass 1: per string, the largest N_t / N_l any step of this call can see (at min f0) ----------
// Also a per-string estimate of the nonlinearity  phi/h^2 Lambda^2  of the first step (from state row n-1): it predicts
// how many block sweeps the string needs, and the host sorts the strings of a launch by it so that the strings
// sharing a warp converge in about the same number of sweeps.
// The results also go to mapped pinned host memory (hNt ... hHam, written by the kernel itself over PCIe): the host reads them
// after one stream synchronisation without a device->host copy, which would queue behind whatever bulk read-back the caller
// has in flight on the copy engine.
template <typename T>
__global__ void sfdtd_prepass_kernel(const __grid_constant__ KArgs A, int32_t *maxNt, int32_t *maxNl, float *est,
                                     int32_t *hNt, int32_t *hNl, float *hEst, uint8_t *hBow, uint8_t *hHam) {
    const int b = blockIdx.x;
    const int Nt = A.a.Nt;
    double fm = INFINITY;
    double dm = 0.0;
    {
        const T *su1 = (const T *)A.a.state_u.ptr + (int64_t)b * A.a.state_u.bs + A.a.state_u.ts;
        for (int i = 1 + threadIdx.x; i < A.a.Nx_t1; i += blockDim.x) dm = fmax(dm, fabs((double)su1[i] - (double)su1[i - 1]));
        for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(FULLMASK, dm, o));
    }
    __shared__ double sd[32];
    if ((threadIdx.x & 31) == 0) sd[threadIdx.x >> 5] = dm;
    if (!A.has_synth && A.a.f0.ts == 0) {
        fm = ldx(A.a.f0, b, 0);
    } else {
        for (int n = 2 + threadIdx.x; n < Nt; n += blockDim.x) { const double v = ctl_f0(A, b, n); fm = v < fm ? v : fm; }
    }
    for (int o = 16; o > 0; o >>= 1) { const double v = __shfl_xor_sync(FULLMASK, fm, o); fm = v < fm ? v : fm; }
    __shared__ double sm[32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = fm;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (blockDim.x >> 5); w++) fm = fm < sm[w] ? fm : sm[w];
        for (int w = 0; w < (blockDim.x >> 5); w++) dm = fmax(dm, sd[w]);
        const double alpha = lds(A.a.alpha, b);
        Derived d = derive(fm, lds(A.a.kappa, b), alpha, A);
        maxNt[b] = clampN(d.Nt); maxNl[b] = clampN(d.Nl);
        const double phi = ((d.gamma * d.gamma) * A.k2) * (alpha * alpha - 1) / 4;
        const double n2 = d.Nt * d.Nt;
        est[b] = (float)(phi * n2 * n2 * dm * dm);
        hNt[b] = maxNt[b]; hNl[b] = maxNl[b]; hEst[b] = est[b];
        hBow[b] = A.a.bow_mask[b]; hHam[b] = A.a.hammer_mask[b];
    }
}

// work-queue words of a call are cleared by a kernel, not by cudaMemsetAsync: a memset may be executed by a copy engine and then
// queues behind whatever bulk transfer the caller has in flight there (measured: +190 ms per call beside a 6 GB read-back)
__global__ void sfdtd_zero_kernel(int32_t *p, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

// ---- prepass 2: batch-max operator widths per (group, step) (misc.cpp:119-127) ------------------------
__global__ void sfdtd_width_kernel(const __grid_constant__ KArgs A, int32_t *Wtab) {
    const int g = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int Nt = A.a.Nt;
    if (n >= Nt) return;
    const int g0 = g * A.a.group_size;
    const int G = min(A.a.group_size, A.a.B - g0);
    int wt = 0, wl = 0;
    for (int s = 0; s < G; s++) {
        const int b = g0 + s;
        const Derived d = derive(ctl_f0(A, b, n), lds(A.a.kappa, b), lds(A.a.alpha, b), A);
        wt = max(wt, clampN(d.Nt)); wl = max(wl, clampN(d.Nl));
    }
    Wtab[(int64_t)g * Nt + n] = (wt + 1) | ((wl + 1) << 16);
}

// ---- warp helpers over the L lanes of one string -------------------------------------------------
template <int L, typename T> __device__ __forceinline__ T shup(T v, int d) { return __shfl_up_sync(FULLMASK, v, d, L); }
template <int L, typename T> __device__ __forceinline__ T shdn(T v, int d) { return __shfl_down_sync(FULLMASK, v, d, L); }
template <int L, typename T> __device__ __forceinline__ T shix(T v, int s) { return __shfl_sync(FULLMASK, v, s, L); }
template <int L, typename T> __device__ __forceinline__ T red_sum(T v) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o, L);
    return v;
}
template <int L> __device__ __forceinline__ int red_or(int v) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v |= __shfl_xor_sync(FULLMASK, v, o, L);
    return v;
}
// any-over-CTA vote.  The barrier reduction returns the same value to every thread, but the compiler does not treat it as
// warp-uniform: a loop that exits on it would count as divergent and every shuffle inside would be compiled with a
// reconvergence sequence.  Passing it through a warp vote makes the uniformity visible.
// ---- hammer contact loop of one string (hammer.cpp:28-53) ---------------------------------------------------------
struct HamIn { double eta1, eta2, wr, base, hm, tol, k2, mhd; };   // wr = w_H^(1+alpha) relu(eta1)^(alpha-1); base = 2 u_H1 - u_H2
// one pass: eta -> contact force, hammer displacement, next eta
__device__ __forceinline__ double ham_pass(const HamIn &h, double eta, double eps_u, double &fo, double &uo) {
    const double fH = (h.wr * (eta + h.eta2)) / 2;
    fo = (h.eta1 > 0) ? fH : 0.0;
    double tt = (h.base - h.k2 * fo) - h.mhd;
    tt = tt > 0 ? tt : (tt != tt ? tt : 0.0);
    uo = tt + h.mhd;
    return (uo - eps_u) * h.hm;
}
// exit tests |eta - eta_estimate| > tol (hammer.cpp:49-51) of the next HAM_HB passes from `eta`, as a bit mask
constexpr int HAM_HB = 6;
__device__ __forceinline__ unsigned ham_mask(const HamIn &h, double eta, double eps_u) {
    unsigned m = 0; double f_, u_;
#pragma unroll
    for (int p = 0; p < HAM_HB; p++) {
        const double en = ham_pass(h, eta, eps_u, f_, u_);
        m |= (fabs(eta - en) > h.tol ? 1u : 0u) << p;
        eta = en;
    }
    return m;
}

// ---- group votes (grouped mode) --------------------------------------------------------------------------------------
// A group (reference batch) runs as one thread-block cluster of 128-thread CTAs.  Its any-over-batch decisions
// (string.cpp:252-253, hammer.cpp:51) are bitwise-OR votes of one 32-bit word per warp: every warp leader writes its word
// into the vote array of every CTA of the cluster through distributed shared memory, one cluster barrier makes them
// visible, every warp ORs the (<= 32) words itself.  The array sits at the start of every CTA's shared memory and is
// double-buffered by phase (a warp can only be one vote ahead of the slowest warp of its group).
constexpr int GB_DOUBLES = 32;                      // 2 phases x 32 words

struct GroupComm {
    unsigned *vw;           // this CTA's vote array
    int cs, rank, phase;    // cluster size, rank of this CTA, vote phase (CTA-uniform)
    __device__ __forceinline__ void init(double *smem_base) {
        vw = (unsigned *)smem_base; phase = 0;
        cg::cluster_group cl = cg::this_cluster();
        cs = (int)cl.num_blocks(); rank = (int)cl.block_rank();
    }
    __device__ __forceinline__ void sync() const {
        if (cs > 1) cg::this_cluster().sync(); else __syncthreads();
    }
    // bitwise OR of `word` over all threads of the group
    __device__ __forceinline__ unsigned vote(unsigned word) {
        const unsigned w = __reduce_or_sync(FULLMASK, word);
        unsigned *slot = vw + 32 * phase;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        int nw;
        if (cs > 1) {
            if (lane < cs) cg::this_cluster().map_shared_rank(slot, lane)[rank * 4 + warp] = w;
            cg::this_cluster().sync();
            nw = cs * 4;
        } else {
            if (lane == 0) slot[warp] = w;
            __syncthreads();
            nw = blockDim.x >> 5;
        }
        const unsigned r = __reduce_or_sync(FULLMASK, lane < nw ? slot[lane] : 0u);
        phase ^= 1;
        return r;
    }
};
template <int L> constexpr int ilog2() { return L <= 1 ? 0 : 1 + ilog2<L / 2>(); }

// |x| as the high word of the double: monotone in |x| for integer compares; NaN and inf sort above every finite value
__device__ __forceinline__ unsigned hi_abs(double v) { return (unsigned)__double2hiint(v) & 0x7fffffffu; }
__device__ __forceinline__ unsigned hi_abs(float v) { return __float_as_uint(v) & 0x7fffffffu; }          // fp32 build: the whole word
template <typename T> __device__ __forceinline__ float hi_to_float(unsigned h) {
    if (sizeof(T) == 8) return (float)__hiloint2double((int)h, 0);                                        // > FLT_MAX -> inf, NaN -> NaN
    return __uint_as_float(h);
}
template <int L> __device__ __forceinline__ unsigned red_maxu(unsigned v) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULLMASK, v, o, L));
    return v;
}
__device__ __forceinline__ double nan0(double v) {   // nan_to_num (string.cpp:225-226)
    if (v != v) return 0.0;
    if (isinf(v)) return v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    return v;
}
__device__ __forceinline__ float nan0(float v) {
    if (v != v) return 0.0f;
    if (isinf(v)) return v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    return v;
}
// arithmetic type of a stepper build: double (SFDTD_F64, the parity mode) or float (SFDTD_F32, the reference's `precision:
// single`).  State rows, the per-step working set, the linear solves and the audio outputs are of that type; the per-step
// scalar table, the bow window, the hammer contact loop and the control curves are evaluated in double in both builds.
template <typename T> struct Real;
template <> struct Real<double> {
    static constexpr float GS_TOL = 1e-13f;          // predicted relative max-norm error of the transverse block that ends the sweeps
    static constexpr float E_FLOOR = 0.f;            // a relative change at or below this is round-off: converged
    static constexpr float RATE_MIN = 0.f;           // the contraction rate is only measured from changes above this
};
template <> struct Real<float> {
    // float32 round-off of one solve is ~2e-7 of the solution's max-norm: the sweeps run down to it (the reference's direct
    // float32 solve is that accurate, and per-step errors accumulate linearly over the run)
    static constexpr float GS_TOL = SFDTD_F32_TOL;
    static constexpr float E_FLOOR = 4e-7f;
    static constexpr float RATE_MIN = 2e-5f;
};
__device__ __forceinline__ float frcp(float x) {     // MUFU.RCP + one Newton step
    const float r = __fdividef(1.0f, x);
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}
template <typename T> struct Vec2;
template <> struct Vec2<double> { typedef double2 type; };
template <> struct Vec2<float> { typedef float2 type; };
__device__ __forceinline__ double t_cospi(double x) { return cospi(x); }
__device__ __forceinline__ float t_cospi(float x) { return cospif(x); }

// ---- partitioned Thomas: local LU of the ET-1 interior rows + PCR over the L interface rows -------
template <typename T, int L, int ET> struct TriSolver {
    static constexpr int M = ET - 1;
    static constexpr int LV = ilog2<L>();
    T inv[M], lw[M], cp[M], V[M], W[M];
    T ae, ce, k1[LV], k2[LV], invB;

    __device__ __forceinline__ void factor(const T (&a)[ET], const T (&b)[ET], const T (&c)[ET], int ln) {
        T cprev = T(0);
#pragma unroll
        for (int r = 0; r < M; r++) {
            const T den = fma(-a[r], cprev, b[r]);
            inv[r] = frcp(den);
            cp[r] = c[r] * inv[r];
            lw[r] = a[r] * inv[r];
            cprev = cp[r];
        }
        // left spike  T_I V = a_0 e_0 ; right spike  T_I W = c_{M-1} e_{M-1}
        V[0] = lw[0];
#pragma unroll
        for (int r = 1; r < M; r++) V[r] = -lw[r] * V[r - 1];
        W[M - 1] = cp[M - 1];
#pragma unroll
        for (int r = M - 2; r >= 0; r--) { V[r] = fma(-cp[r], V[r + 1], V[r]); W[r] = -cp[r] * W[r + 1]; }
        ae = a[ET - 1]; ce = c[ET - 1];
        const T Vn0 = shdn<L>(V[0], 1), Wn0 = shdn<L>(W[0], 1);
        T Ar = -ae * V[M - 1];
        T Br = b[ET - 1] - ae * W[M - 1] - ce * Vn0;
        T Cr = -ce * Wn0;
        if (ln == 0) Ar = T(0);
        if (ln == L - 1) { Cr = T(0); Br = b[ET - 1] - ae * W[M - 1]; }
#pragma unroll
        for (int lv = 0; lv < LV; lv++) {
            const int s = 1 << lv;
            const T iB = frcp(Br);
            const T iBm = shup<L>(iB, s), iBp = shdn<L>(iB, s);
            const T Am = shup<L>(Ar, s), Cm = shup<L>(Cr, s);
            const T Ap = shdn<L>(Ar, s), Cp = shdn<L>(Cr, s);
            const bool hm = ln >= s, hp = ln + s < L;
            const T q1 = hm ? Ar * iBm : T(0), q2 = hp ? Cr * iBp : T(0);
            k1[lv] = q1; k2[lv] = q2;
            Br = Br - (hm ? Cm * q1 : T(0)) - (hp ? Ap * q2 : T(0));
            Ar = hm ? -Am * q1 : T(0);
            Cr = hp ? -Cp * q2 : T(0);
        }
        invB = frcp(Br);
    }

    // d: right-hand side in, solution out
    __device__ __forceinline__ void solve(T (&d)[ET], int ln) const {
        T Y[M];
        Y[0] = d[0] * inv[0];
#pragma unroll
        for (int r = 1; r < M; r++) Y[r] = fma(-lw[r], Y[r - 1], d[r] * inv[r]);
#pragma unroll
        for (int r = M - 2; r >= 0; r--) Y[r] = fma(-cp[r], Y[r + 1], Y[r]);
        T Yn0 = shdn<L>(Y[0], 1);
        // (no boundary select: ce = 0 in the last lane, whose last row has no right neighbour; the shuffle returns its own Y[0])
        T D = d[ET - 1] - ae * Y[M - 1] - ce * Yn0;
#pragma unroll
        for (int lv = 0; lv < LV; lv++) {
            const int s = 1 << lv;
            const T Dm = shup<L>(D, s), Dp = shdn<L>(D, s);
            D = D - k1[lv] * Dm - k2[lv] * Dp;     // k1/k2 are 0 where the neighbour does not exist
        }
        const T xe = D * invB;
        T p = shup<L>(xe, 1);
        // (no boundary select: the left spike V is 0 in lane 0, whose first row has no left neighbour)
#pragma unroll
        for (int r = 0; r < M; r++) d[r] = Y[r] - V[r] * p - W[r] * xe;
        d[ET - 1] = xe;
    }
};

// manufactured-solution forcing (vnv.cpp:11-37) times k^2, split into the per-step part (MsStep) and the per-row part
struct MsStep { double c1, c2, c3, ect, est; };
__device__ __forceinline__ MsStep ms_step(double gamma, double sig0, double K, double p_a, double t, double k2) {
    MsStep m;
    const double sigma = sig0, omega = gamma, mu_sq = M_PI * M_PI;
    m.c1 = sigma * sigma - omega * omega - (2 * sig0) * sigma;
    m.c2 = (2 * mu_sq) * ((4 * (K * K)) * mu_sq + gamma * gamma);
    m.c3 = (2 * omega) * (sigma - sig0);
    const double e = (p_a * exp(-sigma * t)) * k2;
    m.ect = e * cos(omega * t); m.est = e * sin(omega * t);
    return m;
}
// x: position of padded row i on the reference's domain_x axis (misc.cpp:45-52): clamp(i * 2/N_t, 0, 2), then (v - 1) / 2
__device__ __forceinline__ double ms_row(const MsStep &m, int i, double two_ht) {
    const double xv = (fmin(fmax((double)i * two_ht, 0.0), 2.0) - 1.0) * 0.5;
    const double cx = cos(M_PI * xv), c2x = cos((2 * M_PI) * xv);
    return (m.c1 * (cx * cx) + m.c2 * c2x) * m.ect + (m.c3 * (cx * cx)) * m.est;
}

// float32 linear-interpolation row (misc.cpp:78-105; F.interpolate(..., 'linear', align_corners=True) on float32)
__device__ __forceinline__ void interp_row(float s, int o, int in_last, int &i0, int &i1, float &w0, float &w1) {
    const float r = __fmul_rn(s, (float)o);
    int a0 = (int)r;
    a0 = a0 > in_last ? in_last : a0;
    i0 = a0; i1 = a0 + (a0 < in_last ? 1 : 0);
    w1 = __fsub_rn(r, (float)a0);
    w0 = __fsub_rn(1.0f, w1);
}

// gather through a packed pair of BYTE offsets (lo 16 bits | hi 16 bits): one add per address instead of unpack + scale + add
template <typename T> __device__ __forceinline__ T ldb(const T *base, int byte_off) { return *(const T *)((const char *)base + byte_off); }

// select element `slot` of a register array without dynamic indexing
template <int ET, typename T> __device__ __forceinline__ T pick(const T (&v)[ET], int slot) {
    T o = T(0);
#pragma unroll
    for (int r = 0; r < ET; r++) o = (r == slot) ? v[r] : o;
    return o;
}
template <int L, int ET, typename T> __device__ __forceinline__ T fetch_row(const T (&v)[ET], int idx) {
    idx = idx < 0 ? 0 : (idx > L * ET - 1 ? L * ET - 1 : idx);
    const T mine = pick<ET>(v, idx % ET);
    return shix<L>(mine, idx / ET);
}



// shared-memory doubles of the fixed part of a string slot, and of its longitudinal part (W rows incl. guards, W even)
// Two 8-lane strings share a 16-lane shared-memory wavefront: their slots are spaced by 8 (mod 16) doubles, which puts the
// blocked row accesses (lane stride ET+1 doubles, ET = 4 or 6) and the consecutive longitudinal accesses of the two strings
// into disjoint banks.
// (fp32 build: the same holds for two 16-lane strings of 4-byte rows sharing a 32-lane wavefront)
__host__ __device__ inline int slot_spacing(int n, int L, int tsz = 8) {
    n = (n + 1) & ~1;                 // 16-byte aligned slots (int4 / double2 loads)
    if (L <= 8 || (tsz == 4 && L <= 16)) n += (24 - (n & 15)) & 15;      // == 8 (mod 16)
    else if ((n & 15) == 0) n += 2;   // one string per half-warp: the slots only should not start in the same bank
    return n;
}
// (tsz: sizeof of the build's arithmetic type; the row arrays and the longitudinal arrays are of that type)
__host__ __device__ inline int slot_fixed_doubles(int L, int ET, bool grouped, int tsz) {
    const int LE = L * ET;
    const int TBS = grouped ? TBS_G : TBS_I;
    const int rows = (LE + L + 2) + 2 * (LE + L + 6) + (grouped ? LE + L : 0);
    const int n = TBS * (grouped ? NV_G : NV_I) + TBS * NI / 2 + TBS * NOUT + NCONST + (rows * tsz + 7) / 8;
    return slot_spacing(n, L, tsz);
}
__host__ __device__ inline int slot_long_doubles(int W, bool grouped, int L, int tsz) {
    const int n = ((((grouped ? NLA_G : NLA_I) * W + 2 * W) * tsz + 7) / 8 + (W + 1) / 2 + 1) & ~1;
    return (L <= 8 || (tsz == 4 && L <= 16)) ? slot_spacing(n, L, tsz) : n;
}
__host__ __device__ inline int long_rows(int maxNl) { return (maxNl + 1 + WL_MARGIN + 2 + 1) & ~1; }   // + 2 guards, even

// loss parameters of the last step (string.cpp:119-120) -> sig0, sig1; once per string and call, not inlined
__device__ __noinline__ void final_sigmas(const KArgs &A, int b, int Nt) {
    const sfdtd_args &a = A.a;
    const Derived d = derive(ctl_f0(A, b, Nt - 1), lds(a.kappa, b), lds(a.alpha, b), A);
    const double *T60 = (const double *)a.T60.ptr + (int64_t)b * a.T60.bs;
    const double T00 = T60[0], T01 = T60[1], T10 = T60[2], T11 = T60[3];
    const double g2 = d.gamma * d.gamma, g4 = g2 * g2;
    double z1, z2;
    if (d.K > 0) {
        const double w1 = (2 * M_PI) * T00, w2 = (2 * M_PI) * T10;
        z1 = -g2 + sqrt(g4 + (4 * (d.K * d.K)) * (w1 * w1));
        z2 = -g2 + sqrt(g4 + (4 * (d.K * d.K)) * (w2 * w2));
    } else { z1 = (T00 * T00) / g2; z2 = (T10 * T10) / g2; }
    const bool m = (T00 * T01 * T10 * T11) != 0;
    const double s0 = m ? (-z2 / T01 + z1 / T11) : 0.0, s1 = m ? (1 / T01 - 1 / T11) : 0.0;
    const double c6 = 13.815510557964274;
    ((double *)a.sig0)[b] = (c6 * s0) / (z1 - z2); ((double *)a.sig1)[b] = (c6 * s1) / (z1 - z2);
}

// ---- per-step scalars of one string (grid sizes, loss, operator coefficients, bow window, readout weight) ------------
// One call per string and time step, TB steps at a time, one step per lane.  NOT inlined: the divisions, square roots,
// powers and (with sfdtd_synth) the control-curve evaluation are ~15 % of a stepper kernel's code but run once per step
// and lane; one shared copy keeps the kernels' instruction footprint down (three CTAs of different phase share an SM).
struct TabIn {
    int b, NXT, LE, WLa;
    bool bowm, hamm;
    const int32_t *Wrow;
};
template <typename T, bool GROUPED>
__device__ __forceinline__ void fill_table_row_impl(const KArgs &A, const TabIn &in, int n, double *t, int *ti) {
    const sfdtd_args &a = A.a;
    const int b = in.b;
    const double kappa_rel = lds(a.kappa, b), alpha = lds(a.alpha, b);
    const double *T60 = (const double *)a.T60.ptr + (int64_t)b * a.T60.bs;
    const double T00 = T60[0], T01 = T60[1], T10 = T60[2], T11 = T60[3];
    const double alpha2 = alpha * alpha;
    const double xH = lds(a.x_H, b);
    const double exc = 1.0 + (in.hamm ? 1.0 : 0.0) + (in.bowm ? 1.0 : 0.0);
    const double f0 = ctl_f0(A, b, n);
    const Derived d = derive(f0, kappa_rel, alpha, A);
    const int N_t = clampN(d.Nt), N_l = clampN(d.Nl);
    const int32_t w = in.Wrow[n];
    const int Wt = w & 0xffff, Wl = (w >> 16) & 0xffff;
    // loss parameters (string.cpp:100-120)
    const double g2 = d.gamma * d.gamma, g4 = g2 * g2;
    double z1, z2;
    if (d.K > 0) {
        const double w1 = (2 * M_PI) * T00, w2 = (2 * M_PI) * T10;
        z1 = -g2 + sqrt(g4 + (4 * (d.K * d.K)) * (w1 * w1));
        z2 = -g2 + sqrt(g4 + (4 * (d.K * d.K)) * (w2 * w2));
    } else { z1 = (T00 * T00) / g2; z2 = (T10 * T10) / g2; }
    const bool m = (T00 * T01 * T10 * T11) != 0;
    const double s0 = m ? (-z2 / T01 + z1 / T11) : 0.0, s1 = m ? (1 / T01 - 1 / T11) : 0.0;
    const double c6 = 13.815510557964274;   // 6*log(10)
    const double sig0 = (c6 * s0) / (z1 - z2), sig1 = (c6 * s1) / (z1 - z2);
    const double g = g2 * A.k2;
    const double s0k = (2 * sig0) * A.k, s1k = (2 * sig1) * A.k;
    const double phi = (g * (alpha2 - 1)) / 4;
    const double Kk = (d.K * d.K) * A.k2;
    const double iht = d.Nt, ihl = d.Nl;          // 1/h_t = N_t exactly
    const double iht2 = iht * iht, iht4 = iht2 * iht2, ihl2 = ihl * ihl;
    // operator coefficients (string.cpp:138-181; misc.cpp:119-166)
    const double diagA = A.th + s0k + 2 * s1k * iht2, offA = 0.5 * A.omth - s1k * iht2;
    const double kh4 = Kk * iht4;
    const double dA = (1 + s0k) + 2 * s1k * ihl2, eA = -s1k * ihl2, idA = 1.0 / dA;
    // bow window on the Nx_t1-point axis (bow.cpp:32, misc.cpp:20-34)
    const double Nd = (double)in.NXT;
    const double xb = ctl_xb(A, b, n), wd = ctl_wid(A, b, n);
    const double ctr = __ddiv_rn(__dmul_rn(xb, (double)(N_t - 1)), Nd);
    const double wid = __ddiv_rn(__dmul_rn(__dmul_rn(wd, d.ht), (double)(N_t - 1)), Nd);
    const int ic = (int)fmin(fmax(floor((ctr - wid * 0.5) * Nd) - 2, 0.0), 60000.0);
    // rows that are solved: R (see header); forced rows of a bowed string extend it
    int R = N_t + 3;
    if (in.bowm) { const int Rb = (int)fmin(fmax(ceil((ctr + wid * 0.5) * Nd) + 1, 0.0), 60000.0); R = max(R, Rb); }
    if (A.a.flags & SFDTD_MANUFACTURED) R = Wt;          // every padded row is forced (string.cpp:227-232)
    R = min(R, Wt);
    int oob = 0;
    if (R > in.LE) { R = in.LE; oob = 1; }
    // homogeneous tail R..W_t-1 of A11 folded into the pivot of row R-1
    double corr = 0.0;
    {
        const int mt = Wt - R;
        if (mt > 0 && !oob) {
            const double o2 = offA * offA;
            double pv = diagA;
            for (int j = 1; j < mt; j++) { const double pn = diagA - o2 * frcp(pv); if (pn == pv) break; pv = pn; }
            corr = o2 * frcp(pv);
        }
    }
    int WLs = min(N_l + 1 + WL_MARGIN, Wl);
    if (WLs > in.WLa) { WLs = in.WLa; oob = 1; }
    const int keep_flat = N_t + N_l + 2;                         // string.cpp:233
    t[T_IHT] = iht; t[T_IHL] = ihl;
    t[T_OFFA] = offA; t[T_DIAGA] = diagA; t[T_CORR] = corr;
    t[T_OFFC] = 0.5 * A.omth + s1k * iht2; t[T_DIAGC] = A.th - s0k - 2 * s1k * iht2;
    t[T_DIAGB] = -2 * A.th + 2 * g * iht2 + 6 * kh4; t[T_OFF1B] = -A.omth - g * iht2 - 4 * kh4; t[T_KH4] = kh4;
    t[T_PH2] = phi * iht2; t[T_PHL] = (phi != 0.0) ? ihl * d.ht : 0.0;
    t[T_IDA] = idA; t[T_EIDA] = eA * idA;
    t[T_RDW] = ((0.5 * d.ht) * exc) * A.ik;                      // surface-integral weight / k (string.cpp:274-291)
    if (GROUPED) { t[T_TOLT] = pow(d.ht, A.order); t[T_TOLL] = pow(d.hl, A.order); t[T_FB] = ctl_Fb(A, b, n); }
    t[T_CTR] = ctr; t[T_WID] = wid;
    t[T_VB] = ctl_vb(A, b, n);
    t[T_UHPRE] = ctl_uH<T>(A, b, n);
    t[T_S0K] = s0k; t[T_S1K] = s1k; t[T_GA2] = g * alpha2;
    ti[I_NT] = N_t; ti[I_NL] = N_l; ti[I_R] = R | (oob << 30); ti[I_WLS] = WLs;
    ti[I_RK] = min(R, keep_flat); ti[I_KEEPL] = keep_flat - in.NXT;
    ti[I_IDXH] = (int)fmin(fmax(floor(__dmul_rn(xH, (double)(N_t - 1))), 0.0), (double)(in.LE - 1));
    ti[I_IC] = ic;
}
template <typename T, bool GROUPED>
__device__ __noinline__ void fill_table_row(const KArgs &A, const TabIn &in, int n, double *t, int *ti) {
    fill_table_row_impl<T, GROUPED>(A, in, n, t, ti);
}
#ifndef SFDTD_TAB_INLINE_I
#define SFDTD_TAB_INLINE_I 1        // 1: the independent-mode kernels inline the table code (measured: the call costs them 4 %)
#endif

// ======================================================================================================
// PF: build for the reference layout (SFDTD_SAVE_STATE calls: one reference batch per call, latency-bound) -- row n of the state
// histories, read-modified-written at the end of step n, is pulled into L2 at the start of the step (+14 % on the literal
// drop-in batch).  A separate instantiation: the same lines inside the compact kernels cost them 2-3 % (register allocation).
template <typename T, int L, int ET, bool GROUPED, bool MANUF, bool PF = false>
__device__ __forceinline__ void step_body(const KArgs &A, const CtaDesc *cd) {
    constexpr int TSZ = (int)sizeof(T);
    constexpr float GS_TOL = Real<T>::GS_TOL;
    constexpr int TB = GROUPED ? TBS_G : TBS_I;
    constexpr int LE = L * ET;
    constexpr int NLA = GROUPED ? NLA_G : NLA_I;
    constexpr int NV = GROUPED ? NV_G : NV_I;
    // fixed slot layout (offsets in doubles)
    constexpr int O_TABI = TB * NV, O_OST = O_TABI + TB * NI / 2, O_CST = O_OST + TB * NOUT, O_QS = O_CST + NCONST;
    // row arrays (qs, UA, UB, RC) are indexed through PR(): one pad double per ET rows, so that the blocked accesses
    // "lane l touches rows l*ET + c" fall into distinct banks (stride ET+1 doubles instead of ET)
    // (offsets of the row arrays in elements of T from the start of the row region SR = (T *)(S + O_QS))
    constexpr int O_UA = (LE + L + 2) + 4, O_UB = O_UA + (LE + L + 6), O_RC = O_UB + (LE + L + 2);
    auto PR = [](int i) { return i + (i + ET) / ET - 1; };            // i >= -ET
    extern __shared__ __align__(16) double smem[];
    const sfdtd_args &a = A.a;
    const int tid = threadIdx.x;
    const int sl = tid / L, ln = tid % L;
    const int nslots = blockDim.x / L;
    int b; bool valid;
    int wrow = 0;                                // grouped mode: row of the width table
    GroupComm gc;
    if (GROUPED) {
        valid = sl < cd->n;
        const int s_ = valid ? sl : cd->n - 1;   // spare slots shadow the last string and never write, publish or vote
        b = cd->str[s_]; wrow = cd->group;
    } else {
        const int item = blockIdx.x * nslots + sl;
        valid = item < A.n_items;
        b = A.ids[valid ? item : A.n_items - 1];
    }
    const bool queued = !GROUPED && A.queue != nullptr;
    const int Nt = a.Nt, NXT = a.Nx_t1, NXL = a.Nx_l1;
    const bool surf = a.flags & SFDTD_SURFACE_INTEGRAL;
    const bool save_state = a.flags & SFDTD_SAVE_STATE;
    const bool skip_aux = a.flags & SFDTD_SKIP_AUX;
    // the manufactured-solution mode (B = 1 verification runs) has its own kernel: its per-row cosines would triple the code
    // of every other one
    constexpr bool manuf = MANUF;
    uint32_t status = 0;

    // ---- shared memory carve-up: [bow axis][fixed slot parts][longitudinal parts] ----
    if (GROUPED) gc.init(smem);
    constexpr int O_BOARD = GROUPED ? GB_DOUBLES : 0;
    float *xaxs = (float *)(smem + O_BOARD);                          // [NXT] (only when the bow axis is needed)
    const int xoff = O_BOARD + (A.need_xax ? (((NXT + 3) / 4) * 2) : 0);
    double *const S = smem + xoff + (size_t)sl * slot_fixed_doubles(L, ET, GROUPED, TSZ);
    double *Lbd = smem + xoff + (size_t)nslots * slot_fixed_doubles(L, ET, GROUPED, TSZ);
    int WLp;
    if (GROUPED) {
        // per-string longitudinal allocation (sized by the string's own largest N_l), offsets by a prefix sum
        int *ioffs = (int *)Lbd;                                      // [nslots + 2]
        WLp = long_rows(A.maxNl[b]);
        if (ln == 0) ioffs[sl + 1] = slot_long_doubles(WLp, true, L, TSZ);
        __syncthreads();
        if (tid == 0) { ioffs[0] = ((nslots + 2) / 2 + 1) & ~1; for (int q = 0; q < nslots; q++) ioffs[q + 1] += ioffs[q]; }
        __syncthreads();
        const int off = ioffs[sl];
        __syncthreads();
        Lbd += off;
    } else {
        WLp = A.WLp;
        Lbd += (size_t)sl * slot_long_doubles(WLp, false, L, TSZ);
    }
    T *const Lb = (T *)Lbd;
    const int WLa = WLp - 2;                                          // usable longitudinal rows (guards at -1 and WLa)
    double *const tab = S;                                            // [TB][NV]
    int *const tabi = (int *)(S + O_TABI);                            // [TB][NI]
    double *const ost = S + O_OST;                                    // [TB][NOUT]
    double *const cst = S + O_CST;                                    // [NCONST]
    T *const SR = (T *)(S + O_QS);                                    // row region: qs | UA | UB | RC
    T *const qs = SR;                                                 // [LE + L + 2], PR()-indexed
    T *const RC = SR + O_RC;                                          // [LE + L] bow weights (grouped mode only), PR()-indexed
    T *const LW = Lb + NLA * WLp;                                     // [WLp][2] Int_lt weights
    int *const LI = (int *)(LW + 2 * WLp);                            // [WLp]    Int_lt indices i0 | i1 << 16
    // transverse state rows n-1 / n-2 (in S, guards: 2 each side) and longitudinal arrays (in Lb, guard at [-1]); offsets swap
    int u1o = O_UA, u2o = O_UB;
    int z1o = 1, z2o = WLp + 1;
    const int zao = 2 * WLp + 1, zbo = 3 * WLp + 1, rlo = 4 * WLp + 1, zpo = 5 * WLp + 1;
    if (A.need_xax) {
        for (int i = tid; i < NXT; i += blockDim.x) xaxs[i] = a.xax[i];
        __syncthreads();
    }
    if (GROUPED) {
        // A warp without a string leaves (barriers only wait for threads that have not exited); its vote words must read
        // as 0, so every CTA clears its vote array before anybody in the cluster may write into it.
        for (int i = tid; i < 64; i += blockDim.x) gc.vw[i] = 0u;
        gc.sync();
        if (!__any_sync(FULLMASK, valid)) return;
    }

    // Work queue (independent mode): the grid is sized to what is resident at once and every warp pulls one set of 32/L strings
    // after the other (hardest first), so that the load balances at warp granularity instead of in waves of whole CTAs.
    // time slice of the current work item; kept in shared memory (n_lo, n_hi, set, slice index), not in registers
    volatile int *const qst = (volatile int *)(cst + 6);
    if (ln == 0) { qst[0] = A.n_lo; qst[1] = A.n_hi; qst[2] = 0; qst[3] = 0; }
    __syncwarp();
    int n_lo = A.n_lo, n_hi = A.n_hi;                                 // warp-uniform (kernel parameters or derived from the popped item)
    for (bool q_first = true;; q_first = false) {
    if (queued) {
        // Items: first the A.q_full hardest sets, each for the whole call (longest first); then the remaining sets in time slices,
        // slice-major (every set of that tail group advances through slice s before any starts slice s + 1, so the predecessor
        // of a popped item was popped a whole group earlier and is normally long finished).  The slices keep the idle tail at
        // the end of the bucket at a fraction of a slice instead of a fraction of a whole string.
        int k = 0;
        if ((tid & 31) == 0) k = atomicAdd(A.queue, 1);
        k = __shfl_sync(FULLMASK, k, 0);
        const int n_sets = (A.n_items + 32 / L - 1) / (32 / L), n_tail = n_sets - A.q_full;
        int set = k, sidx = 0, lo = 2, hi = Nt;
        if (k >= A.q_full) {
            const int j = k - A.q_full;
            if (n_tail <= 0 || j >= n_tail * A.q_nslices) break;
            sidx = j / n_tail; set = A.q_full + (j - sidx * n_tail);
            lo = 2 + sidx * A.q_slice; hi = min(Nt, lo + A.q_slice);
        }
        const int item = set * (32 / L) + (tid & 31) / L;
        valid = item < A.n_items;
        b = A.ids[valid ? item : A.n_items - 1];
        status = 0; u1o = O_UA; u2o = O_UB; z1o = 1; z2o = WLp + 1;
        __syncwarp();
        if (ln == 0) { qst[0] = lo; qst[1] = hi; qst[2] = set; qst[3] = sidx; }
        n_lo = lo; n_hi = hi;
        if (sidx > 0) {
            // every lane polls the same word and the exit is a warp vote: the loop is convergent by construction
            for (;;) {
                int v;
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(A.done + set) : "memory");
                if (!__any_sync(FULLMASK, v < sidx)) break;
                __nanosleep(200);
            }
        }
        __syncwarp();
    } else if (!q_first) break;

    // ---- per-string constants ----
    const bool bowm = a.bow_mask[b] != 0, hamm = a.hammer_mask[b] != 0;
    const bool forced = bowm || hamm;
    bool group_has_hammer = false, group_has_bow = false;
    if (GROUPED) {
        const unsigned gv = gc.vote((valid && hamm ? 1u : 0u) | (valid && bowm ? 2u : 0u));
        group_has_hammer = gv & 1u; group_has_bow = gv & 2u;
    }
    // CTA-uniform compute switches (they guard shuffles and barriers); per-string output switches
    const bool do_bow = group_has_bow || !skip_aux, do_ham = group_has_hammer || !skip_aux;
    const bool out_bow = bowm || !skip_aux, out_ham = hamm || !skip_aux;
    if (ln == 0) {
        const double aH = lds(a.alpha_H, b);
        const double wH = lds(a.w_H, b) / A.lamc;
        cst[C_WPOW] = pow(wH, 1.0 + aH);
        cst[C_MR] = lds(a.M_r, b) / A.lamc;
        cst[C_AHM1] = aH - 1.0;
        cst[C_PHI0] = lds(a.phi_0, b); cst[C_PHI1] = lds(a.phi_1, b);
        cst[C_RP] = lds(a.pos, b);
    }

    // ---- initial state: rows n-2, n-1 ----
    {
        for (int j = ln; j < 2 + 2 * (LE + L + 6) + (GROUPED ? LE + L : 0); j += L) SR[LE + L + j] = T(0);      // UA, UB (with guards), RC
        for (int j = ln; j < NLA * WLp; j += L) Lb[j] = T(0);
        for (int j = ln; j < WLp; j += L) { LW[2 * j] = T(0); LW[2 * j + 1] = T(0); LI[j] = 0; }
        __syncwarp();
        // rows n_lo-2, n_lo-1: from the (B,Nt,Nx) history with SAVE_STATE, else from the compact (B,2,Nx) carry buffer
        const int64_t row0 = save_state ? (int64_t)(qst[0] - 2) : 0;
        const T *su = (const T *)a.state_u.ptr + (int64_t)b * a.state_u.bs + row0 * a.state_u.ts;
#pragma unroll
        for (int r = 0; r < ET; r++) {
            const int i = ln * ET + r;
            SR[u2o + PR(i)] = (i < NXT) ? su[i] : T(0);
            SR[u1o + PR(i)] = (i < NXT) ? su[a.state_u.ts + i] : T(0);
        }
        const T *sz = (const T *)a.state_z.ptr + (int64_t)b * a.state_z.bs + row0 * a.state_z.ts;
        for (int j = ln; j < WLa; j += L) {
            Lb[z2o + j] = (j < NXL) ? sz[j] : T(0);
            Lb[z1o + j] = (j < NXL) ? sz[a.state_z.ts + j] : T(0);
        }
    }
    double uH1 = 0.0, uH2 = 0.0;
    if (Nt > 2) {
        const int n_lo = qst[0];
        if (a.u_H.ptr || n_lo == 2) { uH2 = ctl_uH<T>(A, b, n_lo - 2); uH1 = ctl_uH<T>(A, b, n_lo - 1); }
        else { uH2 = A.uH_carry[2 * (int64_t)b]; uH1 = A.uH_carry[2 * (int64_t)b + 1]; }     // handed over by the previous time slice
    }
    uint32_t cnt_outer = 0, cnt_sweeps = 0, cnt_ham = 0, cnt_steps = 0;
    // cached interpolation rows of Int_tl for this lane's transverse rows (rebuilt when a grid size changes):
    // indices are stored +1 so that 0 addresses the zero guard (rows beyond N_t)
    int tix[ET]; float twb[ET];
#pragma unroll
    for (int r = 0; r < ET; r++) { tix[r] = 0; twb[r] = 0.f; }
    int curNt = -1, curNl = -1, ext1 = WLa, ext2 = WLa;               // ext: rows of Z1 / Z2 that may be non-zero
    float rho_h = 0.5f;                                               // contraction-rate history of the block iteration
    int s_prev = 0;                                                   // sweeps the first solve of the previous step took (warp-uniform)
    bool have_next = false; unsigned hgm_next = 0;                    // grouped mode: contact-loop exit bits voted ahead for the next step
    __syncwarp();

    for (int n0 = n_lo; n0 < n_hi; n0 += TB) {
        // ================= scalar table for steps n0 .. n0+TB-1 (one step per lane) =================
        {
            TabIn ti_;
            ti_.b = b; ti_.NXT = NXT; ti_.LE = LE; ti_.WLa = WLa; ti_.bowm = bowm; ti_.hamm = hamm;
            ti_.Wrow = A.Wtab + (int64_t)(GROUPED ? wrow : b / a.group_size) * Nt;
            for (int s = ln; s < TB; s += L) {
                const int n = n0 + s;
                if (n >= n_hi) break;
                if (!GROUPED && SFDTD_TAB_INLINE_I) fill_table_row_impl<T, GROUPED>(A, ti_, n, tab + s * NV, tabi + s * NI);
                else fill_table_row<T, GROUPED>(A, ti_, n, tab + s * NV, tabi + s * NI);
            }
        }
        __syncwarp();

        const int jmax = min(TB, n_hi - n0);
        for (int jj = 0; jj < jmax; jj++) {
            const int n = n0 + jj;
            const double *t = tab + jj * NV;
            const int4 tiA = *(const int4 *)(tabi + jj * NI);
            const int N_t = tiA.x, N_l = tiA.y, R = tiA.z & 0xffffff, WLs = tiA.w;
            if (tiA.z >> 30) status |= SFDTD_ST_RANGE;
            if (PF && save_state) {
                const T *pu = (const T *)a.state_u.ptr + (int64_t)b * a.state_u.bs + (int64_t)n * a.state_u.ts + ln * ET;
                if (ln * ET < NXT) asm volatile("prefetch.global.L2 [%0];" :: "l"(pu));
                const T *pz = (const T *)a.state_z.ptr + (int64_t)b * a.state_z.bs + (int64_t)n * a.state_z.ts;
                for (int j = ln * 4; j < NXL; j += L * 4) asm volatile("prefetch.global.L2 [%0];" :: "l"(pz + j));
            }

            // ---- interpolation rows, rebuilt only when a grid size changed (misc.cpp:78-105) ----
            const bool grid_changed = (N_t != curNt) || (N_l != curNl);
            if (__any_sync(FULLMASK, grid_changed)) {
              if (grid_changed) {
                const float s_tl = (N_t > 0) ? __fdiv_rn((float)N_l, (float)N_t) : 0.0f;    // t-row -> l-grid
                const float s_lt = (N_l > 0) ? __fdiv_rn((float)N_t, (float)N_l) : 0.0f;    // l-row -> t-grid
#pragma unroll
                for (int r = 0; r < ET; r++) {
                    const int i = ln * ET + r;
                    int i0, i1; float w0, w1;
                    interp_row(s_tl, i, N_l, i0, i1, w0, w1);
                    i0 = min(i0, WLa - 1) + 1; i1 = min(i1, WLa - 1) + 1;
                    if (i > N_t) { w1 = 0.f; i0 = 0; i1 = 0; }
                    tix[r] = (i0 * TSZ) | ((i1 * TSZ) << 16); twb[r] = w1;      // byte offsets into a longitudinal row
                }
                for (int j = ln; j < WLp; j += L) {
                    int i0 = 0, i1 = 0; float w0 = 0.f, w1 = 0.f;
                    if (j <= N_l) {
                        interp_row(s_lt, j, N_t, i0, i1, w0, w1);
                        i0 = PR(min(i0, LE - 1)); i1 = PR(min(i1, LE - 1));      // qs is PR()-indexed
                    }
                    LW[2 * j] = (T)w0; LW[2 * j + 1] = (T)w1; LI[j] = (i0 * TSZ) | ((i1 * TSZ) << 16);     // byte offsets into qs
                }
                curNt = N_t; curNl = N_l;
              }
              __syncwarp();
            }

            // ---- per-row coefficients, base right-hand side, factorisation ----
            TriSolver<T, L, ET> ts;
            T mu[ET + 1], rt[ET];
            const int i0row = ln * ET;
            const int lnp = ln * (ET + 1);
            auto PRL = [&](int c) { return lnp + c + (c + ET) / ET - 1; };    // PR(i0row + c) with the division folded at compile time
            {
                T sq[ET + 1];
                {
                    // masked previous states (mask_1d, string.cpp:129-132): x * 0 keeps NaN like the reference's multiply
                    T e1[ET + 4], e2[ET + 2];
#pragma unroll
                    for (int r = 0; r < ET + 4; r++) { const int i = i0row + r - 2; e1[r] = SR[u1o + PRL(r - 2)] * ((i <= N_t) ? T(1) : T(0)); }
#pragma unroll
                    for (int r = 0; r < ET + 2; r++) { const int i = i0row + r - 1; e2[r] = SR[u2o + PRL(r - 1)] * ((i <= N_t) ? T(1) : T(0)); }
                    // Lambda = Dxb u1 (string.cpp:152) on the solved rows; mu = phi/h^2 Lambda; sq = phi/h^2 Lambda^2
                    const T iht = (T)t[T_IHT], ph2 = (T)t[T_PH2];
#pragma unroll
                    for (int r = 0; r <= ET; r++) {
                        const int i = i0row + r;
                        const T lam = (i < R) ? (e1[r + 2] - e1[r + 1]) * iht : T(0);
                        mu[r] = ph2 * lam; sq[r] = mu[r] * lam;
                    }
                    // B11 u1 + C11 u2  (string.cpp:223-224)
                    const T diagB = (T)t[T_DIAGB], off1B = (T)t[T_OFF1B], kh4 = (T)t[T_KH4], diagC = (T)t[T_DIAGC], offC = (T)t[T_OFFC];
#pragma unroll
                    for (int r = 0; r < ET; r++) {
                        const int i = i0row + r;
                        T d4 = diagB;
                        if (i == 1 || i == N_t - 1) d4 += kh4;          // Dxxxx_clamped (misc.cpp:146-163)
                        const T Bu = d4 * e1[r + 2] + off1B * (e1[r + 1] + e1[r + 3]) + kh4 * (e1[r] + e1[r + 4]);
                        const T Cu = diagC * e2[r + 1] + offC * (e2[r] + e2[r + 2]) + sq[r] * (e2[r + 1] - e2[r]) - sq[r + 1] * (e2[r + 2] - e2[r + 1]);
                        rt[r] = Bu + Cu;
                    }
                }
                // K_tl (2 z1 + z2)  (B12 = 2 K_tl, C12 = K_tl); the first coupling guess z = 2 z1 - z2 goes to ZA
                {
                    T yy[ET];
#pragma unroll
                    for (int r = 0; r < ET; r++) {
                        const int j0 = tix[r] & 0xffff, j1 = tix[r] >> 16;
                        const float w1f = twb[r];
                        const T w1 = (T)w1f, w0 = (T)__fsub_rn(1.0f, w1f);
                        yy[r] = w0 * (T(2) * ldb(Lb + z1o - 1, j0) + ldb(Lb + z2o - 1, j0)) + w1 * (T(2) * ldb(Lb + z1o - 1, j1) + ldb(Lb + z2o - 1, j1));
                    }
                    for (int j = ln; j < WLs; j += L) Lb[zao + j] = (j <= N_l) ? T(2) * Lb[z1o + j] - Lb[z2o + j] : T(0);
                    T yyl = shup<L>(yy[ET - 1], 1);
                    if (ln == 0) yyl = T(0);
                    T nub[ET + 1];
#pragma unroll
                    for (int r = 0; r < ET; r++) nub[r] = mu[r] * (yy[r] - (r == 0 ? yyl : yy[r - 1]));
                    nub[ET] = shdn<L>(nub[0], 1);
                    if (ln == L - 1) nub[ET] = T(0);
                    const int Rk = tabi[jj * NI + I_RK];
#pragma unroll
                    for (int r = 0; r < ET; r++) rt[r] = (i0row + r < Rk) ? (rt[r] + (nub[r] - nub[r + 1])) : T(0);
                }
                // manufactured solution (string.cpp:227-232): RHS -= f k^2 on every padded row, before the flat mask
                MsStep ms;
                if (manuf) {
                    const float tf = (float)(n + a.n_0) * a.k;                    // float32 product (string.cpp:229)
                    const double gamma = 2.0 * ctl_f0(A, b, n);
                    ms = ms_step(gamma, t[T_S0K] / (2.0 * A.k), gamma * lds(a.kappa, b), lds(a.p_a, b), (double)tf, A.k2);
                    const int Rk = tabi[jj * NI + I_RK];
                    const double two_ht = 2.0 / t[T_IHT];
#pragma unroll
                    for (int r = 0; r < ET; r++) if (i0row + r < Rk) rt[r] -= (T)ms_row(ms, i0row + r, two_ht);
                }
                // A11 (string.cpp:153-162), rows < R, tail folded into row R-1
                {
                    const T offA = (T)t[T_OFFA], diagA = (T)t[T_DIAGA], corr = (T)t[T_CORR];
                    T ca[ET], cb[ET], cc[ET];
#pragma unroll
                    for (int r = 0; r < ET; r++) {
                        const int i = i0row + r;
                        const bool in = i < R;
                        ca[r] = (in && i > 0) ? offA - sq[r] : T(0);
                        cc[r] = (i + 1 < R) ? offA - sq[r + 1] : T(0);
                        T bb = diagA + sq[r] + sq[r + 1];
                        if (i == R - 1) bb -= corr;
                        cb[r] = in ? bb : T(1);
                    }
                    ts.factor(ca, cb, cc, ln);
                }
                // l-block base RHS, only when the flat-index mask leaves any of it (string.cpp:233)
                const int keep_l = tabi[jj * NI + I_KEEPL];
                if (__any_sync(FULLMASK, keep_l > 0)) {
                    // q2 = mu Dxb u2 -> smem, then K_lt u2 by gathers
#pragma unroll
                    for (int r = 0; r < ET; r++) {
                        const int i = i0row + r;
                        const T a2 = SR[u2o + PRL(r)] * ((i <= N_t) ? T(1) : T(0)), b2 = SR[u2o + PRL(r - 1)] * ((i - 1 <= N_t) ? T(1) : T(0));
                        qs[PRL(r)] = mu[r] * (a2 - b2);
                    }
                    __syncwarp();
                    const double ihl = t[T_IHL], ihl2 = ihl * ihl, s0k = t[T_S0K], s1k = t[T_S1K], ga2 = t[T_GA2];
                    const T phl = (T)t[T_PHL];
                    const T dB = (T)(-2 + 2 * ga2 * ihl2), eB = (T)(-ga2 * ihl2);
                    const T dC = (T)((1 - s0k) - 2 * s1k * ihl2), eC = (T)(s1k * ihl2);
                    for (int j = ln; j < WLa; j += L) {
                        T v = T(0);
                        if (j < keep_l && j < WLs) {
                            const T z1c = (j <= N_l) ? Lb[z1o + j] : T(0), z2c = (j <= N_l) ? Lb[z2o + j] : T(0);
                            const T z1l = (j - 1 <= N_l) ? Lb[z1o + j - 1] : T(0), z1r = (j + 1 <= N_l) ? Lb[z1o + j + 1] : T(0);
                            const T z2l = (j - 1 <= N_l) ? Lb[z2o + j - 1] : T(0), z2r = (j + 1 <= N_l) ? Lb[z2o + j + 1] : T(0);
                            const int li0 = LI[j], li1 = LI[j + 1];
                            const T pj = LW[2 * j] * ldb(qs, li0 & 0xffff) + LW[2 * j + 1] * ldb(qs, li0 >> 16);
                            const T pj1 = LW[2 * j + 2] * ldb(qs, li1 & 0xffff) + LW[2 * j + 3] * ldb(qs, li1 >> 16);
                            v = dB * z1c + eB * (z1l + z1r) + dC * z2c + eC * (z2l + z2r) - phl * (pj1 - pj);
                        }
                        // longitudinal rows sit at padded index Nx_t1 + j >= N_t + 1: x clamps to the right end
                        if (manuf && j < keep_l && j < WLs) v -= (T)ms_row(ms, NXT + j, 2.0 / t[T_IHT]);
                        Lb[rlo + j] = v;
                    }
                }
                __syncwarp();
            }

            // ---- block Gauss-Seidel solve of  A w = -(mr ; RL)  -> xs (registers), zfo (offset of the final z) ----
            T xs[ET];
#pragma unroll
            for (int r = 0; r < ET; r++) xs[r] = T(0);
            int zfo = z1o;
            bool capped = false;                 // the block iteration of this step was cut at GS_CAP (diverging string)
            // All strings of a warp sweep until every one of them has converged (extra sweeps only tighten a
            // converged string), so the loop body carries no per-string predication.
            auto gs_solve = [&](const T (&mr)[ET], bool need, bool first) {
                bool conv = !need || !valid;     // (a padding slot shadows a real string: it computes along but never holds the warp's sweep vote)
                int sweeps = 0;
                float e_prev = 0.f, isu = 0.f;
                int zco = first ? zao : zfo;
                const int keep_l = tabi[jj * NI + I_KEEPL];
                // The first solve of a step needs about as many sweeps as the first solve of the previous step did (s_prev,
                // warp-uniform): sweeps 1 .. s_prev-2 run without the stopping test and its bookkeeping (change norms, warp
                // reductions); the test starts one sweep early so that the contraction rate is still measured.
                const int s_skip = first ? s_prev - 2 : 0;
                bool have_su = false, prev_chk = false;
                do {
                    const int zno = (zco == zao) ? zbo : zao;
                    const bool chk = sweeps >= s_skip;                              // warp-uniform
                    T d[ET];
                    {
                        T y[ET];
#pragma unroll
                        for (int r = 0; r < ET; r++) {
                            const int j0 = tix[r] & 0xffff, j1 = tix[r] >> 16;
                            const float w1f = twb[r];
                            y[r] = (T)__fsub_rn(1.0f, w1f) * ldb(Lb + zco - 1, j0) + (T)w1f * ldb(Lb + zco - 1, j1);
                        }
                        T yl = shup<L>(y[ET - 1], 1);
                        if (ln == 0) yl = T(0);
                        T nuc[ET + 1];
#pragma unroll
                        for (int r = 0; r < ET; r++) nuc[r] = mu[r] * (y[r] - (r == 0 ? yl : y[r - 1]));
                        nuc[ET] = shdn<L>(nuc[0], 1);
                        if (ln == L - 1) nuc[ET] = T(0);
#pragma unroll
                        for (int r = 0; r < ET; r++) d[r] = (nuc[r + 1] - nuc[r]) - mr[r];
                    }
                    ts.solve(d, ln);
                    // max-norms of the change and of the solution, tracked on the high words of the doubles
                    // (monotone in |x|, 2^-20 resolution; NaN/inf sort above every finite value)
                    // (grouped mode: a string that needs no new solve in this pass rides along without changing its state --
                    // re-iterating a diverging string would move it and keep the group's fixed-point loop spinning)
                    const bool upd = !GROUPED || need;
                    unsigned du = 0u, su = 0u;
                    float e = 0.f;
                    bool dead = false;
                    if (chk) {
#pragma unroll
                        for (int r = 0; r < ET; r++) {
                            du = max(du, hi_abs(d[r] - xs[r]));
                            if (!have_su) su = max(su, hi_abs(d[r]));
                        }
                    }
#pragma unroll
                    for (int r = 0; r < ET; r++) if (upd) xs[r] = d[r];
                    // q = mu (x_i - x_{i-1})  (the scale phi/h_t^2 and the 1/h_t of Dxb are folded into T_PHL)
                    T xl = shup<L>(xs[ET - 1], 1);
                    if (ln == 0) xl = T(0);
#pragma unroll
                    for (int r = 0; r < ET; r++) qs[PRL(r)] = mu[r] * (xs[r] - (r == 0 ? xl : xs[r - 1]));
                    __syncwarp();
                    // P = Int_lt q is evaluated where it is used (rows j and j+1): one shared-memory round trip and one warp
                    // barrier fewer per sweep than staging P
                    {
                        const T PHL = (T)t[T_PHL], idA = (T)t[T_IDA], eidA = (T)t[T_EIDA];
                        typedef typename Vec2<T>::type T2;
                        if (upd) for (int j = ln; j < WLs; j += L) {
                            const int li0 = LI[j], li1 = LI[j + 1];
                            const T2 w0 = *(const T2 *)(LW + 2 * j), w1 = *(const T2 *)(LW + 2 * j + 2);
                            const T pj = w0.x * ldb(qs, li0 & 0xffff) + w0.y * ldb(qs, li0 >> 16);
                            const T pj1 = w1.x * ldb(qs, li1 & 0xffff) + w1.y * ldb(qs, li1 >> 16);
                            T rhs = PHL * (pj1 - pj);
                            if (j < keep_l) rhs -= Lb[rlo + j];
                            const T zr = (j + 1 < WLs) ? Lb[zco + j + 1] : T(0);
                            Lb[zno + j] = rhs * idA - eidA * (Lb[zco + j - 1] + zr);
                        }
                    }
                    __syncwarp();
                    if (upd) zco = zno;
                    sweeps++;
                    bool ok = false;
                    if (chk) {
                        if (!have_su) {
                            const float suf = hi_to_float<T>(red_maxu<L>(su));
                            isu = __fdividef(1.0f, suf);                            // 1/0 = inf: (0 * inf) = NaN ends the sweeps below
                            dead = !(suf < INFINITY);                               // NaN / inf state: nothing left to converge
                            have_su = true;
                        }
                        if (sweeps > 1) e = hi_to_float<T>(red_maxu<L>(du)) * isu;   // relative change of the transverse block in this sweep
                        if (sweeps == 1) {
                            ok = dead;
                        } else {
                            // predicted error  e * rho / (1 - rho).
                            // The longitudinal block is an affine image of the transverse one (z = A22^-1(-r_l - K_lt x), Jacobi
                            // error contracts by 2e-5 per sweep), so it needs no criterion of its own.
                            float rho = 2.0f * rho_h;
                            if (sweeps >= 3 && prev_chk && e_prev > Real<T>::RATE_MIN) {
                                // measured contraction rate e / e_prev (e_prev = 0: the iteration had already converged)
                                const float rr = __fdividef(e, fmaxf(e_prev, 1e-37f));
                                const float rh = (rr > rho_h) ? rr : 0.5f * (rho_h + rr);
                                rho_h = conv ? rho_h : rh;
                                rho = 1.5f * rho_h;
                            }
                            rho = fminf(rho, 0.9f);
                            // e rho / (1 - rho) <= tol, without the division
                            // (a warm-started re-solve of a later fixed-point pass starts from the previous pass's solution: its
                            // very first change is already a meaningful error measure)
                            ok = (sweeps >= ((keep_l > 0) ? SFDTD_MIN_SW_L : (first ? SFDTD_MIN_SW : 2))) && !(e * rho > GS_TOL * (1.0f - rho));
                            ok = ok || !(e < INFINITY) || dead || (sizeof(T) == 4 && sweeps >= 2 && e <= Real<T>::E_FLOOR);
                            e_prev = e;
                            if (sweeps >= GS_CAP && !ok && !conv) { status |= SFDTD_ST_SOLVER_CAP; capped = true; ok = true; }
                        }
                    }
                    prev_chk = chk;
                    if (!conv) { cnt_sweeps += 1; conv = ok; }
                } while (sweeps < s_skip || __any_sync(FULLMASK, !conv));          // no vote while the test is off
                if (first) s_prev = SFDTD_PREDICT_SWEEPS ? sweeps : 0;
                zfo = zco;
            };

            T vrel = T(0);
            double FH = 0.0, uH = 0.0;
            T bnum = T(0), bden = T(0);
            T nu[ET];

            // ---- bow: raised-cosine weights over the Nx_t1-point axis (bow.cpp:32, misc.cpp:20-34) ----
            // window rows bow_ic + c*L + ln (c = 0, 1); bow_o[c] = this lane's normalised weights; bow_S = normaliser
            T bow_o[2] = {T(0), T(0)}; double bow_S = 1.0; int bow_ic = 0;
            auto bow_window = [&]() {
                const double ctr = t[T_CTR], wid = t[T_WID];
                const double hw = wid * 0.5;
                bow_ic = tabi[jj * NI + I_IC];
                double acc = 0.0, bo[2];
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const int i = bow_ic + c * L + ln;
                    double o = 0.0;
                    if (i < NXT) {
                        const double x = (double)xaxs[i];
                        const double dm = __dsub_rn(__dsub_rn(x, ctr), hw), dp = __dadd_rn(__dsub_rn(x, ctr), hw);
                        const double p = __dmul_rn(-dm, dp);
                        if (p > 0) o = 0.5 * (1 + cos(((2 * M_PI) * (x - ctr)) / wid));
                        else if (p != p) o = p;
                    }
                    if (((c == 1 && ln == L - 1) || i >= LE) && o != 0.0) status |= SFDTD_ST_BOW_WINDOW;
                    bo[c] = o; acc += fabs(o);
                }
                bow_S = red_sum<L>(acc);
                bow_o[0] = (T)(bo[0] / bow_S); bow_o[1] = (T)(bo[1] / bow_S);  // 0/0 -> NaN like the reference
            };
            const double k2 = A.k2;
            const T ik = (T)A.ik;

            if (!GROUPED) {
                // ================= independent mode: unforced string, one solve =================
                // v_r of an un-bowed string (the reference evaluates it for every string, bow.cpp:35-38): the raised-cosine
                // window weights are computed before the solve, branch-free, so that their latency overlaps with it
                T bw0 = T(0), bw1 = T(0);
                if (do_bow) {
                    const double ctr = t[T_CTR], wid = t[T_WID];
                    const double hw = wid * 0.5, iw2 = 2.0 * frcp(wid);
                    bow_ic = tabi[jj * NI + I_IC];
                    auto window = [&](int c) {
                        const int i = bow_ic + c * L + ln;
                        const double x = (double)xaxs[min(i, NXT - 1)];
                        const double dm = __dsub_rn(__dsub_rn(x, ctr), hw), dp = __dadd_rn(__dsub_rn(x, ctr), hw);
                        const double p = __dmul_rn(-dm, dp);
                        // (window membership in double in every build; the weight itself in the build's arithmetic)
                        T o = T(0.5) * (T(1) + t_cospi((T)((x - ctr) * iw2)));
                        o = (p > 0 && i < NXT) ? o : ((p != p && i < NXT) ? (T)p : T(0));
                        if (((c == 1 && ln == L - 1) || i >= LE) && o != T(0)) status |= SFDTD_ST_BOW_WINDOW;
                        return o;
                    };
                    bw0 = window(0);
                    // rows bow_ic + L ... are only inside the window when it is wider than ~L - 4 grid points (x_i = i / (Nx_t1 - 1),
                    // one row of margin for the float32 axis); NaN parameters take the general path
                    const bool second = !((ctr + hw) * (double)NXT + 1.0 <= (double)(bow_ic + L));
                    if (__any_sync(FULLMASK, second)) bw1 = window(1);
                }
                gs_solve(rt, true, true);
                cnt_outer += 2;      // the reference's second pass reproduces the first (residual 0)
#pragma unroll
                for (int r = 0; r < ET; r++) {
                    const int i = i0row + r;
                    const bool keep = (i < N_t) && (i != 0) && (i < R);
                    nu[r] = keep ? xs[r] : xs[r] * T(0);
                }
                if (do_bow) {
                    // v_rel of the last pass: sum rc_i ((u_i - u1_i)/k - v_b), rc = o / sum|o|  ==  (sum o_i d_i) / (k sum o) - v_b
#pragma unroll
                    for (int r = 0; r < ET; r++) { const int i = i0row + r; qs[PRL(r)] = nu[r] - SR[u1o + PRL(r)] * ((i <= N_t) ? T(1) : T(0)); }
                    __syncwarp();
                    const int wi0 = min(bow_ic + ln, LE - 1), wi1 = min(bow_ic + L + ln, LE - 1);
                    bnum = bw0 * qs[PR(wi0)] + bw1 * qs[PR(wi1)]; bden = fabs(bw0) + fabs(bw1);     // reduced with the readout sums below
                    __syncwarp();
                }
                if (do_ham) {
                    // un-hammered string: the contact loop runs once with eta = 0 (hammer.cpp:28-53)
                    const int idxH = tabi[jj * NI + I_IDXH];
                    const double mk = (idxH <= N_t) ? 1.0 : 0.0;
                    const double eta1 = uH1 - (double)SR[u1o + PR(idxH)] * mk;
                    const double eta2 = uH2 - (double)SR[u2o + PR(idxH)] * mk;
                    const double r1 = eta1 > 0 ? eta1 : (eta1 != eta1 ? eta1 : 0.0);
                    const double ex = cst[C_AHM1];
                    const double r1pow = (ex == 2.0) ? r1 * r1 : ((ex == 0.0) ? 1.0 : pow(r1, ex));
                    const double fH = ((cst[C_WPOW] * r1pow) * (0.0 + eta2)) / 2;
                    FH = (eta1 > 0) ? fH : 0.0;
                    double tt = (((2 * uH1) - uH2) - k2 * FH) - A.mhd;
                    tt = tt > 0 ? tt : (tt != tt ? tt : 0.0);
                    uH = tt + A.mhd;
                    cnt_ham += 2;
                }
            } else {
                // ================= grouped mode: fixed-point loop over the forcing (string.cpp:200-258) =================
                const double hm = hamm ? 1.0 : 0.0;
                const int idxH = tabi[jj * NI + I_IDXH];
                const int Rk = tabi[jj * NI + I_RK];
                bool rc_nan = false;
                if (do_bow) {
                    bow_window();
                    rc_nan = (bow_S == 0.0) || (bow_S != bow_S);             // every weight is NaN (0/0) in the reference
#pragma unroll
                    for (int r = 0; r < ET; r++) RC[PRL(r)] = T(0);
                    __syncwarp();
                    if (bow_ic + ln < LE) RC[PR(bow_ic + ln)] = bow_o[0];
                    if (bow_ic + L + ln < LE) RC[PR(bow_ic + L + ln)] = bow_o[1];
                    __syncwarp();
                }
                // hammer: contact point and relative displacements (hammer.cpp:70-74)
                double eta1 = 0.0, eta2 = 0.0, r1pow = 0.0;
                if (do_ham) {
                    const double mk = (idxH <= N_t) ? 1.0 : 0.0;
                    eta1 = uH1 - (double)SR[u1o + PR(idxH)] * mk;
                    eta2 = uH2 - (double)SR[u2o + PR(idxH)] * mk;
                    const double r1 = eta1 > 0 ? eta1 : (eta1 != eta1 ? eta1 : 0.0);
                    const double ex = cst[C_AHM1];
                    r1pow = (ex == 2.0) ? r1 * r1 : ((ex == 0.0) ? 1.0 : pow(r1, ex));
                }
                const double tol_t = GROUPED ? t[T_TOLT] : 0.0;
                const T tol_tT = (T)tol_t, tol_l = GROUPED ? (T)t[T_TOLL] : T(0);
                // the iterate starts as the unmasked state[n-1] (string.cpp:190-191)
#pragma unroll
                for (int r = 0; r < ET; r++) nu[r] = SR[u1o + PRL(r)];
                for (int j = ln; j < WLa; j += L) Lb[zpo + j] = Lb[z1o + j];
                __syncwarp();
                // hammer contact loop (hammer.cpp:28-53).  Every string iterates its own scalar loop; only the exit test is an
                // any-over-batch vote.  The exit tests of the next HAM_HB passes are evaluated ahead of time and travel as a bit
                // mask in one vote word: the group leaves the loop after the first pass whose bit is clear for every string.
                HamIn hin;
                hin.eta1 = eta1; hin.eta2 = eta2; hin.wr = cst[C_WPOW] * r1pow; hin.base = (2 * uH1) - uH2; hin.hm = hm; hin.tol = tol_t;
                hin.k2 = k2; hin.mhd = A.mhd;
                double eps_u = 0.0;
                unsigned hgm = 0;                  // exit-test bits of the coming contact loop, OR-ed over the group
                if (GROUPED && group_has_hammer) {
                    eps_u = (double)fetch_row<L, ET>(nu, idxH);
                    // (normally the mask already came with the last vote of the previous step, see below)
                    if (have_next) hgm = hgm_next;
                    else hgm = gc.vote((valid ? ham_mask(hin, eta1 * hm, eps_u) : 0u) << 1) >> 1;
                    have_next = false;
                }
                int iter = 0;
                bool solved = false;
                while (true) {
                    // bow force (bow.cpp:35-40)
                    double hb = 0.0;
                    if (do_bow) {
                        const T vB = (T)t[T_VB];
                        T acc = T(0);
#pragma unroll
                        for (int r = 0; r < ET; r++) {
                            const int i = i0row + r;
                            const T mk = (i <= N_t) ? T(1) : T(0);
                            const T m1 = SR[u1o + PRL(r)] * mk;
                            const T dd = (iter == 0) ? (m1 - SR[u2o + PRL(r)] * mk) : (nu[r] - m1);
                            const T rcv = rc_nan ? bow_o[0] : RC[PRL(r)];
                            acc += rcv * (dd * ik - vB);
                        }
                        vrel = red_sum<L>(acc);
                        // the friction curve only enters the right-hand side of bowed strings (string.cpp:225)
                        if (bowm) {
                            const double vr = (double)vrel;
                            const double sg = (vr > 0) ? 1.0 : ((vr < 0) ? -1.0 : 0.0);
                            const double phi0 = cst[C_PHI0], phi1 = cst[C_PHI1];
                            hb = (vr != vr) ? vr : sg * (phi1 + (1 - phi1) * exp(-phi0 * fabs(vr)));
                        }
                    }
                    // hammer loop (hammer.cpp:28-53); its exit test is an any-over-batch vote
                    if (do_ham && GROUPED && group_has_hammer) {
                        constexpr unsigned FULL = (1u << HAM_HB) - 1u;
                        double eta = eta1 * hm;
                        int hit = 0;
                        for (;;) {
                            const bool full = (hgm & FULL) == FULL;            // no pass of this chunk ends the loop
                            int np = full ? HAM_HB : __ffs((int)(~hgm & FULL));
                            bool cap = false;
                            if (hit + np >= A.max_iter) { np = A.max_iter - hit; cap = true; }
                            for (int p = 0; p < np; p++) eta = ham_pass(hin, eta, eps_u, FH, uH);
                            hit += np;
                            if (cap) { if (full) status |= SFDTD_ST_HAMMER_CAP; break; }
                            if (!full) break;
                            hgm = gc.vote((valid ? ham_mask(hin, eta, eps_u) : 0u) << 1) >> 1;
                        }
                        cnt_ham += hit;
                    } else if (do_ham) {
                        // no hammered string in the group: eta = 0 for every string, the loop ends after one pass
                        const double fH = (hin.wr * (0.0 + eta2)) / 2;
                        FH = (eta1 > 0) ? fH : 0.0;
                        double tt = (hin.base - k2 * FH) - A.mhd;
                        tt = tt > 0 ? tt : (tt != tt ? tt : 0.0);
                        uH = tt + A.mhd;
                        cnt_ham += 1;
                    }
                    // ---- linear solve  A w = -(RHS)  ----
                    const bool need = !solved || forced;
                    if (__any_sync(FULLMASK, need)) {
                        T mr[ET];
                        const T sB = (T)(-k2 * ((GROUPED ? t[T_FB] : 0.0) * hb) * t[T_IHT]);
                        const T sH = hamm ? (T)nan0(-k2 * (cst[C_MR] * FH)) : T(0);
#pragma unroll
                        for (int r = 0; r < ET; r++) {
                            const int i = i0row + r;
                            T f = T(0);
                            if (bowm) f += nan0(sB * (rc_nan ? bow_o[0] : RC[PRL(r)]));
                            if (hamm && i == idxH) f += sH;
                            mr[r] = (i < Rk) ? rt[r] + f : T(0);
                        }
                        gs_solve(mr, need, !solved);
                        solved = true;
                    }
                    // ---- mask + Dirichlet (string.cpp:240-246), residuals (string.cpp:248-253) ----
                    int nc_t = 0, nan_u = 0;
#pragma unroll
                    for (int r = 0; r < ET; r++) {
                        const int i = i0row + r;
                        const bool keep = (i < N_t) && (i != 0) && (i < R);
                        const T nv = keep ? xs[r] : xs[r] * T(0);
                        const T df = fabs(nu[r] - nv);
                        nan_u |= (df != df);
                        nc_t |= (df > tol_tT);
                        nu[r] = nv;
                    }
                    int nc_l = 0, nan_z = 0;
                    for (int j = ln; j < WLa; j += L) {
                        const bool keep = (j < N_l) && (j != 0) && (j < WLs);
                        const T zs = (j < WLs) ? Lb[zfo + j] : T(0);
                        const T zv = keep ? zs : zs * T(0);
                        const T df = fabs(Lb[zpo + j] - zv);
                        nan_z |= (df != df);
                        nc_l |= (df > tol_l);
                        Lb[zpo + j] = zv;
                    }
                    __syncwarp();
                    nan_u = red_or<L>(nan_u); nan_z = red_or<L>(nan_z);
                    nc_t = red_or<L>(nc_t);
                    nc_l = red_or<L>(nc_l);
                    // a string whose linear iteration diverges has no fixed point to wait for: it does not vote
                    const int not_conv = !capped && ((nc_t && !nan_u) || (nc_l && !nan_z));
                    iter++;
                    unsigned word = (valid && not_conv) ? 1u : 0u;
                    bool spec = false;
                    if (GROUPED && group_has_hammer) {
                        // exit tests of the next pass's contact loop (contact-point displacement of the new iterate) ride along ...
                        eps_u = (double)fetch_row<L, ET>(nu, idxH);
                        if (valid) word |= ham_mask(hin, eta1 * hm, eps_u) << 1;
                        // ... and so do those of the NEXT STEP's first contact loop, in case this pass turns out to be the last
                        // one of the step: everything they need (new state row, hammer displacement of this pass, the next
                        // step's table row) is known here.  Saves the step's own vote (one cluster barrier of three).
                        spec = !save_state && !manuf && (jj + 1 < jmax);
                        if (spec) {
                            const double *tn = tab + (jj + 1) * NV;
                            const int *tin = tabi + (jj + 1) * NI;
                            const int idxHn = tin[I_IDXH];
                            const double mkn = (idxHn <= tin[I_NT]) ? 1.0 : 0.0;
                            const double uH1n = t[T_UHPRE] + (out_ham ? uH : 0.0);
                            const double u1n = (double)fetch_row<L, ET>(nu, idxHn), u2n = (double)SR[u1o + PR(idxHn)];
                            HamIn hn;
                            hn.eta1 = uH1n - u1n * mkn; hn.eta2 = uH1 - u2n * mkn;
                            const double r1n = hn.eta1 > 0 ? hn.eta1 : (hn.eta1 != hn.eta1 ? hn.eta1 : 0.0);
                            const double ex = cst[C_AHM1];
                            hn.wr = cst[C_WPOW] * ((ex == 2.0) ? r1n * r1n : ((ex == 0.0) ? 1.0 : pow(r1n, ex)));
                            hn.base = (2 * uH1n) - uH1; hn.hm = hm; hn.tol = tn[T_TOLT]; hn.k2 = k2; hn.mhd = A.mhd;
                            if (valid) word |= ham_mask(hn, hn.eta1 * hm, u1n) << (1 + HAM_HB);
                        }
                    }
                    const unsigned gvw = gc.vote(word);
                    hgm = gvw >> 1;
                    int more = gvw & 1u;
                    if (iter >= A.max_iter) { if (more) status |= SFDTD_ST_OUTER_CAP; more = 0; }
                    if (!more) { have_next = spec; hgm_next = gvw >> (1 + HAM_HB); break; }
                }
                cnt_outer += iter;
            }
            cnt_steps += 1;

            // ---- save and readout (string.cpp:263-303) ----
            T uo, zo;
            {
                // new longitudinal row (masked, Dirichlet) into the oldest buffer; surface integral on the fly
                const int zno = z2o;
                const T rdw = (T)t[T_RDW];
                T acc = T(0);
                const int hi = max(max(ext1, ext2), WLs);
                if (GROUPED) {
                    for (int j = ln; j < hi; j += L) { const T zv = Lb[zpo + j]; acc += (zv - Lb[z1o + j]) * rdw; Lb[zno + j] = zv; }
                } else {
                    for (int j = ln; j < hi; j += L) {
                        const bool keep = (j < N_l) && (j != 0) && (j < WLs);
                        const T zs = (j < WLs) ? Lb[zfo + j] : T(0);
                        const T zv = keep ? zs : zs * T(0);
                        acc += (zv - Lb[z1o + j]) * rdw; Lb[zno + j] = zv;
                    }
                }
                ext2 = ext1; ext1 = WLs;
                if (surf) {
                    T au = T(0);
#pragma unroll
                    for (int r = 0; r < ET; r++) au += (nu[r] - SR[u1o + PRL(r)]) * rdw;
                    // one butterfly for all sums of the step (independent shuffles overlap)
#pragma unroll
                    for (int o = L / 2; o > 0; o >>= 1) {
                        acc += __shfl_xor_sync(FULLMASK, acc, o, L); au += __shfl_xor_sync(FULLMASK, au, o, L);
                        if (!GROUPED) { bnum += __shfl_xor_sync(FULLMASK, bnum, o, L); bden += __shfl_xor_sync(FULLMASK, bden, o, L); }
                    }
                    zo = acc; uo = au;
                } else {
                    if (!GROUPED) { bnum = red_sum<L>(bnum); bden = red_sum<L>(bden); }
                    __syncwarp();
                    const double rp = cst[C_RP];
                    const int ui = 1 + (int)floor(__dmul_rn((double)N_t, rp));
                    const T uf = (T)(1 + rp * t[T_IHT] - (double)ui);
                    const int zi = 1 + (int)floor(__dmul_rn((double)N_l, rp));
                    const T zf = (T)(1 + rp * t[T_IHL] - (double)zi);
                    const T ua = fetch_row<L, ET>(nu, ui), ub = fetch_row<L, ET>(nu, ui + 1);
                    uo = (T(1) - uf) * ua + uf * ub;
                    const T za = (zi < WLa) ? Lb[zno + zi] : T(0), zb = (zi + 1 < WLa) ? Lb[zno + zi + 1] : T(0);
                    zo = (T(1) - zf) * za + zf * zb;
                }
                if (!GROUPED && do_bow) vrel = (bnum * ik) / bden - (T)t[T_VB];     // empty window: 0/0 = NaN like the reference
                // state rows: state[:, n] += u  (in place, onto pre-loaded content; string.cpp:264-265)
                if (save_state) {
                    T *su = (T *)a.state_u.ptr + (int64_t)b * a.state_u.bs + (int64_t)n * a.state_u.ts;
#pragma unroll
                    for (int r = 0; r < ET; r++) {
                        const int i = i0row + r;
                        // (spare slots shadow a real string: they must not read rows its owner is writing)
                        if (i < NXT && valid) { nu[r] += su[i]; su[i] = nu[r]; }
                    }
                    T *sz = (T *)a.state_z.ptr + (int64_t)b * a.state_z.bs + (int64_t)n * a.state_z.ts;
                    __syncwarp();
                    for (int j = ln; j < WLa; j += L) {
                        if (j < NXL && valid) { const T row = Lb[zno + j] + sz[j]; sz[j] = row; Lb[zno + j] = row; }
                    }
                    ext1 = WLa;
                }
#pragma unroll
                for (int r = 0; r < ET; r++) SR[u2o + PRL(r)] = nu[r];
                { const int tmp = u1o; u1o = u2o; u2o = tmp; }
                z2o = z1o; z1o = zno;
                __syncwarp();
            }
            const double uHtot = t[T_UHPRE] + (out_ham ? uH : 0.0);
            uH2 = uH1; uH1 = uHtot;
            if (ln == 0) {
                double *o = ost + jj * NOUT;
                o[0] = (double)uo; o[1] = (double)zo; o[2] = out_bow ? (double)vrel : 0.0; o[3] = out_ham ? FH : 0.0; o[4] = uHtot;
            }
        }
        // ---- flush staged outputs: lane s writes step n0+s (coalesced rows) ----
        __syncwarp();
        if (valid) {
            const double ik = A.ik;
            for (int s = ln; s < jmax; s += L) {
                const int n = n0 + s;
                const double *o = ost + s * NOUT;
                ((T *)a.uout.ptr)[(int64_t)b * a.uout.bs + (int64_t)n * a.uout.ts] = (T)o[0];
                ((T *)a.zout.ptr)[(int64_t)b * a.zout.bs + (int64_t)n * a.zout.ts] = (T)o[1];
                if (a.v_r.ptr) ((T *)a.v_r.ptr)[(int64_t)b * a.v_r.bs + (int64_t)n * a.v_r.ts] = (T)o[2];
                if (a.F_H.ptr) ((T *)a.F_H.ptr)[(int64_t)b * a.F_H.bs + (int64_t)n * a.F_H.ts] = (T)o[3];
                if (a.u_H.ptr) ((T *)a.u_H.ptr)[(int64_t)b * a.u_H.bs + (int64_t)n * a.u_H.ts] = (T)o[4];
                if (a.u_H_out.ptr) ((T *)a.u_H_out.ptr)[(int64_t)b * a.u_H_out.bs + (int64_t)n * a.u_H_out.ts] = (T)(o[4] * ik);
            }
        }
        __syncwarp();
    }

    // ---- epilogue ----
    const uint32_t status_all = (uint32_t)red_or<L>((int)status);
    if (valid) {
        if (!save_state && Nt > 2) {
            T *su = (T *)a.state_u.ptr + (int64_t)b * a.state_u.bs;
#pragma unroll
            for (int r = 0; r < ET; r++) {
                const int i = ln * ET + r;
                if (i < NXT) { su[i] = SR[u2o + PR(i)]; su[a.state_u.ts + i] = SR[u1o + PR(i)]; }
            }
            T *sz = (T *)a.state_z.ptr + (int64_t)b * a.state_z.bs;
            for (int j = ln; j < WLa; j += L) if (j < NXL) { sz[j] = Lb[z2o + j]; sz[a.state_z.ts + j] = Lb[z1o + j]; }
        }
        // u_H_out / u_H columns 0,1 (simulator.cpp:57 divides the whole tensor)
        if (ln < 2 && ln < Nt && qst[0] == 2 && a.u_H_out.ptr) {
            ((T *)a.u_H_out.ptr)[(int64_t)b * a.u_H_out.bs + (int64_t)ln * a.u_H_out.ts] = (T)(ctl_uH<T>(A, b, ln) * A.ik);
        }
        if (ln == 0 && !a.u_H.ptr && A.uH_carry) { A.uH_carry[2 * (int64_t)b] = uH2; A.uH_carry[2 * (int64_t)b + 1] = uH1; }
        if (ln == 0) {
            if (Nt > 2 && qst[1] == Nt) final_sigmas(A, b, Nt);
            // time slices of one call accumulate (the caller zero-initialises both arrays)
            if (a.status) a.status[b] |= status_all;
            if (a.counters) {
                a.counters[4 * b + 0] += cnt_outer; a.counters[4 * b + 1] += cnt_sweeps;
                a.counters[4 * b + 2] += cnt_ham; a.counters[4 * b + 3] += cnt_steps;
            }
        }
    }
    __syncwarp();
    if (queued) {
        // hand the state rows (and u_H) of this slice over to whichever warp runs the set's next slice
        __threadfence();
        __syncwarp();
        // (every lane stores the same word: a store under a lane predicate at the loop's back edge makes the compiler treat
        // the whole loop body as possibly diverged and wrap every shuffle in a reconvergence sequence)
        asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(A.done + qst[2]), "r"(qst[3] + 1) : "memory");
    }
    }   // work-queue loop
}

// independent mode: strings of unforced groups, any warp of any CTA
template <typename T, int L, int ET, int MAXT, int MINB, bool PF = false>
__global__ void __launch_bounds__(MAXT, MINB) sfdtd_step_kernel(const __grid_constant__ KArgs A) {
    step_body<T, L, ET, false, false, PF>(A, nullptr);
}
// grouped mode: one thread-block cluster of 128-thread CTAs per group; the CTA's descriptor picks the lane/row shape of its
// string slots.  KIND 0: strings of <= 64 rows on 16 lanes x 4 rows, <= 128 rows on 32 lanes x 4 rows (168 registers, three
// CTAs per SM -- of different groups, so that one group's barrier waits are filled by the others);  KIND 1: 32 lanes x 8 rows;
// KIND 2: 32 lanes x 8 rows with the manufactured-solution forcing (vnv.cpp);  KIND 3: 32 lanes x 20 rows (strings of up to
// 640 rows: f0 down to the reference's default floor of 27.5 Hz at 48 kHz; the row arrays of that shape exceed the register
// file and partly live in local memory -- a coverage kernel, not a fast one).
#ifndef SFDTD_GROUP_MINB
#define SFDTD_GROUP_MINB 3
#endif
#ifndef SFDTD_F32_GROUP_MINB
#define SFDTD_F32_GROUP_MINB 3
#endif
template <typename T, int KIND, bool PF = false>
__global__ void __launch_bounds__(128, KIND == 0 ? (sizeof(T) == 4 ? SFDTD_F32_GROUP_MINB : SFDTD_GROUP_MINB) : 1)
sfdtd_group_kernel(const __grid_constant__ KArgs A) {
    const CtaDesc *cd = A.ctas + blockIdx.x;
    if (KIND == 0) {
        if (cd->cls == 0) step_body<T, 16, 4, true, false, PF>(A, cd);
        else step_body<T, 32, 4, true, false, PF>(A, cd);
    } else if (KIND == 1) {
        step_body<T, 32, 8, true, false, PF>(A, cd);
    } else if (KIND == 2) {
        step_body<T, 32, 8, true, true>(A, cd);
    } else {
        step_body<T, 32, 20, true, false>(A, cd);
    }
}


// ---- sfdtd_synth_controls: the curves exactly as the stepper evaluates them -----------------------------------------
__global__ void sfdtd_synth_controls_kernel(const __grid_constant__ KArgs A, sfdtd_array f0, sfdtd_array x_b, sfdtd_array v_b,
                                            sfdtd_array F_b, sfdtd_array u_H) {
    const int b = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= A.a.Nt) return;
    const int g = n + A.sy.t_0;
    auto put = [&](const sfdtd_array &o, double v) { if (o.ptr) ((double *)o.ptr)[(int64_t)b * o.bs + (int64_t)n * o.ts] = v; };
    put(f0, sy_f0(A.sy, b, g)); put(x_b, sy_xb(A.sy, b, g)); put(v_b, sy_vb(A.sy, b, g)); put(F_b, sy_Fb(A.sy, b, g));
    put(u_H, sy_uH(A.sy, b, g));
}

// ---- sfdtd_postprocess: NaN / silence flags, l-infinity gain, PCM quantisation (one CTA per string) ------------------
template <int BITS, typename T>
__global__ void __launch_bounds__(256) sfdtd_postprocess_kernel(sfdtd_array U, sfdtd_array Z, int n0, int ns, double silence_db,
                                                               int normalize, uint8_t *is_nan, uint8_t *is_silent, double *gain_out,
                                                               uint8_t *pu, uint8_t *pz, uint8_t *pw, int64_t pitch) {
    const int b = blockIdx.x;
    const T *u = (const T *)U.ptr + (int64_t)b * U.bs, *z = (const T *)Z.ptr + (int64_t)b * Z.bs;
    double sq = 0.0, mx = 0.0; int nan = 0;
    for (int n = threadIdx.x; n < ns; n += blockDim.x) {
        const double v = (double)u[(int64_t)(n0 + n) * U.ts];
        nan |= (v != v); sq += v * v; mx = fmax(mx, fabs(v));
    }
    __shared__ double s_sq[8], s_mx[8]; __shared__ int s_nan[8];
    for (int o = 16; o > 0; o >>= 1) {
        sq += __shfl_xor_sync(FULLMASK, sq, o); mx = fmax(mx, __shfl_xor_sync(FULLMASK, mx, o)); nan |= __shfl_xor_sync(FULLMASK, nan, o);
    }
    if ((threadIdx.x & 31) == 0) { s_sq[threadIdx.x >> 5] = sq; s_mx[threadIdx.x >> 5] = mx; s_nan[threadIdx.x >> 5] = nan; }
    __syncthreads();
    sq = 0.0; mx = 0.0; nan = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) { sq += s_sq[w]; mx = fmax(mx, s_mx[w]); nan |= s_nan[w]; }
    // a NaN string is zeroed before the silence test (simulate.py:334-335): rms 0 -> -inf dB -> silent
    const double rms = nan ? 0.0 : sqrt(sq / (double)ns);
    const bool silent = 20.0 * log10(rms) <= silence_db;
    double gain = 1.0;
    if (normalize && !nan && mx != 0.0) gain = 1.0 / mx;          // ell_infty_normalize (audio.py:42-48)
    if (threadIdx.x == 0) {
        if (is_nan) is_nan[b] = (uint8_t)nan;
        if (is_silent) is_silent[b] = silent ? 1 : 0;
        if (gain_out) gain_out[b] = gain;
    }
    if (!pu && !pz && !pw) return;
    constexpr int BY = BITS / 8;
    constexpr double SC = BITS == 16 ? 32768.0 : 8388608.0, LO = -SC, HI = SC - 1.0;
    auto q = [&](double v) -> int { v = (v != v) ? 0.0 : v * SC; return __double2int_rn(fmin(fmax(v, LO), HI)); };
    const int64_t row = (int64_t)b * (pitch ? pitch : (int64_t)ns * BY);
    const bool aligned = (pitch % 4) == 0 && pitch > 0;
    // four samples per thread and pass: 4 * BY bytes = BY 32-bit words when the rows are word aligned
    for (int n4 = threadIdx.x * 4; n4 < ns; n4 += blockDim.x * 4) {
        int qu[4], qz[4], qw[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int n = n4 + c;
            double uv = 0.0, zv = 0.0;
            if (n < ns) { uv = (double)u[(int64_t)(n0 + n) * U.ts]; zv = (double)z[(int64_t)(n0 + n) * Z.ts]; }
            qu[c] = q(gain * uv); qz[c] = q(gain * zv); qw[c] = q(gain * uv + gain * zv);
        }
        auto store = [&](uint8_t *base, const int (&v)[4]) {
            if (!base) return;
            uint8_t *o = base + row + (int64_t)n4 * BY;
            if (aligned && n4 + 4 <= ns) {
                uint32_t *w = (uint32_t *)o;
                if (BITS == 16) {
                    w[0] = (uint32_t)(v[0] & 0xffff) | ((uint32_t)v[1] << 16); w[1] = (uint32_t)(v[2] & 0xffff) | ((uint32_t)v[3] << 16);
                } else {
                    const uint32_t a = (uint32_t)v[0] & 0xffffffu, bq = (uint32_t)v[1] & 0xffffffu, c = (uint32_t)v[2] & 0xffffffu, d = (uint32_t)v[3] & 0xffffffu;
                    w[0] = a | (bq << 24); w[1] = (bq >> 8) | (c << 16); w[2] = (c >> 16) | (d << 8);
                }
            } else {
                for (int c = 0; c < 4 && n4 + c < ns; c++)
                    for (int k = 0; k < BY; k++) o[c * BY + k] = (uint8_t)((uint32_t)v[c] >> (8 * k));
            }
        };
        store(pu, qu); store(pz, qz); store(pw, qw);
    }
}

// ---- FMA-pipe peak microbenchmark (roofline denominator; MEASURED_PEAKS.json has no FP64/FP32 FMA figure) ----
template <typename T>
__global__ void __launch_bounds__(256) sfdtd_fma_peak_kernel(T *out, int iters, T seed) {
    T a0 = seed + (T)threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const T m = (T)0.999999, c = (T)1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
            a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
        }
    }
    const T r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == (T)-1) out[0] = r;     // never true: keeps the chains alive
}

// ======================================================================================================
// host side
// ======================================================================================================
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

#ifndef SFDTD_F32_L8
#define SFDTD_F32_L8 0
#endif
#ifndef SFDTD_F32_L8_MINB
#define SFDTD_F32_L8_MINB 3
#endif
#ifndef SFDTD_F32_MINB
#define SFDTD_F32_MINB 4            // CTAs per SM the small fp32 kernels are compiled for (register cap 65536 / (128 x MINB))
#endif
// independent-mode kernels: <=128-thread CTAs, a string needs rows <= L*ET; tier = kernel set
// (tier 2, the default: 16 lanes x 4 rows for every string up to 64 rows, then 32x4, 32x8 -- measured fastest; tier 0 also
// uses the 8-lane kernels, for A/B runs via SFDTD_TIER=0; 2- and 3-row kernels, tighter register caps (128) and a
// REDUX-based max reduction all measured slower).
// kern[dtype]: the fp64 build and (default tier only) the fp32 build; MB32_ = CTAs per SM the fp32 build is compiled for
// kern_pf[dtype]: the build for SFDTD_SAVE_STATE calls (state-history rows prefetched), where one exists
struct Config { int L, ET, tier; void (*kern[2])(const KArgs); void (*kern_pf[2])(const KArgs); };
#define CFG_I(L_, ET_, MB_, TIER_) Config{L_, ET_, TIER_, {sfdtd_step_kernel<double, L_, ET_, 128, MB_>, nullptr}, {nullptr, nullptr}}
#define CFG_D(L_, ET_, MB_, MB32_, TIER_) Config{L_, ET_, TIER_, {sfdtd_step_kernel<double, L_, ET_, 128, MB_>, sfdtd_step_kernel<float, L_, ET_, 128, MB32_>}, {nullptr, nullptr}}
#define CFG_P(L_, ET_, MB_, MB32_, TIER_) Config{L_, ET_, TIER_, {sfdtd_step_kernel<double, L_, ET_, 128, MB_>, sfdtd_step_kernel<float, L_, ET_, 128, MB32_>}, \
                                                 {sfdtd_step_kernel<double, L_, ET_, 128, MB_, true>, sfdtd_step_kernel<float, L_, ET_, 128, MB32_, true>}}
#define CFG_F(L_, ET_, MB32_, TIER_) Config{L_, ET_, TIER_, {nullptr, sfdtd_step_kernel<float, L_, ET_, 128, MB32_>}, {nullptr, nullptr}}
const Config g_configs[] = {   // smallest first
    CFG_I(8, 4, 3, 0), CFG_I(8, 6, 2, 0), CFG_I(16, 4, 3, 0), CFG_I(16, 6, 2, 0), CFG_I(32, 4, 3, 0), CFG_I(32, 8, 1, 0),
#if SFDTD_F32_L8
    // fp32 only: 8 lanes x 8 rows (a float row costs half the registers of a double one: twice the rows per lane, one
    // cyclic-reduction level and half the shuffles per row less)
    CFG_F(8, 8, SFDTD_F32_L8_MINB, 2),
#endif
    CFG_P(16, 4, 3, SFDTD_F32_MINB, 2), CFG_P(32, 4, 3, SFDTD_F32_MINB, 2), CFG_P(32, 8, 1, 2, 2), CFG_D(32, 12, 1, 1, 2), CFG_D(32, 20, 1, 1, 2),
    CFG_I(8, 4, 3, 3), CFG_I(16, 4, 3, 3), CFG_I(32, 4, 3, 3), CFG_I(32, 8, 1, 3),
};
#ifndef SFDTD_DEFAULT_TIER
#define SFDTD_DEFAULT_TIER 2
#endif
constexpr int N_CONFIGS = sizeof(g_configs) / sizeof(g_configs[0]);
// grouped-mode kernels and the slot shapes of their CTA classes
void (*const g_group_kernels[2][4])(const KArgs) = {
    {sfdtd_group_kernel<double, 0>, sfdtd_group_kernel<double, 1>, sfdtd_group_kernel<double, 2>, sfdtd_group_kernel<double, 3>},
    {sfdtd_group_kernel<float, 0>, sfdtd_group_kernel<float, 1>, nullptr, sfdtd_group_kernel<float, 3>}};   // (no fp32 manufactured mode)
void (*const g_group_kernels_pf[2][4])(const KArgs) = {        // SFDTD_SAVE_STATE builds (prefetch), kinds 0 and 1
    {sfdtd_group_kernel<double, 0, true>, sfdtd_group_kernel<double, 1, true>, nullptr, nullptr},
    {sfdtd_group_kernel<float, 0, true>, sfdtd_group_kernel<float, 1, true>, nullptr, nullptr}};
struct GShape { int L, ET; };
const GShape g_gshape[4][2] = {{{16, 4}, {32, 4}}, {{32, 8}, {32, 8}}, {{32, 8}, {32, 8}}, {{32, 20}, {32, 20}}};
constexpr int MAX_ROWS = 640;       // transverse rows of the largest kernel shape (32 lanes x 20 rows)

// longitudinal allocation classes of the independent mode (rows incl. the two guards)
int wl_class(int rows, int c_min = 16) {
    int c = c_min;
    while (c < rows) c *= 2;
    return c;
}
size_t xax_doubles(int NXT, bool need_xax) { return need_xax ? (size_t)((NXT + 3) / 4) * 2 : 0; }
// independent mode: nslots strings with WLp longitudinal rows each
size_t smem_bytes_indep(const Config &c, int nslots, int NXT, int WLp, bool need_xax, int tsz) {
    const size_t dbl = xax_doubles(NXT, need_xax) + (size_t)nslots * (slot_fixed_doubles(c.L, c.ET, false, tsz) + slot_long_doubles(WLp, false, c.L, tsz));
    return dbl * sizeof(double) + 16;
}
int kernel_regs(void (*kern)(const KArgs)) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) { cudaGetLastError(); return 255; }
    return fa.numRegs;
}
int env_int(const char *name, int dflt) { const char *v = getenv(name); return v ? atoi(v) : dflt; }
double env_dbl(const char *name, double dflt) { const char *v = getenv(name); return v ? atof(v) : dflt; }

// side streams so that the launches of different buckets overlap.  Every plan owns its own set for its lifetime (calls of
// different plans in flight on different caller streams must not serialise behind each other); released streams go back to
// a per-device free list (work still queued on them stays ordered).
std::mutex g_mu;
// mapped pinned result buffers of the prepass, pooled per process (one per plan being created; a few hundred KB each; freeing
// pinned memory synchronises the device, so they are kept)
struct PinBuf { void *p; size_t n; bool busy; };
std::vector<PinBuf> g_pin;
void *pin_acquire(size_t n) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (PinBuf &b : g_pin) if (!b.busy && b.n >= n) { b.busy = true; return b.p; }
    PinBuf nb; nb.n = std::max(n, (size_t)1 << 20); nb.busy = true; nb.p = nullptr;
    if (cudaHostAlloc(&nb.p, nb.n, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    g_pin.push_back(nb);
    return nb.p;
}
void pin_release(void *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_mu);
    for (PinBuf &b : g_pin) if (b.p == p) b.busy = false;
}
struct PoolStream { cudaStream_t s; int prio; };
std::map<int, std::vector<PoolStream>> g_free_streams;
std::map<int, bool> g_pool_ready;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { snprintf(g_err, sizeof g_err, "%s: %s", #x, cudaGetErrorString(e_)); rc = SFDTD_ERR_CUDA; goto done; } } while (0)

struct Launch {
    bool grouped = false;
    int cfg = 0;                    // independent: index into g_configs; grouped: kernel kind
    int grid = 0, threads = 0, cluster = 1, WLp = 0, n_items = 0;
    size_t smem = 0, off = 0;       // off: first entry of ids (independent) / of ctas (grouped)
    bool need_xax = false;
    bool queue = false; int q_slice = 0, q_nslices = 0, q_full = 0; size_t q_idx = 0, done_off = 0;
    void (*fn)(const KArgs) = nullptr;      // the kernel of this launch (compact or SAVE_STATE build)
};

int validate(const sfdtd_args *args) {
    if (!args) { snprintf(g_err, sizeof g_err, "args is NULL"); return SFDTD_ERR_ARG; }
    const sfdtd_args &a = *args;
    if (a.abi_version != SFDTD_ABI_VERSION) { snprintf(g_err, sizeof g_err, "abi_version %d != %d", a.abi_version, SFDTD_ABI_VERSION); return SFDTD_ERR_ARG; }
    if (a.dtype != SFDTD_F64 && a.dtype != SFDTD_F32) { snprintf(g_err, sizeof g_err, "dtype must be SFDTD_F64 or SFDTD_F32"); return SFDTD_ERR_UNSUPPORTED; }
    if (a.dtype == SFDTD_F32 && (a.flags & SFDTD_MANUFACTURED)) {
        snprintf(g_err, sizeof g_err, "the manufactured-solution mode is only built for SFDTD_F64"); return SFDTD_ERR_UNSUPPORTED;
    }
    if (a.B <= 0 || a.group_size <= 0 || a.Nt < 0 || a.Nx_t1 <= 0 || a.Nx_l1 <= 0) { snprintf(g_err, sizeof g_err, "bad sizes"); return SFDTD_ERR_ARG; }
    const void *req[] = {a.state_u.ptr, a.state_z.ptr, a.kappa.ptr, a.alpha.ptr, a.pos.ptr, a.T60.ptr, a.phi_0.ptr, a.phi_1.ptr,
                         a.x_H.ptr, a.w_H.ptr, a.M_r.ptr, a.alpha_H.ptr, a.bow_mask, a.hammer_mask, a.xax, a.uout.ptr, a.zout.ptr,
                         a.sig0, a.sig1};
    for (const void *p : req) if (!p) { snprintf(g_err, sizeof g_err, "a required pointer is NULL"); return SFDTD_ERR_ARG; }
    if (a.synth) {
        const sfdtd_synth &y = *a.synth;
        const void *rq[] = {y.f0_a, y.f0_b, y.mod_frq, y.mod_amp, y.vib_t0, y.x_b1, y.x_b2, y.v_b1, y.v_b2, y.F_b1, y.F_b2, y.pulloff, y.wid, y.v_H};
        for (const void *p : rq) if (!p) { snprintf(g_err, sizeof g_err, "a pointer of sfdtd_synth is NULL"); return SFDTD_ERR_ARG; }
        if (y.Nt_full <= 0 || !(y.sr > 0)) { snprintf(g_err, sizeof g_err, "bad sfdtd_synth sizes"); return SFDTD_ERR_ARG; }
    } else {
        const void *rq[] = {a.f0.ptr, a.x_b.ptr, a.v_b.ptr, a.F_b.ptr, a.wid.ptr, a.u_H.ptr};
        for (const void *p : rq) if (!p) { snprintf(g_err, sizeof g_err, "a required pointer is NULL"); return SFDTD_ERR_ARG; }
    }
    if ((a.flags & SFDTD_MANUFACTURED) && !a.p_a.ptr) { snprintf(g_err, sizeof g_err, "p_a is NULL"); return SFDTD_ERR_ARG; }
    return SFDTD_OK;
}

void fill_kargs(KArgs &K, const sfdtd_args &a) {
    memset(&K, 0, sizeof K);
    K.a = a;
    if (a.synth) { K.sy = *a.synth; K.has_synth = 1; }
    K.a.synth = nullptr;
    K.k = (double)a.k; K.ik = 1.0 / (double)a.k; K.k2 = pow((double)a.k, 2.); K.k4 = pow((double)a.k, 4.);
    K.th = (double)a.theta_t;
    { const float om = 1 - a.theta_t; K.omth = (double)om; }                 // float32 (string.cpp:148)
    { const float t1 = 2 * a.theta_t - 1; const float t2 = 2 * t1; K.tt1 = (double)t1; K.tt2 = (double)t2; }   // string.cpp:30-31
    K.lamc = (double)a.lambda_c; K.order = (double)a.relative_order;
    K.mhd = (double)(-0.01f);                                                // hammer.cpp:3
    K.max_iter = a.max_iter > 0 ? a.max_iter : 100;
    K.f32 = a.dtype == SFDTD_F32 ? 1 : 0;
}

// the device the call's memory lives on becomes the current device for the duration of the call
struct DeviceGuard {
    int prev = -1; bool switched = false;
    cudaError_t enter(const void *ptr) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) return e;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeDevice && at.device != prev) {
            e = cudaSetDevice(at.device);
            if (e != cudaSuccess) return e;
            switched = true;
        } else cudaGetLastError();
        return cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

}  // namespace

struct sfdtd_plan {
    int dev = 0, n_sms = 0;
    int32_t B = 0, group_size = 0, Nt = 0, Nx_t1 = 0, Nx_l1 = 0, n_groups = 0, dtype = 0;
    uint32_t flags = 0;
    std::vector<Launch> launches;
    int32_t *d_max = nullptr;        // prepass output [N_t maxima | N_l maxima] (the steppers read the N_l half)
    char *block = nullptr;           // one device allocation: (unused) | ids | ctas | queue words | width table | u_H carry
    int32_t *d_maxNl = nullptr, *d_ids = nullptr, *d_queue = nullptr, *d_wtab = nullptr;
    CtaDesc *d_ctas = nullptr;
    double *d_uH = nullptr;
    size_t queue_words = 0, n_buckets = 0;
    cudaEvent_t fork = nullptr;
    std::vector<cudaEvent_t> joins;
    std::vector<cudaStream_t> side;  // one per bucket (taken from / returned to the device's free list)
    std::vector<int> side_prio;
    bool verbose = false;
};

extern "C" const char *sfdtd_last_error(void) { return g_err; }
extern "C" int sfdtd_abi_version(void) { return SFDTD_ABI_VERSION; }
extern "C" int64_t sfdtd_launch_count(void) { return g_launches.load(); }

// Measures the achievable FMA-pipe rate (TFLOP/s, 2 flops per FMA) of the current device: which = 0 fp64, 1 fp32.
extern "C" int sfdtd_measure_fma_peak(int which, double *tflops) {
    g_err[0] = 0;
    if (!tflops) return SFDTD_ERR_ARG;
    int rc = SFDTD_OK, dev = 0, sms = 0;
    void *buf = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float ms = 0, best = 1e30f;
    const int iters = which == 0 ? 4096 : 8192, blocks_per_sm = 8;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaMalloc(&buf, 64));
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 6; rep++) {
        CK(cudaEventRecord(e0));
        if (which == 0) sfdtd_fma_peak_kernel<double><<<sms * blocks_per_sm, 256>>>((double *)buf, iters, 1.0);
        else sfdtd_fma_peak_kernel<float><<<sms * blocks_per_sm, 256>>>((float *)buf, iters, 1.0f);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    *tflops = 2.0 * 64.0 * iters * 256.0 * sms * blocks_per_sm / (best * 1e-3) / 1e12;
done:
    if (buf) cudaFree(buf);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

extern "C" int sfdtd_plan_destroy(sfdtd_plan *plan, void *cuda_stream) {
    if (!plan) return SFDTD_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != plan->dev) cudaSetDevice(plan->dev);
    if (plan->block) cudaFreeAsync(plan->block, (cudaStream_t)cuda_stream);
    if (plan->d_max) cudaFreeAsync(plan->d_max, (cudaStream_t)cuda_stream);
    if (plan->fork) cudaEventDestroy(plan->fork);
    for (cudaEvent_t e : plan->joins) cudaEventDestroy(e);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        std::vector<PoolStream> &fl = g_free_streams[plan->dev];
        for (size_t i = 0; i < plan->side.size(); i++) fl.push_back(PoolStream{plan->side[i], plan->side_prio[i]});
    }
    if (prev != plan->dev && prev >= 0) cudaSetDevice(prev);
    delete plan;
    return SFDTD_OK;
}

extern "C" int sfdtd_plan_create(const sfdtd_args *args, void *cuda_stream, sfdtd_plan **out) {
    g_err[0] = 0;
    if (!out) { snprintf(g_err, sizeof g_err, "plan pointer is NULL"); return SFDTD_ERR_ARG; }
    *out = nullptr;
    { const int v = validate(args); if (v != SFDTD_OK) return v; }
    const sfdtd_args &a = *args;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    int rc = SFDTD_OK;
    DeviceGuard guard;
    sfdtd_plan *P = new sfdtd_plan();
    int32_t *d_max = nullptr; float *d_est = nullptr;
    const int n_groups = (a.B + a.group_size - 1) / a.group_size;
    // prepass results in mapped pinned memory: h_max = [N_t maxima | N_l maxima], h_est, h_bow, h_ham
    const size_t Bp = ((size_t)a.B + 3) & ~(size_t)3;
    void *pin = nullptr;
    int32_t *h_max = nullptr; float *h_est = nullptr; uint8_t *h_bow = nullptr, *h_ham = nullptr;
    std::vector<int32_t> h_ids;
    std::vector<CtaDesc> h_ctas;
    // run-time knobs (read once per plan)
    // smallest longitudinal allocation class: a coarser one merges buckets (one work queue balances more strings) for more
    // shared memory per string; lane_div: longitudinal rows per lane that still count as a short loop
    const int wl_min = std::max(16, env_int("SFDTD_WLMIN", SFDTD_DEFAULT_WLMIN));
    const int lane_div = std::max(1, env_int("SFDTD_LANE_DIV", SFDTD_DEFAULT_LANE_DIV));
    const int min_lanes_env = env_int("SFDTD_MIN_LANES", 0);
    const bool use_queue = env_int("SFDTD_QUEUE", SFDTD_DEFAULT_QUEUE) != 0;
    const int tier = std::min(3, std::max(0, env_int("SFDTD_TIER", SFDTD_DEFAULT_TIER)));
    const int th_force = env_int("SFDTD_CTA_THREADS", 0);
    const int pad_smem = env_int("SFDTD_PAD_SMEM", 0);
    const int q_want = env_int("SFDTD_QSLICES", 8);
    const double tail_rounds = env_dbl("SFDTD_QTAIL", 1.0), min_rounds = env_dbl("SFDTD_QMIN", 2.0);
    const bool skip_aux = a.flags & SFDTD_SKIP_AUX;
    const bool manuf = a.flags & SFDTD_MANUFACTURED;
    struct IBucket { std::vector<int32_t> ids; };
    struct GBucket { std::vector<CtaDesc> ctas; size_t smem = 0; };
    std::map<std::pair<int, int>, IBucket> ib;                       // (config index, WLp) -> strings
    std::map<std::tuple<int, int, int, int>, GBucket> gb;            // (kind, cluster size, threads, smem class) -> CTAs
    KArgs K;
    fill_kargs(K, a);
    P->verbose = getenv("SFDTD_VERBOSE") != nullptr;
    P->B = a.B; P->group_size = a.group_size; P->Nt = a.Nt; P->Nx_t1 = a.Nx_t1; P->Nx_l1 = a.Nx_l1; P->n_groups = n_groups;
    P->flags = a.flags; P->dtype = a.dtype;
    const int dt = a.dtype == SFDTD_F32 ? 1 : 0, tsz = dt ? 4 : 8;
    const bool save_pf = (a.flags & SFDTD_SAVE_STATE) != 0;

    CK(guard.enter(a.state_u.ptr));
    CK(cudaGetDevice(&P->dev));
    CK(cudaDeviceGetAttribute(&P->n_sms, cudaDevAttrMultiProcessorCount, P->dev));
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!g_pool_ready[P->dev]) {
            // keep freed scratch in the device's default pool instead of returning it to the driver at every synchronisation
            cudaMemPool_t pool; uint64_t thr = UINT64_MAX;
            CK(cudaDeviceGetDefaultMemPool(&pool, P->dev));
            CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
            g_pool_ready[P->dev] = true;
        }
    }
    if (a.Nt <= 2) { *out = P; return SFDTD_OK; }

    // ---- prepass: per-string grid maxima and difficulty estimate; ONE device->host read ----
    CK(cudaMallocAsync((void **)&d_max, sizeof(int32_t) * 2 * (size_t)a.B, stream));
    CK(cudaMallocAsync((void **)&d_est, sizeof(float) * (size_t)a.B, stream));
    pin = pin_acquire(Bp * (2 * sizeof(int32_t) + sizeof(float) + 2));
    if (!pin) { snprintf(g_err, sizeof g_err, "cudaHostAlloc of the prepass result buffer failed"); rc = SFDTD_ERR_CUDA; goto done; }
    h_max = (int32_t *)pin; h_est = (float *)(h_max + 2 * Bp); h_bow = (uint8_t *)(h_est + Bp); h_ham = h_bow + Bp;
    if (dt) sfdtd_prepass_kernel<float><<<a.B, 128, 0, stream>>>(K, d_max, d_max + a.B, d_est, h_max, h_max + Bp, h_est, h_bow, h_ham);
    else sfdtd_prepass_kernel<double><<<a.B, 128, 0, stream>>>(K, d_max, d_max + a.B, d_est, h_max, h_max + Bp, h_est, h_bow, h_ham);
    g_launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(stream));

    // ---- bucketing ----
    for (int g = 0; g < n_groups; g++) {
        const int g0 = g * a.group_size, G = std::min(a.group_size, a.B - g0);
        bool forced = manuf;
        int Wt = 0;
        for (int s = 0; s < G; s++) {
            forced = forced || h_bow[g0 + s] || h_ham[g0 + s];
            Wt = std::max(Wt, h_max[g0 + s] + 1);
        }
        // rows a string can need: its own N_t + 3 (+ the bow window of a bowed string), never more than W_t;
        // every padded row is forced in the manufactured mode (string.cpp:227-232)
        auto rows_of = [&](int b) { return manuf ? Wt : std::min(Wt, h_max[b] + 3 + (h_bow[b] ? 8 : 0)); };
        // lanes that keep the longitudinal loops short
        auto lanes_of = [&](int b) { return std::max(std::min(32, wl_class(long_rows(h_max[Bp + b])) / lane_div), min_lanes_env); };
        if (!forced) {
            for (int s = 0; s < G; s++) {
                const int b = g0 + s, rows = rows_of(b), min_lanes = lanes_of(b);
                int pick = -1;
                for (int c = 0; c < N_CONFIGS && pick < 0; c++)
                    if (g_configs[c].tier == tier && g_configs[c].kern[dt] && rows <= g_configs[c].L * g_configs[c].ET && g_configs[c].L >= min_lanes) pick = c;
                if (pick < 0) {
                    snprintf(g_err, sizeof g_err, "string %d: %d transverse rows are outside the built kernel set", b, rows);
                    rc = SFDTD_ERR_UNSUPPORTED; goto done;
                }
                ib[{pick, wl_class(long_rows(h_max[Bp + b]), wl_min)}].ids.push_back(b);   // the allocation class may be coarser than the lane class
            }
            continue;
        }
        // forced group: one cluster; small strings on 16 lanes (8 per CTA), large ones on 32 lanes (4 per CTA)
        if (G > GB_MAX) {
            snprintf(g_err, sizeof g_err, "group %d: %d strings with bowed / hammered ones (grouped mode holds <= %d per group)", g, G, GB_MAX);
            rc = SFDTD_ERR_UNSUPPORTED; goto done;
        }
        int kind = manuf ? 2 : 0;
        std::vector<int> cls(G);
        for (int s = 0; s < G; s++) {
            const int rows = rows_of(g0 + s);
            if (rows > (manuf ? 256 : MAX_ROWS)) {
                snprintf(g_err, sizeof g_err, "string %d: %d transverse rows are outside the built kernel set", g0 + s, rows);
                rc = SFDTD_ERR_UNSUPPORTED; goto done;
            }
            if (rows > 128 && kind == 0) kind = 1;
            if (rows > 256) kind = 3;
            cls[s] = (rows > 64 || lanes_of(g0 + s) > 16) ? 1 : 0;
        }
        // slots per CTA of the one-shape kinds: 4, or fewer (64- / 32-thread CTAs) when the longitudinal blocks of four strings
        // do not fit one CTA's shared memory and the cluster still holds the group
        std::vector<CtaDesc> ctas;
        int threads = 128;
        size_t smem = 0;
        int CS = 0;
        for (int per_big = 4; per_big >= 1; per_big /= 2) {
            ctas.clear();
            for (int c = 0; c < 2; c++) {
                const int per = (kind != 0) ? per_big : (c == 0 ? 8 : 4);
                CtaDesc cd; memset(&cd, 0, sizeof cd);
                cd.group = g; cd.cls = (kind != 0) ? 0 : c; cd.G = G;
                for (int s = 0; s < G; s++) {
                    if (kind == 0 && cls[s] != c) continue;
                    if (kind != 0 && c == 1) continue;
                    cd.str[cd.n] = g0 + s; cd.gidx[cd.n] = s; cd.n++;
                    if (cd.n == per) { ctas.push_back(cd); cd.n = 0; }
                }
                if (cd.n) ctas.push_back(cd);
            }
            CS = (int)ctas.size();
            threads = (kind != 0) ? 32 * per_big : 128;
            if (CS == 1) { const GShape sh = g_gshape[kind][ctas[0].cls]; threads = std::min(threads, (ctas[0].n * sh.L + 31) / 32 * 32); }
            smem = 0;
            for (const CtaDesc &cd : ctas) {
                const GShape sh = g_gshape[kind][cd.cls];
                const int nslots = threads / sh.L;
                size_t dbl = GB_DOUBLES + xax_doubles(a.Nx_t1, true) + (size_t)nslots * slot_fixed_doubles(sh.L, sh.ET, true, tsz) + (size_t)(nslots + 2) / 2 + 2;
                for (int s = 0; s < nslots; s++) dbl += slot_long_doubles(long_rows(h_max[Bp + cd.str[std::min(s, cd.n - 1)]]), true, sh.L, tsz);
                smem = std::max(smem, dbl * sizeof(double) + 16);
            }
            if (kind == 0 || smem <= 227 * 1024 || per_big == 1) break;
            if ((G + per_big / 2 - 1) / (per_big / 2) > 8) break;       // the smaller CTAs would not fit one cluster
        }
        if (CS > 8) {
            snprintf(g_err, sizeof g_err, "group %d: needs %d CTAs (a cluster holds <= 8)", g, CS);
            rc = SFDTD_ERR_UNSUPPORTED; goto done;
        }
        if (smem > 227 * 1024) {
            snprintf(g_err, sizeof g_err, "group %d needs %zu bytes of shared memory per CTA (> 227 KB)", g, smem);
            rc = SFDTD_ERR_UNSUPPORTED; goto done;
        }
        // shared-memory class: the CTAs of a launch share one dynamic size, so a group with a few huge longitudinal grids
        // must not cost every other group its occupancy (3 CTAs per SM up to 74 KB, 2 up to 112 KB)
        const int sclass = smem <= 74 * 1024 ? 0 : (smem <= 112 * 1024 ? 1 : 2);
        GBucket &bk = gb[std::make_tuple(kind, CS, threads, sclass)];
        bk.ctas.insert(bk.ctas.end(), ctas.begin(), ctas.end());
        bk.smem = std::max(bk.smem, smem);
    }

    // ---- launches: independent buckets (smallest first: every CTA lives for the whole time loop, so a small bucket
    // started late would trail the bulk on a nearly empty GPU), then the grouped ones ----
    {
        typedef std::pair<const std::pair<int, int>, IBucket> IB;
        std::vector<IB *> order;
        for (auto &kv : ib) order.push_back(&kv);
        std::stable_sort(order.begin(), order.end(), [](const IB *x, const IB *y) { return x->second.ids.size() < y->second.ids.size(); });
        for (IB *pkv : order) {
            const Config &cf = g_configs[pkv->first.first];
            std::vector<int32_t> &ids = pkv->second.ids;
            // strings of similar nonlinearity share a warp (they converge in about the same number of sweeps); hardest first
            std::stable_sort(ids.begin(), ids.end(), [&](int32_t x, int32_t y) {
                const float ex = h_est[x], ey = h_est[y];
                return (ex == ex ? ex : INFINITY) > (ey == ey ? ey : INFINITY);
            });
            Launch ln;
            ln.cfg = pkv->first.first; ln.WLp = pkv->first.second; ln.n_items = (int)ids.size(); ln.off = h_ids.size();
            ln.need_xax = !skip_aux;
            h_ids.insert(h_ids.end(), ids.begin(), ids.end());
            // CTA size that keeps the most strings resident per SM (registers and shared memory both bound it)
            ln.fn = (save_pf && cf.kern_pf[dt]) ? cf.kern_pf[dt] : cf.kern[dt];
            const int regs = kernel_regs(ln.fn);
            int best = 32; long best_res = -1;
            for (int th : {128, 96, 64, 32}) {
                if (th % cf.L) continue;
                if (th_force && th != std::max(th_force, cf.L)) continue;
                const size_t b_ = smem_bytes_indep(cf, th / cf.L, a.Nx_t1, ln.WLp, ln.need_xax, tsz);
                if (b_ > 227 * 1024) continue;
                const long by_smem = (long)((227 * 1024) / (b_ + 1024)), by_regs = 65536 / ((long)regs * th);
                long res = std::min(std::min(by_smem, by_regs), 32L) * th;
                if (b_ > 76 * 1024 && th > 32) res /= 2;      // large CTAs cannot share an SM with other buckets: prefer smaller ones
                if (res > best_res) { best_res = res; best = th; }
            }
            ln.threads = best;
            const int per = ln.threads / cf.L;
            ln.grid = (ln.n_items + per - 1) / per;
            ln.smem = std::max(smem_bytes_indep(cf, per, a.Nx_t1, ln.WLp, ln.need_xax, tsz), (size_t)pad_smem);
            if (ln.smem > 227 * 1024) {
                snprintf(g_err, sizeof g_err, "a bucket (L=%d, ET=%d) needs %zu bytes of shared memory (> 227 KB)", cf.L, cf.ET, ln.smem);
                rc = SFDTD_ERR_UNSUPPORTED; goto done;
            }
            CK(cudaFuncSetAttribute(ln.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            // one shared-memory carve-out for every bucket kernel, so that CTAs of different buckets can share an SM
            CK(cudaFuncSetAttribute(ln.fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
            if (use_queue) {
                // persistent grid: what is resident at once; the warps pull their string sets from the bucket's counter
                int per_sm = 0;
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ln.fn, ln.threads, ln.smem));
                const int spw = 32 / cf.L, n_sets = (ln.n_items + spw - 1) / spw, wpc = ln.threads / 32;
                const int resident = std::max(1, per_sm) * P->n_sms;
                // tail group: the sets that would run in the last round of the resident warps, in SFDTD_QSLICES time slices
                // (default 8); a bucket with fewer than two rounds of sets is not sliced (it ends long before the bulk does)
                const int rw = resident * wpc;
                int n_sl = std::max(1, std::min(q_want, (a.Nt - 2) / 512));
                const int q_slice = ((a.Nt - 2 + n_sl - 1) / n_sl + TBS_MAX - 1) / TBS_MAX * TBS_MAX;
                n_sl = (a.Nt - 2 + q_slice - 1) / q_slice;
                const int n_tail = (n_sl > 1 && n_sets >= min_rounds * rw) ? std::max(0, std::min(n_sets, (int)(tail_rounds * rw))) : 0;
                ln.grid = std::min(std::min(ln.grid, resident), (n_sets + wpc - 1) / wpc);
                ln.queue = true; ln.q_slice = q_slice; ln.q_nslices = n_sl; ln.q_full = n_sets - n_tail;
                ln.q_idx = P->launches.size(); ln.done_off = P->queue_words; P->queue_words += n_sets;
            }
            if (P->verbose)
                fprintf(stderr, "[sfdtd] bucket L=%d ET=%d indep WLp=%d items=%d threads=%d grid=%d smem=%zu regs=%d\n", cf.L, cf.ET,
                        ln.WLp, ln.n_items, ln.threads, ln.grid, ln.smem, regs);
            P->launches.push_back(ln);
        }
        for (auto &kv : gb) {
            Launch ln;
            ln.grouped = true; ln.cfg = std::get<0>(kv.first); ln.cluster = std::get<1>(kv.first); ln.threads = std::get<2>(kv.first);
            ln.grid = (int)kv.second.ctas.size(); ln.n_items = ln.grid / ln.cluster; ln.off = h_ctas.size(); ln.need_xax = true;
            ln.smem = std::max(kv.second.smem, (size_t)pad_smem);
            h_ctas.insert(h_ctas.end(), kv.second.ctas.begin(), kv.second.ctas.end());
            ln.fn = (save_pf && g_group_kernels_pf[dt][ln.cfg]) ? g_group_kernels_pf[dt][ln.cfg] : g_group_kernels[dt][ln.cfg];
            CK(cudaFuncSetAttribute(ln.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            CK(cudaFuncSetAttribute(ln.fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
            if (P->verbose)
                fprintf(stderr, "[sfdtd] bucket grouped kind=%d cluster=%d groups=%d threads=%d grid=%d smem=%zu regs=%d\n", ln.cfg,
                        ln.cluster, ln.n_items, ln.threads, ln.grid, ln.smem, kernel_regs(ln.fn));
            P->launches.push_back(ln);
        }
    }
    P->n_buckets = P->launches.size();
    {
        // ---- device block: maxNl | ids | ctas | queue words (one counter per launch + one per set) | width table | u_H carry ----
        auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
        const size_t o_max = 0, o_ids = al(o_max + sizeof(int32_t) * a.B), o_cta = al(o_ids + sizeof(int32_t) * h_ids.size());
        const size_t o_q = al(o_cta + sizeof(CtaDesc) * h_ctas.size());
        P->queue_words += P->n_buckets;
        const size_t o_w = al(o_q + sizeof(int32_t) * P->queue_words), o_uh = al(o_w + sizeof(int32_t) * (size_t)n_groups * a.Nt);
        const size_t total = al(o_uh + sizeof(double) * 2 * (size_t)a.B);
        CK(cudaMallocAsync((void **)&P->block, total, stream));
        P->d_maxNl = (int32_t *)(P->block + o_max); P->d_ids = (int32_t *)(P->block + o_ids); P->d_ctas = (CtaDesc *)(P->block + o_cta);
        P->d_queue = (int32_t *)(P->block + o_q); P->d_wtab = (int32_t *)(P->block + o_w); P->d_uH = (double *)(P->block + o_uh);
        P->d_max = d_max; d_max = nullptr;           // (no device-to-device copy: it would queue on a copy engine like the read-backs)
        P->d_maxNl = P->d_max + a.B;
        if (!h_ids.empty()) CK(cudaMemcpyAsync(P->d_ids, h_ids.data(), sizeof(int32_t) * h_ids.size(), cudaMemcpyHostToDevice, stream));
        if (!h_ctas.empty()) CK(cudaMemcpyAsync(P->d_ctas, h_ctas.data(), sizeof(CtaDesc) * h_ctas.size(), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));        // the host vectors go out of scope; still before any stepper kernel is queued
        CK(cudaEventCreateWithFlags(&P->fork, cudaEventDisableTiming));
        for (size_t i = 0; i < P->n_buckets; i++) {
            cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            P->joins.push_back(e);
        }
        if (P->n_buckets > 1) {
            std::lock_guard<std::mutex> lk(g_mu);
            std::vector<PoolStream> &fl = g_free_streams[P->dev];
            // Stream priorities by bucket size, smallest bucket highest: the persistent grid of ONE large bucket fills every SM
            // (3 CTAs x 128 threads x 168 registers), so a small bucket whose CTAs lose the start-up race against it only runs
            // after it has finished and lengthens the call by its whole duration (measured: +190 ms of 1470 ms whenever the
            // hardware picked the large buckets first, which depended on nothing but the order the streams were created in).
            int pr_least = 0, pr_greatest = 0;
            CK(cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest));
            while (P->side.size() < P->n_buckets) {
                // (the largest bucket lowest, the next one level above it, ...; the smallest ones share the top level)
                const int prio = std::max(pr_greatest, pr_least - (int)(P->n_buckets - 1 - P->side.size()));
                // an IDLE released stream of that priority if there is one (a released stream may still run the call of its last
                // plan: a new call queued behind it would serialise two independent calls), else a new one; beyond 256 streams reuse
                cudaStream_t st = nullptr;
                for (size_t q = 0; q < fl.size() && !st; q++)
                    if (fl[q].prio == prio && cudaStreamQuery(fl[q].s) == cudaSuccess) { st = fl[q].s; fl.erase(fl.begin() + q); }
                cudaGetLastError();
                if (!st && fl.size() >= 256)
                    for (size_t q = 0; q < fl.size() && !st; q++)
                        if (fl[q].prio == prio) { st = fl[q].s; fl.erase(fl.begin() + q); }
                if (!st) CK(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio));
                P->side.push_back(st); P->side_prio.push_back(prio);
            }
        }
    }
done:
    pin_release(pin);
    if (d_max) cudaFreeAsync(d_max, stream);
    if (d_est) cudaFreeAsync(d_est, stream);
    if (rc != SFDTD_OK) { sfdtd_plan_destroy(P, stream); return rc; }
    *out = P;
    return rc;
}

extern "C" int sfdtd_forward_plan(sfdtd_plan *P, const sfdtd_args *args, void *cuda_stream) {
    g_err[0] = 0;
    if (!P) { snprintf(g_err, sizeof g_err, "plan is NULL"); return SFDTD_ERR_ARG; }
    { const int v = validate(args); if (v != SFDTD_OK) return v; }
    const sfdtd_args &a = *args;
    if (a.B != P->B || a.group_size != P->group_size || a.Nt != P->Nt || a.Nx_t1 != P->Nx_t1 || a.Nx_l1 != P->Nx_l1 || a.flags != P->flags ||
        a.dtype != P->dtype) {
        snprintf(g_err, sizeof g_err, "args do not match the plan (B, group_size, Nt, Nx_t1, Nx_l1, flags, dtype)");
        return SFDTD_ERR_ARG;
    }
    if (a.Nt <= 2) return SFDTD_OK;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    int rc = SFDTD_OK;
    DeviceGuard guard;
    KArgs K;
    fill_kargs(K, a);
    const int dt = a.dtype == SFDTD_F32 ? 1 : 0;
    std::vector<cudaEvent_t> t0, t1;
    {
    CK(guard.enter(a.state_u.ptr));
    sfdtd_width_kernel<<<dim3((a.Nt + 127) / 128, P->n_groups), 128, 0, stream>>>(K, P->d_wtab);
    g_launches++;
    CK(cudaGetLastError());
    K.Wtab = P->d_wtab; K.maxNl = P->d_maxNl; K.uH_carry = P->d_uH;
    if (P->queue_words) {
        sfdtd_zero_kernel<<<(unsigned)((P->queue_words + 255) / 256), 256, 0, stream>>>(P->d_queue, P->queue_words);
        g_launches++;
        CK(cudaGetLastError());
    }
    const size_t nb = P->n_buckets;
    const std::vector<cudaStream_t> &ss = P->side;
    if (nb > 1) CK(cudaEventRecord(P->fork, stream));
    for (size_t bi = 0; bi < nb; bi++) {
        const Launch &ln = P->launches[bi];
        cudaStream_t s = (nb > 1) ? ss[bi] : stream;
        if (nb > 1) CK(cudaStreamWaitEvent(s, P->fork, 0));
        K.n_items = ln.n_items; K.WLp = ln.WLp; K.need_xax = ln.need_xax ? 1 : 0;
        K.n_lo = 2; K.n_hi = a.Nt;
        K.queue = nullptr; K.done = nullptr; K.ids = nullptr; K.ctas = nullptr;
        if (P->verbose) {
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            t0.push_back(e0); t1.push_back(e1);
            CK(cudaEventRecord(e0, s));
        }
        if (!ln.grouped) {
            K.ids = P->d_ids + ln.off;
            if (ln.queue) {
                K.queue = P->d_queue + ln.q_idx; K.done = P->d_queue + P->n_buckets + ln.done_off;
                K.q_slice = ln.q_slice; K.q_nslices = ln.q_nslices; K.q_full = ln.q_full;
            }
            ln.fn<<<(unsigned)ln.grid, ln.threads, ln.smem, s>>>(K);
        } else {
            K.ctas = P->d_ctas + ln.off;
            cudaLaunchConfig_t lc; memset(&lc, 0, sizeof lc);
            lc.gridDim = dim3((unsigned)ln.grid); lc.blockDim = dim3((unsigned)ln.threads); lc.dynamicSmemBytes = ln.smem; lc.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)ln.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            lc.attrs = at; lc.numAttrs = 1;
            CK(cudaLaunchKernelEx(&lc, ln.fn, K));
        }
        g_launches++;
        CK(cudaGetLastError());
        if (P->verbose) CK(cudaEventRecord(t1.back(), s));
        if (nb > 1) { CK(cudaEventRecord(P->joins[bi], s)); CK(cudaStreamWaitEvent(stream, P->joins[bi], 0)); }
    }
    }
    if (P->verbose) {
        CK(cudaStreamSynchronize(stream));
        for (size_t bi = 0; bi < t0.size(); bi++) {
            float ms = 0, ms0 = 0;
            cudaEventElapsedTime(&ms, t0[bi], t1[bi]); cudaEventElapsedTime(&ms0, t0[0], t1[bi]);
            const Launch &ln = P->launches[bi];
            fprintf(stderr, "[sfdtd] timing bucket %zu (%s cfg=%d WLp=%d cluster=%d items=%d grid=%d x %d): %.1f ms (ends at %.1f ms)\n", bi,
                    ln.grouped ? "grouped" : "indep", ln.cfg, ln.WLp, ln.cluster, ln.n_items, ln.grid, ln.threads, ms, ms0);
        }
    }
done:
    for (cudaEvent_t e : t0) cudaEventDestroy(e);
    for (cudaEvent_t e : t1) cudaEventDestroy(e);
    return rc;
}

extern "C" int sfdtd_forward(const sfdtd_args *args, void *cuda_stream) {
    g_err[0] = 0;
    { const int v = validate(args); if (v != SFDTD_OK) return v; }
    if (args->Nt <= 2) return SFDTD_OK;
    sfdtd_plan *P = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    int rc = sfdtd_plan_create(args, cuda_stream, &P);
    if (rc != SFDTD_OK) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    rc = sfdtd_forward_plan(P, args, cuda_stream);
    if (getenv("SFDTD_VERBOSE"))
        fprintf(stderr, "[sfdtd] host: plan %.2f ms, launch %.2f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    char keep[sizeof g_err]; memcpy(keep, g_err, sizeof keep);
    sfdtd_plan_destroy(P, cuda_stream);               // stream-ordered: the scratch is released after the kernels
    memcpy(g_err, keep, sizeof keep);
    return rc;
}

extern "C" int sfdtd_synth_controls(const sfdtd_args *args, const sfdtd_array *f0, const sfdtd_array *x_b, const sfdtd_array *v_b,
                                    const sfdtd_array *F_b, const sfdtd_array *u_H, void *cuda_stream) {
    g_err[0] = 0;
    if (!args || !args->synth) { snprintf(g_err, sizeof g_err, "args->synth is NULL"); return SFDTD_ERR_ARG; }
    if (args->B <= 0 || args->Nt <= 0) { snprintf(g_err, sizeof g_err, "bad sizes"); return SFDTD_ERR_ARG; }
    int rc = SFDTD_OK;
    DeviceGuard guard;
    KArgs K;
    fill_kargs(K, *args);
    const sfdtd_array none = {nullptr, 0, 0};
    CK(guard.enter(args->synth->f0_a));
    sfdtd_synth_controls_kernel<<<dim3((args->Nt + 127) / 128, args->B), 128, 0, (cudaStream_t)cuda_stream>>>(
        K, f0 ? *f0 : none, x_b ? *x_b : none, v_b ? *v_b : none, F_b ? *F_b : none, u_H ? *u_H : none);
    g_launches++;
    CK(cudaGetLastError());
done:
    return rc;
}

namespace {
int postprocess_impl(int dtype, const sfdtd_array *uout, const sfdtd_array *zout, int32_t B, int32_t n0, int32_t n_samples,
                     double silence_db, int32_t normalize, int32_t bits, int64_t pcm_pitch, uint8_t *is_nan,
                     uint8_t *is_silent, double *gain, void *pcm_u, void *pcm_z, void *pcm_w, void *cuda_stream) {
    g_err[0] = 0;
    if (!uout || !zout || !uout->ptr || !zout->ptr || B <= 0 || n_samples <= 0 || n0 < 0 || (bits != 16 && bits != 24)) {
        snprintf(g_err, sizeof g_err, "bad arguments"); return SFDTD_ERR_ARG;
    }
    int rc = SFDTD_OK;
    DeviceGuard guard;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    uint8_t *pu = (uint8_t *)pcm_u, *pz = (uint8_t *)pcm_z, *pw = (uint8_t *)pcm_w;
    CK(guard.enter(uout->ptr));
#define PP_LAUNCH(BITS_, T_) sfdtd_postprocess_kernel<BITS_, T_><<<B, 256, 0, st>>>(*uout, *zout, n0, n_samples, silence_db, normalize, \
                                                                                    is_nan, is_silent, gain, pu, pz, pw, pcm_pitch)
    if (dtype == SFDTD_F32) { if (bits == 16) PP_LAUNCH(16, float); else PP_LAUNCH(24, float); }
    else { if (bits == 16) PP_LAUNCH(16, double); else PP_LAUNCH(24, double); }
#undef PP_LAUNCH
    g_launches++;
    CK(cudaGetLastError());
done:
    return rc;
}
}  // namespace

extern "C" int sfdtd_postprocess(const sfdtd_array *uout, const sfdtd_array *zout, int32_t B, int32_t n0, int32_t n_samples,
                                 double silence_db, int32_t normalize, int32_t bits, int64_t pcm_pitch, uint8_t *is_nan,
                                 uint8_t *is_silent, double *gain, void *pcm_u, void *pcm_z, void *pcm_w, void *cuda_stream) {
    return postprocess_impl(SFDTD_F64, uout, zout, B, n0, n_samples, silence_db, normalize, bits, pcm_pitch, is_nan, is_silent, gain,
                            pcm_u, pcm_z, pcm_w, cuda_stream);
}
extern "C" int sfdtd_postprocess_f32(const sfdtd_array *uout, const sfdtd_array *zout, int32_t B, int32_t n0, int32_t n_samples,
                                     double silence_db, int32_t normalize, int32_t bits, int64_t pcm_pitch, uint8_t *is_nan,
                                     uint8_t *is_silent, double *gain, void *pcm_u, void *pcm_z, void *pcm_w, void *cuda_stream) {
    return postprocess_impl(SFDTD_F32, uout, zout, B, n0, n_samples, silence_db, normalize, bits, pcm_pitch, is_nan, is_silent, gain,
                            pcm_u, pcm_z, pcm_w, cuda_stream);
}
