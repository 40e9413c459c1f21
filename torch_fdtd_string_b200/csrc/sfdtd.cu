// sfdtd.cu -- B200 (sm_100a) time-loop-fused StringFDTD stepper behind the C ABI of include/sfdtd.h.
//
// One CTA owns one "group" (= one reference batch: the strings that share the batch-max operator
// widths, reference misc.cpp:119-127, and the any-over-batch convergence votes, string.cpp:252-253,
// hammer.cpp:51).  Inside the CTA every string is owned by L lanes of a warp for the whole run:
//   * transverse block: blocked layout, ET consecutive grid rows per lane, u^{n-1}, u^{n-2} and all
//     per-step vectors live in registers; stencil halos move with warp shuffles;
//   * the implicit system  [A11 K_tl; K_lt A22] w = -RHS  (string.cpp:162-181,238) is solved
//     matrix-free: block Gauss-Seidel over the transverse/longitudinal blocks; A11 (tridiagonal, varies
//     with Lambda(u^{n-1})) by a register-resident partitioned Thomas factorisation (local LU of the
//     ET-1 interior rows per lane + parallel cyclic reduction over the L interface rows via shuffles);
//     A22 (constant-coefficient, off/diag ~1e-5) by Jacobi sweeps folded into the same iteration;
//   * longitudinal block: shared memory, rows distributed cyclically over the string's lanes; the
//     linear interpolation operators Int_tl / Int_lt (misc.cpp:78-105) are gathers from shared memory;
//   * per-step scalars (grid sizes, loss, tolerances; string.cpp:16-41,96-120) are computed L steps at
//     a time, one time step per lane, into a shared-memory table; outputs are staged there too and
//     flushed as coalesced 128-byte rows.
// No tensor cores: no step is a dense contraction.  HBM traffic: controls in, audio out.
//
// Reference citations are to /root/reference/src/model/cpp/*.cpp (see DESIGN.md for the derivation).

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <vector>
#include <map>
#include <algorithm>
#include <atomic>

#include "sfdtd.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define FULLMASK 0xffffffffu

namespace {

constexpr int WL_MARGIN = 2;        // ghost rows of the longitudinal block kept beyond N_l (decay (e/d)^m, e/d ~ 1e-5)
constexpr int NV = 20;              // doubles per time step in the per-string scalar table
constexpr int NOUT = 5;             // staged outputs per step
constexpr int GS_CAP = 200;         // cap on block Gauss-Seidel sweeps per solve
constexpr double GS_TOL = 1e-13;    // relative max-norm change that ends the sweeps

// table slots
enum { T_NT = 0, T_NL, T_HT, T_HL, T_S0K, T_S1K, T_G, T_PHI, T_KK, T_TOLT, T_TOLL, T_XB, T_VB, T_FB, T_WID, T_UHPRE, T_SIG0, T_SIG1 };

struct KArgs {
    sfdtd_args a;
    double k, k2, k4, th, omth, tt1, tt2, lamc, order, mhd;
    const int32_t *maxNt, *maxNl;   // per string, over this call (prepass)
    const int32_t *group_ids;       // groups handled by this launch
    int32_t max_iter;
};

__device__ __forceinline__ double ldx(const sfdtd_array &A, int b, int n) {
    return ((const double *)A.ptr)[(int64_t)b * A.bs + (int64_t)n * A.ts];
}
__device__ __forceinline__ double lds(const sfdtd_array &A, int b) {
    return ((const double *)A.ptr)[(int64_t)b * A.bs];
}

// ---- get_derived_vars (string.cpp:16-41), reference operation order, no FMA contraction ----------
struct Derived { double gamma, K, Nt, ht, Nl, hl; };
__device__ __forceinline__ Derived derive(double f0, double kappa_rel, double alpha, const KArgs &A) {
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

// ---- prepass: per string, the largest N_t / N_l any step of this call can see (at min f0) ----------
__global__ void sfdtd_prepass_kernel(const __grid_constant__ KArgs A, int32_t *maxNt, int32_t *maxNl) {
    const int b = blockIdx.x;
    const int Nt = A.a.Nt;
    double fm = INFINITY;
    if (A.a.f0.ts == 0) {
        fm = ldx(A.a.f0, b, 0);
    } else {
        for (int n = 2 + threadIdx.x; n < Nt; n += blockDim.x) { const double v = ldx(A.a.f0, b, n); fm = v < fm ? v : fm; }
    }
    for (int o = 16; o > 0; o >>= 1) { const double v = __shfl_xor_sync(FULLMASK, fm, o); fm = v < fm ? v : fm; }
    __shared__ double sm[32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = fm;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (blockDim.x >> 5); w++) fm = fm < sm[w] ? fm : sm[w];
        Derived d = derive(fm, lds(A.a.kappa, b), lds(A.a.alpha, b), A);
        double nt = d.Nt, nl = d.Nl;
        if (!(nt >= 0)) nt = 0; if (!(nl >= 0)) nl = 0;
        if (nt > 1e6) nt = 1e6; if (nl > 1e6) nl = 1e6;
        maxNt[b] = (int32_t)nt; maxNl[b] = (int32_t)nl;
    }
}

// ---- warp helpers over the L lanes of one string -------------------------------------------------
template <int L> __device__ __forceinline__ double shup(double v, int d) { return __shfl_up_sync(FULLMASK, v, d, L); }
template <int L> __device__ __forceinline__ double shdn(double v, int d) { return __shfl_down_sync(FULLMASK, v, d, L); }
template <int L> __device__ __forceinline__ double shix(double v, int s) { return __shfl_sync(FULLMASK, v, s, L); }
template <int L> __device__ __forceinline__ double red_sum(double v) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o, L);
    return v;
}
template <int L> __device__ __forceinline__ float red_maxf(float v) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULLMASK, v, o, L));
    return v;
}
template <int L> __device__ __forceinline__ int red_or(int v) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v |= __shfl_xor_sync(FULLMASK, v, o, L);
    return v;
}
template <int L> constexpr int ilog2() { return L <= 1 ? 0 : 1 + ilog2<L / 2>(); }

__device__ __forceinline__ double nan0(double v) {   // nan_to_num (string.cpp:225-226)
    if (v != v) return 0.0;
    if (isinf(v)) return v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    return v;
}

// ---- partitioned Thomas: local LU of the ET-1 interior rows + PCR over the L interface rows -------
template <int L, int ET> struct TriSolver {
    static constexpr int M = ET - 1;
    static constexpr int LV = ilog2<L>();
    double inv[M], lw[M], cp[M], V[M], W[M];
    double ae, ce, k1[LV], k2[LV], invB;

    __device__ __forceinline__ void factor(const double (&a)[ET], const double (&b)[ET], const double (&c)[ET], int ln) {
        double cprev = 0.0;
#pragma unroll
        for (int r = 0; r < M; r++) {
            const double den = b[r] - a[r] * cprev;
            inv[r] = __drcp_rn(den);
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
        for (int r = M - 2; r >= 0; r--) { V[r] = V[r] - cp[r] * V[r + 1]; W[r] = -cp[r] * W[r + 1]; }
        ae = a[ET - 1]; ce = c[ET - 1];
        const double Vn0 = shdn<L>(V[0], 1), Wn0 = shdn<L>(W[0], 1);
        double Ar = -ae * V[M - 1];
        double Br = b[ET - 1] - ae * W[M - 1] - ce * Vn0;
        double Cr = -ce * Wn0;
        if (ln == 0) Ar = 0.0;
        if (ln == L - 1) { Cr = 0.0; Br = b[ET - 1] - ae * W[M - 1]; }
#pragma unroll
        for (int lv = 0; lv < LV; lv++) {
            const int s = 1 << lv;
            const double iB = __drcp_rn(Br);
            const double iBm = shup<L>(iB, s), iBp = shdn<L>(iB, s);
            const double Am = shup<L>(Ar, s), Cm = shup<L>(Cr, s);
            const double Ap = shdn<L>(Ar, s), Cp = shdn<L>(Cr, s);
            const bool hm = ln >= s, hp = ln + s < L;
            const double q1 = hm ? Ar * iBm : 0.0, q2 = hp ? Cr * iBp : 0.0;
            k1[lv] = q1; k2[lv] = q2;
            Br = Br - (hm ? Cm * q1 : 0.0) - (hp ? Ap * q2 : 0.0);
            Ar = hm ? -Am * q1 : 0.0;
            Cr = hp ? -Cp * q2 : 0.0;
        }
        invB = __drcp_rn(Br);
    }

    // d: right-hand side in, solution out
    __device__ __forceinline__ void solve(double (&d)[ET], int ln) const {
        double Y[M];
        Y[0] = d[0] * inv[0];
#pragma unroll
        for (int r = 1; r < M; r++) Y[r] = d[r] * inv[r] - lw[r] * Y[r - 1];
#pragma unroll
        for (int r = M - 2; r >= 0; r--) Y[r] = Y[r] - cp[r] * Y[r + 1];
        double Yn0 = shdn<L>(Y[0], 1);
        if (ln == L - 1) Yn0 = 0.0;
        double D = d[ET - 1] - ae * Y[M - 1] - ce * Yn0;
#pragma unroll
        for (int lv = 0; lv < LV; lv++) {
            const int s = 1 << lv;
            const double Dm = shup<L>(D, s), Dp = shdn<L>(D, s);
            D = D - k1[lv] * Dm - k2[lv] * Dp;     // k1/k2 are 0 where the neighbour does not exist
        }
        const double xe = D * invB;
        double p = shup<L>(xe, 1);
        if (ln == 0) p = 0.0;
#pragma unroll
        for (int r = 0; r < M; r++) d[r] = Y[r] - V[r] * p - W[r] * xe;
        d[ET - 1] = xe;
    }
};

// float32 linear-interpolation row (misc.cpp:78-105; F.interpolate(..., 'linear', align_corners=True) on float32)
__device__ __forceinline__ void interp_row(float s, int o, int in_last, int &i0, int &i1, double &w0, double &w1) {
    const float r = __fmul_rn(s, (float)o);
    int a0 = (int)r;
    a0 = a0 > in_last ? in_last : a0;
    i0 = a0; i1 = a0 + (a0 < in_last ? 1 : 0);
    const float l1 = __fsub_rn(r, (float)a0);
    const float l0 = __fsub_rn(1.0f, l1);
    w0 = (double)l0; w1 = (double)l1;
}

// select element `slot` of a register array without dynamic indexing
template <int ET> __device__ __forceinline__ double pick(const double (&v)[ET], int slot) {
    double o = 0.0;
#pragma unroll
    for (int r = 0; r < ET; r++) o = (r == slot) ? v[r] : o;
    return o;
}
template <int L, int ET> __device__ __forceinline__ double fetch_row(const double (&v)[ET], int idx) {
    idx = idx < 0 ? 0 : (idx > L * ET - 1 ? L * ET - 1 : idx);
    const double mine = pick<ET>(v, idx % ET);
    return shix<L>(mine, idx / ET);
}

// ======================================================================================================
template <int L, int ET, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) sfdtd_step_kernel(const __grid_constant__ KArgs A) {
    constexpr int TB = L;   // time steps per scalar-table block
    extern __shared__ double smem[];
    const sfdtd_args &a = A.a;
    const int tid = threadIdx.x;
    const int sl = tid / L, ln = tid % L;
    const int nslots = blockDim.x / L;
    const int gid = A.group_ids[blockIdx.x];
    const int g0 = gid * a.group_size;
    const int G = min(a.group_size, a.B - g0);
    const bool valid = sl < G;
    const int b = g0 + (valid ? sl : G - 1);     // spare slots shadow the last string and never write or vote
    const int Nt = a.Nt, NXT = a.Nx_t1, NXL = a.Nx_l1;
    const bool surf = a.flags & SFDTD_SURFACE_INTEGRAL;
    const bool save_state = a.flags & SFDTD_SAVE_STATE;
    const bool skip_aux = a.flags & SFDTD_SKIP_AUX;
    const double kk = A.k, k2 = A.k2;
    uint32_t status = 0;

    // ---- shared memory carve-up ----
    const int WLa = A.maxNl[b] + 1 + WL_MARGIN;              // rows allocated for this string's longitudinal block
    int *ioffs = (int *)smem;                                // [nslots+1] l-block offsets (doubles)
    int *gN = ioffs + (nslots + 2);                          // [2][TB][nslots]
    int *gW = gN + 2 * TB * nslots;                          // [2][TB]
    float *xaxs = (float *)(gW + 2 * TB);                    // [NXT]
    size_t cur_off = ((size_t)((char *)(xaxs + NXT) - (char *)smem) + 7) / 8;
    double *tab_all = smem + cur_off; cur_off += (size_t)nslots * TB * NV;
    double *ost_all = smem + cur_off; cur_off += (size_t)nslots * TB * (NOUT + 1);
    double *qs_all = smem + cur_off;  cur_off += (size_t)nslots * (L * ET + 2);
    double *lblk_all = smem + cur_off;
    if (ln == 0) ioffs[sl + 1] = 6 * (WLa + 2);
    for (int i = tid; i < NXT; i += blockDim.x) xaxs[i] = a.xax[i];
    __syncthreads();
    if (tid == 0) { ioffs[0] = 0; for (int s = 0; s < nslots; s++) ioffs[s + 1] += ioffs[s]; }
    __syncthreads();
    double *tab = tab_all + (size_t)sl * TB * NV;
    double *ost = ost_all + (size_t)sl * TB * (NOUT + 1);
    double *qs = qs_all + (size_t)sl * (L * ET + 2);
    double *lb = lblk_all + ioffs[sl];
    const int WLp = WLa + 2;
    double *Z1 = lb, *Z2 = lb + WLp, *ZP = lb + 2 * WLp, *ZA = lb + 3 * WLp, *ZB = lb + 4 * WLp, *RL = lb + 5 * WLp;

    // ---- per-string constants ----
    const double kappa_rel = lds(a.kappa, b), alpha = lds(a.alpha, b), rp = lds(a.pos, b);
    const double *T60 = (const double *)a.T60.ptr + (int64_t)b * a.T60.bs;
    const double T00 = T60[0], T01 = T60[1], T10 = T60[2], T11 = T60[3];
    const double phi0 = lds(a.phi_0, b), phi1 = lds(a.phi_1, b);
    const double xH = lds(a.x_H, b), aH = lds(a.alpha_H, b);
    const double wH = lds(a.w_H, b) / A.lamc, Mr = lds(a.M_r, b) / A.lamc;
    const double wpow = pow(wH, 1.0 + aH);
    const bool bowm = a.bow_mask[b] != 0, hamm = a.hammer_mask[b] != 0;
    const double bm = bowm ? 1.0 : 0.0, hm = hamm ? 1.0 : 0.0;
    const bool forced = bowm || hamm;
    const int group_has_hammer = __syncthreads_or(valid && hamm);
    const int group_has_bow = __syncthreads_or(valid && bowm);
    // CTA-uniform compute switches (they guard shuffles and barriers); per-string output switches
    const bool do_bow = group_has_bow || !skip_aux, do_ham = group_has_hammer || !skip_aux;
    const bool out_bow = bowm || !skip_aux, out_ham = hamm || !skip_aux;
    const double alpha2 = alpha * alpha;

    // ---- initial state: rows n-2, n-1 ----
    double u1[ET], u2[ET];
    {
        const double *su = (const double *)a.state_u.ptr + (int64_t)b * a.state_u.bs;
#pragma unroll
        for (int r = 0; r < ET; r++) {
            const int i = ln * ET + r;
            u2[r] = (i < NXT) ? su[i] : 0.0;
            u1[r] = (i < NXT) ? su[a.state_u.ts + i] : 0.0;
        }
        const double *sz = (const double *)a.state_z.ptr + (int64_t)b * a.state_z.bs;
        for (int j = ln; j < WLp; j += L) {
            Z2[j] = (j < NXL && j < WLa) ? sz[j] : 0.0;
            Z1[j] = (j < NXL && j < WLa) ? sz[a.state_z.ts + j] : 0.0;
            ZP[j] = 0.0; ZA[j] = 0.0; ZB[j] = 0.0; RL[j] = 0.0;
        }
    }
    double uH1 = 0.0, uH2 = 0.0;
    if (Nt > 2) { uH2 = ldx(a.u_H, b, 0); uH1 = ldx(a.u_H, b, 1); }
    double sig0_last = 0.0, sig1_last = 0.0;
    int64_t cnt_outer = 0, cnt_sweeps = 0, cnt_ham = 0, cnt_steps = 0;
    __syncwarp();

    for (int n0 = 2; n0 < Nt; n0 += TB) {
        // ================= scalar table for steps n0 .. n0+TB-1 (one step per lane) =================
        {
            const int n = n0 + ln;
            double *t = tab + ln * NV;
            int iNt = 0, iNl = 0;
            if (n < Nt) {
                const double f0 = ldx(a.f0, b, n);
                const Derived d = derive(f0, kappa_rel, alpha, A);
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
                const double g = g2 * k2;
                t[T_NT] = d.Nt; t[T_NL] = d.Nl; t[T_HT] = d.ht; t[T_HL] = d.hl;
                t[T_S0K] = (2 * sig0) * kk; t[T_S1K] = (2 * sig1) * kk;
                t[T_G] = g; t[T_PHI] = (g * (alpha2 - 1)) / 4; t[T_KK] = (d.K * d.K) * k2;
                t[T_TOLT] = pow(d.ht, A.order); t[T_TOLL] = pow(d.hl, A.order);
                t[T_XB] = ldx(a.x_b, b, n); t[T_VB] = ldx(a.v_b, b, n); t[T_FB] = ldx(a.F_b, b, n); t[T_WID] = ldx(a.wid, b, n);
                t[T_UHPRE] = ldx(a.u_H, b, n);
                t[T_SIG0] = sig0; t[T_SIG1] = sig1;
                iNt = (int)fmin(fmax(d.Nt, 0.0), 1e6); iNl = (int)fmin(fmax(d.Nl, 0.0), 1e6);
            }
            gN[(0 * TB + ln) * nslots + sl] = valid ? iNt : 0;
            gN[(1 * TB + ln) * nslots + sl] = valid ? iNl : 0;
        }
        __syncthreads();
        for (int q = tid; q < 2 * TB; q += blockDim.x) {
            int mx = 0;
            for (int s = 0; s < nslots; s++) mx = max(mx, gN[q * nslots + s]);
            gW[q] = mx + 1;                         // W_t / W_l = batch-max width (misc.cpp:119-127)
        }
        __syncthreads();

        const int jmax = min(TB, Nt - n0);
        for (int jj = 0; jj < jmax; jj++) {
            const int n = n0 + jj;
            const double *t = tab + jj * NV;
            int Wt = gW[jj], Wl = gW[TB + jj];
            const int N_t = (int)t[T_NT], N_l = (int)t[T_NL];
            const double ht = t[T_HT], iht = t[T_NT], ihl = t[T_NL];
            const double s0k = t[T_S0K], s1k = t[T_S1K], g = t[T_G], phi = t[T_PHI], Kk = t[T_KK];
            const double tol_t = t[T_TOLT], tol_l = t[T_TOLL];
            if (Wt > L * ET) { Wt = L * ET; status |= SFDTD_ST_RANGE; }
            int WLs = min(N_l + 1 + WL_MARGIN, Wl);
            if (WLs > WLa) { WLs = WLa; status |= SFDTD_ST_RANGE; }
            const int keep_flat = N_t + N_l + 2;            // string.cpp:233
            const int keep_l = keep_flat - NXT;             // l rows j < keep_l keep their base RHS
            const double iht2 = iht * iht, iht4 = iht2 * iht2, ihl2 = ihl * ihl;
            const double diagA = A.th + s0k + 2 * s1k * iht2, offA = 0.5 * A.omth - s1k * iht2;
            const double diagC = A.th - s0k - 2 * s1k * iht2, offC = 0.5 * A.omth + s1k * iht2;
            const double kh4 = Kk * iht4;
            const double diagB = -2 * A.th + 2 * g * iht2 + 6 * kh4, off1B = -A.omth - g * iht2 - 4 * kh4, off2B = kh4;
            const double ph2 = phi * iht2;
            const double dA = (1 + s0k) + 2 * s1k * ihl2, eA = -s1k * ihl2, idA = 1.0 / dA;
            const bool coupled = (phi != 0.0);
            const float s_tl = (N_t > 0) ? __fdiv_rn((float)N_l, (float)N_t) : 0.0f;    // t-row -> l-grid
            const float s_lt = (N_l > 0) ? __fdiv_rn((float)N_t, (float)N_l) : 0.0f;    // l-row -> t-grid

            // ---- masked previous states (mask_1d, string.cpp:129-132) and halos ----
            double m1[ET], m2[ET];
#pragma unroll
            for (int r = 0; r < ET; r++) {
                const int i = ln * ET + r;
                m1[r] = (i <= N_t) ? u1[r] : u1[r] * 0.0;
                m2[r] = (i <= N_t) ? u2[r] : u2[r] * 0.0;
            }
            double e1[ET + 4];
            {
                double l2 = shup<L>(m1[ET - 2], 1), l1 = shup<L>(m1[ET - 1], 1);
                double r0 = shdn<L>(m1[0], 1), r1 = shdn<L>(m1[1], 1);
                if (ln == 0) { l2 = 0.0; l1 = 0.0; }
                if (ln == L - 1) { r0 = 0.0; r1 = 0.0; }
                e1[0] = l2; e1[1] = l1; e1[ET + 2] = r0; e1[ET + 3] = r1;
#pragma unroll
                for (int r = 0; r < ET; r++) e1[r + 2] = m1[r];
            }
            double m2l = shup<L>(m2[ET - 1], 1), m2r = shdn<L>(m2[0], 1);
            if (ln == 0) m2l = 0.0;
            if (ln == L - 1) m2r = 0.0;
            // Lambda = Dxb u1 (string.cpp:152), rows < W_t
            double lam[ET + 1];
#pragma unroll
            for (int r = 0; r < ET; r++) {
                const int i = ln * ET + r;
                lam[r] = (i < Wt) ? (e1[r + 2] - e1[r + 1]) * iht : 0.0;
            }
            lam[ET] = shdn<L>(lam[0], 1);
            if (ln == L - 1) lam[ET] = 0.0;

            // t-row interpolation parameters onto the l grid (Int_tl)
            int ti0[ET], ti1[ET]; float tw0[ET], tw1[ET];
#pragma unroll
            for (int r = 0; r < ET; r++) {
                const int i = ln * ET + r;
                double w0, w1;
                interp_row(s_tl, i, N_l, ti0[r], ti1[r], w0, w1);
                tw0[r] = (float)w0; tw1[r] = (float)w1;
                if (i > N_t) { tw0[r] = 0.0f; tw1[r] = 0.0f; ti0[r] = 0; ti1[r] = 0; }
                ti0[r] = min(ti0[r], WLp - 1); ti1[r] = min(ti1[r], WLp - 1);
            }

            // ---- A11 (string.cpp:153-162) ----
            double ca[ET], cb[ET], cc[ET];
#pragma unroll
            for (int r = 0; r < ET; r++) {
                const int i = ln * ET + r;
                const double l2 = lam[r] * lam[r], lp2 = lam[r + 1] * lam[r + 1];
                const bool in = i < Wt;
                ca[r] = (in && i > 0) ? offA - ph2 * l2 : 0.0;
                cc[r] = (in && i + 1 < Wt) ? offA - ph2 * lp2 : 0.0;
                cb[r] = in ? diagA + ph2 * (l2 + lp2) : 1.0;
            }
            TriSolver<L, ET> ts;
            ts.factor(ca, cb, cc, ln);

            // ---- base RHS  B w1 + C w2  (string.cpp:223-224) ----
            double rt[ET];
            // K_tl (2 z1 + z2): stage zz in ZA
            for (int j = ln; j < WLp; j += L) ZA[j] = (j <= N_l && j < WLa) ? 2.0 * Z1[j] + Z2[j] : 0.0;
            __syncwarp();
            {
                double y[ET];
#pragma unroll
                for (int r = 0; r < ET; r++) y[r] = coupled ? (double)tw0[r] * ZA[ti0[r]] + (double)tw1[r] * ZA[ti1[r]] : 0.0;
                double yl = shup<L>(y[ET - 1], 1);
                if (ln == 0) yl = 0.0;
                double q[ET + 1];
#pragma unroll
                for (int r = 0; r < ET; r++) q[r] = lam[r] * ((y[r] - (r == 0 ? yl : y[r - 1])) * iht);
                q[ET] = shdn<L>(q[0], 1);
                if (ln == L - 1) q[ET] = 0.0;
#pragma unroll
                for (int r = 0; r < ET; r++) {
                    const int i = ln * ET + r;
                    const double l2 = lam[r] * lam[r], lp2 = lam[r + 1] * lam[r + 1];
                    double d4 = diagB;
                    if (i == 1 || i == N_t - 1) d4 += kh4;          // Dxxxx_clamped (misc.cpp:146-163)
                    const double Bu = d4 * e1[r + 2] + off1B * (e1[r + 1] + e1[r + 3]) + off2B * (e1[r] + e1[r + 4]);
                    const double m2m = (r == 0) ? m2l : m2[r - 1], m2p = (r == ET - 1) ? m2r : m2[r + 1];
                    const double Cu = (diagC + ph2 * (l2 + lp2)) * m2[r] + (offC - ph2 * l2) * m2m + (offC - ph2 * lp2) * m2p;
                    const double Kz = -phi * ((q[r + 1] - q[r]) * iht);
                    rt[r] = (i < Wt && i < keep_flat) ? (Bu + Cu + Kz) : 0.0;
                }
            }
            // l-block base RHS, only when the flat-index mask leaves any of it (string.cpp:233)
            const bool has_rl = keep_l > 0;
            if (__any_sync(FULLMASK, has_rl)) {
                // q2 = Lam Dxb u2 -> smem, then K_lt u2 by gathers
#pragma unroll
                for (int r = 0; r < ET; r++) qs[ln * ET + r] = lam[r] * ((m2[r] - (r == 0 ? m2l : m2[r - 1])) * iht);
                __syncwarp();
                const double dB = -2 + 2 * (g * alpha2) * ihl2, eB = -(g * alpha2) * ihl2;
                const double dC = (1 - s0k) - 2 * s1k * ihl2, eC = s1k * ihl2;
                for (int j = ln; j < WLs; j += L) {
                    double v = 0.0;
                    if (has_rl && j < keep_l) {
                        const double z1c = (j <= N_l) ? Z1[j] : 0.0, z2c = (j <= N_l) ? Z2[j] : 0.0;
                        const double z1l = (j > 0 && j - 1 <= N_l) ? Z1[j - 1] : 0.0, z1r = (j + 1 <= N_l && j + 1 < WLa) ? Z1[j + 1] : 0.0;
                        const double z2l = (j > 0 && j - 1 <= N_l) ? Z2[j - 1] : 0.0, z2r = (j + 1 <= N_l && j + 1 < WLa) ? Z2[j + 1] : 0.0;
                        double pj = 0.0, pj1 = 0.0;
                        if (coupled) {
                            int i0, i1; double w0, w1;
                            if (j <= N_l) { interp_row(s_lt, j, N_t, i0, i1, w0, w1); pj = w0 * qs[i0] + w1 * qs[i1]; }
                            if (j + 1 <= N_l) { interp_row(s_lt, j + 1, N_t, i0, i1, w0, w1); pj1 = w0 * qs[i0] + w1 * qs[i1]; }
                        }
                        v = dB * z1c + eB * (z1l + z1r) + dC * z2c + eC * (z2l + z2r) - phi * ((pj1 - pj) * ihl);
                    }
                    RL[j] = v;
                }
                __syncwarp();
            }

            // ---- bow: raised-cosine weights over the Nx_t1-point axis (bow.cpp:32, misc.cpp:20-34) ----
            double rc[ET];
            double rc_extra = 0.0;
            const double vB = t[T_VB], FB = t[T_FB];
            if (do_bow) {
                const double Nd = (double)NXT;
                const double ctr = __ddiv_rn(__dmul_rn(t[T_XB], (double)(N_t - 1)), Nd);
                const double wid = __ddiv_rn(__dmul_rn(__dmul_rn(t[T_WID], ht), (double)(N_t - 1)), Nd);
                const double hw = wid * 0.5;
                int ic = (int)floor((ctr - hw) * Nd) - 2;
                ic = ic < 0 ? 0 : ic;
                const int i = ic + ln;
                double o = 0.0;
                if (i < NXT) {
                    const double x = (double)xaxs[i];
                    const double dm = __dsub_rn(__dsub_rn(x, ctr), hw), dp = __dadd_rn(__dsub_rn(x, ctr), hw);
                    const double p = __dmul_rn(-dm, dp);
                    if (p > 0) o = 0.5 * (1 + cos(((2 * M_PI) * (x - ctr)) / wid));
                    else if (p != p) o = p;
                }
                if (ln == L - 1 && o != 0.0) status |= SFDTD_ST_BOW_WINDOW;
                const double S = red_sum<L>(fabs(o));
                o = o / S;                               // 0/0 -> NaN like the reference
                rc_extra = red_sum<L>((i >= L * ET && i < NXT) ? o : 0.0);
#pragma unroll
                for (int r = 0; r < ET; r++) {
                    const int w = ln * ET + r - ic;
                    const double v = shix<L>(o, w & (L - 1));
                    rc[r] = (w >= 0 && w < L) ? v : ((S == 0.0 || S != S) ? v * 0.0 : 0.0);
                }
            } else {
#pragma unroll
                for (int r = 0; r < ET; r++) rc[r] = 0.0;
            }

            // ---- hammer: contact point and relative displacements (hammer.cpp:70-74) ----
            const int idxH = (int)floor(__dmul_rn(xH, (double)(N_t - 1)));
            double eta1 = 0.0, eta2 = 0.0, r1pow = 0.0;
            if (do_ham) {
                eta1 = uH1 - fetch_row<L, ET>(m1, idxH);
                eta2 = uH2 - fetch_row<L, ET>(m2, idxH);
                const double r1 = eta1 > 0 ? eta1 : (eta1 != eta1 ? eta1 : 0.0);
                const double ex = aH - 1.0;
                r1pow = (ex == 2.0) ? r1 * r1 : ((ex == 0.0) ? 1.0 : pow(r1, ex));
            }

            // ---- fixed-point loop over the forcing (string.cpp:200-258) ----
            double uit[ET], nu[ET], xs[ET];
#pragma unroll
            for (int r = 0; r < ET; r++) { uit[r] = u1[r]; xs[r] = m1[r]; nu[r] = 0.0; }
            __syncwarp();      // all gathers of the staged 2 z1 + z2 are done before ZA is reused
            for (int j = ln; j < WLp; j += L) { ZP[j] = Z1[j]; ZA[j] = (j <= N_l) ? Z1[j] : 0.0; }
            __syncwarp();
            int zc = 0;                       // current GS z buffer: 0 -> ZA, 1 -> ZB
            double vrel = 0.0, FH = 0.0, uH = 0.0;
            int iter = 0;
            bool solved = false;
            while (true) {
                // bow force (bow.cpp:35-40)
                double hb = 0.0;
                if (do_bow) {
                    double acc = 0.0;
#pragma unroll
                    for (int r = 0; r < ET; r++) {
                        const double dd = (iter == 0) ? (m1[r] - m2[r]) : (uit[r] - m1[r]);
                        acc += rc[r] * (dd / kk - vB);
                    }
                    vrel = red_sum<L>(acc) + rc_extra * (0.0 / kk - vB);
                    const double sg = (vrel > 0) ? 1.0 : ((vrel < 0) ? -1.0 : 0.0);
                    hb = (vrel != vrel) ? vrel : sg * (phi1 + (1 - phi1) * exp(-phi0 * fabs(vrel)));
                }
                // hammer loop (hammer.cpp:28-53), votes over the group
                if (do_ham) {
                    const double eps_u = fetch_row<L, ET>(uit, idxH);
                    double eta_est = eta1 * hm;
                    int hit = 0, more;
                    do {
                        const double eta = eta_est;
                        const double fH = ((wpow * r1pow) * (eta + eta2)) / 2;
                        FH = (eta1 > 0) ? fH : 0.0;
                        double v = ((2 * uH1) - uH2) - k2 * FH;
                        double tt = v - A.mhd;
                        tt = tt > 0 ? tt : (tt != tt ? tt : 0.0);
                        uH = tt + A.mhd;
                        eta_est = (uH - eps_u) * hm;
                        const int nc = fabs(eta - eta_est) > tol_t;
                        hit++;
                        more = group_has_hammer ? __syncthreads_or(valid && nc) : nc;
                        if (hit >= A.max_iter) { if (more) status |= SFDTD_ST_HAMMER_CAP; more = 0; }
                    } while (more);
                    cnt_ham += hit;
                }
                // ---- linear solve  A w = -(RHS)  ----
                const bool need = !solved || forced;
                if (__any_sync(FULLMASK, need)) {
                    double mr[ET];
                    const double sB = -k2 * (FB * hb) * iht;
                    const double sH = hamm ? nan0(-k2 * (Mr * FH)) : 0.0;
#pragma unroll
                    for (int r = 0; r < ET; r++) {
                        const int i = ln * ET + r;
                        double f = 0.0;
                        if (bowm) f += nan0(sB * rc[r]);
                        if (hamm && i == idxH) f += sH;
                        mr[r] = (i < Wt && i < keep_flat) ? rt[r] + f : 0.0;
                    }
                    // block Gauss-Seidel: u <- A11^-1(-r_t - K_tl z) ; z <- Jacobi(A22, -r_l - K_lt u)
                    bool conv = !need;
                    int sweeps = 0;
                    do {
                        const bool act = !conv;
                        const double *zcur = zc ? ZB : ZA;
                        double *znew = zc ? ZA : ZB;
                        double d[ET];
                        {
                            double y[ET];
#pragma unroll
                            for (int r = 0; r < ET; r++) y[r] = coupled ? (double)tw0[r] * zcur[ti0[r]] + (double)tw1[r] * zcur[ti1[r]] : 0.0;
                            double yl = shup<L>(y[ET - 1], 1);
                            if (ln == 0) yl = 0.0;
                            double q[ET + 1];
#pragma unroll
                            for (int r = 0; r < ET; r++) q[r] = lam[r] * ((y[r] - (r == 0 ? yl : y[r - 1])) * iht);
                            q[ET] = shdn<L>(q[0], 1);
                            if (ln == L - 1) q[ET] = 0.0;
#pragma unroll
                            for (int r = 0; r < ET; r++) d[r] = -mr[r] + phi * ((q[r + 1] - q[r]) * iht);
                        }
                        ts.solve(d, ln);
                        float du = 0.f, su = 0.f;
#pragma unroll
                        for (int r = 0; r < ET; r++) {
                            du = fmaxf(du, (float)fabs(d[r] - xs[r]));
                            su = fmaxf(su, (float)fabs(d[r]));
                            if (act) xs[r] = d[r];
                        }
                        float dz = 0.f, sz = 0.f;
                        {
                            double xl = shup<L>(xs[ET - 1], 1);
                            if (ln == 0) xl = 0.0;
                            if (act) {
#pragma unroll
                                for (int r = 0; r < ET; r++) qs[ln * ET + r] = lam[r] * ((xs[r] - (r == 0 ? xl : xs[r - 1])) * iht);
                            }
                            __syncwarp();
                            if (act) {
                                for (int j = ln; j < WLs; j += L) {
                                    double pj = 0.0, pj1 = 0.0;
                                    if (coupled) {
                                        int i0, i1; double w0, w1;
                                        if (j <= N_l) { interp_row(s_lt, j, N_t, i0, i1, w0, w1); pj = w0 * qs[i0] + w1 * qs[i1]; }
                                        if (j + 1 <= N_l) { interp_row(s_lt, j + 1, N_t, i0, i1, w0, w1); pj1 = w0 * qs[i0] + w1 * qs[i1]; }
                                    }
                                    const double rhs = -((has_rl && j < keep_l) ? RL[j] : 0.0) + phi * ((pj1 - pj) * ihl);
                                    const double zl = (j > 0) ? zcur[j - 1] : 0.0, zr = (j + 1 < WLs) ? zcur[j + 1] : 0.0;
                                    const double zn = (rhs - eA * (zl + zr)) * idA;
                                    dz = fmaxf(dz, (float)fabs(zn - zcur[j]));
                                    sz = fmaxf(sz, (float)fabs(zn));
                                    znew[j] = zn;
                                }
                                for (int j = WLs + ln; j < WLp; j += L) znew[j] = 0.0;
                            }
                            __syncwarp();
                            if (act) zc ^= 1;
                        }
                        du = red_maxf<L>(du); su = red_maxf<L>(su); dz = red_maxf<L>(dz); sz = red_maxf<L>(sz);
                        sweeps++;
                        bool ok = !(du > (float)GS_TOL * su) && !(dz > (float)GS_TOL * sz);
                        if (!coupled && !has_rl) ok = true;              // uncoupled and no l-RHS: the first solve is exact
                        if (act && sweeps >= GS_CAP && !ok) { status |= SFDTD_ST_SOLVER_CAP; ok = true; }
                        if (act) { cnt_sweeps += 1; conv = ok; }
                    } while (__any_sync(FULLMASK, !conv));
                    solved = true;
                }
                // ---- mask + Dirichlet (string.cpp:240-246), residuals (string.cpp:248-253) ----
                int nc_t = 0, nan_u = 0;
#pragma unroll
                for (int r = 0; r < ET; r++) {
                    const int i = ln * ET + r;
                    const bool keep = (i <= N_t) && (i != 0) && (i != N_t) && (i < Wt);
                    nu[r] = keep ? xs[r] : xs[r] * 0.0;
                    const double df = fabs(uit[r] - nu[r]);
                    nan_u |= (df != df);
                    nc_t |= (df > tol_t);
                    uit[r] = nu[r];
                }
                const double *zsol = zc ? ZB : ZA;
                int nc_l = 0, nan_z = 0;
                for (int j = ln; j < WLp; j += L) {
                    const bool keep = (j <= N_l) && (j != 0) && (j != N_l) && (j < WLs);
                    const double zv = keep ? zsol[j] : zsol[j] * 0.0;
                    const double df = fabs(ZP[j] - zv);
                    nan_z |= (df != df);
                    nc_l |= (df > tol_l);
                    ZP[j] = zv;
                }
                __syncwarp();
                nan_u = red_or<L>(nan_u); nan_z = red_or<L>(nan_z);
                nc_t = red_or<L>(nc_t);
                nc_l = red_or<L>(nc_l);
                const int not_conv = (nc_t && !nan_u) || (nc_l && !nan_z);
                iter++;
                int more = __syncthreads_or(valid && not_conv);
                if (iter >= A.max_iter) { if (more) status |= SFDTD_ST_OUTER_CAP; more = 0; }
                if (!more) break;
            }
            cnt_outer += iter; cnt_steps += 1;

            // ---- save and readout (string.cpp:263-303) ----
            double uo, zo;
            if (surf) {
                const double rw = 0.5 * ht;
                const double wgt = rw * 1.0 + rw * hm + rw * bm;
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < ET; r++) acc += ((nu[r] - u1[r]) * wgt) / kk;
                uo = red_sum<L>(acc);
                acc = 0.0;
                for (int j = ln; j < WLp; j += L) acc += ((ZP[j] - Z1[j]) * wgt) / kk;
                zo = red_sum<L>(acc);
            } else {
                const int ui = 1 + (int)floor(__dmul_rn((double)N_t, rp));
                const double uf = 1 + rp / ht - (double)ui;
                const int zi = 1 + (int)floor(__dmul_rn((double)N_l, rp));
                const double zf = 1 + rp / t[T_HL] - (double)zi;
                const double ua = fetch_row<L, ET>(nu, ui), ub = fetch_row<L, ET>(nu, ui + 1);
                uo = (1 - uf) * ua + uf * ub;
                const double za = (zi < WLp) ? ZP[zi] : 0.0, zb = (zi + 1 < WLp) ? ZP[zi + 1] : 0.0;
                zo = (1 - zf) * za + zf * zb;
            }
            // state rows: state[:, n] += u  (in place, onto pre-loaded content; string.cpp:264-265)
            {
                double *su = (double *)a.state_u.ptr + (int64_t)b * a.state_u.bs + (int64_t)n * a.state_u.ts;
#pragma unroll
                for (int r = 0; r < ET; r++) {
                    const int i = ln * ET + r;
                    double row = nu[r];
                    if (save_state && i < NXT) { row += su[i]; if (valid) su[i] = row; }
                    u2[r] = u1[r]; u1[r] = row;
                }
                double *sz = (double *)a.state_z.ptr + (int64_t)b * a.state_z.bs + (int64_t)n * a.state_z.ts;
                double *Zn = Z2;                       // recycle the oldest row buffer
                for (int j = ln; j < WLp; j += L) {
                    double row = ZP[j];
                    if (save_state && j < NXL && j < WLa) { row += sz[j]; if (valid) sz[j] = row; }
                    Zn[j] = row;
                }
                Z2 = Z1; Z1 = Zn;
                __syncwarp();
            }
            const double uHtot = t[T_UHPRE] + (out_ham ? uH : 0.0);
            uH2 = uH1; uH1 = uHtot;
            sig0_last = t[T_SIG0]; sig1_last = t[T_SIG1];
            if (ln == 0) {
                double *o = ost + jj * (NOUT + 1);
                o[0] = uo; o[1] = zo; o[2] = out_bow ? vrel : 0.0; o[3] = out_ham ? FH : 0.0; o[4] = uHtot;
            }
        }
        // ---- flush staged outputs: lane j writes step n0+j (coalesced rows) ----
        __syncwarp();
        if (valid && ln < jmax) {
            const int n = n0 + ln;
            const double *o = ost + ln * (NOUT + 1);
            ((double *)a.uout.ptr)[(int64_t)b * a.uout.bs + (int64_t)n * a.uout.ts] = o[0];
            ((double *)a.zout.ptr)[(int64_t)b * a.zout.bs + (int64_t)n * a.zout.ts] = o[1];
            ((double *)a.v_r.ptr)[(int64_t)b * a.v_r.bs + (int64_t)n * a.v_r.ts] = o[2];
            ((double *)a.F_H.ptr)[(int64_t)b * a.F_H.bs + (int64_t)n * a.F_H.ts] = o[3];
            ((double *)a.u_H.ptr)[(int64_t)b * a.u_H.bs + (int64_t)n * a.u_H.ts] = o[4];
            ((double *)a.u_H_out.ptr)[(int64_t)b * a.u_H_out.bs + (int64_t)n * a.u_H_out.ts] = o[4] / kk;
        }
        __syncthreads();
    }

    // ---- epilogue ----
    const uint32_t status_all = (uint32_t)red_or<L>((int)status);
    if (valid) {
        if (!save_state && Nt > 2) {
            double *su = (double *)a.state_u.ptr + (int64_t)b * a.state_u.bs;
#pragma unroll
            for (int r = 0; r < ET; r++) {
                const int i = ln * ET + r;
                if (i < NXT) { su[i] = u2[r]; su[a.state_u.ts + i] = u1[r]; }
            }
            double *sz = (double *)a.state_z.ptr + (int64_t)b * a.state_z.bs;
            for (int j = ln; j < WLp; j += L) if (j < NXL && j < WLa) { sz[j] = Z2[j]; sz[a.state_z.ts + j] = Z1[j]; }
        }
        // u_H_out / u_H columns 0,1 (simulator.cpp:57 divides the whole tensor)
        if (ln < 2 && ln < Nt) {
            ((double *)a.u_H_out.ptr)[(int64_t)b * a.u_H_out.bs + (int64_t)ln * a.u_H_out.ts] = ldx(a.u_H, b, ln) / kk;
        }
        if (ln == 0) {
            if (Nt > 2) { ((double *)a.sig0)[b] = sig0_last; ((double *)a.sig1)[b] = sig1_last; }
            if (a.status) a.status[b] = status_all;
            if (a.counters) {
                a.counters[4 * b + 0] = cnt_outer; a.counters[4 * b + 1] = cnt_sweeps;
                a.counters[4 * b + 2] = cnt_ham; a.counters[4 * b + 3] = cnt_steps;
            }
        }
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

struct Config { int L, ET, MAXT; void (*kern)(const KArgs); };
#define CFG(L_, ET_, MT_) Config{L_, ET_, MT_, sfdtd_step_kernel<L_, ET_, MT_>}
// smallest first; a group needs  W_t <= L*ET  and  ceil32(G*L) <= MAXT
const Config g_configs[] = {
    CFG(16, 6, 128), CFG(16, 6, 384), CFG(16, 6, 1024),
    CFG(32, 4, 256), CFG(32, 4, 1024),
    CFG(32, 8, 128), CFG(32, 8, 512),
};
constexpr int N_CONFIGS = sizeof(g_configs) / sizeof(g_configs[0]);

size_t smem_bytes(const Config &c, int nslots, int NXT, size_t lblk_doubles) {
    const int TB = c.L;
    size_t bytes = sizeof(int) * ((nslots + 2) + 2 * TB * nslots + 2 * TB) + sizeof(float) * NXT;
    bytes = (bytes + 7) / 8 * 8;
    bytes += sizeof(double) * ((size_t)nslots * TB * NV + (size_t)nslots * TB * (NOUT + 1) + (size_t)nslots * (c.L * c.ET + 2) + lblk_doubles);
    return bytes + 64;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { snprintf(g_err, sizeof g_err, "%s: %s", #x, cudaGetErrorString(e_)); rc = SFDTD_ERR_CUDA; goto done; } } while (0)

}  // namespace

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

extern "C" int sfdtd_forward(const sfdtd_args *args, void *cuda_stream) {
    g_err[0] = 0;
    if (!args) { snprintf(g_err, sizeof g_err, "args is NULL"); return SFDTD_ERR_ARG; }
    const sfdtd_args &a = *args;
    if (a.abi_version != SFDTD_ABI_VERSION) { snprintf(g_err, sizeof g_err, "abi_version %d != %d", a.abi_version, SFDTD_ABI_VERSION); return SFDTD_ERR_ARG; }
    if (a.dtype != SFDTD_F64) { snprintf(g_err, sizeof g_err, "only SFDTD_F64 is built"); return SFDTD_ERR_UNSUPPORTED; }
    if (a.flags & SFDTD_MANUFACTURED) { snprintf(g_err, sizeof g_err, "SFDTD_MANUFACTURED is not built yet"); return SFDTD_ERR_UNSUPPORTED; }
    if (a.B <= 0 || a.group_size <= 0 || a.Nt < 0 || a.Nx_t1 <= 0 || a.Nx_l1 <= 0) { snprintf(g_err, sizeof g_err, "bad sizes"); return SFDTD_ERR_ARG; }
    const void *req[] = {a.state_u.ptr, a.state_z.ptr, a.kappa.ptr, a.alpha.ptr, a.f0.ptr, a.pos.ptr, a.T60.ptr, a.x_b.ptr, a.v_b.ptr,
                         a.F_b.ptr, a.wid.ptr, a.phi_0.ptr, a.phi_1.ptr, a.x_H.ptr, a.w_H.ptr, a.M_r.ptr, a.alpha_H.ptr, a.u_H.ptr,
                         a.bow_mask, a.hammer_mask, a.xax, a.uout.ptr, a.zout.ptr, a.v_r.ptr, a.F_H.ptr, a.u_H_out.ptr, a.sig0, a.sig1};
    for (const void *p : req) if (!p) { snprintf(g_err, sizeof g_err, "a required pointer is NULL"); return SFDTD_ERR_ARG; }
    if (a.Nt <= 2) return SFDTD_OK;

    cudaStream_t stream = (cudaStream_t)cuda_stream;
    int rc = SFDTD_OK;
    int32_t *d_max = nullptr, *d_gids = nullptr;
    std::vector<int32_t> h_max(2 * (size_t)a.B);
    const int n_groups = (a.B + a.group_size - 1) / a.group_size;
    std::map<int, std::vector<int32_t>> buckets;      // config index -> group ids
    std::map<int, size_t> bucket_smem;
    std::vector<int32_t> h_gids;

    KArgs K;
    memset(&K, 0, sizeof K);
    K.a = a;
    K.k = (double)a.k; K.k2 = pow((double)a.k, 2.); K.k4 = pow((double)a.k, 4.);
    K.th = (double)a.theta_t;
    { const float om = 1 - a.theta_t; K.omth = (double)om; }                 // float32 (string.cpp:148)
    { const float t1 = 2 * a.theta_t - 1; const float t2 = 2 * t1; K.tt1 = (double)t1; K.tt2 = (double)t2; }   // string.cpp:30-31
    K.lamc = (double)a.lambda_c; K.order = (double)a.relative_order;
    K.mhd = (double)(-0.01f);                                                // hammer.cpp:3
    K.max_iter = a.max_iter > 0 ? a.max_iter : 1000;

    CK(cudaMalloc(&d_max, sizeof(int32_t) * 2 * (size_t)a.B));
    K.maxNt = d_max; K.maxNl = d_max + a.B;
    sfdtd_prepass_kernel<<<a.B, 128, 0, stream>>>(K, d_max, d_max + a.B);
    g_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_max.data(), d_max, sizeof(int32_t) * 2 * (size_t)a.B, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));

    for (int g = 0; g < n_groups; g++) {
        const int g0 = g * a.group_size, G = std::min(a.group_size, a.B - g0);
        int Wt = 0; size_t lblk = 0;
        for (int s = 0; s < G; s++) {
            Wt = std::max(Wt, h_max[g0 + s] + 1);
            lblk += 6 * (size_t)(h_max[a.B + g0 + s] + 1 + WL_MARGIN + 2);
        }
        int pick = -1;
        for (int c = 0; c < N_CONFIGS && pick < 0; c++) {
            const Config &cf = g_configs[c];
            const int threads = (a.group_size * cf.L + 31) / 32 * 32;
            if (Wt <= cf.L * cf.ET && threads <= cf.MAXT) pick = c;
        }
        if (pick < 0) {
            snprintf(g_err, sizeof g_err, "group %d: W_t=%d with %d strings is outside the built kernel set", g, Wt, G);
            rc = SFDTD_ERR_UNSUPPORTED; goto done;
        }
        const Config &cf = g_configs[pick];
        const int threads = (a.group_size * cf.L + 31) / 32 * 32;
        // spare slots shadow the last string: account for their l-blocks too
        const int nslots = threads / cf.L;
        const size_t last = 6 * (size_t)(h_max[a.B + g0 + G - 1] + 1 + WL_MARGIN + 2);
        const size_t sm = smem_bytes(cf, nslots, a.Nx_t1, lblk + (size_t)(nslots - G) * last);
        if (sm > 227 * 1024) {
            snprintf(g_err, sizeof g_err, "group %d needs %zu bytes of shared memory (> 227 KB)", g, sm);
            rc = SFDTD_ERR_UNSUPPORTED; goto done;
        }
        buckets[pick * 4096 + threads].push_back(g);
        bucket_smem[pick * 4096 + threads] = std::max(bucket_smem[pick * 4096 + threads], sm);
    }
    for (auto &kv : buckets) h_gids.insert(h_gids.end(), kv.second.begin(), kv.second.end());
    CK(cudaMalloc(&d_gids, sizeof(int32_t) * h_gids.size()));
    CK(cudaMemcpyAsync(d_gids, h_gids.data(), sizeof(int32_t) * h_gids.size(), cudaMemcpyHostToDevice, stream));
    {
        size_t off = 0;
        for (auto &kv : buckets) {
            const Config &cf = g_configs[kv.first / 4096];
            const int threads = kv.first % 4096;
            const size_t sm = bucket_smem[kv.first];
            CK(cudaFuncSetAttribute(cf.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            K.group_ids = d_gids + off;
            cf.kern<<<(unsigned)kv.second.size(), threads, sm, stream>>>(K);
            g_launches++;
            CK(cudaGetLastError());
            off += kv.second.size();
        }
    }
    CK(cudaStreamSynchronize(stream));
done:
    if (d_max) cudaFree(d_max);
    if (d_gids) cudaFree(d_gids);
    return rc;
}
