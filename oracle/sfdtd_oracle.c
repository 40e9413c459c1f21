/*
 * sfdtd_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or
 * executed from the product path (torch_fdtd_string_b200/); only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may call it, and only
 * as the checker.
 *
 * Plain-C, CPU, fp64 restatement of the reference's batched time stepper
 *   forward_fn   (reference src/model/cpp/simulator.cpp:14-59)
 *   string_step  (reference src/model/cpp/string.cpp:43-306)
 * and the helpers they use (bow.cpp, hammer.cpp, misc.cpp, vnv.cpp).  The dense
 * (B, Nx_t1+Nx_l1, Nx_t1+Nx_l1) operators of the reference are restated
 * matrix-free per string; the linear solve is a dense LU with partial pivoting on
 * the (W_t+W_l) unpadded block system -- numerically equivalent to the reference's
 * linalg_inv + matmul (string.cpp:175,238).
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks this file against golden
 * vectors produced by the compiled, unmodified reference (tests/golden/make_golden.py).
 *
 * Each function cites the reference lines it follows.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct {
    /* sizes */
    int32_t B, Nt, Nx_t1, Nx_l1;
    /* in/out states, contiguous (B, Nt, Nx_t1) / (B, Nt, Nx_l1) */
    double *state_u, *state_z;
    /* string params: kappa(B) alpha(B) p_a(B) f0(B,Nt) pos(B) T60(B,2,2) */
    const double *kappa, *alpha, *p_a, *f0, *pos, *T60;
    /* bow params: x_b v_b F_b wid (B,Nt); phi_0 phi_1 (B) */
    const double *x_b, *v_b, *F_b, *wid, *phi_0, *phi_1;
    /* hammer params: x_H w_H M_r alpha_H (B); u_H (B,Nt) in/out */
    const double *x_H, *w_H, *M_r, *alpha_H;
    double *u_H;
    const uint8_t *bow_mask, *hammer_mask; /* (B) */
    /* constants: float32 on purpose (reference simulator.cpp:22, string.cpp:80) */
    float k, theta_t, lambda_c, relative_order;
    int32_t surface_integral, manufactured, n_0;
    /* outputs (B,Nt): columns 0,1 stay zero; sig0,sig1 (B) */
    double *uout, *zout, *v_r, *F_H, *u_H_out, *sig0, *sig1;
    /* diagnostics (may be NULL): iteration counters [outer_total, outer_max,
       hammer_total, hammer_max, steps] */
    int64_t *stats;
    int32_t max_iter; /* safety cap on the (uncapped in the reference) loops */
    /* 0: dense LU (faithful to inv+matmul).  1: matrix-free block Gauss-Seidel on the
       t/l blocks with Thomas solves -- the algorithm of the CUDA kernel, kept here so
       that it can be validated against the goldens on the CPU.
       stats[5]=sweeps total, stats[6]=sweeps max, stats[7]=solves */
    int32_t solver;
} sfdtd_oracle_args;

/* ---- per-string, per-step derived quantities --------------------------------
 * get_derived_vars (reference string.cpp:16-41) and loss parameters
 * (string.cpp:96-120).  Operation order and the float32 sub-expressions
 * (2*theta_t-1), 2*(2*theta_t-1) follow the reference.                         */
typedef struct {
    double gamma, K, h_t, h_l, sig0, sig1, tol_t, tol_l;
    int N_t, N_l;
} derived_t;

static void derive(const sfdtd_oracle_args *a, int b, int n, derived_t *d)
{
    const double k = (double)a->k;
    const double k2 = pow(k, 2.), k4 = pow(k, 4.);
    const float tt1f = 2 * a->theta_t - 1;          /* float32 */
    const float tt2f = 2 * tt1f;                    /* float32 */
    const double tt1 = (double)tt1f, tt2 = (double)tt2f;
    const double lam = (double)a->lambda_c;
    const double f0 = a->f0[(size_t)b * a->Nt + n];
    const double gamma = 2 * f0;
    const double kappa = gamma * a->kappa[b];
    const double t0 = (M_PI * kappa) / gamma;
    const double IHP = t0 * t0;
    const double K = sqrt(IHP) * (gamma / M_PI);
    const double g2 = gamma * gamma, g4 = pow(gamma, 4.);
    const double h1 = lam * sqrt((g2 * k2 + sqrt(g4 * k4 + ((16 * (K * K)) * k2) * tt1)) / tt2);
    const double N_t = floor(1 / h1);
    const double h2 = ((lam * gamma) * a->alpha[b]) * k;
    const double N_l = floor(1 / h2);
    d->gamma = gamma; d->K = K;
    d->N_t = (int)N_t; d->h_t = 1 / N_t;
    d->N_l = (int)N_l; d->h_l = 1 / N_l;
    d->tol_t = pow(d->h_t, (double)a->relative_order);
    d->tol_l = pow(d->h_l, (double)a->relative_order);
    /* loss (string.cpp:100-120); T60[b] = [[f1,t1],[f2,t2]] */
    const double *T = a->T60 + (size_t)b * 4;
    const double T00 = T[0], T01 = T[1], T10 = T[2], T11 = T[3];
    double z1, z2;
    if (K > 0) {
        double w1 = (2 * M_PI) * T00, w2 = (2 * M_PI) * T10;
        z1 = -g2 + sqrt(g4 + (4 * (K * K)) * (w1 * w1));
        z2 = -g2 + sqrt(g4 + (4 * (K * K)) * (w2 * w2));
    } else {
        z1 = (T00 * T00) / g2;
        z2 = (T10 * T10) / g2;
    }
    const int m = (T00 * T01 * T10 * T11) != 0;
    double s0 = m ? (-z2 / T01 + z1 / T11) : 0.0;
    double s1 = m ? (1 / T01 - 1 / T11) : 0.0;
    const double c = 6 * log(10);
    d->sig0 = (c * s0) / (z1 - z2);
    d->sig1 = (c * s1) / (z1 - z2);
}

/* Linear-interpolation operator rows, float32 arithmetic cast to double
 * (reference misc.cpp:78-105: F.interpolate(eye(in), size=out, 'linear',
 * align_corners=True), computed on the default float32 dtype).  Row o of the
 * (out x in) matrix has weights l0 at i0 and l1 at i1.                          */
static void interp_row(int in, int out, int o, int *i0, int *i1, double *l0, double *l1)
{
    const float s = (out > 1) ? (float)(in - 1) / (float)(out - 1) : 0.0f;
    const float r = s * (float)o;
    int a0 = (int)r;
    if (a0 > in - 1) a0 = in - 1;
    const int a1 = a0 + (a0 < in - 1 ? 1 : 0);
    const float w1 = r - (float)a0;
    const float w0 = 1.0f - w1;
    *i0 = a0; *i1 = a1; *l0 = (double)w0; *l1 = (double)w1;
}

/* float32 linspace(h, 1, N) as torch computes it (reference misc.cpp:26-27;
 * ATen RangeFactories: step=(end-start)/(steps-1); first half start+step*i,
 * second half end-step*(steps-1-i)), cast to double.                           */
static void linspace_f32(int N, double *x)
{
    const float h = (float)(1. / N);
    const float step = (N > 1) ? (1.0f - h) / (float)(N - 1) : 0.0f;
    const int half = N / 2;
    for (int i = 0; i < N; i++) {
        /* verified bit-exact against torch.linspace (float32, CPU) for N=2..1300 */
        float v = (i < half) ? fmaf(step, (float)i, h) : fmaf(-step, (float)(N - 1 - i), 1.0f);
        x[i] = (double)v;
    }
}

/* dense LU with partial pivoting: solves M x = rhs in place (M is n x n,
 * row-major, destroyed).  Stands in for linalg_inv + matmul (string.cpp:175,238). */
static void lu_solve(double *M, double *x, int n)
{
    for (int c = 0; c < n; c++) {
        int p = c; double best = fabs(M[(size_t)c * n + c]);
        for (int r = c + 1; r < n; r++) {
            double v = fabs(M[(size_t)r * n + c]);
            if (v > best) { best = v; p = r; }
        }
        if (p != c) {
            for (int j = 0; j < n; j++) {
                double t = M[(size_t)c * n + j]; M[(size_t)c * n + j] = M[(size_t)p * n + j]; M[(size_t)p * n + j] = t;
            }
            double t = x[c]; x[c] = x[p]; x[p] = t;
        }
        const double piv = M[(size_t)c * n + c];
        for (int r = c + 1; r < n; r++) {
            double f = M[(size_t)r * n + c];
            if (f == 0.0) continue;
            f /= piv;
            M[(size_t)r * n + c] = 0.0;
            for (int j = c + 1; j < n; j++) M[(size_t)r * n + j] -= f * M[(size_t)c * n + j];
            x[r] -= f * x[c];
        }
    }
    for (int r = n - 1; r >= 0; r--) {
        double s = x[r];
        for (int j = r + 1; j < n; j++) s -= M[(size_t)r * n + j] * x[j];
        x[r] = s / M[(size_t)r * n + r];
    }
}

/* per-string workspace */
typedef struct {
    double *u1, *u2, *z1, *z2;      /* masked previous states (mask_1d, string.cpp:129-132) */
    double *lam;                    /* Lam = Dxb u1 (string.cpp:152) */
    double *A;                      /* dense (Wt+Wl)^2 */
    double *Ktl, *Klt;              /* dense Wt x Wl, Wl x Wt */
    double *rb;                     /* base RHS (B w1 + C w2), length Wt+Wl */
    double *rhs, *u, *z, *unew, *znew, *rc, *tmp, *tmp2;
    derived_t d;
    double FH, uH, vrel;
} ws_t;

static double *dalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

/* K_tl y = -phi * Dxf Lam Dxb Int_tl y  (string.cpp:158), width Wt/Wl operators */
static void build_coupling(const ws_t *w, int Wt, int Wl, double phi, double *Ktl, double *Klt)
{
    const int Nt_ = w->d.N_t, Nl_ = w->d.N_l;
    const double ht = w->d.h_t, hl = w->d.h_l;
    /* Int_tl: (Wt x Wl), rows o<=N_t from Interpolator(in=N_l+1,out=N_t+1) */
    /* Int_lt: (Wl x Wt), rows o<=N_l from Interpolator(in=N_t+1,out=N_l+1) */
    double *Itl = dalloc((size_t)Wt * Wl), *Ilt = dalloc((size_t)Wl * Wt);
    for (int o = 0; o <= Nt_ && o < Wt; o++) {
        int i0, i1; double l0, l1;
        interp_row(Nl_ + 1, Nt_ + 1, o, &i0, &i1, &l0, &l1);
        if (i0 < Wl) Itl[(size_t)o * Wl + i0] += l0;
        if (i1 < Wl) Itl[(size_t)o * Wl + i1] += l1;
    }
    for (int o = 0; o <= Nl_ && o < Wl; o++) {
        int i0, i1; double l0, l1;
        interp_row(Nt_ + 1, Nl_ + 1, o, &i0, &i1, &l0, &l1);
        if (i0 < Wt) Ilt[(size_t)o * Wt + i0] += l0;
        if (i1 < Wt) Ilt[(size_t)o * Wt + i1] += l1;
    }
    /* Ktl = -phi * Dxf * (Lam * (Dxb * Itl)) */
    double *q = dalloc((size_t)Wt * Wl);
    for (int i = 0; i < Wt; i++)
        for (int j = 0; j < Wl; j++) {
            double y = Itl[(size_t)i * Wl + j] / ht - (i > 0 ? Itl[(size_t)(i - 1) * Wl + j] / ht : 0.0);
            q[(size_t)i * Wl + j] = w->lam[i] * y;
        }
    for (int i = 0; i < Wt; i++)
        for (int j = 0; j < Wl; j++) {
            double v = -q[(size_t)i * Wl + j] / ht + (i + 1 < Wt ? q[(size_t)(i + 1) * Wl + j] / ht : 0.0);
            Ktl[(size_t)i * Wl + j] = -phi * v;
        }
    free(q);
    /* Klt = -phi * Dxf_l * (Int_lt * (Lam * Dxb))   (string.cpp:159) */
    double *LD = dalloc((size_t)Wt * Wt);       /* Lam * Dxb */
    for (int i = 0; i < Wt; i++) {
        LD[(size_t)i * Wt + i] = w->lam[i] * (1 / ht);
        if (i > 0) LD[(size_t)i * Wt + i - 1] = w->lam[i] * (-1 / ht);
    }
    double *P = dalloc((size_t)Wl * Wt);
    for (int j = 0; j < Wl; j++)
        for (int c = 0; c < Wt; c++) {
            double s = 0;
            /* Ilt row j has <=2 nonzeros, but stay generic */
            for (int m = 0; m < Wt; m++) {
                double e = Ilt[(size_t)j * Wt + m];
                if (e != 0.0) s += e * LD[(size_t)m * Wt + c];
            }
            P[(size_t)j * Wt + c] = s;
        }
    for (int j = 0; j < Wl; j++)
        for (int c = 0; c < Wt; c++) {
            double v = -P[(size_t)j * Wt + c] / hl + (j + 1 < Wl ? P[(size_t)(j + 1) * Wt + c] / hl : 0.0);
            Klt[(size_t)j * Wt + c] = -phi * v;
        }
    free(P); free(LD); free(Itl); free(Ilt);
}

/* raised_cosine (reference misc.cpp:20-34) over all Nx_t1 points, normalised */
static void raised_cosine(const double *xax, int N, double n, double ctr, double wid, double *out)
{
    ctr = (ctr * n) / N;
    wid = (wid * n) / N;
    double s = 0;
    for (int i = 0; i < N; i++) {
        double p = -(xax[i] - ctr - wid / 2) * (xax[i] - ctr + wid / 2);
        double r = p > 0 ? p : 0.0;                /* relu (NaN -> NaN) */
        if (p != p) r = p;
        double ind = (r > 0) - (r < 0);            /* sign; NaN handled below */
        if (r != r) ind = r;
        double o = (0.5 * ind) * (1 + cos(((2 * M_PI) * (xax[i] - ctr)) / wid));
        out[i] = o;
        s += fabs(o);
    }
    for (int i = 0; i < N; i++) out[i] = out[i] / s;   /* 0/0 -> NaN like the reference */
}

static double nan_to_num(double v)
{
    if (v != v) return 0.0;
    if (isinf(v)) return v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    return v;
}

/* manufactured-solution forcing (reference vnv.cpp:11-37), B == 1 only */
static double msf(double gamma, double sig0, double K, double p_a, double x, double t)
{
    const double sigma = sig0, omega = gamma, mu = M_PI, mu_sq = pow(M_PI, 2);
    const double cx = cos(mu * x);
    const double coeff_1 = (sigma * sigma - omega * omega - (2 * sig0) * sigma) * (cx * cx);
    const double coeff_2 = ((2 * mu_sq) * ((4 * (K * K)) * mu_sq + gamma * gamma)) * cos((2 * mu) * x);
    const double coeff_3 = ((2 * omega) * (sigma - sig0)) * (cx * cx);
    const double cos_term = (coeff_1 + coeff_2) * cos(omega * t);
    const double sin_term = coeff_3 * sin(omega * t);
    return (p_a * (cos_term + sin_term)) * exp((-1 * sigma) * t);
}


/* ---- matrix-free operators (solver == 1) ------------------------------------ */
/* out = K_tl z = -phi Dxf Lam Dxb Int_tl z   (string.cpp:158) */
static void apply_Ktl(const ws_t *w, int Wt, int Wl, double phi, const double *z, double *out, double *y)
{
    const int Nt_ = w->d.N_t, Nl_ = w->d.N_l; const double ht = w->d.h_t;
    for (int i = 0; i < Wt; i++) {
        if (i <= Nt_) {
            int i0, i1; double l0, l1;
            interp_row(Nl_ + 1, Nt_ + 1, i, &i0, &i1, &l0, &l1);
            y[i] = (i0 < Wl ? l0 * z[i0] : 0.0) + (i1 < Wl ? l1 * z[i1] : 0.0);
        } else y[i] = 0.0;
    }
    double qprev = 0.0; /* q_i = lam_i (y_i - y_{i-1})/ht */
    double q0 = w->lam[0] * (y[0] / ht);
    qprev = q0;
    for (int i = 0; i < Wt; i++) {
        double qn = (i + 1 < Wt) ? w->lam[i + 1] * ((y[i + 1] - y[i]) / ht) : 0.0;
        out[i] = -phi * ((qn - qprev) / ht);
        qprev = qn;
    }
}
/* out = K_lt u = -phi Dxf_l Int_lt Lam Dxb u   (string.cpp:159) */
static void apply_Klt(const ws_t *w, int Wt, int Wl, double phi, const double *u, double *out, double *q)
{
    const int Nt_ = w->d.N_t, Nl_ = w->d.N_l; const double ht = w->d.h_t, hl = w->d.h_l;
    for (int i = 0; i < Wt; i++) q[i] = w->lam[i] * ((u[i] - (i > 0 ? u[i - 1] : 0.0)) / ht);
    double pprev = 0.0;
    for (int j = 0; j <= Wl; j++) {
        double p = 0.0;
        if (j < Wl && j <= Nl_) {
            int i0, i1; double l0, l1;
            interp_row(Nt_ + 1, Nl_ + 1, j, &i0, &i1, &l0, &l1);
            p = (i0 < Wt ? l0 * q[i0] : 0.0) + (i1 < Wt ? l1 * q[i1] : 0.0);
        }
        if (j > 0) out[j - 1] = -phi * ((p - pprev) / hl);
        pprev = p;
    }
}
/* Thomas solve, tridiagonal (a sub, b diag, c super), cp = scratch */
static void thomas(const double *a, const double *b, const double *c, double *d, double *cp, int n)
{
    double den = b[0];
    cp[0] = c[0] / den; d[0] = d[0] / den;
    for (int i = 1; i < n; i++) {
        den = b[i] - a[i] * cp[i - 1];
        cp[i] = c[i] / den;
        d[i] = (d[i] - a[i] * d[i - 1]) / den;
    }
    for (int i = n - 2; i >= 0; i--) d[i] -= cp[i] * d[i + 1];
}

int sfdtd_oracle_forward(sfdtd_oracle_args *a)
{
    const int B = a->B, Nt = a->Nt, NXT = a->Nx_t1, NXL = a->Nx_l1;
    const double k = (double)a->k, k2 = pow(k, 2.);
    const double th = (double)a->theta_t;
    const float omthf = 1 - a->theta_t;              /* float32 (string.cpp:148) */
    const double omth = (double)omthf;
    const double lamc = (double)a->lambda_c;
    const double M_HD = (double)(-0.01f);            /* hammer.cpp:3 */
    const int max_iter = a->max_iter > 0 ? a->max_iter : 1000;
    int status = 0;

    double *xax = dalloc(NXT);
    linspace_f32(NXT, xax);
    ws_t *W = (ws_t *)calloc(B, sizeof(ws_t));
    const int NTOT = NXT + NXL;
    for (int b = 0; b < B; b++) {
        ws_t *w = &W[b];
        w->u1 = dalloc(NXT); w->u2 = dalloc(NXT); w->z1 = dalloc(NXL); w->z2 = dalloc(NXL);
        w->lam = dalloc(NXT + 1);
        w->A = dalloc((size_t)NTOT * NTOT);
        w->Ktl = dalloc((size_t)NXT * NXL); w->Klt = dalloc((size_t)NXT * NXL);
        w->rb = dalloc(NTOT); w->rhs = dalloc(NTOT);
        w->u = dalloc(NXT); w->z = dalloc(NXL); w->unew = dalloc(NXT); w->znew = dalloc(NXL);
        w->rc = dalloc(NXT); w->tmp = dalloc(NTOT); w->tmp2 = dalloc((size_t)NTOT * NTOT + 16 * (size_t)NTOT + 64);
    }
    if (a->stats) memset(a->stats, 0, 8 * sizeof(int64_t));

    for (int n = 2; n < Nt; n++) {                        /* simulator.cpp:40 */
        /* ---- derived vars, group-max operator widths (misc.cpp:119-127) ---- */
        int Wt = 0, Wl = 0;
        for (int b = 0; b < B; b++) {
            derive(a, b, n, &W[b].d);
            if (W[b].d.N_t + 1 > Wt) Wt = W[b].d.N_t + 1;
            if (W[b].d.N_l + 1 > Wl) Wl = W[b].d.N_l + 1;
        }
        if (Wt > NXT || Wl > NXL || Wt < 1 || Wl < 1) { status = -2; goto done; }
        const int nw = Wt + Wl;

        for (int b = 0; b < B; b++) {
            ws_t *w = &W[b];
            const derived_t *d = &w->d;
            const int N_t = d->N_t, N_l = d->N_l;
            const double ht = d->h_t, hl = d->h_l;
            const double *su1 = a->state_u + ((size_t)b * Nt + (n - 1)) * NXT;
            const double *su2 = a->state_u + ((size_t)b * Nt + (n - 2)) * NXT;
            const double *sz1 = a->state_z + ((size_t)b * Nt + (n - 1)) * NXL;
            const double *sz2 = a->state_z + ((size_t)b * Nt + (n - 2)) * NXL;
            for (int i = 0; i < NXT; i++) {              /* mask_1d: keep i <= N_t */
                w->u1[i] = (i <= N_t) ? su1[i] : su1[i] * 0.0;
                w->u2[i] = (i <= N_t) ? su2[i] : su2[i] * 0.0;
                w->u[i] = su1[i];                        /* iterate starts unmasked (string.cpp:190) */
            }
            for (int j = 0; j < NXL; j++) {
                w->z1[j] = (j <= N_l) ? sz1[j] : sz1[j] * 0.0;
                w->z2[j] = (j <= N_l) ? sz2[j] : sz2[j] * 0.0;
                w->z[j] = sz1[j];
            }
            /* Lam_i = (Dxb u1)_i, width Wt (string.cpp:152) */
            for (int i = 0; i < Wt; i++)
                w->lam[i] = w->u1[i] / ht - (i > 0 ? w->u1[i - 1] / ht : 0.0);
            w->lam[Wt] = 0.0;

            const double g = (d->gamma * d->gamma) * k2;                 /* gamma_k */
            const double phi = (g * (a->alpha[b] * a->alpha[b] - 1)) / 4; /* phi_pow */
            const double s0k = (2 * d->sig0) * k, s1k = (2 * d->sig1) * k;
            const double ht2 = ht * ht, hl2 = hl * hl, ht4 = pow(ht, 4.);
            const double Kk = (d->K * d->K) * k2;

            build_coupling(w, Wt, Wl, phi, w->Ktl, w->Klt);

            /* ---- dense A (string.cpp:153-174) and base RHS B w1 + C w2 (string.cpp:223-224) ---- */
            memset(w->A, 0, (size_t)nw * nw * sizeof(double));
            memset(w->rb, 0, (size_t)nw * sizeof(double));
            for (int i = 0; i < Wt; i++) {
                /* tridiagonal pieces at row i: columns i-1,i,i+1 (and i+-2 for D4) */
                for (int dj = -2; dj <= 2; dj++) {
                    int j = i + dj;
                    if (j < 0 || j >= Wt) continue;
                    double Id = (dj == 0), Mx = (dj == 1 || dj == -1) ? 0.5 : 0.0;
                    double Dxx = (dj == 0 ? -2.0 : ((dj == 1 || dj == -1) ? 1.0 : 0.0)) / ht2;
                    double d4c = (dj == 0 ? 6.0 : ((dj == 1 || dj == -1) ? -4.0 : 1.0));
                    if (dj == 0 && (i == 1 || i == N_t - 1)) d4c += 1.0;    /* Dxxxx_clamped (misc.cpp:146-163) */
                    double D4 = d4c / ht4;
                    double Theta = th * Id + omth * Mx;
                    /* V = -phi * Dxf Lam^2 Dxb */
                    double li2 = w->lam[i] * w->lam[i];
                    double lp2 = (i + 1 < Wt) ? w->lam[i + 1] * w->lam[i + 1] : 0.0;
                    double Vp = 0.0;
                    if (dj == -1) Vp = li2 / ht2;
                    else if (dj == 0) Vp = -(li2 + lp2) / ht2;
                    else if (dj == 1) Vp = lp2 / ht2;
                    double V = -phi * Vp;
                    double Qp = Theta + s0k * Id - s1k * Dxx;
                    double Qm = Theta - s0k * Id + s1k * Dxx;
                    double A1 = Qp + V, C1 = Qm + V;
                    double B1 = -2 * Theta - g * Dxx + Kk * D4;
                    if (dj >= -1 && dj <= 1) w->A[(size_t)i * nw + j] = A1;
                    w->rb[i] += B1 * w->u1[j];
                    if (dj >= -1 && dj <= 1) w->rb[i] += C1 * w->u2[j];
                }
                for (int j = 0; j < Wl; j++) {
                    double kt = w->Ktl[(size_t)i * Wl + j];
                    w->A[(size_t)i * nw + Wt + j] = kt;
                    w->rb[i] += (2 * kt) * w->z1[j] + kt * w->z2[j];   /* B_2 = 2 K_tl, C_2 = K_tl */
                }
            }
            for (int j = 0; j < Wl; j++) {
                for (int dj = -1; dj <= 1; dj++) {
                    int c = j + dj;
                    if (c < 0 || c >= Wl) continue;
                    double Id = (dj == 0);
                    double Dxx = (dj == 0 ? -2.0 : 1.0) / hl2;
                    double Qp = (1 + s0k) * Id - s1k * Dxx;
                    double Qm = (1 - s0k) * Id + s1k * Dxx;
                    double B4 = -2 * Id - (g * (a->alpha[b] * a->alpha[b])) * Dxx;
                    w->A[(size_t)(Wt + j) * nw + Wt + c] = Qp;
                    w->rb[Wt + j] += B4 * w->z1[c] + Qm * w->z2[c];
                }
                for (int c = 0; c < Wt; c++) {
                    double kl = w->Klt[(size_t)j * Wt + c];
                    w->A[(size_t)(Wt + j) * nw + c] = kl;
                    w->rb[Wt + j] += kl * w->u2[c];                    /* B_3 = 0, C_3 = K_lt */
                }
            }
            /* raised cosine for this step (bow.cpp:32) */
            {
                const double widl = a->wid[(size_t)b * Nt + n] * ht;      /* bow_wid_length (string.cpp:88) */
                raised_cosine(xax, NXT, (double)(N_t - 1), a->x_b[(size_t)b * Nt + n], widl, w->rc);
            }
        }

        /* ---- fixed-point loop over the forcing (string.cpp:200-258) ---- */
        int iter = 0, nc = 1;
        while (nc) {
            /* hammer: group-coupled inner loop (hammer.cpp:33-52) */
            double eta1[B > 0 ? B : 1], eta2[B > 0 ? B : 1], epsu[B > 0 ? B : 1],
                   eta_est[B > 0 ? B : 1], wH[B > 0 ? B : 1], Mr[B > 0 ? B : 1];
            int idxH[B > 0 ? B : 1];
            for (int b = 0; b < B; b++) {
                ws_t *w = &W[b];
                const double uH1 = a->u_H[(size_t)b * Nt + n - 1], uH2 = a->u_H[(size_t)b * Nt + n - 2];
                const double fi = floor(a->x_H[b] * (double)(w->d.N_t - 1));
                int idx = (int)fi;
                idxH[b] = idx;
                const int ok = (idx >= 0 && idx < NXT);
                epsu[b] = ok ? w->u[idx] : 0.0;
                eta1[b] = uH1 - (ok ? w->u1[idx] : 0.0);
                eta2[b] = uH2 - (ok ? w->u2[idx] : 0.0);
                wH[b] = a->w_H[b] / lamc; Mr[b] = a->M_r[b] / lamc;
                eta_est[b] = eta1[b] * (double)a->hammer_mask[b];
            }
            int hnc = 1, hit = 0;
            while (hnc) {
                hnc = 0;
                for (int b = 0; b < B; b++) {
                    ws_t *w = &W[b];
                    const double uH1 = a->u_H[(size_t)b * Nt + n - 1], uH2 = a->u_H[(size_t)b * Nt + n - 2];
                    const double eta = eta_est[b];
                    const double r1 = eta1[b] > 0 ? eta1[b] : (eta1[b] != eta1[b] ? eta1[b] : 0.0);
                    const double fH = ((pow(wH[b], 1 + a->alpha_H[b]) * pow(r1, a->alpha_H[b] - 1)) * (eta + eta2[b])) / 2;
                    w->FH = (eta1[b] > 0) ? fH : 0.0;
                    double uH = 2 * uH1 - uH2 - k2 * w->FH;
                    double t = uH - M_HD;
                    t = t > 0 ? t : (t != t ? t : 0.0);
                    w->uH = t + M_HD;
                    eta_est[b] = (w->uH - epsu[b]) * (double)a->hammer_mask[b];
                    if (fabs(eta - eta_est[b]) > w->d.tol_t) hnc = 1;
                }
                hit++;
                if (hit >= max_iter) { status = 1; break; }
            }
            if (a->stats) { a->stats[2] += hit; if (hit > a->stats[3]) a->stats[3] = hit; }

            nc = 0;
            for (int b = 0; b < B; b++) {
                ws_t *w = &W[b];
                const derived_t *d = &w->d;
                const int N_t = d->N_t, N_l = d->N_l;
                /* bow (bow.cpp:17-41) */
                double vrel = 0.0;
                const double vB = a->v_b[(size_t)b * Nt + n];
                for (int i = 0; i < NXT; i++) {
                    double dd = (iter == 0) ? (w->u1[i] - w->u2[i]) : (w->u[i] - w->u1[i]);
                    vrel += w->rc[i] * (dd / k - vB);
                }
                w->vrel = vrel;
                const double sg = (vrel > 0) - (vrel < 0);
                const double hb = (vrel != vrel) ? vrel
                    : sg * (a->phi_1[b] + (1 - a->phi_1[b]) * exp(-a->phi_0[b] * fabs(vrel)));
                const double FB = a->F_b[(size_t)b * Nt + n];
                /* RHS (string.cpp:223-233) */
                for (int i = 0; i < nw; i++) w->rhs[i] = w->rb[i];
                if (a->bow_mask[b])
                    for (int i = 0; i < Wt; i++) {
                        double GB = -k2 * (((w->rc[i] / d->h_t) * FB) * hb);
                        w->rhs[i] += nan_to_num(GB);
                    }
                if (a->hammer_mask[b]) {
                    int idx = idxH[b];
                    if (idx >= 0 && idx < Wt) w->rhs[idx] += nan_to_num(-k2 * ((1.0 * Mr[b]) * w->FH));
                }
                if (a->manufactured) {
                    /* domain_x (misc.cpp:45-52): sequential cumsum of 2/N_t, clamp, (v-1)/2 */
                    const float tf = (float)(n + a->n_0) * a->k;     /* float32 product (string.cpp:229) */
                    const double t = (double)tf;
                    const double v = 2 / (double)N_t;
                    double cs = 0;
                    for (int i = 0; i < NTOT; i++) {
                        cs += v;
                        double xv = cs - v;
                        xv = xv < 0 ? 0 : (xv > 2 ? 2 : xv);
                        xv = (xv - 1) / 2;
                        double f = msf(d->gamma, d->sig0, d->K, a->p_a[b], xv, t) * k2;
                        int row = (i < NXT) ? (i < Wt ? i : -1) : ((i - NXT) < Wl ? Wt + (i - NXT) : -1);
                        if (row >= 0) w->rhs[row] -= f;
                    }
                }
                /* flat-index mask (string.cpp:233): keep padded index < N_t+N_l+2 */
                const int keep = N_t + N_l + 2;
                for (int i = 0; i < Wt; i++) if (!(i < keep)) w->rhs[i] *= 0.0;
                for (int j = 0; j < Wl; j++) if (!(NXT + j < keep)) w->rhs[Wt + j] *= 0.0;
                /* solve A w = -RHS */
                if (a->solver == 0) {
                    memcpy(w->tmp2, w->A, (size_t)nw * nw * sizeof(double));
                    for (int i = 0; i < nw; i++) w->tmp[i] = -w->rhs[i];
                    lu_solve(w->tmp2, w->tmp, nw);
                } else if (a->solver == 2) {
                    /* CPU prototype of the CUDA kernel's solve (DESIGN.md "linear solve"):
                     *  - transverse rows trimmed to R = min(W_t, max(N_t+3, last forced row+1)); the homogeneous
                     *    Toeplitz ghost tail R..W_t-1 is folded exactly into the pivot of row R-1 (continued fraction);
                     *  - longitudinal rows trimmed to min(N_l+3, W_l), A22 by Jacobi sweeps folded into the block iteration;
                     *  - initial coupling guess z = 2 z1 - z2; rate-based stopping rule on the predicted error. */
                    const double g_ = (d->gamma * d->gamma) * k2;
                    const double phi_ = (g_ * (a->alpha[b] * a->alpha[b] - 1)) / 4;
                    double *ta = w->tmp2, *tb = ta + nw, *tc = tb + nw, *cp = tc + nw,
                           *sc1 = cp + nw, *uu = sc1 + nw + 2, *zc = uu + nw, *zn = zc + nw, *uo = zn + nw, *kz = uo + nw, *kl = kz + nw;
                    int last = -1;
                    for (int i = 0; i < Wt; i++) if (w->rhs[i] != 0.0) last = i;
                    int R = N_t + 3; if (last + 1 > R) R = last + 1; if (R > Wt) R = Wt;
                    const double offA = w->A[(size_t)(Wt - 1) * nw + Wt - 2];   /* Toeplitz tail coefficients */
                    const double diagA = (Wt - 1 >= N_t + 2) ? w->A[(size_t)(Wt - 1) * nw + Wt - 1] : 0.0;
                    for (int i = 0; i < R; i++) {
                        ta[i] = (i > 0) ? w->A[(size_t)i * nw + i - 1] : 0.0;
                        tb[i] = w->A[(size_t)i * nw + i];
                        tc[i] = (i + 1 < R) ? w->A[(size_t)i * nw + i + 1] : 0.0;
                    }
                    if (Wt > R) {
                        const int m = Wt - R;
                        double pv = diagA;
                        for (int j = 1; j < m; j++) { double pn = diagA - (offA * offA) / pv; if (pn == pv) break; pv = pn; }
                        tb[R - 1] -= (offA * offA) / pv;
                    }
                    int WLs = N_l + 3; if (WLs > Wl) WLs = Wl;
                    const double dA = w->A[(size_t)(Wt) * nw + Wt], eA = (Wl > 1) ? w->A[(size_t)(Wt) * nw + Wt + 1] : 0.0;
                    for (int j = 0; j < Wl; j++) { zc[j] = w->z1[j]; zn[j] = 0.0; }
                    for (int i = 0; i < Wt; i++) uu[i] = 0.0;
                    {
                        const int g2 = getenv("SFDTD_GUESS2") ? atoi(getenv("SFDTD_GUESS2")) : 2;
                        const double *sz3 = (n >= 3) ? a->state_z + ((size_t)b * Nt + (n - 3)) * NXL : NULL;
                        for (int j = 0; j < Wl; j++) {
                            if (g2 == 4) kl[j] = 0.0;
                            else if (g2 == 3 && sz3) kl[j] = 3 * w->z1[j] - 3 * w->z2[j] + ((j <= N_l) ? sz3[j] : 0.0);
                            else if (g2 == 1) kl[j] = w->z1[j];
                            else kl[j] = 2 * w->z1[j] - w->z2[j];
                        }
                    }
                    if (getenv("SFDTD_GUESS2") && atoi(getenv("SFDTD_GUESS2")) == 4) {
                        /* slaved guess: z0 = A22^-1 (-r_l - K_lt (2 u1 - u2)) (one Jacobi application around z1) */
                        for (int i = 0; i < Wt; i++) uu[i] = (i < R) ? 2 * w->u1[i] - w->u2[i] : 0.0;
                        if (phi_ != 0.0) apply_Klt(w, Wt, Wl, phi_, uu, kz, sc1); else for (int j = 0; j < Wl; j++) kz[j] = 0.0;
                        for (int j = 0; j < WLs; j++) {
                            const double zl = j > 0 ? w->z1[j - 1] : 0.0, zr = (j + 1 < WLs) ? w->z1[j + 1] : 0.0;
                            kl[j] = ((-w->rhs[Wt + j] - kz[j]) - eA * (zl + zr)) / dA;
                        }
                        for (int j = WLs; j < Wl; j++) kl[j] = 0.0;
                        for (int j = 0; j < Wl; j++) zc[j] = kl[j];
                        for (int i = 0; i < Wt; i++) uu[i] = 0.0;
                    }
                    int sweeps = 0; double du_prev = 0, su_ = 0;
                    const double TOL = getenv("SFDTD_TOL2") ? atof(getenv("SFDTD_TOL2")) : 1e-13;
                    static __thread double rho_hist[4096];
                    double rho_h = (n == 2) ? 0.5 : rho_hist[b & 4095];
                    for (;;) {
                        for (int i = 0; i < R; i++) uo[i] = uu[i];
                        if (phi_ != 0.0) apply_Ktl(w, Wt, Wl, phi_, sweeps == 0 ? kl : zc, kz, sc1);
                        else for (int i = 0; i < Wt; i++) kz[i] = 0.0;
                        for (int i = 0; i < R; i++) uu[i] = -w->rhs[i] - kz[i];
                        thomas(ta, tb, tc, uu, cp, R);
                        for (int i = R; i < Wt; i++) uu[i] = 0.0;
                        if (getenv("SFDTD_AA")) {
                            /* experiment: Anderson(1) mixing of the transverse iterate */
                            static __thread double fprev[4096], gprev[4096];
                            double num = 0, den = 0;
                            if (sweeps >= 1) {
                                for (int i = 0; i < R; i++) { const double f = uu[i] - uo[i], df = f - fprev[i]; num += f * df; den += df * df; }
                            }
                            const double gam = (sweeps >= 1 && den > 0) ? num / den : 0.0;
                            for (int i = 0; i < R; i++) {
                                const double g_ = uu[i], f = uu[i] - uo[i];
                                if (sweeps >= 1) uu[i] = g_ - gam * (g_ - gprev[i]);
                                fprev[i] = f; gprev[i] = g_;
                            }
                        }
                        if (phi_ != 0.0) apply_Klt(w, Wt, Wl, phi_, uu, kl, sc1);
                        else for (int j = 0; j < Wl; j++) kl[j] = 0.0;
                        for (int j = 0; j < WLs; j++) {
                            const double zl = j > 0 ? zc[j - 1] : 0.0, zr = (j + 1 < WLs) ? zc[j + 1] : 0.0;
                            zn[j] = ((-w->rhs[Wt + j] - kl[j]) - eA * (zl + zr)) / dA;
                        }
                        for (int j = WLs; j < Wl; j++) zn[j] = 0.0;
                        { double *t_ = zc; zc = zn; zn = t_; }
                        sweeps++;
                        double du = 0;
                        for (int i = 0; i < R; i++) { double e = fabs(uu[i] - uo[i]); if (e > du || e != e) du = e; }
                        if (sweeps == 1) for (int i = 0; i < R; i++) { double e = fabs(uu[i]); if (e > su_) su_ = e; }
                        const int has_rl = (N_t + N_l + 2 - NXT) > 0;
                        const int minS = has_rl ? 4 : (phi_ != 0.0 ? 2 : 1);
                        double rho = rho_h * 2;
                        if (sweeps >= 3 && du_prev > 0) { double r_ = du / du_prev; if (r_ > rho_h) rho_h = r_; else rho_h = 0.5 * (rho_h + r_); rho = 1.5 * rho_h; }
                        if (rho > 0.9) rho = 0.9;
                        const double est = du * rho / (1 - rho);
                        du_prev = du;
                        double dz = 0, sz_ = 0;
                        for (int j = 0; j < WLs; j++) { double e = fabs(zc[j] - zn[j]); if (e > dz) dz = e; e = fabs(zc[j]); if (e > sz_) sz_ = e; }
                        const double estz = dz * rho / (1 - rho);
                        if (getenv("SFDTD_TRACE") && n == atoi(getenv("SFDTD_TRACE")) && b < 6)
                            fprintf(stderr, "n=%d b=%d sweep=%d du/su=%.2e dz/sz=%.2e rho=%.3f est=%.2e estz=%.2e\n", n, b, sweeps, du / su_, dz / sz_, rho, est / su_, estz / sz_);
                        if (getenv("SFDTD_NOZ")) { if (sweeps >= 3 && !(est > TOL * su_)) break; }
                        else
                        if (sweeps >= minS && !(est > TOL * su_) && !(estz > TOL * sz_)) break;      /* also exits on NaN */
                        if (sweeps >= 500) { status = 2; break; }
                    }
                    rho_hist[b & 4095] = rho_h;
                    if (a->stats) { a->stats[5] += sweeps; if (sweeps > a->stats[6]) a->stats[6] = sweeps; a->stats[7] += 1; }
                    for (int i = 0; i < Wt; i++) w->tmp[i] = uu[i];
                    for (int j = 0; j < Wl; j++) w->tmp[Wt + j] = zc[j];
                } else {
                    /* block Gauss-Seidel: A11 u = -r_t - K_tl z ;  A22 z = -r_l - K_lt u */
                    const double g_ = (d->gamma * d->gamma) * k2;
                    const double phi_ = (g_ * (a->alpha[b] * a->alpha[b] - 1)) / 4;
                    double *ta = w->tmp2, *tb = ta + nw, *tc = tb + nw, *cp = tc + nw,
                           *la = cp + nw, *lb = la + nw, *lc = lb + nw, *sc1 = lc + nw, *sc2 = sc1 + nw + 2,
                           *uu = sc2 + nw + 2, *zz = uu + nw, *uo = zz + nw, *zo = uo + nw;
                    for (int i = 0; i < Wt; i++) {
                        ta[i] = (i > 0) ? w->A[(size_t)i * nw + i - 1] : 0.0;
                        tb[i] = w->A[(size_t)i * nw + i];
                        tc[i] = (i + 1 < Wt) ? w->A[(size_t)i * nw + i + 1] : 0.0;
                    }
                    for (int j = 0; j < Wl; j++) {
                        la[j] = (j > 0) ? w->A[(size_t)(Wt + j) * nw + Wt + j - 1] : 0.0;
                        lb[j] = w->A[(size_t)(Wt + j) * nw + Wt + j];
                        lc[j] = (j + 1 < Wl) ? w->A[(size_t)(Wt + j) * nw + Wt + j + 1] : 0.0;
                    }
                    const double GS_TOL = getenv("SFDTD_GSTOL") ? atof(getenv("SFDTD_GSTOL")) : 1e-13;
                    const int guess = getenv("SFDTD_GUESS") ? atoi(getenv("SFDTD_GUESS")) : 0;
                    for (int j = 0; j < Wl; j++) zz[j] = guess == 0 ? 0.0 : (guess == 1 ? w->z1[j] : 2 * w->z1[j] - w->z2[j]);
                    for (int i = 0; i < Wt; i++) uu[i] = 0.0;
                    int sweeps = 0;
                    for (;;) {
                        for (int i = 0; i < Wt; i++) uo[i] = uu[i];
                        for (int j = 0; j < Wl; j++) zo[j] = zz[j];
                        if (phi_ != 0.0) apply_Ktl(w, Wt, Wl, phi_, zz, uu, sc1);
                        else for (int i = 0; i < Wt; i++) uu[i] = 0.0;
                        for (int i = 0; i < Wt; i++) uu[i] = -w->rhs[i] - uu[i];
                        thomas(ta, tb, tc, uu, cp, Wt);
                        if (phi_ != 0.0) apply_Klt(w, Wt, Wl, phi_, uu, zz, sc1);
                        else for (int j = 0; j < Wl; j++) zz[j] = 0.0;
                        for (int j = 0; j < Wl; j++) zz[j] = -w->rhs[Wt + j] - zz[j];
                        thomas(la, lb, lc, zz, cp, Wl);
                        sweeps++;
                        double du = 0, su_ = 0, dz = 0, sz_ = 0;
                        for (int i = 0; i < Wt; i++) { double e = fabs(uu[i] - uo[i]); if (e > du) du = e; e = fabs(uu[i]); if (e > su_) su_ = e; }
                        for (int j = 0; j < Wl; j++) { double e = fabs(zz[j] - zo[j]); if (e > dz) dz = e; e = fabs(zz[j]); if (e > sz_) sz_ = e; }
                        if (!(du > GS_TOL * su_) && !(dz > GS_TOL * sz_)) break;   /* also exits on NaN */
                        if (getenv("SFDTD_DEBUG") && sweeps >= 20 && sweeps < 24)
                            fprintf(stderr, "n=%d b=%d it=%d sweep=%d du=%.3e su=%.3e dz=%.3e sz=%.3e\n", n, b, iter, sweeps, du, su_, dz, sz_);
                        if (sweeps >= 500) { status = 2; break; }
                    }
                    if (a->stats) { a->stats[5] += sweeps; if (sweeps > a->stats[6]) a->stats[6] = sweeps; a->stats[7] += 1; }
                    for (int i = 0; i < Wt; i++) w->tmp[i] = uu[i];
                    for (int j = 0; j < Wl; j++) w->tmp[Wt + j] = zz[j];
                }
                /* mask + Dirichlet (string.cpp:240-246) */
                double res_u = 0, res_z = 0; int nanu = 0, nanz = 0;
                for (int i = 0; i < NXT; i++) {
                    double v = (i < Wt) ? w->tmp[i] : 0.0;
                    if (!(i <= N_t)) v *= 0.0;
                    if (i == 0 || i == N_t) v *= 0.0;
                    w->unew[i] = v;
                    double r = fabs(w->u[i] - v);
                    if (r != r) nanu = 1; else if (r > res_u) res_u = r;
                }
                for (int j = 0; j < NXL; j++) {
                    double v = (j < Wl) ? w->tmp[Wt + j] : 0.0;
                    if (!(j <= N_l)) v *= 0.0;
                    if (j == 0 || j == N_l) v *= 0.0;
                    w->znew[j] = v;
                    double r = fabs(w->z[j] - v);
                    if (r != r) nanz = 1; else if (r > res_z) res_z = r;
                }
                /* torch max() propagates NaN; NaN > tol is false */
                if (!nanu && res_u > d->tol_t) nc = 1;
                if (!nanz && res_z > d->tol_l) nc = 1;
            }
            for (int b = 0; b < B; b++) {
                memcpy(W[b].u, W[b].unew, NXT * sizeof(double));
                memcpy(W[b].z, W[b].znew, NXL * sizeof(double));
            }
            iter++;
            if (iter >= max_iter) { status = 1; break; }
        }
        if (a->stats) { a->stats[0] += iter; if (iter > a->stats[1]) a->stats[1] = iter; a->stats[4] += 1; }

        /* ---- save and readout (string.cpp:263-303) ---- */
        for (int b = 0; b < B; b++) {
            ws_t *w = &W[b];
            const derived_t *d = &w->d;
            double *su = a->state_u + ((size_t)b * Nt + n) * NXT;
            double *sz = a->state_z + ((size_t)b * Nt + n) * NXL;
            const double *su1 = a->state_u + ((size_t)b * Nt + (n - 1)) * NXT;
            const double *sz1 = a->state_z + ((size_t)b * Nt + (n - 1)) * NXL;
            for (int i = 0; i < NXT; i++) su[i] += w->u[i];
            for (int j = 0; j < NXL; j++) sz[j] += w->z[j];
            double uo, zo;
            if (a->surface_integral) {
                const double rw = 0.5 * d->h_t;
                const double wgt = rw * 1.0 + rw * (double)a->hammer_mask[b] + rw * (double)a->bow_mask[b];
                uo = 0; zo = 0;
                for (int i = 0; i < NXT; i++) uo += ((w->u[i] - su1[i]) * wgt) / k;
                for (int j = 0; j < NXL; j++) zo += ((w->z[j] - sz1[j]) * wgt) / k;
            } else {
                const double rp = a->pos[b];
                const long ui = 1 + (long)floor((double)d->N_t * rp);
                const double uf = 1 + rp / d->h_t - (double)ui;
                const long zi = 1 + (long)floor((double)d->N_l * rp);
                const double zf = 1 + rp / d->h_l - (double)zi;
                if (ui + 1 >= NXT || zi + 1 >= NXL || ui < 0 || zi < 0) { status = -3; goto done; }
                uo = (1 - uf) * w->u[ui] + uf * w->u[ui + 1];
                zo = (1 - zf) * w->z[zi] + zf * w->z[zi + 1];
            }
            a->uout[(size_t)b * Nt + n] = uo;
            a->zout[(size_t)b * Nt + n] = zo;
            a->v_r[(size_t)b * Nt + n] = w->vrel;
            a->F_H[(size_t)b * Nt + n] = w->FH;
            a->u_H[(size_t)b * Nt + n] += w->uH;
            a->sig0[b] = d->sig0; a->sig1[b] = d->sig1;
        }
    }
    /* u_H / k (simulator.cpp:57) */
    for (size_t i = 0; i < (size_t)B * Nt; i++) a->u_H_out[i] = a->u_H[i] / k;

done:
    for (int b = 0; b < B; b++) {
        ws_t *w = &W[b];
        free(w->u1); free(w->u2); free(w->z1); free(w->z2); free(w->lam); free(w->A);
        free(w->Ktl); free(w->Klt); free(w->rb); free(w->rhs); free(w->u); free(w->z);
        free(w->unew); free(w->znew); free(w->rc); free(w->tmp); free(w->tmp2);
    }
    free(W); free(xax);
    return status;
}
