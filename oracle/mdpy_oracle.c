/* oracle/mdpy_oracle.c — CPU oracle for the mdpy nonbonded hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  A plain-C restatement of the reference's algorithm
 * (mdpy v0.2.x, /root/reference) plus float64 brute-force / Ewald truth for the
 * parts of the north star that have no reference code (erfc direct space, CHARMM
 * switch, PME — "parity unpinned" for those).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call it; the product library never links it.
 *
 * Part 1 (reference restatement, float and double): mdpy_oracle_impl.h
 *   ora_wrap_positions_*   <- mdpy/utils/pbc.py:28-36
 *   ora_cell_index_*       <- mdpy/core/cell_list.py:83-104
 *   ora_cell_fill          <- mdpy/core/cell_list.py:106-116
 *   ora_lj_cell_*          <- mdpy/constraint/charmm_nonbonded_constraint.py:64-108
 *   ora_coulomb_allpairs_* <- mdpy/constraint/electrostatic_constraint.py:52-79
 * Part 2 (float64 truth, this file):
 *   ora_nonbonded_bruteforce  all-pairs minimum-image LJ (+CHARMM switch) + erfc/bare Coulomb
 *   ora_pair_set_f32          the canonical fp32 in-cutoff pair set
 *   ora_ewald_recip           explicit k-space Ewald sum (structure factors)
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -o liboracle.so mdpy_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define SUF _f32
#include "mdpy_oracle_impl.h"
#undef REAL
#undef SUF
#define REAL double
#define SUF _f64
#include "mdpy_oracle_impl.h"
#undef REAL
#undef SUF

/* Restates the second half of CellList.kernel (mdpy/core/cell_list.py:106-116):
 * dense [nx,ny,nz,P] table filled with -1 (:107-110), atoms scattered in index order
 * (:112-115). */
void ora_cell_fill(int n, const int *pci, const int *ncell, int P, int *cell_list) {
    size_t total = (size_t)ncell[0] * ncell[1] * ncell[2];
    for (size_t i = 0; i < total * (size_t)P; ++i) cell_list[i] = -1;
    int *cur = (int *)calloc(total, sizeof(int));
    for (int i = 0; i < n; ++i) {
        size_t lin = ((size_t)pci[3 * i] * ncell[1] + pci[3 * i + 1]) * ncell[2] + pci[3 * i + 2];
        cell_list[lin * P + cur[lin]] = i;
        ++cur[lin];
    }
    free(cur);
}

static inline int in_row(const int *row, int width, int id) {
    for (int k = 0; k < width; ++k)
        if (row[k] == id) return 1;
    return 0;
}

static inline double minimg(double d, double L) { return d - L * rint(d / L); }

/* Float64 all-pairs truth for the nonbonded direct-space terms.
 *   pair set: {(i,j): |minimg(x_j - x_i)| <= rc, j not in bonded[i]}   
 *   LJ: the reference's formula (charmm_nonbonded_constraint.py:98-107) times the CHARMM
 *       energy switch S(r) on (r_on, rc]; r_on >= rc reproduces the reference's plain cut.
 *   coul_mode 0: none; 1: erfc(alpha r)/r inside rc_c + erf correction for excluded pairs
 *       (SURVEY §8c SPME spec); 2: bare q1 q2/(4 pi eps0 r), minimum image, no cutoff
 *       (electrostatic_constraint.py:52-79).
 * Forces are accumulated on atoms i in [i0, i1) only, over ALL j != i (full shell), so a
 * sub-range can be checked on large systems; energies are half the ordered-pair sums over
 * that range (== the total when the range is [0, n)).
 * energies[0]=E_lj  [1]=E_coul (direct or bare)  [2]=E_excluded-pair correction
 * [3]=sum of |LJ pair energies|  [4]=sum of |Coulomb pair energies| (the scale a cancelling total is judged on)
 * counts[0]=ordered in-cutoff LJ pairs, counts[1]=ordered in-cutoff Coulomb pairs */
void ora_nonbonded_bruteforce(int n, const double *pos, const double *box, const double *params,
                              const double *charges, const int *bonded, int wb,
                              const int *scaling, int ws, double rc_lj, double r_on,
                              int coul_mode, double k_e, double alpha, double rc_c, int i0,
                              int i1, double *f_lj, double *f_coul, double *energies,
                              long long *counts) {
    const double two_over_sqrtpi = 1.1283791670955126;
    double e_lj = 0, e_c = 0, e_x = 0, e_lj_abs = 0, e_c_abs = 0;
    long long n_lj = 0, n_c = 0;
    const double a2 = rc_lj * rc_lj, b2 = r_on * r_on;
    const int use_switch = (r_on < rc_lj);
    const double inv_ab3 = use_switch ? 1.0 / ((a2 - b2) * (a2 - b2) * (a2 - b2)) : 0.0;
    for (int i = i0; i < i1; ++i) {
        const int *bi = bonded + (size_t)i * wb;
        const int *si = scaling + (size_t)i * ws;
        double fl[3] = {0, 0, 0}, fc[3] = {0, 0, 0};
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            double d[3];
            for (int a = 0; a < 3; ++a) d[a] = minimg(pos[3 * j + a] - pos[3 * i + a], box[a]);
            double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            double r = sqrt(r2);
            int excluded = in_row(bi, wb, j);
            double fscal = 0.0; /* force on i = -fscal_* d  (d points i -> j) */
            if (!excluded && rc_lj > 0 && r <= rc_lj) {
                int is14 = in_row(si, ws, j);
                double e1 = params[4 * i + (is14 ? 2 : 0)], g1 = params[4 * i + (is14 ? 3 : 1)];
                double e2 = params[4 * j + (is14 ? 2 : 0)], g2 = params[4 * j + (is14 ? 3 : 1)];
                double eps = sqrt(e1 * e2), sig = 0.5 * (g1 + g2);
                double sr2 = sig * sig / r2, sr6 = sr2 * sr2 * sr2, sr12 = sr6 * sr6;
                double e = 4 * eps * (sr12 - sr6);
                double dedr = -24 * eps * (2 * sr12 - sr6) / r; /* dE/dr */
                if (use_switch && r > r_on) {
                    double S = (a2 - r2) * (a2 - r2) * (a2 + 2 * r2 - 3 * b2) * inv_ab3;
                    double dS = 12 * r * (a2 - r2) * (b2 - r2) * inv_ab3;
                    dedr = dedr * S + e * dS;
                    e *= S;
                }
                e_lj += 0.5 * e; e_lj_abs += 0.5 * fabs(e);
                ++n_lj;
                for (int a = 0; a < 3; ++a) fl[a] += dedr * d[a] / r; /* F_i = +dE/dr * rhat(i->j) */
                (void)fscal;
            }
            double qq = k_e * charges[i] * charges[j];
            if (coul_mode == 2) {
                if (!excluded) {
                    e_c += 0.5 * qq / r; e_c_abs += 0.5 * fabs(qq / r);
                    ++n_c;
                    for (int a = 0; a < 3; ++a) fc[a] += -qq / r2 * d[a] / r;
                }
            } else if (coul_mode == 1) {
                if (!excluded) {
                    if (r <= rc_c) {
                        double ar = alpha * r;
                        double erfc_ar = erfc(ar);
                        e_c += 0.5 * qq * erfc_ar / r; e_c_abs += 0.5 * fabs(qq * erfc_ar / r);
                        ++n_c;
                        double dedr = -qq * (erfc_ar / r2 + two_over_sqrtpi * alpha * exp(-ar * ar) / r);
                        for (int a = 0; a < 3; ++a) fc[a] += dedr * d[a] / r;
                    }
                } else {
                    /* excluded pair: remove the reciprocal-space image of the pair,
                     * E = -qq erf(alpha r)/r */
                    double ar = alpha * r;
                    double erf_ar = erf(ar);
                    e_x += 0.5 * (-qq * erf_ar / r);
                    double dedr = -qq * (two_over_sqrtpi * alpha * exp(-ar * ar) / r - erf_ar / r2);
                    for (int a = 0; a < 3; ++a) fc[a] += dedr * d[a] / r;
                }
            }
        }
        for (int a = 0; a < 3; ++a) {
            if (f_lj) f_lj[3 * i + a] = fl[a];
            if (f_coul) f_coul[3 * i + a] = fc[a];
        }
    }
    energies[0] = e_lj; energies[1] = e_c; energies[2] = e_x; energies[3] = e_lj_abs; energies[4] = e_c_abs;
    counts[0] = n_lj; counts[1] = n_c;
}

/* The canonical fp32 pair-set criterion (the definition both the CUDA kernels and this
 * checker evaluate, bit for bit; SURVEY §8a Q1 asks for one un-ambiguous fp32 expression):
 *     d   = x_j - x_i                                   (fp32 subtract)
 *     t   = fmaf(d, invL, 12582912.f) - 12582912.f       (== rintf(d*invL), one rounding)
 *     d   = fmaf(-L, t, d)
 *     r2  = fmaf(dz, dz, fmaf(dy, dy, dx*dx))
 *     in  = r2 <= rc*rc                                  (fp32 product)
 * with invL = 1.0f/L.  The reference tests sqrt(r2) <= rc on its own (BLAS-dependent)
 * rounding (charmm_nonbonded_constraint.py:85-90); the two differ only for pairs within
 * an ulp of the cutoff.  Pairs are emitted i<j, row-major, excluded (bonded) pairs
 * skipped.  Returns the pair count (and writes at most `cap`). */
long long ora_pair_set_f32_range(int n, const float *pos, const float *box, float rc, const int *bonded,
                                 int wb, int i0, int i1, int *out_i, int *out_j, long long cap);

long long ora_pair_set_f32(int n, const float *pos, const float *box, float rc, const int *bonded,
                           int wb, int *out_i, int *out_j, long long cap) {
    return ora_pair_set_f32_range(n, pos, box, rc, bonded, wb, 0, n, out_i, out_j, cap);
}

/* rows i in [i0, i1) of the same set (the rows are independent: host threads take slices) */
long long ora_pair_set_f32_range(int n, const float *pos, const float *box, float rc, const int *bonded,
                                 int wb, int i0, int i1, int *out_i, int *out_j, long long cap) {
    const float MAGIC = 12582912.0f;
    const float rc2 = rc * rc;
    float invL[3] = {1.0f / box[0], 1.0f / box[1], 1.0f / box[2]};
    long long cnt = 0;
    for (int i = i0; i < i1; ++i) {
        const int *bi = bonded + (size_t)i * wb;
        for (int j = i + 1; j < n; ++j) {
            float d[3];
            for (int a = 0; a < 3; ++a) {
                float dd = pos[3 * j + a] - pos[3 * i + a];
                volatile float t0 = fmaf(dd, invL[a], MAGIC);
                float t = t0 - MAGIC;
                d[a] = fmaf(-box[a], t, dd);
            }
            float r2 = fmaf(d[2], d[2], fmaf(d[1], d[1], d[0] * d[0]));
            if (!(r2 <= rc2)) continue;
            if (in_row(bi, wb, j)) continue;
            if (cnt < cap) { out_i[cnt] = i; out_j[cnt] = j; }
            ++cnt;
        }
    }
    return cnt;
}

/* Explicit reciprocal-space Ewald sum (float64), orthorhombic box:
 *   E_rec = k_e/(2 pi V) sum_{m != 0} exp(-pi^2 m^2/alpha^2)/m^2 |S(m)|^2,
 *   S(m) = sum_j q_j exp(2 pi i m.r_j),  m = (mx/Lx, my/Ly, mz/Lz), |m_a| <= kmax[a]
 *   F_i  = 2 k_e q_i / V sum_m A(m) m (C sin(th_i) - D cos(th_i)),  S = C + iD
 * (SURVEY §8c spec).  energies[0]=E_rec, [1]=E_self=-k_e alpha/sqrt(pi) sum q^2,
 * [2]=E_background=-k_e pi (sum q)^2/(2 V alpha^2). */
void ora_ewald_recip(int n, const double *pos, const double *q, const double *box, double alpha,
                     const int *kmax, double k_e, double *forces, double *energies) {
    const double PI = 3.14159265358979323846;
    double V = box[0] * box[1] * box[2];
    int K[3] = {kmax[0], kmax[1], kmax[2]};
    /* per-atom phase tables e^{2 pi i m x / L}, m = 0..K */
    double *tab[3][2];
    for (int a = 0; a < 3; ++a) {
        tab[a][0] = (double *)malloc(sizeof(double) * (size_t)n * (K[a] + 1));
        tab[a][1] = (double *)malloc(sizeof(double) * (size_t)n * (K[a] + 1));
        for (int i = 0; i < n; ++i)
            for (int m = 0; m <= K[a]; ++m) {
                double th = 2 * PI * m * pos[3 * i + a] / box[a];
                tab[a][0][(size_t)i * (K[a] + 1) + m] = cos(th);
                tab[a][1][(size_t)i * (K[a] + 1) + m] = sin(th);
            }
    }
    memset(forces, 0, sizeof(double) * 3 * (size_t)n);
    double *cs = (double *)malloc(sizeof(double) * n), *sn = (double *)malloc(sizeof(double) * n);
    double e_rec = 0;
    for (int mx = 0; mx <= K[0]; ++mx)
    for (int my = (mx == 0 ? 0 : -K[1]); my <= K[1]; ++my)
    for (int mz = ((mx == 0 && my == 0) ? 1 : -K[2]); mz <= K[2]; ++mz) {
        /* half space: mx>0, or mx==0&&my>0, or mx==my==0&&mz>0; weight 2 */
        if (mx == 0 && my < 0) continue;
        double mv[3] = {mx / box[0], my / box[1], mz / box[2]};
        double m2 = mv[0] * mv[0] + mv[1] * mv[1] + mv[2] * mv[2];
        double A = exp(-PI * PI * m2 / (alpha * alpha)) / m2;
        if (A < 1e-300) continue;
        double C = 0, D = 0;
        int ay = my < 0 ? -my : my, az = mz < 0 ? -mz : mz;
        for (int i = 0; i < n; ++i) {
            double cx = tab[0][0][(size_t)i * (K[0] + 1) + mx], sx = tab[0][1][(size_t)i * (K[0] + 1) + mx];
            double cy = tab[1][0][(size_t)i * (K[1] + 1) + ay], sy = tab[1][1][(size_t)i * (K[1] + 1) + ay];
            double cz = tab[2][0][(size_t)i * (K[2] + 1) + az], sz = tab[2][1][(size_t)i * (K[2] + 1) + az];
            if (my < 0) sy = -sy;
            if (mz < 0) sz = -sz;
            double cxy = cx * cy - sx * sy, sxy = sx * cy + cx * sy;
            double c = cxy * cz - sxy * sz, s = sxy * cz + cxy * sz;
            cs[i] = c; sn[i] = s;
            C += q[i] * c; D += q[i] * s;
        }
        e_rec += 2.0 * A * (C * C + D * D);
        for (int i = 0; i < n; ++i) {
            double w = 2.0 * A * q[i] * (C * sn[i] - D * cs[i]);
            forces[3 * i] += w * mv[0]; forces[3 * i + 1] += w * mv[1]; forces[3 * i + 2] += w * mv[2];
        }
    }
    double pref_f = 2.0 * k_e / V;
    for (int i = 0; i < 3 * n; ++i) forces[i] *= pref_f;
    double sq = 0, sq2 = 0;
    for (int i = 0; i < n; ++i) { sq += q[i]; sq2 += q[i] * q[i]; }
    energies[0] = k_e / (2 * PI * V) * e_rec;
    energies[1] = -k_e * alpha / sqrt(PI) * sq2;
    energies[2] = -k_e * PI * sq * sq / (2 * V * alpha * alpha);
    free(cs); free(sn);
    for (int a = 0; a < 3; ++a) { free(tab[a][0]); free(tab[a][1]); }
}
