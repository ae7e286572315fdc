/* oracle/mdpy_oracle_impl.h — type-generic body, included twice by mdpy_oracle.c
 * (REAL=float, SUF=_f32) and (REAL=double, SUF=_f64).
 *
 * TEST INFRASTRUCTURE ONLY: a CPU restatement of the reference's algorithm
 * (mdpy v0.2.x, /root/reference).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may call it.  The product
 * (mdpy_b200/) never links, imports or executes anything in oracle/.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* v (row vector) times 3x3 row-major matrix m — numpy `np.dot(vec, m)` for a 1-D vec. */
static inline void FN(vecmat)(const REAL v[3], const REAL *m, REAL out[3]) {
    for (int c = 0; c < 3; ++c)
        out[c] = v[0] * m[0 * 3 + c] + v[1] * m[1 * 3 + c] + v[2] * m[2 * 3 + c];
}

/* Restates mdpy/utils/pbc.py:38-44 (unwrap_vec): minimum image through the
 * scaled coordinates `vec . pbc_inv`, np.round_ (= round-half-even = rint), back
 * through `. pbc_matrix`. */
static inline void FN(unwrap_vec)(const REAL vec[3], const REAL *pbc, const REAL *pbc_inv,
                                  REAL out[3]) {
    REAL s[3];
    FN(vecmat)(vec, pbc_inv, s);
    for (int c = 0; c < 3; ++c) s[c] = s[c] - (REAL)rint((double)s[c]);
    FN(vecmat)(s, pbc, out);
}

/* Restates mdpy/utils/pbc.py:28-36 (wrap_positions): x - round(x . pbc_inv) . pbc;
 * returns the number of atoms that moved >= 2 images (the reference raises
 * ParticleLossError in that case, pbc.py:30-34); first offender in *first_lost. */
int FN(ora_wrap_positions)(int n, const REAL *pos_in, const REAL *pbc, const REAL *pbc_inv,
                           REAL *pos_out, int *first_lost) {
    int lost = 0;
    for (int i = 0; i < n; ++i) {
        REAL s[3], mv[3], sh[3];
        FN(vecmat)(pos_in + 3 * i, pbc_inv, s);
        int bad = 0;
        for (int c = 0; c < 3; ++c) {
            mv[c] = -(REAL)rint((double)s[c]);
            if (fabs((double)mv[c]) >= 2) bad = 1;
        }
        if (bad) { if (!lost && first_lost) *first_lost = i; ++lost; }
        FN(vecmat)(mv, pbc, sh);
        for (int c = 0; c < 3; ++c) pos_out[3 * i + c] = pos_in[3 * i + c] + sh[c];
    }
    return lost;
}

/* Restates mdpy/core/cell_list.py:83-104 (CellList.update + first half of kernel):
 * shift by the per-axis minimum (:86), cell = floor(pos . cell_inv) (:101), count the
 * fullest cell (:102-106).  Returns max particles per cell. */
int FN(ora_cell_index)(int n, const REAL *pos, const REAL *cell_inv, const int *ncell,
                       int *pci /* [n,3] */) {
    REAL mn[3] = {pos[0], pos[1], pos[2]};
    for (int i = 1; i < n; ++i)
        for (int c = 0; c < 3; ++c)
            if (pos[3 * i + c] < mn[c]) mn[c] = pos[3 * i + c];
    int total = ncell[0] * ncell[1] * ncell[2];
    int *cnt = (int *)calloc((size_t)total, sizeof(int));
    int maxp = 0;
    for (int i = 0; i < n; ++i) {
        REAL p[3] = {pos[3 * i] - mn[0], pos[3 * i + 1] - mn[1], pos[3 * i + 2] - mn[2]};
        REAL s[3];
        FN(vecmat)(p, cell_inv, s);
        int ix = (int)floor((double)s[0]), iy = (int)floor((double)s[1]), iz = (int)floor((double)s[2]);
        pci[3 * i] = ix; pci[3 * i + 1] = iy; pci[3 * i + 2] = iz;
        int lin = (ix * ncell[1] + iy) * ncell[2] + iz;
        if (lin >= 0 && lin < total) { if (++cnt[lin] > maxp) maxp = cnt[lin]; }
    }
    free(cnt);
    return maxp;
}

static inline int FN(in_list)(const int *row, int width, int id) {
    /* the CPU kernel filters `!= -1` first (charmm_nonbonded_constraint.py:75-76),
     * so every non-padding entry counts, wherever it sits in the row */
    for (int k = 0; k < width; ++k)
        if (row[k] == id) return 1;
    return 0;
}

/* Restates CharmmNonbondedConstraint.cpu_kernel
 * (mdpy/constraint/charmm_nonbonded_constraint.py:64-108): for every atom, the 27
 * stencil cells (`>= n` wraps by -n at :79-80, negative indices wrap the python way),
 * every listed j that is not i / not padding / not in bonded[i] (:83), minimum image
 * (:85-88), inclusive cutoff `r <= rc` (:90), 1-4 parameters when j is in scaling[i]
 * (:92-97), Lorentz-Berthelot mixing (:98-101), half the pair force to i and minus
 * half to j (:103-106), half the pair energy (:107).  Every pair is therefore visited
 * twice.  Returns the number of (ordered) in-cutoff visits. */
long long FN(ora_lj_cell)(int n, const REAL *pos, const REAL *params /* [n,4] */, const REAL *pbc,
                          const REAL *pbc_inv, REAL rc, const int *bonded, int wb,
                          const int *scaling, int ws, const int *pci, const int *cell_list,
                          const int *ncell, int P, int i0, int i1, REAL *forces, double *energy) {
    /* [i0, i1): the outer-loop slice this call evaluates (the reference runs [0, n) on one
     * thread; bench.py's reference arm runs disjoint slices on the host threads and adds up) */
    memset(forces, 0, sizeof(REAL) * 3 * (size_t)n);
    double e_acc = 0.0;
    long long visits = 0;
    for (int id1 = i0; id1 < i1; ++id1) {
        const int *b1 = bonded + (size_t)id1 * wb;
        const int *s1 = scaling + (size_t)id1 * ws;
        for (int ti = -1; ti <= 1; ++ti)
        for (int tj = -1; tj <= 1; ++tj)
        for (int tk = -1; tk <= 1; ++tk) {
            int c[3] = {pci[3 * id1] + ti, pci[3 * id1 + 1] + tj, pci[3 * id1 + 2] + tk};
            for (int a = 0; a < 3; ++a) {
                if (c[a] >= ncell[a]) c[a] -= ncell[a];
                if (c[a] < 0) c[a] += ncell[a];   /* python negative index */
            }
            const int *cl = cell_list + (((size_t)c[0] * ncell[1] + c[1]) * ncell[2] + c[2]) * P;
            for (int s = 0; s < P; ++s) {
                int id2 = cl[s];
                if (id2 == -1 || id2 == id1) continue;
                if (FN(in_list)(b1, wb, id2)) continue;
                REAL d[3] = {pos[3 * id2] - pos[3 * id1], pos[3 * id2 + 1] - pos[3 * id1 + 1],
                             pos[3 * id2 + 2] - pos[3 * id1 + 2]};
                REAL v[3];
                FN(unwrap_vec)(d, pbc, pbc_inv, v);
                REAL r = (REAL)sqrt((double)(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
                if (!(r <= rc)) continue;
                ++visits;
                REAL e1, g1, e2, g2;
                if (FN(in_list)(s1, ws, id2)) {
                    e1 = params[4 * id1 + 2]; g1 = params[4 * id1 + 3];
                    e2 = params[4 * id2 + 2]; g2 = params[4 * id2 + 3];
                } else {
                    e1 = params[4 * id1 + 0]; g1 = params[4 * id1 + 1];
                    e2 = params[4 * id2 + 0]; g2 = params[4 * id2 + 1];
                }
                REAL eps = (REAL)sqrt((double)(e1 * e2));
                REAL sig = (g1 + g2) / 2;
                REAL sr = sig / r;
                REAL sr2 = sr * sr, sr6 = sr2 * sr2 * sr2, sr12 = sr6 * sr6;
                REAL fval = -(2 * sr12 - sr6) / r * eps * 24;
                for (int a = 0; a < 3; ++a) {
                    REAL f = v[a] / r * fval / 2;
                    forces[3 * id1 + a] += f;
                    forces[3 * id2 + a] -= f;
                }
                e_acc += (double)(4 * eps * (sr12 - sr6) / 2);
            }
        }
    }
    *energy = e_acc;
    return visits;
}

/* Restates ElectrostaticConstraint.cpu_kernel
 * (mdpy/constraint/electrostatic_constraint.py:52-79): ALL pairs id1 < id2 not in
 * bonded[id1] (:63-64), minimum image (:67-70), f = -q1 q2 / k / r^2 on id1 (:73-75),
 * E += q1 q2 / k / r (:78), k = 4 pi eps0 (:60).  No cutoff, no 1-4 scaling. */
void FN(ora_coulomb_allpairs)(int n, const REAL *pos, const REAL *charges, const int *bonded,
                              int wb, const REAL *pbc, const REAL *pbc_inv, double k, int i0, int i1,
                              REAL *forces, double *energy) {
    memset(forces, 0, sizeof(REAL) * 3 * (size_t)n);
    double e_acc = 0.0;
    for (int id1 = i0; id1 < i1; ++id1) {
        const int *b1 = bonded + (size_t)id1 * wb;
        for (int id2 = id1 + 1; id2 < n; ++id2) {
            if (FN(in_list)(b1, wb, id2)) continue;
            REAL d[3] = {pos[3 * id2] - pos[3 * id1], pos[3 * id2 + 1] - pos[3 * id1 + 1],
                         pos[3 * id2 + 2] - pos[3 * id1 + 2]};
            REAL v[3];
            FN(unwrap_vec)(d, pbc, pbc_inv, v);
            REAL r = (REAL)sqrt((double)(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
            double qq = (double)charges[id1] * (double)charges[id2];
            REAL fval = (REAL)(-qq / k / ((double)r * (double)r));
            for (int a = 0; a < 3; ++a) {
                REAL f = v[a] / r * fval;
                forces[3 * id1 + a] += f;
                forces[3 * id2 + a] -= f;
            }
            e_acc += qq / k / (double)r;
        }
    }
    *energy = e_acc;
}

#undef FN
#undef CAT
#undef CAT_
