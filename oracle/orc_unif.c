/*
 * orc_unif.c -- oracle for piece 3 (uniformized matrices, R^k cache, expm, branch retyping).
 * TEST INFRASTRUCTURE ONLY (see mtt_oracle.h).  Restates, from the Java sources:
 *   evolution/tree/SCMigrationModel.java:150-201 (updateMatrices), :404-473 (getRpowN,
 *       RpowSteadyN)
 *   operators/UniformizationRetypeOperator.java:74-105 (drawEventCount), :116-225
 *       (retypeBranch), :233-275 (getBranchTypeProb)
 * and, from outside the reference tree [EXT], jblas 1.2.5 MatrixFunctions.expm /
 * DoubleMatrix.mmul / Solve.solve (org.jblas:jblas:1.2.5, jar missing from the checkout):
 * the published Pade-13 scaling-and-squaring algorithm.  No reference test pins any value
 * at that boundary => PARITY UNPINNED; tests check against scipy.linalg.expm and
 * numpy.linalg.matrix_power.
 *
 * Arithmetic contract shared bit-for-bit with the CUDA kernels (k3_unif.cu): every matrix
 * product is the naive triple loop  c_ij = (((0 + a_i0*b_0j) + a_i1*b_1j) + ...)  with
 * separately rounded multiply and add (no FMA contraction; this file is compiled with
 * -ffp-contract=off, the kernels use __dmul_rn/__dadd_rn).  Matrices are row-major here;
 * jblas is column-major, which does not change any element's value.
 */
#include "mtt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ small dense helpers */

static void mat_mul(int d, const double *A, const double *B, double *C)
{
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
            double s = 0.0;
            for (int k = 0; k < d; k++)
                s = s + A[i * d + k] * B[k * d + j];
            C[i * d + j] = s;
        }
}

/* ------------------------------------------------------------------ updateMatrices */

/* SCMigrationModel.updateMatrices (:150-201).  Q_ii accumulated by repeated subtraction for
 * j ascending (:166-168); mu = max_i(-Q_ii) (:175-176); R = Q.mul(1.0/mu).add(I) (:183):
 * multiply by the reciprocal, then add. */
void orc_update_matrices(int d, const double *mDense, double *Q, double *Qsym, double *mu,
                         double *muSym, double *R, double *Rsym)
{
    double m = 0.0, ms = 0.0;
    for (int i = 0; i < d; i++) {
        Q[i * d + i] = 0.0;
        Qsym[i * d + i] = 0.0;
        for (int j = 0; j < d; j++) {
            if (i != j) {
                Q[i * d + j] = mDense[i * d + j];
                Q[i * d + i] = Q[i * d + i] - Q[i * d + j];
                Qsym[i * d + j] = 0.5 * (mDense[i * d + j] + mDense[j * d + i]);
                Qsym[i * d + i] = Qsym[i * d + i] - Qsym[i * d + j];
            }
        }
        if (-Q[i * d + i] > m)
            m = -Q[i * d + i];
        if (-Qsym[i * d + i] > ms)
            ms = -Qsym[i * d + i];
    }
    double rm = 1.0 / m, rms = 1.0 / ms;
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
            R[i * d + j] = Q[i * d + j] * rm + (i == j ? 1.0 : 0.0);
            Rsym[i * d + j] = Qsym[i * d + j] * rms + (i == j ? 1.0 : 0.0);
        }
    *mu = m;
    *muSym = ms;
}

/* ------------------------------------------------------------------ R^k cache */

struct orc_rpow {
    int d;
    double *R;
    double *pow;   /* size * d*d */
    int size, cap;
    int steady;
};

orc_rpow *orc_rpow_new(int d, const double *R)
{
    orc_rpow *c = malloc(sizeof(orc_rpow));
    c->d = d;
    c->R = malloc(sizeof(double) * d * d);
    memcpy(c->R, R, sizeof(double) * d * d);
    c->cap = 16;
    c->pow = malloc(sizeof(double) * d * d * c->cap);
    for (int i = 0; i < d * d; i++)               /* RpowN = [I]  (:193) */
        c->pow[i] = (i / d == i % d) ? 1.0 : 0.0;
    c->size = 1;
    c->steady = 0;
    return c;
}

void orc_rpow_free(orc_rpow *c)
{
    if (!c)
        return;
    free(c->R);
    free(c->pow);
    free(c);
}

/* SCMigrationModel.getRpowN (:404-452) */
const double *orc_rpow_get(orc_rpow *c, int n)
{
    int dd = c->d * c->d;
    if (n >= c->size) {
        if (c->steady)                                          /* :422-426 */
            return c->pow + (size_t)(c->size - 1) * dd;
        int startN = c->size;
        for (int i = startN; i <= n; i++) {                     /* :428-448 */
            if (c->size == c->cap) {
                c->cap *= 2;
                c->pow = realloc(c->pow, sizeof(double) * dd * c->cap);
            }
            mat_mul(c->d, c->pow + (size_t)(i - 1) * dd, c->R, c->pow + (size_t)i * dd);
            c->size = i + 1;
            if (i % 10 == 0) {                                  /* :435-447 */
                double maxDiff = 0.0;
                const double *a = c->pow + (size_t)i * dd, *b = c->pow + (size_t)(i - 1) * dd;
                for (int e = 0; e < dd; e++) {
                    double v = fabs(a[e] - b[e]);
                    /* Java Math.max propagates NaN */
                    if (isnan(v) || isnan(maxDiff))
                        maxDiff = NAN;
                    else if (v > maxDiff)
                        maxDiff = v;
                }
                if (!(maxDiff > 0)) {
                    c->steady = 1;
                    return c->pow + (size_t)i * dd;
                }
            }
        }
    }
    return c->pow + (size_t)n * dd;                             /* :451 */
}

int orc_rpow_steady_n(const orc_rpow *c)
{
    return c->steady ? c->size : -1;                            /* :460-473 */
}

int orc_rpow_size(const orc_rpow *c)
{
    return c->size;
}

/* ------------------------------------------------------------------ expm (jblas) [EXT] */

/* LU with partial pivoting, column by column (LAPACK dgesv semantics: pivot = first row of
 * maximal |value| in the column; unblocked right-looking elimination), then forward /
 * backward substitution for every right-hand-side column.  D and N row-major, F = D^-1 N. */
static void lu_solve(int d, double *D, double *N)
{
    int piv[ORC_MAX_D];
    for (int k = 0; k < d; k++) {
        int p = k;
        double best = fabs(D[k * d + k]);
        for (int i = k + 1; i < d; i++) {
            double v = fabs(D[i * d + k]);
            if (v > best) {
                best = v;
                p = i;
            }
        }
        piv[k] = p;
        if (p != k) {
            for (int j = 0; j < d; j++) {
                double tmp = D[k * d + j]; D[k * d + j] = D[p * d + j]; D[p * d + j] = tmp;
                tmp = N[k * d + j]; N[k * d + j] = N[p * d + j]; N[p * d + j] = tmp;
            }
        }
        double pivot = D[k * d + k];
        for (int i = k + 1; i < d; i++) {
            double l = D[i * d + k] / pivot;
            D[i * d + k] = l;
            for (int j = k + 1; j < d; j++)
                D[i * d + j] = D[i * d + j] - l * D[k * d + j];
            for (int j = 0; j < d; j++)                 /* forward substitution fused */
                N[i * d + j] = N[i * d + j] - l * N[k * d + j];
        }
    }
    (void)piv;
    for (int k = d - 1; k >= 0; k--) {                  /* back substitution */
        double pivot = D[k * d + k];
        for (int j = 0; j < d; j++)
            N[k * d + j] = N[k * d + j] / pivot;
        for (int i = 0; i < k; i++) {
            double u = D[i * d + k];
            for (int j = 0; j < d; j++)
                N[i * d + j] = N[i * d + j] - u * N[k * d + j];
        }
    }
}

/* Pade-13 coefficients as printed in jblas MatrixFunctions.expm [EXT, recalled] */
static const double PADE[14] = {
    1.0, 0.5, 0.12, 0.01833333333333333, 0.0019927536231884053, 1.630434782608695E-4,
    1.0351966873706E-5, 5.175983436853E-7, 2.0431513566525E-8, 6.306022705717593E-10,
    1.4837700484041396E-11, 2.5291534915979653E-13, 2.8101705462199615E-15,
    1.5440497506703084E-17
};

void orc_expm(int d, const double *A, double *out)
{
    int dd = d * d;
    double As[ORC_MAX_D * ORC_MAX_D] = { 0 }, A2[ORC_MAX_D * ORC_MAX_D], A4[ORC_MAX_D * ORC_MAX_D],
        A6[ORC_MAX_D * ORC_MAX_D], U[ORC_MAX_D * ORC_MAX_D], V[ORC_MAX_D * ORC_MAX_D],
        T[ORC_MAX_D * ORC_MAX_D], AV[ORC_MAX_D * ORC_MAX_D];

    /* j = max(0, 1 + floor(log(normmax(A)) / log(2))) */
    double nm = 0.0;
    for (int e = 0; e < dd; e++)
        if (fabs(A[e]) > nm)
            nm = fabs(A[e]);
    int j = 0;
    if (nm > 0.0) {
        double f = floor(log(nm) / log(2.0));
        if (f > -1.0)
            j = 1 + (int)f;
    }
    double scale = pow(2.0, (double)j);
    for (int e = 0; e < dd; e++)
        As[e] = A[e] / scale;

    mat_mul(d, As, As, A2);
    mat_mul(d, A2, A2, A4);
    mat_mul(d, A4, A2, A6);

    /* U = c0 I + c2 A2 + c4 A4 + (c6 I + c8 A2 + c10 A4 + c12 A6) A6   (left to right) */
    for (int e = 0; e < dd; e++) {
        double eye = (e / d == e % d) ? 1.0 : 0.0;
        T[e] = ((eye * PADE[6] + A2[e] * PADE[8]) + A4[e] * PADE[10]) + A6[e] * PADE[12];
    }
    mat_mul(d, T, A6, U);
    for (int e = 0; e < dd; e++) {
        double eye = (e / d == e % d) ? 1.0 : 0.0;
        U[e] = ((eye * PADE[0] + A2[e] * PADE[2]) + A4[e] * PADE[4]) + U[e];
    }
    /* V = c1 I + c3 A2 + c5 A4 + (c7 I + c9 A2 + c11 A4 + c13 A6) A6 */
    for (int e = 0; e < dd; e++) {
        double eye = (e / d == e % d) ? 1.0 : 0.0;
        T[e] = ((eye * PADE[7] + A2[e] * PADE[9]) + A4[e] * PADE[11]) + A6[e] * PADE[13];
    }
    mat_mul(d, T, A6, V);
    for (int e = 0; e < dd; e++) {
        double eye = (e / d == e % d) ? 1.0 : 0.0;
        V[e] = ((eye * PADE[1] + A2[e] * PADE[3]) + A4[e] * PADE[5]) + V[e];
    }

    mat_mul(d, As, V, AV);
    double Nn[ORC_MAX_D * ORC_MAX_D], D[ORC_MAX_D * ORC_MAX_D];
    for (int e = 0; e < dd; e++) {
        Nn[e] = U[e] + AV[e];
        D[e] = U[e] - AV[e];
    }
    lu_solve(d, D, Nn);                         /* F = D^-1 N, left in Nn */

    for (int k = 0; k < j; k++) {               /* square j times */
        mat_mul(d, Nn, Nn, T);
        memcpy(Nn, T, sizeof(double) * dd);
    }
    memcpy(out, Nn, sizeof(double) * dd);
}

/* ------------------------------------------------------------------ retypeBranch */

static int cmp_double(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* drawEventCount (:74-105).  Returns status; *nVirtOut valid on ORC_ST_OK.  On
 * ORC_ST_NEEDS_POISSON_TAIL *nVirtOut = thresh (:96). */
static int draw_event_count(const orc_unif_model *um, int typeStart, int typeEnd, double muL,
                            double Pba, const double *u, int64_t nU, int64_t *pos,
                            int *nVirtOut)
{
    int d = um->d;
    int nVirt = 0;
    if (*pos >= nU)
        return ORC_ST_NEED_MORE_UNIFORMS;
    double uu = u[(*pos)++];                                    /* :79 */
    double P_low_given_ab = 0.0;
    double acc = -muL - log(Pba);                               /* :81 */
    double log_muL = log(muL);                                  /* :82 */

    do {
        const double *Rn = orc_rpow_get(um->rpow, nVirt);
        P_low_given_ab += exp(log(Rn[typeStart * d + typeEnd]) + acc);   /* :86 */
        if (P_low_given_ab > uu) {                              /* :88-89 */
            *nVirtOut = nVirt;
            return ORC_ST_OK;
        }
        nVirt += 1;
        acc += log_muL - log((double)nVirt);                    /* :92 */
    } while (orc_rpow_steady_n(um->rpow) < 0 || nVirt < orc_rpow_steady_n(um->rpow));

    *nVirtOut = nVirt;                                          /* thresh (:96) */
    return ORC_ST_NEEDS_POISSON_TAIL;
}

int orc_retype_branch(const orc_unif_model *um, double tBottom, double tTop, int typeBottom,
                      int typeTop, const double *u, int64_t nU, int nVirtForced, int cap,
                      int *nChg, double *chgTime, int32_t *chgType, double *logProb,
                      int64_t *uUsed, int *nVirtOut)
{
    int d = um->d;
    int64_t pos = 0;
    *nChg = 0;
    *logProb = -INFINITY;
    *uUsed = 0;
    *nVirtOut = -1;

    double L = tTop - tBottom;                                  /* :124 */
    double muL = um->mu * L;                                    /* :130 */

    double QL[ORC_MAX_D * ORC_MAX_D], E[ORC_MAX_D * ORC_MAX_D];
    for (int e = 0; e < d * d; e++)
        QL[e] = um->Q[e] * L;                                   /* Q.mul(L) :133-134 */
    orc_expm(d, QL, E);
    double Pba = E[typeBottom * d + typeTop];

    if (Pba == 0.0)                                             /* :137-138 */
        return ORC_ST_NO_VALID_PATH;
    if (Pba > 1.0 || Pba < 0.0)                                 /* :141-144 */
        return ORC_ST_PBA_RUBBISH;

    int nVirt = 0;
    if (nVirtForced >= 0) {
        nVirt = nVirtForced;
    } else {
        int st = draw_event_count(um, typeBottom, typeTop, muL, Pba, u, nU, &pos, &nVirt);
        *uUsed = pos;
        *nVirtOut = nVirt;
        if (st != ORC_ST_OK)
            return st;
    }
    *nVirtOut = nVirt;

    if (pos + 2 * (int64_t)nVirt > nU)
        return ORC_ST_NEED_MORE_UNIFORMS;

    double *times = malloc(sizeof(double) * (nVirt > 0 ? nVirt : 1));
    int *types = malloc(sizeof(int) * (nVirt > 0 ? nVirt : 1));
    for (int i = 0; i < nVirt; i++) {
        double prod = u[pos++] * L;                             /* :156, multiply ... */
        times[i] = prod + tBottom;                              /* ... then add */
    }
    qsort(times, nVirt, sizeof(double), cmp_double);            /* :157 */

    int prevType = typeBottom;
    int status = ORC_ST_OK;
    for (int i = 1; i <= nVirt; i++) {                          /* :163-194 */
        const double *Rhi = orc_rpow_get(um->rpow, nVirt - i + 1);
        double u2 = u[pos++] * Rhi[prevType * d + typeTop];     /* :165-166 */
        const double *Rlo = orc_rpow_get(um->rpow, nVirt - i);
        int c;
        int fellThrough = 1;
        for (c = 0; c < d; c++) {
            double prod = um->R[prevType * d + c] * Rlo[c * d + typeTop];
            u2 = u2 - prod;                                     /* :170-171 */
            if (u2 < 0.0) {
                fellThrough = 0;
                break;
            }
        }
        if (fellThrough) {                                      /* :179-190 */
            status = ORC_ST_FB_FELL_THROUGH;
            break;
        }
        types[i - 1] = c;
        prevType = c;
    }
    *uUsed = pos;
    if (status != ORC_ST_OK) {
        free(times); free(types);
        return status;
    }

    double lp = 0.0;                                            /* :196-221 */
    prevType = typeBottom;
    double prevTime = tBottom;
    int n = 0;
    for (int i = 0; i < nVirt; i++) {
        if (types[i] != prevType) {
            if (n < cap) {
                chgTime[n] = times[i];
                chgType[n] = types[i];
            }
            n++;
            lp += um->Q[prevType * d + prevType] * (times[i] - prevTime)
                + log(um->Q[prevType * d + types[i]]);
            prevType = types[i];
            prevTime = times[i];
        }
    }
    lp += um->Q[prevType * d + prevType] * (tTop - prevTime);
    lp -= log(Pba);
    *nChg = n;
    *logProb = lp;
    free(times); free(types);
    return ORC_ST_OK;
}

/* getBranchTypeProb (:233-275) */
int orc_branch_type_prob(const orc_unif_model *um, double tBottom, double tTop, int typeBottom,
                         int nChg, const double *chgTime, const int32_t *chgType, int typeTop,
                         double *logProb)
{
    int d = um->d;
    double lp = 0.0;
    double L = tTop - tBottom;
    double lastTime = tBottom;
    int lastCol = typeBottom;
    for (int i = 0; i < nChg; i++) {                            /* :249-258 */
        double thisTime = chgTime[i];
        int thisCol = chgType[i];
        lp += (thisTime - lastTime) * um->Q[lastCol * d + lastCol]
            + log(um->Q[lastCol * d + thisCol]);
        lastTime = thisTime;
        lastCol = thisCol;
    }
    lp += (tTop - lastTime) * um->Q[lastCol * d + lastCol];     /* :259 */

    double QL[ORC_MAX_D * ORC_MAX_D], E[ORC_MAX_D * ORC_MAX_D];
    for (int e = 0; e < d * d; e++)
        QL[e] = um->Q[e] * L;
    orc_expm(d, QL, E);                                         /* :263-264 */
    double Pba = E[typeBottom * d + typeTop];
    if (Pba > 1.0 || Pba < 0.0) {                               /* :267-270 */
        *logProb = -INFINITY;
        return ORC_ST_PBA_RUBBISH;
    }
    lp -= log(Pba);                                             /* :272 */
    *logProb = lp;
    return ORC_ST_OK;
}
