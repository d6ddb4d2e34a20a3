/*
 * orc_peel.c -- oracle for piece 2 (Felsenstein peeling as BEAST 2's TreeLikelihood runs it).
 * TEST INFRASTRUCTURE ONLY (see mtt_oracle.h).
 *
 * The arithmetic of this piece is NOT in the reference tree: the reference only wires BEAST 2
 * core's beast.base.evolution.likelihood.TreeLikelihood in XML
 * (examples/structuredCoalescent.xml:568-572, fxtemplates/MultiTypeTree.xml:184-197).
 * Third-party dependency: BEAST.base >= 2.7.2 (version.xml:2-3; built from GitHub master by
 * build.xml:27-30, unpinned).  This file restates the published algorithm of
 * TreeLikelihood / BeerLikelihoodCore / HKY / SiteModel [EXT, recalled -- SURVEY Appendix C
 * and F.3].  No reference test pins any peeling value => PARITY UNPINNED; tests check this
 * file against brute-force enumeration of internal states with scipy.linalg.expm matrices.
 *
 * Layouts follow BeerLikelihoodCore: partials[c*P*S + p*S + i], matrices[c*S*S + i*S + j]
 * with row i = parent state; 4 states.
 */
#include "mtt_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define S 4

/* HKY.setupMatrix + getTransitionProbabilities [EXT]; state order A,C,G,T */
void orc_hky_matrix(double kappa, const double *freqs, double distance, double *matrix)
{
    double freqA = freqs[0], freqC = freqs[1], freqG = freqs[2], freqT = freqs[3];
    double freqR = freqA + freqG;
    double freqY = freqC + freqT;

    double tab1A = freqA * ((1 / freqR) - 1);
    double tab3A = freqA / freqR;
    double tab2A = 1 - tab3A;
    double tab1C = freqC * ((1 / freqY) - 1);
    double tab3C = freqC / freqY;
    double tab2C = 1 - tab3C;
    double tab1G = freqG * ((1 / freqR) - 1);
    double tab3G = tab2A;
    double tab2G = tab3A;
    double tab1T = freqT * ((1 / freqY) - 1);
    double tab3T = tab2C;
    double tab2T = tab3C;

    double beta = -1.0 / (2.0 * (freqR * freqY + kappa * (freqA * freqG + freqC * freqT)));
    double A_R = 1.0 + freqR * (kappa - 1);
    double A_Y = 1.0 + freqY * (kappa - 1);

    double xx = beta * distance;
    double bbR = exp(xx * A_R);
    double bbY = exp(xx * A_Y);
    double aa = exp(xx);
    double oneminusa = 1 - aa;

    double t1Aaa = tab1A * aa;
    matrix[0] = freqA + t1Aaa + (tab2A * bbR);
    matrix[1] = freqC * oneminusa;
    double t1Gaa = tab1G * aa;
    matrix[2] = freqG + t1Gaa - (tab3G * bbR);
    matrix[3] = freqT * oneminusa;

    matrix[4] = freqA * oneminusa;
    double t1Caa = tab1C * aa;
    matrix[5] = freqC + t1Caa + (tab2C * bbY);
    matrix[6] = freqG * oneminusa;
    double t1Taa = tab1T * aa;
    matrix[7] = freqT + t1Taa - (tab3T * bbY);

    matrix[8] = freqA + t1Aaa - (tab3A * bbR);
    matrix[9] = matrix[1];
    matrix[10] = freqG + t1Gaa + (tab2G * bbR);
    matrix[11] = matrix[3];

    matrix[12] = matrix[4];
    matrix[13] = freqC + t1Caa - (tab3C * bbY);
    matrix[14] = matrix[6];
    matrix[15] = freqT + t1Taa + (tab2T * bbY);
}

typedef struct {
    int nLeaves, nNodes, P, C;
    const int32_t *tipStates;
    const int32_t *left, *right;
    const double *mats;       /* [node][C][16] */
    double *partials;         /* [node - nLeaves][C*P*S] */
    double *scaling;          /* [node - nLeaves][P] (log scale factors), NULL = off */
} peel_ctx;

static double *node_partials(peel_ctx *x, int node)
{
    return x->partials + (size_t)(node - x->nLeaves) * x->C * x->P * S;
}

/* BeerLikelihoodCore.calculateStatesStatesPruning [EXT] */
static void states_states(peel_ctx *x, int n1, int n2, int n3)
{
    const int32_t *s1 = x->tipStates + (size_t)n1 * x->P, *s2 = x->tipStates + (size_t)n2 * x->P;
    double *p3 = node_partials(x, n3);
    int v = 0;
    for (int l = 0; l < x->C; l++) {
        const double *m1 = x->mats + ((size_t)n1 * x->C + l) * 16;
        const double *m2 = x->mats + ((size_t)n2 * x->C + l) * 16;
        for (int k = 0; k < x->P; k++) {
            int state1 = s1[k], state2 = s2[k];
            int w = 0;
            if (state1 < S && state2 < S) {
                for (int i = 0; i < S; i++) {
                    p3[v++] = m1[w + state1] * m2[w + state2];
                    w += S;
                }
            } else if (state1 < S) {
                for (int i = 0; i < S; i++) {
                    p3[v++] = m1[w + state1];
                    w += S;
                }
            } else if (state2 < S) {
                for (int i = 0; i < S; i++) {
                    p3[v++] = m2[w + state2];
                    w += S;
                }
            } else {
                for (int i = 0; i < S; i++)
                    p3[v++] = 1.0;
            }
        }
    }
}

/* BeerLikelihoodCore.calculateStatesPartialsPruning [EXT]: n1 tip, n2 internal */
static void states_partials(peel_ctx *x, int n1, int n2, int n3)
{
    const int32_t *s1 = x->tipStates + (size_t)n1 * x->P;
    const double *p2 = node_partials(x, n2);
    double *p3 = node_partials(x, n3);
    int u = 0, v = 0;
    for (int l = 0; l < x->C; l++) {
        const double *m1 = x->mats + ((size_t)n1 * x->C + l) * 16;
        const double *m2 = x->mats + ((size_t)n2 * x->C + l) * 16;
        for (int k = 0; k < x->P; k++) {
            int state1 = s1[k];
            int w = 0;
            if (state1 < S) {
                for (int i = 0; i < S; i++) {
                    double tmp = m1[w + state1];
                    double sum = 0.0;
                    for (int j = 0; j < S; j++) {
                        sum += m2[w] * p2[v + j];
                        w++;
                    }
                    p3[u++] = tmp * sum;
                }
            } else {
                for (int i = 0; i < S; i++) {
                    double sum = 0.0;
                    for (int j = 0; j < S; j++) {
                        sum += m2[w] * p2[v + j];
                        w++;
                    }
                    p3[u++] = sum;
                }
            }
            v += S;
        }
    }
}

/* BeerLikelihoodCore.calculatePartialsPartialsPruning [EXT] */
static void partials_partials(peel_ctx *x, int n1, int n2, int n3)
{
    const double *p1 = node_partials(x, n1), *p2 = node_partials(x, n2);
    double *p3 = node_partials(x, n3);
    int u = 0, v = 0;
    for (int l = 0; l < x->C; l++) {
        const double *m1 = x->mats + ((size_t)n1 * x->C + l) * 16;
        const double *m2 = x->mats + ((size_t)n2 * x->C + l) * 16;
        for (int k = 0; k < x->P; k++) {
            int w = 0;
            for (int i = 0; i < S; i++) {
                double sum1 = 0.0, sum2 = 0.0;
                for (int j = 0; j < S; j++) {
                    sum1 += m1[w] * p1[v + j];
                    sum2 += m2[w] * p2[v + j];
                    w++;
                }
                p3[u++] = sum1 * sum2;
            }
            v += S;
        }
    }
}

/* BeerLikelihoodCore.scalePartials [EXT]: per pattern, max over categories and states;
 * rescale when the max falls below the threshold 1e-100. */
static void scale_partials(peel_ctx *x, int n3)
{
    double *p3 = node_partials(x, n3);
    double *sf = x->scaling + (size_t)(n3 - x->nLeaves) * x->P;
    int u = 0;
    for (int i = 0; i < x->P; i++) {
        double scaleFactor = 0.0;
        int v = u;
        for (int k = 0; k < x->C; k++) {
            for (int j = 0; j < S; j++) {
                if (p3[v] > scaleFactor)
                    scaleFactor = p3[v];
                v++;
            }
            v += (x->P - 1) * S;
        }
        if (scaleFactor < 1.0E-100) {
            v = u;
            for (int k = 0; k < x->C; k++) {
                for (int j = 0; j < S; j++) {
                    p3[v] /= scaleFactor;
                    v++;
                }
                v += (x->P - 1) * S;
            }
            sf[i] = log(scaleFactor);
        } else {
            sf[i] = 0.0;
        }
        u += S;
    }
}

/* TreeLikelihood.traverse [EXT]: post-order, child 1 then child 2 then the node. */
static void traverse(peel_ctx *x, int root)
{
    /* iterative post-order to survive caterpillar trees of 5000 tips */
    int *stack = malloc(sizeof(int) * 2 * x->nNodes);
    char *state = calloc(x->nNodes, 1);
    int sp = 0;
    stack[sp++] = root;
    while (sp > 0) {
        int node = stack[sp - 1];
        if (node < x->nLeaves) {
            sp--;
            continue;
        }
        if (state[node] == 0) {
            state[node] = 1;
            stack[sp++] = x->right[node];
            stack[sp++] = x->left[node];       /* left is processed first */
            continue;
        }
        sp--;
        int c1 = x->left[node], c2 = x->right[node];
        int t1 = c1 < x->nLeaves, t2 = c2 < x->nLeaves;
        if (t1 && t2)
            states_states(x, c1, c2, node);
        else if (t1)
            states_partials(x, c1, c2, node);
        else if (t2)
            states_partials(x, c2, c1, node);
        else
            partials_partials(x, c1, c2, node);
        if (x->scaling)
            scale_partials(x, node);
    }
    free(stack);
    free(state);
}

static double likelihood_core(int nLeaves, int P, int C, const int32_t *tipStates,
                              const int32_t *weights, const int32_t *parent, const double *mats,
                              const double *catProps, const double *freqs, double pInv,
                              int scaling, double *patternLogL)
{
    int nNodes = 2 * nLeaves - 1;
    int32_t *left = malloc(sizeof(int32_t) * nNodes), *right = malloc(sizeof(int32_t) * nNodes);
    int root = -1;
    for (int i = 0; i < nNodes; i++)
        left[i] = right[i] = -1;
    for (int i = 0; i < nNodes; i++) {
        int p = parent[i];
        if (p < 0) { root = i; continue; }
        if (left[p] < 0) left[p] = i; else right[p] = i;
    }
    peel_ctx x = { nLeaves, nNodes, P, C, tipStates, left, right, mats, NULL, NULL };
    x.partials = malloc(sizeof(double) * (size_t)(nLeaves - 1) * C * P * S);
    if (scaling)
        x.scaling = calloc((size_t)(nLeaves - 1) * P, sizeof(double));

    traverse(&x, root);

    /* integratePartials [EXT] */
    const double *rp = node_partials(&x, root);
    double *rootPartials = malloc(sizeof(double) * P * S);
    for (int e = 0; e < P * S; e++)
        rootPartials[e] = rp[e] * catProps[0];
    for (int l = 1; l < C; l++)
        for (int e = 0; e < P * S; e++)
            rootPartials[e] += rp[(size_t)l * P * S + e] * catProps[l];

    /* proportion invariant [EXT]: TreeLikelihood.calcConstantPatternIndices -- pattern p is
     * constant for state k when every unambiguous tip shows k (ambiguous codes ignored). */
    if (pInv > 0) {
        for (int p = 0; p < P; p++) {
            int inv[S] = { 1, 1, 1, 1 };
            for (int tip = 0; tip < nLeaves; tip++) {
                int st = tipStates[(size_t)tip * P + p];
                if (st < S)
                    for (int k = 0; k < S; k++)
                        if (k != st)
                            inv[k] = 0;
            }
            for (int k = 0; k < S; k++)
                if (inv[k])
                    rootPartials[p * S + k] += pInv;
        }
    }

    /* calculateLogLikelihoods [EXT] + TreeLikelihood.calcLogP */
    double logP = 0.0;
    for (int k = 0; k < P; k++) {
        double sum = 0.0;
        for (int i = 0; i < S; i++)
            sum += freqs[i] * rootPartials[k * S + i];
        double ls = 0.0;
        if (scaling)
            for (int n = 0; n < nLeaves - 1; n++)
                ls += x.scaling[(size_t)n * P + k];
        double ll = log(sum) + ls;
        if (patternLogL)
            patternLogL[k] = ll;
        logP += ll * weights[k];
    }
    free(rootPartials);
    free(x.partials);
    free(x.scaling);
    free(left);
    free(right);
    return logP;
}

double orc_tree_likelihood_mats(int nLeaves, int P, int C, const int32_t *tipStates,
                                const int32_t *weights, const int32_t *parent,
                                const double *mats, const double *catProps, const double *freqs,
                                double pInv, int useScaling, double *patternLogL)
{
    if (useScaling == 2) {
        /* TreeLikelihood.calculateLogP [EXT]: scaling is switched on after a -inf result */
        double lp = likelihood_core(nLeaves, P, C, tipStates, weights, parent, mats, catProps,
                                    freqs, pInv, 0, patternLogL);
        if (lp == -INFINITY)
            lp = likelihood_core(nLeaves, P, C, tipStates, weights, parent, mats, catProps,
                                 freqs, pInv, 1, patternLogL);
        return lp;
    }
    return likelihood_core(nLeaves, P, C, tipStates, weights, parent, mats, catProps, freqs,
                           pInv, useScaling, patternLogL);
}

double orc_tree_likelihood(int nLeaves, int P, int C, const int32_t *tipStates,
                           const int32_t *weights, const int32_t *parent, const double *height,
                           const double *catRates, const double *catProps, const double *freqs,
                           double kappa, double branchRate, double pInv, int useScaling,
                           double *patternLogL)
{
    int nNodes = 2 * nLeaves - 1;
    double *mats = calloc((size_t)nNodes * C * 16, sizeof(double));
    /* TreeLikelihood.traverse [EXT]: jointBranchRate = categoryRate * branchRate;
     * distance = (parentHeight - nodeHeight) * jointBranchRate */
    for (int i = 0; i < nNodes; i++) {
        if (parent[i] < 0)
            continue;
        for (int c = 0; c < C; c++) {
            double jointBranchRate = catRates[c] * branchRate;
            double distance = (height[parent[i]] - height[i]) * jointBranchRate;
            orc_hky_matrix(kappa, freqs, distance, mats + ((size_t)i * C + c) * 16);
        }
    }
    double lp = orc_tree_likelihood_mats(nLeaves, P, C, tipStates, weights, parent, mats,
                                         catProps, freqs, pInv, useScaling, patternLogL);
    free(mats);
    return lp;
}

typedef struct {
    int B, nLeaves, P, C, useScaling, tid, nThreads;
    const int32_t *tipStates, *weights, *parent;
    const double *height, *catRates, *catProps, *freqs;
    double kappa, branchRate, pInv;
    double *logP;
} peel_job;

static void *peel_worker(void *arg)
{
    peel_job *j = arg;
    int nNodes = 2 * j->nLeaves - 1;
    for (int b = j->tid; b < j->B; b += j->nThreads)
        j->logP[b] = orc_tree_likelihood(j->nLeaves, j->P, j->C, j->tipStates, j->weights,
                                         j->parent + (size_t)b * nNodes,
                                         j->height + (size_t)b * nNodes, j->catRates,
                                         j->catProps, j->freqs, j->kappa, j->branchRate,
                                         j->pInv, j->useScaling, NULL);
    return NULL;
}

void orc_tree_likelihood_batch(int B, int nLeaves, int P, int C, const int32_t *tipStates,
                               const int32_t *weights, const int32_t *parent,
                               const double *height, const double *catRates,
                               const double *catProps, const double *freqs, double kappa,
                               double branchRate, double pInv, int useScaling, int nThreads,
                               double *logP)
{
    if (nThreads < 1)
        nThreads = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * nThreads);
    peel_job *jobs = malloc(sizeof(peel_job) * nThreads);
    for (int i = 0; i < nThreads; i++) {
        peel_job j = { B, nLeaves, P, C, useScaling, i, nThreads, tipStates, weights, parent,
                       height, catRates, catProps, freqs, kappa, branchRate, pInv, logP };
        jobs[i] = j;
        if (nThreads == 1)
            peel_worker(&jobs[i]);
        else
            pthread_create(&th[i], NULL, peel_worker, &jobs[i]);
    }
    if (nThreads > 1)
        for (int i = 0; i < nThreads; i++)
            pthread_join(th[i], NULL);
    free(th);
    free(jobs);
}
