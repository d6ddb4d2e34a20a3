/* orc_untyped.c -- TEST INFRASTRUCTURE ONLY (see mtt_oracle.h).
 *
 * CPU restatement of distributions/StructuredCoalescentUntypedTreeDensity.java:
 *   calculateLogP        :116-207   importance-sampling estimate over nParticles typings
 *   colourFirstLineage   :216-290   unconditioned CTMC walk from the first leaf to the root
 *   colourLineage        :292-316   leaf -> first already-typed ancestor
 *   addTypeChanges       :386-512   uniformization + forward-backward path (the same sampler as
 *                                   UniformizationRetypeOperator.retypeBranch -> orc_retype_branch)
 *                                   plus typing of the internal nodes the path passes
 *
 * Randomness: the reference draws from BEAST's global Randomizer.  Here one particle consumes
 *   - a leaf order (Randomizer.shuffled, :129) supplied by the caller, and
 *   - uniforms u[0..nU) in the reference's order: per first-lineage step one for
 *     nextExponential (-log(1-u)/aTot, [EXT]) and one for the type choice (:265); per other
 *     lineage the stream of retypeBranch.
 * PARITY UNPINNED by any reference test (the reference only has statistical MCMC tests, and this
 * density is itself a Monte-Carlo estimate); the restatement is checked against the exact
 * two-leaf marginal density (tests/test_untyped_gpu.py).
 * The steady-state Poisson tail (:371-377) is drawn from the same window by orc_poisson_window
 * (multiplicative method below 40, inversion by summation above): distributionally exact,
 * not stream-compatible with a JVM -- BEAST's own nextPoisson is [EXT] and unpinned.
 */
#include "mtt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { UK_SAMPLE = 0, UK_COALESCE = 1, UK_MIGRATE = 2 };

typedef struct {
    double time;
    int kind, type, dest, seq;
} u_event;

int orc_poisson_window(double lam, const double *u, int64_t nU, int64_t *pos, int *out)
{
    if (lam < 40.0) {
        double p = exp(-lam), r = 1.0;
        int n = 0;
        for (;;) {
            if (*pos >= nU)
                return ORC_ST_NEED_MORE_UNIFORMS;
            r *= u[(*pos)++];
            if (r < p) {
                *out = n;
                return ORC_ST_OK;
            }
            n++;
        }
    }
    if (*pos >= nU)
        return ORC_ST_NEED_MORE_UNIFORMS;
    double uu = u[(*pos)++];
    int n = 0;
    double logp = -lam, cdf = exp(logp), loglam = log(lam);
    while (cdf < uu && n < 10.0 * lam + 100.0) {
        n++;
        logp += loglam - log((double)n);
        cdf += exp(logp);
    }
    *out = n;
    return ORC_ST_OK;
}

static int cmp_event(const void *a, const void *b)
{
    const u_event *x = a, *y = b;
    if (x->time < y->time)
        return -1;
    if (x->time > y->time)
        return 1;
    return (x->seq > y->seq) - (x->seq < y->seq);      /* Collections.sort is stable (:143) */
}

typedef struct {
    u_event *ev;
    int n, cap;
} ev_list;

static void push(ev_list *L, int kind, int type, int dest, double time)
{
    if (L->n == L->cap) {
        L->cap = L->cap ? 2 * L->cap : 256;
        L->ev = realloc(L->ev, sizeof(u_event) * (size_t)L->cap);
    }
    u_event *e = &L->ev[L->n];
    e->kind = kind;
    e->type = type;
    e->dest = dest;
    e->time = time;
    e->seq = L->n;
    L->n++;
}

/* One particle.  Returns ORC_ST_OK or ORC_ST_NEED_MORE_UNIFORMS.  *logW is the particle's log
 * weight (:145-186 minus the proposal log-probabilities :134-137); may be +-inf / NaN exactly
 * where the Java arithmetic produces them. */
/* The typing is also returned when chgNodeOut != NULL: the changes in the order they are
 * added (node below the change, type entered, time), *nChgOut of them (at most chgCap stored) --
 * the typed tree MultiTypeTreeFromUntypedNewick.java:104-134 builds with the same colouring; a
 * lineage without a valid path then ends the particle with that lineage's status (:128-131). */
static int particle_impl(const orc_unif_model *um, const double *N, int nNodes,
                         const int32_t *parent, const double *height, const uint8_t *leafType,
                         const int32_t *leafOrder, const double *u, int64_t nU, double *logW,
                         int64_t *uUsed, int32_t *nodeTypeOut, int32_t *nEventsOut, int chgCap,
                         int32_t *nChgOut, int32_t *chgNodeOut, int32_t *chgTypeOut,
                         double *chgTimeOut)
{
    const int d = um->d;
    const int nLeaves = (nNodes + 1) / 2;
    int *nodeType = malloc(sizeof(int) * (size_t)nNodes);
    for (int i = 0; i < nNodes; i++)
        nodeType[i] = (i < nLeaves) ? (int)leafType[i] : -1;      /* :124-126 */
    ev_list L = {0, 0, 0};
    int64_t pos = 0;
    double w = 0.0;
    int rc = ORC_ST_OK;
    int nAdded = 0;
    const int cap = 4096;
    double *ct = malloc(sizeof(double) * cap);
    int32_t *cty = malloc(sizeof(int32_t) * cap);

    for (int li = 0; li < nLeaves && rc == ORC_ST_OK; li++) {
        const int leaf = leafOrder[li];
        if (li == 0) {
            /* ---- colourFirstLineage (:216-290) */
            double time = height[leaf];
            int type = nodeType[leaf];
            int next = parent[leaf];
            int thisNode = leaf;
            push(&L, UK_SAMPLE, type, 0, height[leaf]);
            double thisLogP = 0.0;
            for (;;) {
                double aTot = 0;
                for (int c = 0; c < d; c++)
                    if (c != type)
                        aTot += um->Q[type * d + c];              /* getBackwardRate(type, c) */
                if (pos >= nU) {
                    rc = ORC_ST_NEED_MORE_UNIFORMS;
                    break;
                }
                double newTime = time + (-1.0 * log(1 - u[pos++]) / aTot);   /* nextExponential */
                while (next >= 0 && height[next] < newTime) {
                    nodeType[next] = type;
                    push(&L, UK_COALESCE, type, 0, height[next]);
                    thisLogP += -aTot * (height[next] - time);
                    time = height[next];
                    thisNode = next;
                    next = parent[next];
                }
                if (next < 0)
                    break;
                thisLogP += -aTot * (newTime - time);
                time = newTime;
                if (pos >= nU) {
                    rc = ORC_ST_NEED_MORE_UNIFORMS;
                    break;
                }
                double uu = u[pos++] * aTot;
                int newType;
                for (newType = 0; newType < d; newType++) {
                    if (newType == type)
                        continue;
                    uu -= um->Q[type * d + newType];
                    if (uu < 0.0)
                        break;
                }
                if (newType >= d)       /* rounding fell through: Java would throw; take the last */
                    newType = (type == d - 1) ? d - 2 : d - 1;
                push(&L, UK_MIGRATE, type, newType, time);
                if (chgNodeOut) {                                 /* thisNode.addChange */
                    if (nAdded < chgCap) {
                        chgNodeOut[nAdded] = thisNode;
                        chgTypeOut[nAdded] = newType;
                        chgTimeOut[nAdded] = time;
                    }
                    nAdded++;
                }
                thisLogP += log(um->Q[type * d + newType]);
                type = newType;
            }
            w -= thisLogP;                                        /* :134 */
        } else {
            /* ---- colourLineage (:292-316) + addTypeChanges (:386-512) */
            push(&L, UK_SAMPLE, nodeType[leaf], 0, height[leaf]);
            int anc = parent[leaf];
            while (nodeType[anc] < 0)
                anc = parent[anc];
            const double t0 = height[leaf], t1 = height[anc];
            const int a = nodeType[leaf], b = nodeType[anc];
            int nChg = 0, nVirt = -1, forced = -1;
            double lp = 0.0;
            int64_t used = 0;
            int st = orc_retype_branch(um, t0, t1, a, b, u + pos, nU - pos, -1, cap, &nChg, ct, cty,
                                       &lp, &used, &nVirt);
            if (st == ORC_ST_NEEDS_POISSON_TAIL) {
                pos += used;
                const double muL = um->mu * (t1 - t0);
                const int thresh = nVirt;
                do {
                    st = orc_poisson_window(muL, u, nU, &pos, &forced);
                } while (st == ORC_ST_OK && forced < thresh);
                if (st == ORC_ST_OK)
                    st = orc_retype_branch(um, t0, t1, a, b, u + pos, nU - pos, forced, cap, &nChg,
                                           ct, cty, &lp, &used, &nVirt);
            }
            if (st == ORC_ST_NEED_MORE_UNIFORMS) {
                rc = st;
                break;
            }
            pos += used;
            if (st != ORC_ST_OK && chgNodeOut) {
                rc = st;
                break;
            }
            if (st != ORC_ST_OK) {
                /* NoValidPath / rubbish / FB failure: colourLineage or addTypeChanges returns
                 * -Infinity (:311-313, :399-403, :438-441) and the weight takes "-= -inf" */
                w -= -INFINITY;
                continue;
            }
            int prevType = a, prevNode = leaf, below = leaf;
            for (int i = 0; i < nChg && i < cap; i++) {           /* :452-486 */
                push(&L, UK_MIGRATE, prevType, cty[i], ct[i]);
                while (height[prevNode] < ct[i]) {
                    if (prevNode >= nLeaves) {
                        nodeType[prevNode] = prevType;
                        push(&L, UK_COALESCE, prevType, 0, height[prevNode]);
                    }
                    below = prevNode;
                    prevNode = parent[prevNode];
                }
                if (chgNodeOut) {                                 /* prevNode.addChange */
                    if (nAdded < chgCap) {
                        chgNodeOut[nAdded] = below;
                        chgTypeOut[nAdded] = cty[i];
                        chgTimeOut[nAdded] = ct[i];
                    }
                    nAdded++;
                }
                prevType = cty[i];
            }
            while (height[prevNode] < t1) {                       /* :490-502 */
                if (prevNode >= nLeaves) {
                    nodeType[prevNode] = prevType;
                    push(&L, UK_COALESCE, prevType, 0, height[prevNode]);
                }
                prevNode = parent[prevNode];
            }
            w -= lp;                                              /* :136 */
        }
    }

    if (rc == ORC_ST_OK) {
        /* ---- structured-coalescent density of the typed tree (:143-186) */
        qsort(L.ev, (size_t)L.n, sizeof(u_event), cmp_event);
        int cnt[ORC_MAX_D];
        for (int c = 0; c < d; c++)
            cnt[c] = (L.ev[0].type == c) ? 1 : 0;                 /* :145-146 */
        for (int i = 1; i < L.n; i++) {
            const u_event *e = &L.ev[i];
            const double dt = e->time - L.ev[i - 1].time;
            if (dt > 0) {                                         /* :154-169 */
                double lambda = 0.0;
                for (int c = 0; c < d; c++) {
                    const int k = cnt[c];
                    lambda += k * (k - 1) / (2.0 * N[c]);
                    for (int cp = 0; cp < d; cp++)
                        if (cp != c)
                            lambda += k * um->Q[c * d + cp];
                }
                w += -dt * lambda;
            }
            switch (e->kind) {                                    /* :172-190 */
            case UK_COALESCE:
                w += log(1.0 / N[e->type]);
                cnt[e->type] -= 1;
                break;
            case UK_MIGRATE:
                w += log(um->Q[e->type * d + e->dest]);
                cnt[e->type] -= 1;
                cnt[e->dest] += 1;
                break;
            default:
                cnt[e->type] += 1;
                break;
            }
        }
    }
    *logW = w;
    *uUsed = pos;
    if (nodeTypeOut)
        for (int i = 0; i < nNodes; i++)
            nodeTypeOut[i] = nodeType[i];
    if (nEventsOut)
        *nEventsOut = L.n;
    if (nChgOut)
        *nChgOut = nAdded;
    free(nodeType); free(L.ev); free(ct); free(cty);
    return rc;
}

int orc_untyped_particle(const orc_unif_model *um, const double *N, int nNodes,
                         const int32_t *parent, const double *height, const uint8_t *leafType,
                         const int32_t *leafOrder, const double *u, int64_t nU, double *logW,
                         int64_t *uUsed, int32_t *nodeTypeOut, int32_t *nEventsOut)
{
    return particle_impl(um, N, nNodes, parent, height, leafType, leafOrder, u, nU, logW, uUsed,
                         nodeTypeOut, nEventsOut, 0, NULL, NULL, NULL, NULL);
}

int orc_untyped_particle_tree(const orc_unif_model *um, const double *N, int nNodes,
                              const int32_t *parent, const double *height,
                              const uint8_t *leafType, const int32_t *leafOrder, const double *u,
                              int64_t nU, double *logW, int64_t *uUsed, int32_t *nodeTypeOut,
                              int chgCap, int32_t *nChgOut, int32_t *chgNodeOut,
                              int32_t *chgTypeOut, double *chgTimeOut)
{
    return particle_impl(um, N, nNodes, parent, height, leafType, leafOrder, u, nU, logW, uUsed,
                         nodeTypeOut, NULL, chgCap, nChgOut, chgNodeOut, chgTypeOut, chgTimeOut);
}

/* logP = log(sum_p exp(w_p - max) / P) + max   (:193-201) */
double orc_untyped_combine(const double *logW, int nParticles)
{
    double maxLogWeight = -INFINITY;
    for (int p = 0; p < nParticles; p++) {                /* Math.max: NaN wins */
        if (isnan(logW[p]) || isnan(maxLogWeight))
            maxLogWeight = NAN;
        else if (logW[p] > maxLogWeight)
            maxLogWeight = logW[p];
    }
    double sumScaledWeights = 0;
    for (int p = 0; p < nParticles; p++)
        sumScaledWeights += exp(logW[p] - maxLogWeight);
    return log(sumScaledWeights / nParticles) + maxLogWeight;
}
