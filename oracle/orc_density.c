/*
 * orc_density.c -- oracle for piece 1 (structured-coalescent tree density).
 * TEST INFRASTRUCTURE ONLY (see mtt_oracle.h).  Restates, from the Java sources:
 *   distributions/StructuredCoalescentTreeDensity.java:86-248
 *   evolution/tree/SCMigrationModel.java:119-133, 252-303, 324-345, 373-375
 *   evolution/tree/MultiTypeTree.java:396-429
 */
#include "mtt_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ migration model */

/* SCMigrationModel.getArrayOffset (:324-345).  Encoding detection (:119-133):
 * dim == d*d square; dim == d*(d-1) row-major without diagonal; dim == d*(d-1)/2 symmetric. */
int orc_array_offset(const orc_model *m, int i, int j)
{
    int d = m->nTypes;
    if (m->rateDim == d * d)
        return i * d + j;
    if (m->rateDim != d * (d - 1)) {          /* symmetric lower-triangular */
        if (j < i)
            return i * (i - 1) / 2 + j;
        return j * (j - 1) / 2 + i;
    }
    if (j > i)
        j -= 1;
    return i * (d - 1) + j;
}

double orc_pop_size(const orc_model *m, int i)
{
    return m->popScale * m->popSizes[i];      /* :373-375 */
}

double orc_forward_rate(const orc_model *m, int i, int j);

/* SCMigrationModel.getBackwardRate (:252-268) */
double orc_backward_rate(const orc_model *m, int i, int j)
{
    if (i == j)
        return 0;
    if (m->useForward)
        return orc_forward_rate(m, j, i) * orc_pop_size(m, j) / orc_pop_size(m, i);
    int off = orc_array_offset(m, i, j);
    if (m->rateFlags && !m->rateFlags[off])
        return 0.0;
    return m->rateScale * m->rateMatrix[off];
}

/* SCMigrationModel.getForwardRate (:285-303) */
double orc_forward_rate(const orc_model *m, int i, int j)
{
    if (i == j)
        return 0.0;
    if (m->useForward) {
        int off = orc_array_offset(m, i, j);
        if (m->rateFlags && !m->rateFlags[off])
            return 0.0;
        return m->rateScale * m->rateMatrix[off];
    }
    return orc_backward_rate(m, j, i) * orc_pop_size(m, j) / orc_pop_size(m, i);
}

void orc_resolve_dense(const orc_model *m, double *mDense, double *N)
{
    int d = m->nTypes;
    for (int i = 0; i < d; i++) {
        N[i] = orc_pop_size(m, i);
        for (int j = 0; j < d; j++)
            mDense[i * d + j] = orc_backward_rate(m, i, j);
    }
}

/* ------------------------------------------------------------------ tree helpers */

/* children derived from parent[]: left = lower node index */
static void build_children(const orc_tree *t, int32_t *left, int32_t *right, int *root)
{
    for (int i = 0; i < t->nNodes; i++)
        left[i] = right[i] = -1;
    *root = -1;
    for (int i = 0; i < t->nNodes; i++) {
        int p = t->parent[i];
        if (p < 0) {
            *root = i;
            continue;
        }
        if (left[p] < 0)
            left[p] = i;
        else
            right[p] = i;
    }
}

static int change_count(const orc_tree *t, int node)
{
    return t->chgOff[node + 1] - t->chgOff[node];
}

/* MultiTypeNode.getFinalType (MultiTypeNode.java:73-78) */
static int final_type(const orc_tree *t, int node)
{
    int n = change_count(t, node);
    if (n > 0)
        return t->chgType[t->chgOff[node] + n - 1];
    return t->nodeType[node];
}

/* MultiTypeTree.isValid = timesAreValid && typesAreValid (MultiTypeTree.java:396-429).
 * The recursion visits every (node, child) pair; a flat loop over non-root nodes is the
 * same set of checks. */
int orc_is_valid(const orc_tree *t)
{
    for (int c = 0; c < t->nNodes; c++) {
        int p = t->parent[c];
        if (p < 0)
            continue;
        double lastHeight = t->height[p];                      /* :401 */
        for (int idx = change_count(t, c) - 1; idx >= 0; idx--) {
            double thisHeight = t->chgTime[t->chgOff[c] + idx];
            if (thisHeight > lastHeight)                       /* :404-405 */
                return 0;
            lastHeight = thisHeight;
        }
        if (t->height[c] > lastHeight)                         /* :408-409 */
            return 0;
    }
    for (int c = 0; c < t->nNodes; c++) {
        int p = t->parent[c];
        if (p < 0)
            continue;
        if (t->nodeType[p] != final_type(t, c))                /* :421-422 */
            return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------ faithful path */

enum { EV_COALESCE = 0, EV_MIGRATE = 1, EV_SAMPLE = 2 };

typedef struct {
    double time;
    int type, destType, kind, node;
} sc_event;

double orc_sc_logp_faithful(const orc_tree *t, const orc_model *m, int checkValidity)
{
    if (checkValidity && !orc_is_valid(t))                     /* :90-91 */
        return -INFINITY;

    int d = m->nTypes;
    int nNodes = t->nNodes;
    int M = t->chgOff[nNodes] - t->chgOff[0];
    int maxEv = nNodes + M;

    int32_t *left = malloc(sizeof(int32_t) * nNodes);
    int32_t *right = malloc(sizeof(int32_t) * nNodes);
    int root;
    build_children(t, left, right, &root);

    /* updateEventSequence (:153-248).  The Java HashMap<Node,Integer> "changeIdx" is
     * restated as an unordered active list; iteration order is arbitrary in Java too. */
    int *actNode = malloc(sizeof(int) * nNodes);
    int *actIdx = malloc(sizeof(int) * nNodes);
    int nAct = 0;
    actNode[nAct] = root;
    actIdx[nAct++] = -1;                                       /* :161-162 */

    sc_event *events = malloc(sizeof(sc_event) * (maxEv + 1));
    int *counts = malloc(sizeof(int) * (size_t)(maxEv + 1) * d);
    int *lineageCount = malloc(sizeof(int) * d);
    for (int c = 0; c < d; c++)                                /* :165-170 */
        lineageCount[c] = (c == t->nodeType[root]) ? 1 : 0;

    int nEv = 0;
    while (nAct > 0) {                                         /* :173 */
        sc_event next;
        next.time = -INFINITY;
        next.node = root;
        next.kind = -1;
        next.type = next.destType = 0;
        int slot = 0;
        for (int a = 0; a < nAct; a++) {                       /* :180-212 */
            int node = actNode[a];
            if (actIdx[a] < 0) {
                if (t->height[node] > next.time) {
                    next.time = t->height[node];
                    next.kind = (left[node] < 0) ? EV_SAMPLE : EV_COALESCE;
                    next.type = t->nodeType[node];
                    next.node = node;
                    slot = a;
                }
            } else {
                int ci = actIdx[a];
                double thisChangeTime = t->chgTime[t->chgOff[node] + ci];
                if (thisChangeTime > next.time) {
                    next.time = thisChangeTime;
                    next.kind = EV_MIGRATE;
                    next.destType = t->chgType[t->chgOff[node] + ci];
                    if (ci > 0)
                        next.type = t->chgType[t->chgOff[node] + ci - 1];
                    else
                        next.type = t->nodeType[node];
                    next.node = node;
                    slot = a;
                }
            }
        }
        if (next.kind < 0)      /* NaN heights: Java would NPE on the switch; bail out */
            break;

        switch (next.kind) {                                   /* :215-237 */
        case EV_COALESCE: {
            int l = left[next.node], r = right[next.node];
            actNode[slot] = l;
            actIdx[slot] = change_count(t, l) - 1;
            actNode[nAct] = r;
            actIdx[nAct++] = change_count(t, r) - 1;
            lineageCount[next.type]++;
            break;
        }
        case EV_SAMPLE:
            actNode[slot] = actNode[nAct - 1];
            actIdx[slot] = actIdx[nAct - 1];
            nAct--;
            lineageCount[next.type]--;
            break;
        case EV_MIGRATE:
            lineageCount[next.destType]--;
            lineageCount[next.type]++;
            actIdx[slot] -= 1;
            break;
        }
        events[nEv] = next;                                    /* :240-241 */
        memcpy(counts + (size_t)nEv * d, lineageCount, sizeof(int) * d);
        nEv++;
        if (nEv > maxEv)
            break;
    }

    /* calculateLogP (:97-146) on the reversed lists (:245-246): reversed index e maps to
     * stored index nEv-1-e. */
    double logP = 0;
    for (int e = 1; e < nEv; e++) {
        const sc_event *ev = &events[nEv - 1 - e];
        const sc_event *prev = &events[nEv - e];
        const int *lc = counts + (size_t)(nEv - 1 - e) * d;
        double delta_t = ev->time - prev->time;

        if (delta_t > 0) {                                     /* :108-124 */
            double lambda = 0.0;
            for (int c = 0; c < d; c++) {
                int k = lc[c];
                double Nc = orc_pop_size(m, c);
                lambda += k * (k - 1) / (2.0 * Nc);
                for (int cp = 0; cp < d; cp++) {
                    if (cp == c)
                        continue;
                    double mr = orc_backward_rate(m, c, cp);
                    lambda += k * mr;
                }
            }
            logP += -delta_t * lambda;
        }

        switch (ev->kind) {                                    /* :127-143 */
        case EV_COALESCE: {
            double N = orc_pop_size(m, ev->type);
            logP += log(1.0 / N);
            break;
        }
        case EV_MIGRATE: {
            double mr = orc_backward_rate(m, ev->type, ev->destType);
            logP += log(mr);
            break;
        }
        default:
            break;
        }
    }

    free(left); free(right); free(actNode); free(actIdx);
    free(events); free(counts); free(lineageCount);
    return logP;
}

/* ------------------------------------------------------------------ sorted path */

typedef struct {
    double time;
    int kind, type, destType, seq;
} flat_event;

static int cmp_flat(const void *a, const void *b)
{
    const flat_event *x = a, *y = b;
    if (x->time < y->time) return -1;
    if (x->time > y->time) return 1;
    return (x->seq > y->seq) - (x->seq < y->seq);
}

/* Flat extraction + ascending sort: for a valid tree the event
 * multiset {(height, SAMPLE|COALESCE, nodeType)} U {(changeTime, MIGRATE, from, to)} sorted
 * by time reproduces the reversed list of updateEventSequence; going tip->root the count
 * updates are SAMPLE k[type]++, COALESCE k[type]--, MIGRATE k[type]--, k[destType]++. */
double orc_sc_logp_sorted(const orc_tree *t, int d, const double *mDense, const double *N,
                          int checkValidity)
{
    if (checkValidity && !orc_is_valid(t))
        return -INFINITY;

    int nNodes = t->nNodes;
    int nEv = 0;
    int M = t->chgOff[nNodes] - t->chgOff[0];
    flat_event *ev = malloc(sizeof(flat_event) * (size_t)(nNodes + M));
    char *hasChild = calloc(nNodes, 1);
    for (int i = 0; i < nNodes; i++)
        if (t->parent[i] >= 0)
            hasChild[t->parent[i]] = 1;

    for (int i = 0; i < nNodes; i++) {
        ev[nEv].time = t->height[i];
        ev[nEv].kind = hasChild[i] ? EV_COALESCE : EV_SAMPLE;
        ev[nEv].type = t->nodeType[i];
        ev[nEv].destType = 0;
        ev[nEv].seq = nEv;
        nEv++;
        if (t->parent[i] < 0)
            continue;                          /* root-branch changes ignored (:161-162) */
        int prevType = t->nodeType[i];
        for (int k = t->chgOff[i]; k < t->chgOff[i + 1]; k++) {
            ev[nEv].time = t->chgTime[k];
            ev[nEv].kind = EV_MIGRATE;
            ev[nEv].type = prevType;
            ev[nEv].destType = t->chgType[k];
            ev[nEv].seq = nEv;
            prevType = t->chgType[k];
            nEv++;
        }
    }
    free(hasChild);
    qsort(ev, nEv, sizeof(flat_event), cmp_flat);

    int *k = calloc(d, sizeof(int));
    double logP = 0;
    for (int e = 0; e < nEv; e++) {
        if (e > 0) {
            double delta_t = ev[e].time - ev[e - 1].time;
            if (delta_t > 0) {
                double lambda = 0.0;
                for (int c = 0; c < d; c++) {
                    lambda += k[c] * (k[c] - 1) / (2.0 * N[c]);
                    for (int cp = 0; cp < d; cp++)
                        if (cp != c)
                            lambda += k[c] * mDense[c * d + cp];
                }
                logP += -delta_t * lambda;
            }
            if (ev[e].kind == EV_COALESCE)
                logP += log(1.0 / N[ev[e].type]);
            else if (ev[e].kind == EV_MIGRATE)
                logP += log(mDense[ev[e].type * d + ev[e].destType]);
        }
        if (ev[e].kind == EV_SAMPLE)
            k[ev[e].type]++;
        else if (ev[e].kind == EV_COALESCE)
            k[ev[e].type]--;
        else {
            k[ev[e].type]--;
            k[ev[e].destType]++;
        }
    }
    free(k);
    free(ev);
    return logP;
}

/* ------------------------------------------------------------------ sufficient statistics */

typedef struct { double time; int type; int delta; } cnt_event;

static int cmp_cnt(const void *a, const void *b)
{
    const cnt_event *x = a, *y = b;
    return (x->time > y->time) - (x->time < y->time);
}

/* logP = -sum_c T_c mu_c - sum_intervals dt sum_c k_c(k_c-1)/(2 N_c)
 *        + sum_{c!=c'} n_cc' log m(c,c') - sum_c C_c log N_c          
 * T_c, n_cc', C_c are accumulated per branch (no global event order); only the coalescent
 * integral uses a sort, and of a different object (per-type +1/-1 count steps). */
double orc_sc_logp_suffstat(const orc_tree *t, int d, const double *mDense, const double *N)
{
    int nNodes = t->nNodes;
    int M = t->chgOff[nNodes] - t->chgOff[0];
    double *T = calloc(d, sizeof(double));
    double *nChg = calloc((size_t)d * d, sizeof(double));
    double *C = calloc(d, sizeof(double));
    char *hasChild = calloc(nNodes, 1);
    cnt_event *ce = malloc(sizeof(cnt_event) * (size_t)(2 * (nNodes + M) + 2));
    int nce = 0;

    for (int i = 0; i < nNodes; i++)
        if (t->parent[i] >= 0)
            hasChild[t->parent[i]] = 1;
    for (int i = 0; i < nNodes; i++) {
        if (hasChild[i])
            C[t->nodeType[i]] += 1.0;
        int p = t->parent[i];
        if (p < 0)
            continue;
        double segStart = t->height[i];
        int type = t->nodeType[i];
        for (int k = t->chgOff[i]; k <= t->chgOff[i + 1]; k++) {
            double segEnd = (k < t->chgOff[i + 1]) ? t->chgTime[k] : t->height[p];
            T[type] += segEnd - segStart;
            ce[nce].time = segStart; ce[nce].type = type; ce[nce].delta = +1; nce++;
            ce[nce].time = segEnd;   ce[nce].type = type; ce[nce].delta = -1; nce++;
            if (k < t->chgOff[i + 1]) {
                nChg[type * d + t->chgType[k]] += 1.0;
                type = t->chgType[k];
                segStart = segEnd;
            }
        }
    }
    qsort(ce, nce, sizeof(cnt_event), cmp_cnt);

    double logP = 0;
    for (int c = 0; c < d; c++) {
        double muc = 0;
        for (int cp = 0; cp < d; cp++)
            if (cp != c)
                muc += mDense[c * d + cp];
        logP -= T[c] * muc;
        logP -= C[c] * log(N[c]);
        for (int cp = 0; cp < d; cp++)
            if (cp != c && nChg[c * d + cp] > 0)
                logP += nChg[c * d + cp] * log(mDense[c * d + cp]);
    }
    int *k = calloc(d, sizeof(int));
    double coal = 0;
    for (int e = 0; e < nce; e++) {
        if (e > 0) {
            double dt = ce[e].time - ce[e - 1].time;
            if (dt > 0)
                for (int c = 0; c < d; c++)
                    coal += dt * (k[c] * (k[c] - 1) / (2.0 * N[c]));
        }
        k[ce[e].type] += ce[e].delta;
    }
    logP -= coal;
    free(k); free(ce); free(hasChild); free(T); free(nChg); free(C);
    return logP;
}

/* ------------------------------------------------------------------ batch driver */

/* Batch layout = the C-ABI's compact device layout (include/mtt_b200.h, mtt_sc_logp_batch):
 * nodeOff[B+1] node offsets; height f64, parent i32, nodeType u8, chgCount u16 per node;
 * chgBase[B+1] offsets into chgTime f64 / chgType u8 (changes of a tree stored node by node). */
typedef struct {
    int B, d, nModels, checkValidity, variant, tid, nThreads;
    const double *mDense, *N;
    const int64_t *nodeOff, *chgBase;
    const double *height, *chgTime;
    const int32_t *parent;
    const uint8_t *nodeType, *chgType;
    const uint16_t *chgCount;
    double *logP;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = arg;
    int d = j->d;
    for (int b = j->tid; b < j->B; b += j->nThreads) {
        int64_t n0 = j->nodeOff[b];
        int nNodes = (int)(j->nodeOff[b + 1] - n0);
        int64_t c0 = j->chgBase[b];
        int M = (int)(j->chgBase[b + 1] - c0);
        int32_t *off = malloc(sizeof(int32_t) * (nNodes + 1));
        int32_t *ntype = malloc(sizeof(int32_t) * nNodes);
        int32_t *ctype = malloc(sizeof(int32_t) * (M > 0 ? M : 1));
        off[0] = 0;
        for (int i = 0; i < nNodes; i++) {
            off[i + 1] = off[i] + j->chgCount[n0 + i];
            ntype[i] = j->nodeType[n0 + i];
        }
        for (int k = 0; k < M; k++)
            ctype[k] = j->chgType[c0 + k];
        orc_tree t = { nNodes, j->height + n0, j->parent + n0, ntype, off,
                       j->chgTime + c0, ctype };
        const double *mD = j->mDense + (j->nModels > 1 ? (size_t)b * d * d : 0);
        const double *Nn = j->N + (j->nModels > 1 ? (size_t)b * d : 0);
        if (j->variant == 0) {
            /* the faithful path resolves rates per call from the raw Inputs; a dense
             * square rateMatrix (diagonal ignored, :119-121) is one of the 3 encodings */
            orc_model m = { d, mD, d * d, Nn, 1.0, 1.0, NULL, 0 };
            j->logP[b] = orc_sc_logp_faithful(&t, &m, j->checkValidity);
        } else if (j->variant == 1) {
            j->logP[b] = orc_sc_logp_sorted(&t, d, mD, Nn, j->checkValidity);
        } else {
            j->logP[b] = orc_sc_logp_suffstat(&t, d, mD, Nn);
        }
        free(off); free(ntype); free(ctype);
    }
    return NULL;
}

void orc_sc_logp_batch(int B, int d, int nModels, const double *mDense, const double *N,
                       const int64_t *nodeOff, const double *height, const int32_t *parent,
                       const uint8_t *nodeType, const uint16_t *chgCount,
                       const int64_t *chgBase, const double *chgTime, const uint8_t *chgType,
                       int checkValidity, int variant, int nThreads, double *logP)
{
    if (nThreads < 1)
        nThreads = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * nThreads);
    batch_job *jobs = malloc(sizeof(batch_job) * nThreads);
    for (int i = 0; i < nThreads; i++) {
        batch_job j = { B, d, nModels, checkValidity, variant, i, nThreads, mDense, N,
                        nodeOff, chgBase, height, chgTime, parent, nodeType, chgType,
                        chgCount, logP };
        jobs[i] = j;
        if (nThreads == 1)
            batch_worker(&jobs[i]);
        else
            pthread_create(&th[i], NULL, batch_worker, &jobs[i]);
    }
    if (nThreads > 1)
        for (int i = 0; i < nThreads; i++)
            pthread_join(th[i], NULL);
    free(th);
    free(jobs);
}
