/*
 * mtt_oracle.h -- CPU restatement of the MultiTypeTree likelihood hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: it may be
 * imported / linked / executed only by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py, and there only as the checker or
 * the timed CPU baseline -- never as a fallback for the CUDA path.
 *
 * Every function cites the reference file:line (relative to the upstream
 * tgvaughan/MultiTypeTree v8.1.0 tree) whose algorithm it restates.  The reference is
 * Java and cannot be compiled or run in the build container (no JVM, no BEAST 2 jar, no
 * jblas jar), so this restatement is pinned by:
 *   piece 1: the reference's own known-answer test
 *            test/multitypetree/coalescent/SCLikelihoodTest.java:43-73 (-16.52831, 1e-5)
 *            plus an independent sufficient-statistics identity and mpmath re-evaluation;
 *   piece 2 (BEAST 2 core TreeLikelihood, source NOT under the reference tree) and
 *   piece 3 (jblas expm / mmul, BEAST Randomizer; sources NOT under the reference tree):
 *            PARITY UNPINNED by any reference test -- checked against brute-force state
 *            enumeration, scipy.linalg.expm and numpy.linalg.matrix_power instead.
 *
 * Tree layout ("flattened MultiTypeTree", reference MultiTypeNode.java:33-36 fields as SoA):
 *   nNodes = 2*nLeaves-1; nodes [0,nLeaves) are leaves (BEAST numbering,
 *   NodeRetype.java:36-37 relies on it); height[nNodes]; parent[nNodes] (-1 = root);
 *   nodeType[nNodes]; chgOff[nNodes+1] CSR into chgTime[]/chgType[] holding the type changes
 *   on the branch ABOVE each node, ordered youngest -> oldest (MultiTypeTree.java:590-592).
 */
#ifndef MTT_ORACLE_H
#define MTT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- migration model */

/* Raw Inputs of SCMigrationModel (SCMigrationModel.java:38-66), resolved per call exactly
 * as the reference does (getBackwardRate :252-268, getPopSize :373-375). */
typedef struct {
    int nTypes;
    const double *rateMatrix;          /* rateDim values */
    int rateDim;                       /* d*d | d*(d-1) | d*(d-1)/2  (:119-133) */
    const double *popSizes;            /* nTypes values */
    double rateScale;                  /* rateMatrixScaleFactor (1 if absent) */
    double popScale;                   /* popSizesScaleFactor (1 if absent) */
    const unsigned char *rateFlags;    /* rateMatrixFlags, NULL = all on */
    int useForward;                    /* useForwardMigrationRateMatrix */
} orc_model;

int    orc_array_offset(const orc_model *m, int i, int j);   /* :324-345 */
double orc_backward_rate(const orc_model *m, int i, int j);  /* :252-268 */
double orc_forward_rate(const orc_model *m, int i, int j);   /* :285-303 */
double orc_pop_size(const orc_model *m, int i);              /* :373-375 */
/* dense m[d*d] (row-major, diagonal 0) and N[d]: what the C-ABI receives */
void   orc_resolve_dense(const orc_model *m, double *mDense, double *N);

/* ---------------------------------------------------------------- piece 1: density */

typedef struct {
    int nNodes;
    const double  *height;
    const int32_t *parent;
    const int32_t *nodeType;
    const int32_t *chgOff;   /* nNodes+1, absolute offsets into chgTime/chgType */
    const double  *chgTime;
    const int32_t *chgType;
} orc_tree;

/* MultiTypeTree.isValid (MultiTypeTree.java:396-429) */
int orc_is_valid(const orc_tree *t);

/* Faithful restatement: updateEventSequence (StructuredCoalescentTreeDensity.java:153-248,
 * O(E x active lineages) selection loop) + calculateLogP (:86-147) with per-call rate
 * lookups.  This is the "reference CPU path" that bench.py times. */
double orc_sc_logp_faithful(const orc_tree *t, const orc_model *m, int checkValidity);

/* Same value for valid trees via flat event extraction + sort. */
double orc_sc_logp_sorted(const orc_tree *t, int d, const double *mDense, const double *N,
                          int checkValidity);

/* Independent cross-check: sufficient-statistics identity (the
 * statistics are those of util/TypeLengths.java:69-92, util/TypeChangeCounts.java:76-104,
 * util/NodeTypeCounts.java:83-103). */
double orc_sc_logp_suffstat(const orc_tree *t, int d, const double *mDense, const double *N);

/* Batch driver over trees in the C-ABI's compact layout (include/mtt_b200.h):
 * nodeOff[B+1]; per node height f64 / parent i32 / nodeType u8 / chgCount u16;
 * chgBase[B+1] into chgTime f64 / chgType u8.  variant 0 = faithful, 1 = sorted,
 * 2 = sufficient statistics.  nThreads host threads over independent trees.
 * nModels is 1 (shared) or B (dense m/N per tree). */
void orc_sc_logp_batch(int B, int d, int nModels, const double *mDense, const double *N,
                       const int64_t *nodeOff, const double *height, const int32_t *parent,
                       const uint8_t *nodeType, const uint16_t *chgCount,
                       const int64_t *chgBase, const double *chgTime, const uint8_t *chgType,
                       int checkValidity, int variant, int nThreads, double *logP);

/* ---------------------------------------------------------------- piece 3: uniformization */

#define ORC_MAX_D 32

/* SCMigrationModel.updateMatrices (:150-201): Q, Qsym, mu, muSym, R, Rsym (row-major d*d). */
void orc_update_matrices(int d, const double *mDense, double *Q, double *Qsym, double *mu,
                         double *muSym, double *R, double *Rsym);

/* Lazy R^k cache with steady-state detection (SCMigrationModel.getRpowN :404-452,
 * RpowSteadyN :460-473). */
typedef struct orc_rpow orc_rpow;
orc_rpow     *orc_rpow_new(int d, const double *R);
void          orc_rpow_free(orc_rpow *c);
const double *orc_rpow_get(orc_rpow *c, int n);      /* row-major d*d */
int           orc_rpow_steady_n(const orc_rpow *c);  /* -1 if not (yet known) steady */
int           orc_rpow_size(const orc_rpow *c);

/* jblas 1.2.5 MatrixFunctions.expm restated (Pade-13 scaling and squaring + LU solve with
 * partial pivoting).  [EXT -- source not under the reference tree; call sites
 * UniformizationRetypeOperator.java:132-134, 263-264.]  A, out row-major d*d. */
void orc_expm(int d, const double *A, double *out);

/* status codes shared with the C-ABI (include/mtt_b200.h MTT_ST_*) */
enum {
    ORC_ST_OK = 0,
    ORC_ST_NO_VALID_PATH = 1,      /* Pba == 0 (NoValidPathException, :137-138) */
    ORC_ST_PBA_RUBBISH = 2,        /* Pba > 1 || Pba < 0 (:141-144, :267-270) */
    ORC_ST_FB_FELL_THROUGH = 3,    /* :179-190 */
    ORC_ST_NEEDS_POISSON_TAIL = 4, /* :96-102 reached; host Randomizer must finish */
    ORC_ST_NEED_MORE_UNIFORMS = 5,
    ORC_ST_NEED_MORE_POWERS = 6,   /* GPU only */
    ORC_ST_SKIPPED = 7             /* earlier branch of the same job aborted */
};

typedef struct {
    int d;
    const double *Q;   /* getQ(sym) */
    const double *R;   /* getR(sym) */
    double mu;         /* getMu(sym) */
    orc_rpow *rpow;
} orc_unif_model;

/* UniformizationRetypeOperator.retypeBranch (:116-225) with drawEventCount (:74-105).
 * Uniforms are consumed from u[0..nU) in the reference's order.
 * nVirtForced >= 0 skips the draw of the event count (used after a host-side Poisson tail).
 * Outputs: nChg, chgTime/chgType (capacity cap), logProb, uUsed, nVirt.  Returns status. */
int orc_retype_branch(const orc_unif_model *um, double tBottom, double tTop, int typeBottom,
                      int typeTop, const double *u, int64_t nU, int nVirtForced, int cap,
                      int *nChg, double *chgTime, int32_t *chgType, double *logProb,
                      int64_t *uUsed, int *nVirtOut);

/* UniformizationRetypeOperator.getBranchTypeProb (:233-275). Returns status. */
int orc_branch_type_prob(const orc_unif_model *um, double tBottom, double tTop, int typeBottom,
                         int nChg, const double *chgTime, const int32_t *chgType, int typeTop,
                         double *logProb);

/* ---------------------------------------------------------------- next row: untyped density
 * distributions/StructuredCoalescentUntypedTreeDensity.java (orc_untyped.c). */
int orc_poisson_window(double lam, const double *u, int64_t nU, int64_t *pos, int *out);
int orc_untyped_particle(const orc_unif_model *um, const double *N, int nNodes,
                         const int32_t *parent, const double *height, const uint8_t *leafType,
                         const int32_t *leafOrder, const double *u, int64_t nU, double *logW,
                         int64_t *uUsed, int32_t *nodeTypeOut, int32_t *nEventsOut);
/* the same particle returned as a typed tree (evolution/tree/MultiTypeTreeFromUntypedNewick.java
 * :104-134, :150-357): node types and the changes in the order they are added */
int orc_untyped_particle_tree(const orc_unif_model *um, const double *N, int nNodes,
                              const int32_t *parent, const double *height,
                              const uint8_t *leafType, const int32_t *leafOrder, const double *u,
                              int64_t nU, double *logW, int64_t *uUsed, int32_t *nodeTypeOut,
                              int chgCap, int32_t *nChgOut, int32_t *chgNodeOut,
                              int32_t *chgTypeOut, double *chgTimeOut);
double orc_untyped_combine(const double *logW, int nParticles);

/* ---------------------------------------------------------------- piece 2: peeling */

/* BEAST 2 HKY.getTransitionProbabilities restated [EXT].
 * P row-major 4x4, row = parent state; distance = (tParent - tNode) * rate. */
void orc_hky_matrix(double kappa, const double *freqs, double distance, double *P);

/* BEAST 2 TreeLikelihood.calculateLogP with BeerLikelihoodCore semantics [EXT]
 * (SURVEY Appendix C), 4 states.  tipStates[nLeaves*P] (state >= 4 = gap/ambiguous),
 * weights[P], C categories (rates, proportions already adjusted for pInv by the SiteModel),
 * branchRate = strict clock rate, pInv = proportion invariant (0 = none).
 * useScaling: 0 never, 1 always (threshold 1e-100), 2 BEAST behaviour (retry on -inf).
 * patternLogL (may be NULL) receives per-pattern log-likelihoods.  Returns logP. */
double orc_tree_likelihood(int nLeaves, int P, int C, const int32_t *tipStates,
                           const int32_t *weights, const int32_t *parent, const double *height,
                           const double *catRates, const double *catProps, const double *freqs,
                           double kappa, double branchRate, double pInv, int useScaling,
                           double *patternLogL);

/* same with explicit per-node transition matrices mats[node][C][16] (root entry unused) */
double orc_tree_likelihood_mats(int nLeaves, int P, int C, const int32_t *tipStates,
                                const int32_t *weights, const int32_t *parent,
                                const double *mats, const double *catProps, const double *freqs,
                                double pInv, int useScaling, double *patternLogL);

void orc_tree_likelihood_batch(int B, int nLeaves, int P, int C, const int32_t *tipStates,
                               const int32_t *weights, const int32_t *parent,
                               const double *height, const double *catRates,
                               const double *catProps, const double *freqs, double kappa,
                               double branchRate, double pInv, int useScaling, int nThreads,
                               double *logP);

#ifdef __cplusplus
}
#endif
#endif
