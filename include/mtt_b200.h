/*
 * mtt_b200.h -- plain-C ABI of the B200-native MultiTypeTree likelihood hot path.
 *
 * This is the drop-in boundary: what a BEAST 2 plugin binds through Panama FFM / JNI (see
 * INTEGRATION.md and java/) in place of the reference's Java + jblas code.  Plain pointers
 * and sizes only; no torch / CUDA / JNI types.  Each entry point cites the reference
 * interface (file:line, relative to tgvaughan/MultiTypeTree v8.1.0 src/multitypetree/) it
 * replaces.
 *
 * Conventions
 *   - return 0 (MTT_OK) on success, < 0 on API / CUDA failure (text via mtt_last_error);
 *     never aborts, never throws, no callbacks.
 *   - numerical failure is a VALUE, as in the reference: logP may be -inf, and per-item
 *     status codes MTT_ST_* report what the Java code signals with -Infinity / exceptions.
 *   - caller owns every host buffer for the duration of the call only; the library owns all
 *     device memory, tied to the ctx.  One ctx per chain/thread (a ctx is not re-entrant);
 *     separate ctxs are independent.  Calls block until results are on the host unless the
 *     name ends in _dev (device pointers, asynchronous on the ctx stream).
 *   - there is NO CPU fallback: without a CUDA device mtt_ctx_create fails.
 *
 * Flattened MultiTypeTree ("tree SoA"; replaces the object graph of
 * evolution/tree/MultiTypeNode.java:33-36 + MultiTypeTree.java:256-264).  A batch of B trees
 * is stored back to back:
 *   nodeOff [B+1]  int64   node offsets; tree b owns nodes [nodeOff[b], nodeOff[b+1]),
 *                          nNodes = 2*nLeaves-1, nodes [0,nLeaves) are the leaves (BEAST Nr
 *                          numbering, relied on by operators/NodeRetype.java:36-37)
 *   height  [nodes] f64    Node.getHeight()
 *   parent  [nodes] int32  tree-local parent Nr, -1 for the root (may be NULL when neither
 *                          validity checking nor peeling is requested)
 *   nodeType[nodes] uint8  MultiTypeNode.getNodeType()           (MultiTypeNode.java:100-102)
 *   chgBase [B+1]  int64   offsets into chgTime/chgType/chgNode; a tree's changes are stored
 *                          node by node in Nr order, each branch youngest -> oldest
 *   chgTime [changes] f64  MultiTypeNode.getChangeTime(idx)      (:65-67)
 *   chgType [changes] uint8 MultiTypeNode.getChangeType(idx)     (:54-56)
 *   which branch a change sits on, in one of two encodings:
 *   chgNode [changes] uint16 tree-local Nr of the node below the change (OWNER layout,
 *                          mtt_sc_logp_owner_batch*: what the kernels read), or
 *   chgCount[nodes] uint16 MultiTypeNode.getChangeCount()        (:43-45) (COUNTS layout,
 *                          mtt_sc_logp_batch*: expanded to the owner layout on the device).
 *                          Root-branch changes must not be listed unless parent[] is given
 *                          (they are ignored, distributions/StructuredCoalescentTreeDensity
 *                          .java:161-162)
 * Compact integer widths are part of the contract: the density path is HBM/PCIe bound and
 * these arrays ARE its traffic (12 bytes per event algorithmic).
 */
#ifndef MTT_B200_H
#define MTT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTT_ABI_VERSION 1

#define MTT_OK                0
#define MTT_ERR_INVALID_ARG  -1
#define MTT_ERR_CUDA         -2
#define MTT_ERR_NO_DEVICE    -3
#define MTT_ERR_TOO_LARGE    -4   /* a tree / model exceeds a documented kernel limit */
#define MTT_ERR_STATE        -5   /* call order violated (e.g. logl before set_*) */

#define MTT_MAX_DEMES        16   /* register/smem-resident d x d kernels */
#define MTT_MAX_TREE_EVENTS  16384 /* nodes + changes per tree (density kernel, smem sort) */

/* per-item status codes (values, not errors) */
#define MTT_ST_OK                 0
#define MTT_ST_NO_VALID_PATH      1  /* Pba == 0: NoValidPathException,
                                        operators/UniformizationRetypeOperator.java:137-138 */
#define MTT_ST_PBA_RUBBISH        2  /* Pba > 1 || Pba < 0 (:141-144, :267-270) -> -inf */
#define MTT_ST_FB_FELL_THROUGH    3  /* forward-backward failure (:179-190) -> -inf */
#define MTT_ST_NEEDS_POISSON_TAIL 4  /* steady-state tail (:96-102): the host Randomizer must
                                        draw nVirt >= nVirt_out and resubmit with nVirtForced */
#define MTT_ST_NEED_MORE_UNIFORMS 5  /* uniform window exhausted: resubmit with a longer one */
#define MTT_ST_NEED_MORE_POWERS   6  /* internal (handled inside the library) */
#define MTT_ST_SKIPPED            7  /* an earlier branch of the same job aborted */

typedef struct mtt_ctx mtt_ctx;

/* ------------------------------------------------------------------ context */

/* device: CUDA ordinal, or -1 for the current device. */
int mtt_ctx_create(int device, mtt_ctx **out);
int mtt_ctx_destroy(mtt_ctx *ctx);
/* last error text of this ctx (or of the failed mtt_ctx_create when ctx == NULL) */
const char *mtt_last_error(const mtt_ctx *ctx);
int mtt_abi_version(void);
/* Run this ctx's work on an existing CUDA stream (a cudaStream_t passed as void*), e.g.
 * the caller's torch stream, instead of the ctx's own stream.  NULL restores the own stream. */
int mtt_ctx_set_stream(mtt_ctx *ctx, void *cuda_stream);
int mtt_ctx_synchronize(mtt_ctx *ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
int64_t mtt_ctx_launch_count(const mtt_ctx *ctx);
/* page-lock / unlock a caller buffer so that host<->device copies run at full PCIe rate */
int mtt_host_register(void *ptr, int64_t bytes);
int mtt_host_unregister(void *ptr);

/* ------------------------------------------------------------------ piece 1: density
 * Replaces distributions/StructuredCoalescentTreeDensity.java:86-147 (calculateLogP),
 * :153-248 (updateEventSequence) and evolution/tree/MultiTypeTree.java:396-429 (isValid),
 * batched over B trees.  The migration model arrives already resolved to dense form (what
 * evolution/tree/SCMigrationModel.java:252-268 getBackwardRate and :373-375 getPopSize
 * return): m[nModels][d*d] row-major backward rates (diagonal ignored), N[nModels][d];
 * nModels is 1 (shared by all trees) or B (one per tree).
 *   logP [B]  out: log density; -inf when checkValidity != 0 and the tree is invalid (:90-91)
 *   valid[B]  out (may be NULL): isValid() result, computed whenever parent != NULL
 * Event extraction, time ordering and the interval scan all happen on the device. */
int mtt_sc_logp_batch(mtt_ctx *ctx, int64_t B, int d, int64_t nModels, const double *m,
                      const double *N, const int64_t *nodeOff, const double *height,
                      const int32_t *parent, const uint8_t *nodeType, const uint16_t *chgCount,
                      const int64_t *chgBase, const double *chgTime, const uint8_t *chgType,
                      int checkValidity, double *logP, uint8_t *valid);

/* Same with every pointer a DEVICE pointer; asynchronous on the ctx stream.  Each array must
 * be readable 16 bytes past its logical end (bulk-copy granularity). */
int mtt_sc_logp_batch_dev(mtt_ctx *ctx, int64_t B, int d, int64_t nModels, const double *m,
                          const double *N, const int64_t *nodeOff, const double *height,
                          const int32_t *parent, const uint8_t *nodeType,
                          const uint16_t *chgCount, const int64_t *chgBase,
                          const double *chgTime, const uint8_t *chgType, int checkValidity,
                          double *logP, uint8_t *valid);
/* The same two calls with the change lists in the OWNER layout, which is what the kernels read
 * (the calls above expand their per-node counts into it on the device first):
 *   chgNode[changes]   tree-local number of the node each change sits above; the changes of one
 *                      node are adjacent and ordered child -> parent, nodes in increasing order.
 * 2 bytes per change instead of 2 bytes per node: about 12 % less to move for a 100-tip tree
 * with 40 changes, and no prefix sum on the device. */
int mtt_sc_logp_owner_batch(mtt_ctx *ctx, int64_t B, int d, int64_t nModels, const double *m,
                            const double *N, const int64_t *nodeOff, const double *height,
                            const int32_t *parent, const uint8_t *nodeType,
                            const uint16_t *chgNode, const int64_t *chgBase,
                            const double *chgTime, const uint8_t *chgType, int checkValidity,
                            double *logP, uint8_t *valid);
int mtt_sc_logp_owner_batch_dev(mtt_ctx *ctx, int64_t B, int d, int64_t nModels,
                                const double *m, const double *N, const int64_t *nodeOff,
                                const double *height, const int32_t *parent,
                                const uint8_t *nodeType, const uint16_t *chgNode,
                                const int64_t *chgBase, const double *chgTime,
                                const uint8_t *chgType, int checkValidity, double *logP,
                                uint8_t *valid);

/* Trees over ONE taxon set.  The tips are data -- tip dates and tip types are fixed for a whole
 * analysis -- so they are installed once, like the alignment of the peeling handle, and a call
 * moves only what a proposal can change: the heights and types of the nLeaves-1 internal nodes
 * (node nLeaves + i is entry i) and the change lists in the owner layout.  For a 100-tip tree
 * with 40 changes that is 1.35 KB instead of 2.26 KB over the host link, which is what bounds
 * the host-buffer entries.  Results are bit-identical to mtt_sc_logp_owner_batch on the same
 * trees.  leafHeight[nLeaves], leafType[nLeaves]; intHeight / intType [B][nLeaves-1];
 * parent [B][2*nLeaves-1] (NULL without validity checking). */
int mtt_sc_set_tips(mtt_ctx *ctx, int nLeaves, const double *leafHeight, const uint8_t *leafType);
int mtt_sc_logp_tips_batch(mtt_ctx *ctx, int64_t B, int d, int64_t nModels, const double *m,
                           const double *N, const double *intHeight, const int32_t *parent,
                           const uint8_t *intType, const uint16_t *chgNode,
                           const int64_t *chgBase, const double *chgTime,
                           const uint8_t *chgType, int checkValidity, double *logP,
                           uint8_t *valid);

/* Optional hint for the _dev entries: an upper bound on nodes + changes of any tree in later
 * batches (0 = unknown; it stays set on the ctx until changed).  With a bound <= 1022 the batch
 * is ONE kernel launch sized for that bound.  The bound is a promise: a tree that exceeds it is
 * not evaluated and its logP is NaN (valid = 0) -- an error the caller can see, never a
 * silently missing result.
 * Malformed input is reported the same way: a node or change type >= d, or a parent index
 * outside the tree, gives logP = NaN.  A change that names an owner outside the tree is
 * ignored (the tree is invalid: -infinity under checkValidity). */
int mtt_sc_set_max_tree_events(mtt_ctx *ctx, int64_t maxEvents);

/* ------------------------------------------------------------------ piece 3: uniformization
 * Replaces evolution/tree/SCMigrationModel.java:150-201 (updateMatrices), :404-473 (getRpowN /
 * RpowSteadyN: the jblas mmul chain) and operators/UniformizationRetypeOperator.java:74-275
 * (drawEventCount, retypeBranch, getBranchTypeProb: jblas expm + forward-backward sampling).
 *
 * A ctx holds nModels uniformization models (one per chain).  m is the dense backward-rate
 * matrix as above; sym selects the symmetrised variant (useSymmetrizedRates,
 * UniformizationRetypeOperator.java:41-43,118).  Setting a model drops its cached powers,
 * like SCMigrationModel.java:187-198. */
int mtt_unif_set_models(mtt_ctx *ctx, int nModels, int d, const double *m, int sym);

/* Q, R (row-major d*d) and mu of one model, as getQ/getR/getMu(sym) return them. */
int mtt_unif_get_matrices(mtt_ctx *ctx, int model, double *Q, double *R, double *mu);

/* R^0 .. R^kmax of one model (out: (kmax+1)*d*d doubles, row-major), computed on the device
 * with the reference's chain R^i = R^(i-1) * R and steady-state rule (every 10th power,
 * SCMigrationModel.java:435-447).  *steadyN receives RpowSteadyN() (-1 if not steady within
 * kmax); powers beyond the steady index repeat the last matrix (:422-426). */
int mtt_unif_rpow(mtt_ctx *ctx, int model, int kmax, double *out, int *steadyN);

/* expm(Q*L) for a batch of branch lengths (out: n*d*d doubles): the jblas
 * MatrixFunctions.expm call of UniformizationRetypeOperator.java:132-134. */
int mtt_unif_expm_batch(mtt_ctx *ctx, int64_t n, const int32_t *model, const double *L,
                        double *out);

/* getBranchTypeProb (:233-275) for nBranches branches.  Branch i runs from (tBottom,
 * typeBottom) up to (tTop, typeTop) with the existing changes
 * chgTime/chgType[chgOff[i] .. chgOff[i+1]).  status: MTT_ST_OK | MTT_ST_PBA_RUBBISH. */
int mtt_unif_branch_logp_batch(mtt_ctx *ctx, int64_t nBranches, const int32_t *model,
                               const double *tBottom, const double *tTop,
                               const uint8_t *typeBottom, const uint8_t *typeTop,
                               const int64_t *chgOff, const double *chgTime,
                               const uint8_t *chgType, double *logProb, int32_t *status);

/* retypeBranch (:116-225) for nJobs jobs.  A job is the sequence of branches
 * [brOff[j], brOff[j+1]) retyped in order from ONE uniform stream
 * uniforms[uOff[j] .. uOff[j+1]) -- e.g. the <= 3 branches of one NodeRetype proposal
 * (operators/NodeRetype.java:52-56).  Uniforms are consumed in the reference's order
 * (per branch: 1 for the event count, nVirt times, nVirt forward-backward draws).
 * nVirtForced (may be NULL): >= 0 skips the event-count draw of that branch.
 * Outputs per branch: nChg, changes written at outCap*branch (at most outCap of them),
 * logProb, status, nVirt; per job: uUsed.  A branch with status NO_VALID_PATH /
 * PBA_RUBBISH / FB_FELL_THROUGH / NEEDS_POISSON_TAIL / NEED_MORE_UNIFORMS stops its job;
 * the job's later branches get MTT_ST_SKIPPED. */
int mtt_unif_retype_batch(mtt_ctx *ctx, int64_t nJobs, const int64_t *brOff,
                          const int32_t *model, const double *tBottom, const double *tTop,
                          const uint8_t *typeBottom, const uint8_t *typeTop,
                          const int32_t *nVirtForced, const double *uniforms,
                          const int64_t *uOff, int outCap, int32_t *nChg, double *outChgTime,
                          uint8_t *outChgType, double *logProb, int32_t *status,
                          int32_t *nVirt, int64_t *uUsed);

/* NodeRetype.proposal (operators/NodeRetype.java:31-63) in ONE device call: the probability of
 * the current typing of nOld branches (:40-44, arguments and outputs of
 * mtt_unif_branch_logp_batch) and the retyping of nJobs jobs (:52-56, arguments and outputs of
 * mtt_unif_retype_batch).  The first draws no random numbers, so it runs beside the second on
 * another stream; results are those of the two separate calls, bit for bit. */
int mtt_unif_node_retype_batch(mtt_ctx *ctx, int64_t nOld, const int32_t *modelOld,
                               const double *tBottomOld, const double *tTopOld,
                               const uint8_t *typeBottomOld, const uint8_t *typeTopOld,
                               const int64_t *chgOff, const double *chgTime,
                               const uint8_t *chgType, double *logProbOld, int32_t *statusOld,
                               int64_t nJobs, const int64_t *brOff, const int32_t *model,
                               const double *tBottom, const double *tTop,
                               const uint8_t *typeBottom, const uint8_t *typeTop,
                               const int32_t *nVirtForced, const double *uniforms,
                               const int64_t *uOff, int outCap, int32_t *nChg,
                               double *outChgTime, uint8_t *outChgType, double *logProb,
                               int32_t *status, int32_t *nVirt, int64_t *uUsed);

/* ------------------------------------------------------------------ next row: untyped density
 * Replaces distributions/StructuredCoalescentUntypedTreeDensity.java:116-207 (calculateLogP)
 * with :216-290 (colourFirstLineage), :292-316 (colourLineage), :386-512 (addTypeChanges):
 * the importance-sampling estimate of the structured-coalescent density of an UNTYPED tree,
 * batched over nTrees trees x nParticles particles (one warp per particle).
 * Models: those installed by mtt_unif_set_models with sym = 0 (backward rates = Q's
 * off-diagonals); model[nTrees] (may be NULL = model 0); N[nModels][d] population sizes.
 * Tree t = nodes nodeOff[t]..nodeOff[t+1] in BEAST numbering (leaves first, odd count <= 1024),
 * parent -1 = root; leafType[node] is read for leaves only.
 * Randomness (caller's Randomizer): leafOrder holds, tree-major then particle-major, one
 * permutation of the tree's leaves per particle (Randomizer.shuffled, :129); particle
 * q = t * nParticles + p consumes uniforms[uOff[q] .. uOff[q+1]) in the reference's order.
 * Outputs: logP[nTrees]; optional per particle logW, status (MTT_ST_OK or
 * MTT_ST_NEED_MORE_UNIFORMS: resubmit with longer windows), uUsed. */
int mtt_untyped_logp_batch(mtt_ctx *ctx, int64_t nTrees, int nParticles, const int32_t *model,
                           const double *N, const int64_t *nodeOff, const double *height,
                           const int32_t *parent, const uint8_t *leafType,
                           const int32_t *leafOrder, const double *uniforms,
                           const int64_t *uOff, double *logP, double *logW, int32_t *status,
                           int64_t *uUsed);

/* The typings themselves: every particle of mtt_untyped_logp_batch returned as a typed tree.
 * Replaces evolution/tree/MultiTypeTreeFromUntypedNewick.java:104-134 (the lineage-by-lineage
 * colouring of an untyped tree: colourFirstLineage :150-199, colourLineage :202-212,
 * addTypeChanges :272-357 -- the same sampler as the density's), batched over trees x
 * particles; inputs as above.  Outputs per particle q = t * nParticles + p:
 *   nodeType[nodeOff[t] * nParticles + p * n .. + n)  the type of every node (255 = not reached:
 *       status != MTT_ST_OK);
 *   nChg[q] type changes, the first min(nChg[q], chgCap) of them in chgTime / chgType /
 *       chgNode[q * chgCap ..): chgNode = the node BELOW the change (MultiTypeNode.addChange),
 *       chgType = the type entered going rootward; the changes of one node appear in increasing
 *       time.  nChg[q] > chgCap: resubmit with more room.  Times along the conditioned paths
 *       (u * L + t0) are bit-reproducible against the CPU restatement; the waiting times of the
 *       first, unconditioned lineage go through log() (Randomizer.nextExponential) and agree to
 *       the last place.
 *   logW (optional): the particle's importance weight, as in mtt_untyped_logp_batch.
 * status: MTT_ST_OK, MTT_ST_NEED_MORE_UNIFORMS, or the status of the lineage that has no valid
 * path (MTT_ST_NO_VALID_PATH / _PBA_RUBBISH / _FB_FELL_THROUGH): the reference starts the whole
 * colouring over with a new leaf order (:128-131). */
int mtt_untyped_sample_batch(mtt_ctx *ctx, int64_t nTrees, int nParticles, const int32_t *model,
                             const double *N, const int64_t *nodeOff, const double *height,
                             const int32_t *parent, const uint8_t *leafType,
                             const int32_t *leafOrder, const double *uniforms,
                             const int64_t *uOff, int chgCap, uint8_t *nodeType, int32_t *nChg,
                             double *chgTime, uint8_t *chgType, int32_t *chgNode, double *logW,
                             int32_t *status, int64_t *uUsed);

/* ------------------------------------------------------------------ piece 2: peeling
 * Replaces BEAST 2 core beast.base.evolution.likelihood.TreeLikelihood.calculateLogP +
 * BeerLikelihoodCore as wired by the reference at examples/structuredCoalescent.xml:568-572
 * and fxtemplates/MultiTypeTree.xml:184-197 (Inputs data, tree, siteModel, branchRateModel).
 * 4-state (nucleotide) data. */
typedef struct mtt_peel mtt_peel;

/* nTips taxa, nPatterns site patterns owned by THIS handle (a pattern shard when sharded
 * across GPUs), nCat site-rate categories. */
int mtt_peel_create(mtt_ctx *ctx, int nTips, int64_t nPatterns, int nCat, mtt_peel **out);
int mtt_peel_destroy(mtt_peel *h);
/* states[nTips][nPatterns] (Alignment.getPattern; state >= 4 = gap / ambiguity code, treated
 * as missing) and integer pattern weights (Alignment.getPatternWeight). */
int mtt_peel_set_tip_states(mtt_peel *h, const int32_t *states);
int mtt_peel_set_pattern_weights(mtt_peel *h, const int32_t *weights);
/* SiteModel: category rates and proportions (already adjusted for pInv as BEAST's SiteModel
 * does), root frequencies, proportion invariant; strict-clock branchRate. */
int mtt_peel_set_site_model(mtt_peel *h, const double *catRates, const double *catProps,
                            const double *freqs, double pInv, double branchRate);
/* HKY substitution model (examples/structuredCoalescent.xml:534-539): transition matrices are
 * evaluated on the device from node heights. */
int mtt_peel_set_hky(mtt_peel *h, double kappa);
/* log-likelihoods of B trees over this handle's patterns:
 * parent/height [B][2*nTips-1] (tree SoA above).  logL[B] = sum_p weight_p * logL_p. */
int mtt_peel_logl_batch(mtt_peel *h, int64_t B, const int32_t *parent, const double *height,
                        double *logL);
/* same with a substitution / clock parameter set PER TREE -- kappa[B] and branchRate[B] (NULL =
 * the handle's rate): B parameter proposals (BEAST ScaleOperator on hky.kappa / mutationRate,
 * examples/structuredCoalescent.xml:629-640) or B chains with their own parameters in one call. */
int mtt_peel_logl_batch_params(mtt_peel *h, int64_t B, const int32_t *parent,
                               const double *height, const double *kappa,
                               const double *branchRate, double *logL);
/* same under a relaxed clock (UCRelaxedClockModel / RandomLocalClockModel of
 * fxtemplates/MTTClockModels.xml:40-120): branchRates[B][2*nTips-1] holds
 * branchRateModel.getRateForBranch(node) for the branch above every node (the root's entry is
 * unused) and replaces the handle's strict-clock rate. */
int mtt_peel_logl_batch_rates(mtt_peel *h, int64_t B, const int32_t *parent, const double *height,
                              const double *branchRates, double *logL);
/* same with explicit transition matrices mats[B][2*nTips-1][nCat][16] (row = parent state),
 * for substitution models other than HKY (root entry unused). */
int mtt_peel_logl_batch_mats(mtt_peel *h, int64_t B, const int32_t *parent, const double *mats,
                             double *logL);
/* device-pointer variant of mtt_peel_logl_batch (async on the ctx stream) */
int mtt_peel_logl_batch_dev(mtt_peel *h, int64_t B, const int32_t *parent, const double *height,
                            double *logL);
/* Pattern shards with a proportion of invariant sites: BEAST's TreeLikelihood returns the
 * evaluation without partial rescaling unless ANY pattern of the alignment underflowed in it
 * (then the rescaled evaluation, which treats pInv differently).  A shard cannot make that
 * choice alone, so this entry returns, per tree, parts[0][B] = the unscaled sum over this
 * handle's patterns, parts[1][B] = the rescaled sum, parts[2][B] = number of tiles that saw an
 * underflow; add the three over the shards (one all-reduce of 3 B doubles) and take
 * parts[2] > 0 ? parts[1] : parts[0].  With pInv = 0 both sums are the same number.
 * The _dev entry takes device pointers and is asynchronous on the ctx stream. */
int mtt_peel_logl_parts_batch(mtt_peel *h, int64_t B, const int32_t *parent, const double *height,
                              double *parts);
int mtt_peel_logl_parts_batch_dev(mtt_peel *h, int64_t B, const int32_t *parent,
                                  const double *height, double *parts);
/* per-pattern log-likelihoods of ONE tree (TreeLikelihood.getPatternLogLikelihoods) */
int mtt_peel_pattern_logl(mtt_peel *h, const int32_t *parent, const double *height,
                          double *patternLogL);

/* Resident chains: incremental peeling.  Replaces what TreeLikelihood [EXT] does between MCMC
 * steps -- BeerLikelihoodCore keeps two banks of partials per node
 * (currentPartialsIndex / storedPartialsIndex), TreeLikelihood.traverse recomputes only the
 * nodes above a dirty one, store() / restore() flip the banks -- for nChains chains that share
 * the alignment of `h`.  The partials of every internal node of every chain stay in HBM
 * (nChains x 2 x (nTips-1) x (4 nCat + 1) x patterns x 8 bytes; TOO_LARGE when that does not fit).
 *   update  evaluates one proposal per chain: parent / height [nChains][2 nTips - 1]; optional
 *           branchRates (same shape, relaxed clock), kappa[nChains], clockRate[nChains] (NULL =
 *           the handle's values).  The proposal is compared with the chain's ACCEPTED tree and
 *           only the internal nodes above a changed branch are recomputed (all of them the first
 *           time, or when kappa / the strict clock rate / the site model / the alignment changed).
 *           dirty (optional, [nChains][2 nTips - 1]): the nodes that MAY differ, as
 *           Node.isDirty() != IS_CLEAN -- only those are compared, which makes a step
 *           O(dirty x depth) on the host; NULL compares every node.  A node that differs but is
 *           not flagged is the caller's error, as in BEAST.  logL[nChains]: the proposal's
 *           log-likelihood (the accepted one where nothing changed; NaN for a malformed tree,
 *           which leaves the chain untouched).
 *   accept  accept[k] != 0 (or accept == NULL): the proposal becomes chain k's accepted state
 *           (BEAST: the next store()); 0: it is dropped (restore()).  A proposal still unanswered
 *           at the next update counts as rejected.  Host-side only.
 * Results agree with mtt_peel_logl_batch on the same trees to rounding (the per-pattern log
 * scale factors are summed per subtree instead of along the traversal).  */
typedef struct mtt_peel_chains mtt_peel_chains;
int mtt_peel_chains_create(mtt_peel *h, int nChains, mtt_peel_chains **out);
int mtt_peel_chains_destroy(mtt_peel_chains *s);
int mtt_peel_chains_update(mtt_peel_chains *s, const int32_t *parent, const double *height,
                           const double *branchRates, const double *kappa,
                           const double *clockRate, const uint8_t *dirty, double *logL);
/* update in sparse form: chain k's proposal is the list of nodes node[off[k] .. off[k+1]) that
 * may differ from its accepted tree, with their new parent / height (/ branchRates, may be NULL
 * under a strict clock) given PER LIST ENTRY -- what a BEAST shim has in hand (the dirty nodes)
 * without flattening the whole tree.  Nodes not listed keep their accepted values; a node may be
 * listed twice.  MTT_ERR_STATE when a chain has no accepted tree yet or its kappa / strict clock
 * rate differs from the accepted one (the whole tree is dirty then: use mtt_peel_chains_update). */
int mtt_peel_chains_update_sparse(mtt_peel_chains *s, const int64_t *off, const int32_t *node,
                                  const int32_t *parent, const double *height,
                                  const double *branchRates, const double *kappa,
                                  const double *clockRate, double *logL);
/* update on a pattern shard (the handle holds a contiguous slice of the patterns, the chains'
 * trees are replicated on every rank): parts[3][nChains] as mtt_peel_logl_parts_batch -- summed
 * over the shards, parts[2] > 0 ? parts[1] : parts[0] is the proposal's log-likelihood; every
 * rank then answers with the same accept mask */
int mtt_peel_chains_update_parts(mtt_peel_chains *s, const int32_t *parent, const double *height,
                                 const double *branchRates, const double *kappa,
                                 const double *clockRate, const uint8_t *dirty, double *parts);
int mtt_peel_chains_accept(mtt_peel_chains *s, const uint8_t *accept);
/* nodes recomputed by the last update, over all chains */
int64_t mtt_peel_chains_last_entries(const mtt_peel_chains *s);
/* integration aid: with on != 0, an update that is given a dirty mask still compares every node
 * and fails with MTT_ERR_INVALID_ARG (naming the chain and the node) when one differs from the
 * accepted tree without being flagged -- the mistake that otherwise goes unnoticed, here as in
 * BEAST.  Costs the O(nodes) comparison the mask exists to avoid: for tests of a binding. */
int mtt_peel_chains_set_verify(mtt_peel_chains *s, int on);
/* measurement: bracket the kernel launches of every update with CUDA events on the ctx stream;
 * last_device_ms = the last update's device time (0 when nothing was launched) */
int mtt_peel_chains_set_timing(mtt_peel_chains *s, int on);
double mtt_peel_chains_last_device_ms(const mtt_peel_chains *s);

/* ------------------------------------------------------------------ next row: tree simulation
 * Replaces evolution/tree/StructuredCoalescentMultiTypeTree.java:191-442 (simulateTree,
 * updatePropensities, getNextEvent, updateTree): B structured-coalescent multi-type trees
 * simulated on the device (one thread per tree), written as the flattened tree SoA above in
 * the owner layout -- synthetic inputs for the density / peeling entries without a trip through
 * the host, and the ground truth of the reference's operator tests
 * (util/UtilMethods.java:29-47, 10^5 replicates).  m, N, leafTypes[nLeaves],
 * leafTimes[nLeaves] are host arrays; tree b uses the stream seeded with seed + b.  The event
 * loop and the draw order are those of the host generator (mtt_sim_run): same seed, same tree
 * (times equal to rounding).  Nodes: leaves 0..nLeaves-1 as given, internal nodes in order of
 * creation (= increasing height), root last. */
typedef struct mtt_sim mtt_sim;
int mtt_sim_trees_create(mtt_ctx *ctx, int64_t B, int nLeaves, int d, const double *m,
                         const double *N, const uint8_t *leafTypes, const double *leafTimes,
                         uint64_t seed, mtt_sim **out);
int mtt_sim_trees_destroy(mtt_sim *h);
int64_t mtt_sim_trees_total_changes(const mtt_sim *h);
/* device pointers of the simulated batch (owned by the handle); any argument may be NULL */
int mtt_sim_trees_device(const mtt_sim *h, const int64_t **nodeOff, const double **height,
                         const int32_t **parent, const uint8_t **nodeType,
                         const uint16_t **chgNode, const int64_t **chgBase,
                         const double **chgTime, const uint8_t **chgType);
/* copy to host arrays sized B+1 / B*(2*nLeaves-1) / total_changes; any argument may be NULL */
int mtt_sim_trees_export(const mtt_sim *h, int64_t *nodeOff, double *height, int32_t *parent,
                         uint8_t *nodeType, uint16_t *chgNode, int64_t *chgBase, double *chgTime,
                         uint8_t *chgType);

#ifdef __cplusplus
}
#endif
#endif
