/*
 * sfb200.h -- C ABI of libsfb200.so: the B200 (sm_100a) implementation of Sailfish's quantification hot path
 *             (read batch -> quasi-mapping -> equivalence-class counts -> EM / VBEM / bootstrap / Gibbs).
 *
 * The reference (kingsfordgroup/sailfish v0.10.0) has no plugin / FFI layer: the path sits behind five C++ call
 * sites inside src/SailfishQuantify.cpp.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference tree); sailfish_b200/host/ holds C++ adaptors with the reference's own
 * class / method names on top of this ABI, and INTEGRATION.md shows the patch a Sailfish maintainer would apply.
 *
 * Conventions: plain C, no exceptions cross the boundary.  Every call returns 0 on success or a negative
 * SFB200_E* code; sfb200_last_error(ctx) gives the message.  Host buffers are caller-owned.  A context is bound
 * to one CUDA device; one call in flight per context.  There is NO CPU fallback: without a CUDA device every
 * call fails with SFB200_ENODEV.
 */
#ifndef SFB200_H
#define SFB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFB200_OK          0
#define SFB200_ENODEV     -1   /* no usable CUDA device */
#define SFB200_ECUDA      -2   /* CUDA runtime error (message has the call) */
#define SFB200_EINVAL     -3   /* bad argument / call order */
#define SFB200_ENOACTIVE  -4   /* optimize(): "no transcripts are expressed" (CollapsedEMOptimizer.cpp:794-798) */
#define SFB200_ESMALLSUM  -5   /* optimize(): "Total alpha weight was too small" (CollapsedEMOptimizer.cpp:877-881) */
#define SFB200_EFULL      -6   /* equivalence-class table / label arena exhausted */
#define SFB200_ENCCL      -7   /* NCCL error or NCCL not loadable */
#define SFB200_ECALLBACK  -8   /* a row callback returned non-zero */

typedef struct sfb200_ctx sfb200_ctx;

/* ---- context ------------------------------------------------------------------------------------------------ */
int  sfb200_version(void);
int  sfb200_ctx_create(int device, sfb200_ctx** out);
void sfb200_ctx_destroy(sfb200_ctx* ctx);
const char* sfb200_last_error(const sfb200_ctx* ctx);
/* Launch on a caller-owned stream (e.g. torch's current stream, so the caller's CUDA events bracket the work);
 * NULL restores the context's own stream. */
int  sfb200_ctx_set_stream(sfb200_ctx* ctx, void* cuda_stream);
int  sfb200_ctx_sync(sfb200_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t sfb200_launch_count(const sfb200_ctx* ctx);
/* pinned host memory for read batches (so H2D overlaps mapping) */
void* sfb200_host_alloc(size_t bytes);
void  sfb200_host_free(void* p);
/* One process per GPU on a multi-socket host: restrict the CALLING thread (and the threads it creates afterwards) to the CPUs of the
 * NUMA node the device hangs off, so that page-locked read buffers allocated afterwards are node-local and the H2D copies of several
 * ranks do not cross the socket interconnect.  Returns the number of CPUs in the new affinity mask, 0 when nothing was changed
 * (single node, node or device unknown, or SFB200_NO_BIND=1), < 0 on error.  Call before sfb200_host_alloc / before starting parser threads.
 * (The reference has no counterpart: its parser and mapper threads share one address space, src/SailfishQuantify.cpp:445-520.) */
int  sfb200_bind_host_near_device(int device);

/* ---- multi-GPU: one process per GPU, reads sharded, one all-reduce of the per-transcript vector per EM iteration -
 * The reference is single-process; this is the exchange step BASELINE.json's north_star adds.
 * id: 128-byte ncclUniqueId made on rank 0 by sfb200_comm_unique_id and broadcast by the launcher. */
int sfb200_comm_unique_id(uint8_t id_out[128]);
int sfb200_comm_init(sfb200_ctx* ctx, int n_ranks, int rank, const uint8_t id[128]);

/* ---- index ---------------------------------------------------------------------------------------------------
 * Replaces what ReadExperiment takes from RapMapSAIndex<IndexT> after SailfishIndex::load
 * (include/SailfishIndex.hpp:28-43,104-144; fields seq / txpOffsets / txpLens, include/ReadExperiment.hpp:103-116):
 * the device index (2-bit text, k-mer-bucketed suffix array, k-mer hash table, presence filter) is built on the GPU
 * from the transcript sequences.  seq: ASCII, transcript t = seq[txp_off[t] .. txp_off[t]+txp_len[t]).  k odd, <= 31
 * (src/SailfishIndexer.cpp:79,199-205). */
int sfb200_index_build(sfb200_ctx* ctx, const char* seq, const uint64_t* txp_off, const uint32_t* txp_len,
                       uint32_t n_txp, int k);
/* stats[0] text_len [1] n_suffixes [2] n_kmers [3] table_slots [4] index bytes in HBM [5] max bucket */
int sfb200_index_stats(const sfb200_ctx* ctx, uint64_t stats[8]);
/* copy the index back to the host (tests, and bench.py's CPU arm): words[text_len/32+2], sa_pos[n_suffixes],
 * sa_tid[n_suffixes]; any pointer may be NULL */
int sfb200_index_export(sfb200_ctx* ctx, uint64_t* words, uint32_t* sa_pos, uint32_t* sa_tid);
/* the k-mer table: table_slots entries of {k-mer (u64), first entry (u32), entry count (u32)}; empty = all-ones k-mer;
 * slot = 2 * (mix(k-mer) & (table_slots/2 - 1)), linear probing (mix = sfb_kmer_mix, sailfish_b200/csrc/common.cuh) */
int sfb200_index_export_table(sfb200_ctx* ctx, void* table16);
/* The device index as ONE file (what SailfishIndex::load reads from the index directory, include/SailfishIndex.hpp:80-144: the
 * RapMap structures ready to use): a header (magic, format version, k, sizes) followed by the arrays as they lie in HBM -- packed
 * text, transcript starts / lengths, suffix entries, k-mer table, m-mer bitmap.  save streams HBM -> file in 64 MB pieces, load file
 * -> HBM; a file written for another format version or a truncated one is refused (SFB200_EINVAL).  Building the index from the
 * sequences takes 0.1 - 0.6 s on a B200 (200 k - 1 M transcripts), so the drivers rebuild by default and use the file on request. */
int sfb200_index_save(sfb200_ctx* ctx, const char* path);
int sfb200_index_load(sfb200_ctx* ctx, const char* path);

/* ---- mapping + equivalence classes ---------------------------------------------------------------------------
 * Replaces processReadsQuasi<IndexT> (src/SailfishQuantify.cpp:105-452 paired, :458-646 single) together with the
 * EquivalenceClassBuilder it feeds (include/EquivalenceClassBuilder.hpp:62-110).  Fields = the SailfishOpts members
 * processReadsQuasi reads (include/SailfishOpts.hpp:9-41) plus rl.format(). */
typedef struct {
    uint32_t max_read_occs;     /* sfOpts.maxReadOccs      (200)   */
    uint32_t max_frag_len;      /* sfOpts.maxFragLen       (1000)  */
    int32_t  num_frag_samples;  /* sfOpts.numFragSamples   (10000) */
    int32_t  lib_format_id;     /* rl.format().formatID()  (include/LibraryFormat.hpp:89-98) */
    int32_t  strict_intersect;  /* sfOpts.strictIntersect  */
    int32_t  allow_orphans;     /* sfOpts.allowOrphans     */
    int32_t  allow_dovetail;    /* sfOpts.allowDovetail    */
    int32_t  ignore_compat;     /* sfOpts.ignoreLibCompat  */
    int32_t  enforce_compat;    /* sfOpts.enforceLibCompat */
    uint32_t max_interval;      /* k-mer buckets larger than this are not used as seeds (1000) */
} sfb200_map_opts;

/* == eqBuilder.start() (SailfishQuantify.cpp:1322): resets the class table, counters and the FLD sampler */
int sfb200_map_begin(sfb200_ctx* ctx, const sfb200_map_opts* opts);
/* One batch of reads from HOST memory (what a parser job holds, PairSequenceParser.hpp:16-24, concatenated):
 * read i = bases[off[i] .. off[i+1]).  bases2/off2 NULL for a single-end library.  Includes the H2D copy. */
int sfb200_map_batch(sfb200_ctx* ctx, const char* bases1, const uint64_t* off1, const char* bases2,
                     const uint64_t* off2, uint64_t n_reads);
/* The same for reads of ONE length per mate, stored back to back (read i of mate 1 = bases1[i*len1 .. (i+1)*len1)): no offsets array
 * -- the usual sequencer output; 8 of ~84 bytes per read less over the host link.  bases2 NULL for a single-end library. */
int sfb200_map_batch_fixed(sfb200_ctx* ctx, const char* bases1, uint32_t len1, const char* bases2, uint32_t len2, uint64_t n_reads);
/* Same with buffers already resident in device memory. */
int sfb200_map_batch_device(sfb200_ctx* ctx, const char* d_bases1, const uint64_t* d_off1, const char* d_bases2,
                            const uint64_t* d_off2, uint64_t n_reads);
/* Read ingestion on the device (replaces the parser threads of quasiMapReads, src/SailfishQuantify.cpp:882-898,
 * 996-1005, include/PairSequenceParser.hpp:28-191, for plain four-line FASTQ and two-line FASTA reads -- the first character of
 * text1 says which).  text1 (and text2 for a paired library) = a block of FASTQ TEXT in HOST memory that starts at a record boundary; its complete records -- at most max_records (0 = no limit), the same
 * number from both mates -- are extracted on the GPU and mapped as by sfb200_map_batch.  *n_records = how many, *consumed1/2 = the
 * bytes of text they cover: the caller keeps text[consumed ..) and puts it in front of what it reads next.  A final record
 * without a trailing newline needs one appended.  Gzipped input is inflated by the caller first.
 * Parity: tests/test_gpu_map.py::test_map_fastq_equals_map_batch (green on a B200, profiles/r02a_experimental_gpu.txt); CPU check of the
 * arithmetic: tests/fastq_core_test.cpp. */
int sfb200_map_fastq(sfb200_ctx* ctx, const char* text1, uint64_t n1, const char* text2, uint64_t n2, uint64_t max_records,
                     uint64_t* n_records, uint64_t* consumed1, uint64_t* consumed2);
/* --biasCorrect / --gcBiasCorrect: collect, while mapping, what the reference collects in processReadsQuasi
 * (src/SailfishQuantify.cpp:255-287 and :555-583: the 6-mer context around the start of each read's first hit that has one,
 * readBias().update, for the first num_bias_samples such reads in read order -- sfOpts.numBiasSamples, the reference at -p 1;
 * :372-389: observedGC()[gcFrac(start, stop)]++ for every properly paired hit inside its transcript).
 * Call after map_begin and before the first batch.  Parity: tests/test_gpu_map.py. */
int sfb200_map_set_bias(sfb200_ctx* ctx, int seq_bias, int gc_bias, int32_t num_bias_samples);
/* readExp.readBias().counts (4096 bins) and readExp.observedGC() (101 bins), both with the reference's initial count of 1 per
 * bin -- the arrays sfb200_bias_model takes.  With a communicator: summed over ranks. */
int sfb200_map_get_bias(sfb200_ctx* ctx, uint32_t* read_bias, uint32_t* observed_gc);

/* == thread join + eqBuilder.finish() (SailfishQuantify.cpp:942-947,1328).
 * counters: [0] numObservedFragments [1] numMappedFragments [2] numFragHits [3] upperBoundHits [4] numFwd [5] numRC
 * (ReadExperiment.hpp:74-97); fld_hist[max_frag_len] = flMap (SailfishQuantify.cpp:867).  With a communicator the
 * counters are summed over ranks, fld_hist holds the first num_frag_samples eligible fragments in global read order, and
 * the classes of all ranks are merged (every rank then holds the global class set; set SFB200_MULTI_EM_ALLREDUCE=1 to keep
 * them rank-local and all-reduce the per-transcript vector every EM iteration instead). */
int sfb200_map_finish(sfb200_ctx* ctx, uint64_t counters[6], uint32_t* fld_hist, uint64_t* n_classes, uint64_t* nnz);
/* device time of all mapping-kernel launches between map_begin and map_finish, in milliseconds (CUDA events) */
double sfb200_last_map_kernel_ms(const sfb200_ctx* ctx);
/* bytes sfb200_map_batch[_fixed] has sent to the device since map_begin (bench.py's h2d_bytes_per_step).  Host batches travel in
 * pieces of SFB200_HOST_PIECE reads (512 k), the copy of a piece overlapping the kernels of the one before. */
uint64_t sfb200_map_h2d_bytes(const sfb200_ctx* ctx);
/* Mates longer than 256 bases are mapped by their first 256 (mapping spec v1; the reference maps the whole read,
 * SailfishQuantify.cpp:192-213): how many mates were cut since map_begin.  The drivers print a warning when it is not zero. */
uint64_t sfb200_map_clipped(sfb200_ctx* ctx);
/* == eqVec() (EquivalenceClassBuilder.hpp:110) as CSR in canonical order (label-lexicographic);
 * this is the content of aux/eq_classes.txt (src/GZipWriter.cpp:51-92) */
int sfb200_eq_export(sfb200_ctx* ctx, uint64_t* row_ptr, uint32_t* labels, uint64_t* counts);
/* inverse: run inference without mapping (the commented-out loadEquivClasses, SailfishQuantify.cpp:1444-1495) */
int sfb200_eq_import(sfb200_ctx* ctx, uint32_t n_txp, uint64_t n_classes, const uint64_t* row_ptr,
                     const uint32_t* labels, const uint64_t* counts);

/* ---- inference ------------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t  use_vb;        /* sopt.useVBOpt */
    double   prior_alpha;   /* 0.01  (CollapsedEMOptimizer.cpp:786) */
    double   tol;           /* relDiffTolerance, 0.01 (SailfishQuantify.cpp:1343) */
    uint32_t min_iter;      /* 50    (CollapsedEMOptimizer.cpp:716) */
    uint32_t max_iter;      /* 10000 (SailfishQuantify.cpp:1343) */
    uint32_t fixed_iters;   /* >0: exactly this many iterations, convergence ignored */
    double   check_cutoff;  /* 1e-2  (CollapsedEMOptimizer.cpp:811) */
    double   min_alpha;     /* 1e-8  (CollapsedEMOptimizer.cpp:810) */
} sfb200_em_opts;
void sfb200_em_default_opts(sfb200_em_opts* o);

/* Replaces CollapsedEMOptimizer::optimize (include/CollapsedEMOptimizer.hpp:25-28, src/CollapsedEMOptimizer.cpp:711-893).
 * eff_lens[t] = noEffectiveLengthCorrection ? RefLength : EffectiveLength; num_mapped = numMappedFragments() (global).
 * alphas_out[t] -> Transcript::setEstCount; mass = alphas/sum.  Classes come from map_finish or eq_import.
 * With a communicator and rank-local classes: one all-reduce(sum) of the T-vector per iteration. */
int sfb200_em_run(sfb200_ctx* ctx, const double* eff_lens, uint32_t n_txp, uint64_t num_mapped,
                  const sfb200_em_opts* opts, double* alphas_out, uint32_t* iters_out, double* max_rel_diff_out);
/* device time of the iteration loop of the last em_run / bootstrap in milliseconds (CUDA events on the launch stream) */
double sfb200_last_em_loop_ms(const sfb200_ctx* ctx);
/* which iteration loop the last em_run / bootstrap used: 0 binned layout, global-memory scatter (k_em_persistent);
 * 1 CTA-partitioned, shared-memory scatter (k_em_part); 2 CTA-partitioned, atomic-free gather form (k_em_gather);
 * 3 one launch per phase (rank-local classes with a per-iteration all-reduce, or SFB200_EM_MODE=steps);
 * 4 one thread per connected component of the class structure (k_em_dense);
 * 5 k_em_dense with a pool loop: small components on component threads, the rest (large components, classes that cross CTA ranges)
 *   as an independent sub-problem on CTAs of their own */
int sfb200_last_em_kernel(const sfb200_ctx* ctx);
/* how k_em_dense ran (0 for the other kernels): bit 0 = class counts / base / 1/effLen streamed from global memory (the slice of a CTA
 * did not fit in shared memory), bit 1 = lagged stopping rule */
int sfb200_last_em_variant(const sfb200_ctx* ctx);

/* Not called by the quantification drivers yet (parity-tested on its own: tests/test_gpu_bias.py).
 * Replaces sailfish::utils::updateEffectiveLengths (src/SailfishUtils.cpp:611-926): effective lengths corrected for
 * sequence-specific (--biasCorrect) or fragment-GC (--gcBiasCorrect) bias from the current abundances.  The model is what the
 * reference reads from ReadExperiment: readBias().counts (include/ReadKmerDist.hpp:16-24, pseudo-counts included),
 * observedGC(), numFwd()/numRC(), and fragLengthDist() as its float cdf table (cdf(x) = 1 for x >= n_cdf) and maxValue().
 * eff_model[t] = Transcript::EffectiveLength (the fragment-length model's value), eff_in = the optimizer's current vector. */
typedef struct {
    int32_t  mode;              /* 1 = --biasCorrect, 2 = --gcBiasCorrect */
    uint32_t gc_samp;           /* sopt.pdfSampFactor (--gcSpeedSamp), 1 */
    int64_t  num_fwd, num_rc;
    const uint32_t* read_bias;  /* 4096 (mode 1) */
    const uint32_t* observed_gc;/* 101  (mode 2) */
    const float* fld_cdf; uint32_t n_cdf; uint32_t fld_max;
} sfb200_bias_model;
int sfb200_bias_eff_lens(sfb200_ctx* ctx, const sfb200_bias_model* model, const double* eff_model, const double* eff_in,
                         const double* alphas, uint32_t n_txp, double* eff_out);
/* Replaces CollapsedEMOptimizer::optimize when sopt.biasCorrect or sopt.gcBiasCorrect is set (src/CollapsedEMOptimizer.cpp:820-840):
 * as sfb200_em_run, and at the top of iterations 50, 500 and 1000 the effective lengths are recomputed from the current alphas
 * (sfb200_bias_eff_lens) and the class weights with them (updateEqClassWeights, :527-556).  eff_lens[t] = Transcript::EffectiveLength;
 * eff_out (optional) = the lengths the run ended with, which the reference writes to quant.sf (:888).  The index must be resident
 * (the correction reads the transcript sequences).  Single rank, or classes merged over ranks.
 * Parity: tests/test_gpu_bias.py (needs SFB200_EXPERIMENTAL=1 until it has had its first GPU run). */
int sfb200_em_run_bias(sfb200_ctx* ctx, const double* eff_lens, uint32_t n_txp, uint64_t num_mapped, const sfb200_em_opts* opts,
                       const sfb200_bias_model* model, double* alphas_out, double* eff_out, uint32_t* iters_out,
                       double* max_rel_diff_out);

typedef int (*sfb200_f64_row_cb)(void* user, const double* row, size_t n);
typedef int (*sfb200_i32_row_cb)(void* user, const int32_t* row, size_t n);
/* Replaces CollapsedEMOptimizer::gatherBootstraps (src/CollapsedEMOptimizer.cpp:557-709): cb == writeBootstrap(alphas) */
int sfb200_bootstrap_run(sfb200_ctx* ctx, const double* eff_lens, uint32_t n_txp, const sfb200_em_opts* opts,
                         uint32_t n_boot, uint64_t seed, sfb200_f64_row_cb cb, void* user);
/* doBootstrap's EM on caller-supplied resampled counts (counts in eq_export order); parity hook for tests */
int sfb200_bootstrap_em(sfb200_ctx* ctx, const double* eff_lens, uint32_t n_txp, const uint64_t* samp_counts,
                        const sfb200_em_opts* opts, double* alphas_out, uint32_t* iters_out);
/* Replaces CollapsedGibbsSampler::sample (include/CollapsedGibbsSampler.hpp:27-31, src/CollapsedGibbsSampler.cpp:199-291):
 * masses[t] = Transcript::mass() after optimize; cb == writeSample(counts) */
int sfb200_gibbs_run(sfb200_ctx* ctx, const double* eff_lens, const double* masses, uint32_t n_txp, uint64_t num_mapped,
                     uint32_t n_samples, uint64_t seed, sfb200_i32_row_cb cb, void* user);

/* hashing primitive exposed for known-answer tests: XXH64 (src/xxhash.c:346-455) of n messages computed ON THE DEVICE;
 * message i = data[off[i]..off[i+1]) (lengths must be multiples of 4, as labels are) */
int sfb200_xxh64_device(sfb200_ctx* ctx, const uint8_t* data, const uint64_t* off, uint64_t n, uint64_t seed, uint64_t* out);
/* device digamma (VBEM) for accuracy tests */
int sfb200_digamma_device(sfb200_ctx* ctx, const double* x, uint64_t n, double* out);
/* exp(digamma(x)) as the VBEM kernels evaluate it (series without the logarithm, csrc/vb_math.hpp) */
int sfb200_exp_digamma_device(sfb200_ctx* ctx, const double* x, uint64_t n, double* out);

#ifdef __cplusplus
}
#endif
#endif
