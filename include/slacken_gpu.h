/*
 * slacken_gpu.h -- C ABI of libslacken_gpu.so: the B200 (sm_100a) replacement for the Kraken-2-style
 * build/classify hot path of JNP-Solutions/Slacken.
 *
 * The reference has no native seam (it is pure Scala inside Spark closures), so each entry point names the
 * reference code it replaces; paths are relative to src/main/scala/com/jnpersson/ of the reference.
 * INTEGRATION.md shows the JNI / Panama binding a Slacken maintainer would add on the Scala side.
 *
 * Conventions
 *  - every call returns 0 (SLK_OK) or a negative SLK_E_* code and never throws or aborts;
 *    slk_last_error() returns a thread-local message for the last failure on the calling thread;
 *  - the caller owns all host buffers (ideally pinned: slk_host_alloc / slk_host_register), the library
 *    owns all device memory behind opaque handles;
 *  - sequences are ASCII, one byte per base, WITHOUT line breaks (InputFragment "does not contain
 *    whitespace", kmers/minimizer/MinSplitter.scala:23-32; KeyValueIndex.getSpans requires the same,
 *    slacken/KeyValueIndex.scala:161-162). A,C,G,T,U in either case are bases, every other byte is ambiguous;
 *  - sequence i of a batch occupies bases[off[i] .. off[i+1]);
 *  - taxon ids are the raw ids of the taxonomy (Taxonomy.parents is indexed by them, NONE = 0, ROOT = 1,
 *    slacken/Taxonomy.scala:30-31,159-160); minimizers are the left-aligned priority words that Slacken stores
 *    in the Parquet column id1 (kmers/util/NTBitArray.scala:124-125), as signed 64-bit integers;
 *  - handles are immutable after creation and may be shared by threads; a slk_classifier owns streams and
 *    scratch and must be used by one thread at a time (create one per Spark task thread).
 *  There is no CPU fallback: without a CUDA device every call fails with SLK_E_CUDA.
 */
#ifndef SLACKEN_GPU_H
#define SLACKEN_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SLK_API __attribute__((visibility("default")))
#else
#define SLK_API
#endif

#define SLK_OK 0
#define SLK_E_INVALID (-1)    /* bad argument / unsupported parameter combination */
#define SLK_E_CUDA (-2)       /* CUDA runtime failure (including: no device) */
#define SLK_E_NOMEM (-3)      /* host or device allocation failed */
#define SLK_E_NOSPACE (-4)    /* an output buffer is too small; the required size is reported */
#define SLK_E_UNSUPPORTED (-5)/* input outside the supported envelope ("limits") */

typedef struct slk_ctx slk_ctx;
typedef struct slk_tax slk_tax;
typedef struct slk_index slk_index;
typedef struct slk_builder slk_builder;
typedef struct slk_classifier slk_classifier;
typedef struct slk_counts slk_counts;

/* IndexParams / SplitterFormat of an index (kmers/IndexParams.scala:63-91, kmers/SplitterFormat.scala:55-77):
 * k, m, minimizerSpaces, XORmask, canonical. Only the `randomXOR` splitter exists in Slacken. */
typedef struct {
  int32_t k, m, spaces, canonical;
  uint64_t toggle_mask;
} slk_params;

/* One merged hit of a read: TaxonCounts after TaxonCounts.fromHits (slacken/TaxonCounts.scala:31-48).
 * taxon = raw taxon id, 0 for a minimizer that is not in the library, -1 for an ambiguous span ("A:n"),
 * -2 for the mate-pair border ("|:|", count = -(k-1)). */
typedef struct {
  int32_t taxon;
  int32_t count;
} slk_hit;

/* Per-read details needed for ClassifiedRead.outputLine (slacken/Classifier.scala:39-45). */
typedef struct {
  uint64_t hit_off;      /* index of the read's first merged hit in hits_out */
  uint32_t hit_cnt;      /* number of merged hits */
  uint32_t len1;         /* lengthString part 1: sum of span k-mers + (k-1) (slacken/TaxonCounts.scala:114-121) */
  uint32_t len2;         /* same for mate 2; 0xFFFFFFFF for single-end */
  uint32_t num_distinct; /* hit groups: spans with distinct && taxon != NONE (slacken/Classifier.scala:94) */
} slk_read_detail;

#define SLK_READ_CLASSIFIED 1u /* ClassifiedRead.classified */
#define SLK_READ_HAS_SPAN 2u   /* 0: the read yields no span, so Slacken emits no line and does not count it */

/* ClassifyParams (slacken/Classifier.scala:47-63), the per-call part. */
typedef struct {
  double confidence;      /* one of ClassifyParams.thresholds */
  int32_t min_hit_groups; /* ClassifyParams.minHitGroups */
  int32_t reserved;
} slk_classify_opts;

/* ClassifyParams with several thresholds (ClassifyParams.thresholds, slacken/Classifier.scala:47-63): Slacken keeps the
 * hits of a read and classifies them once per threshold (Classifier.scala:156-170); so does one call with these options. */
#define SLK_MAX_THRESHOLDS 8
typedef struct {
  uint32_t n_thresholds;  /* 1 .. SLK_MAX_THRESHOLDS */
  int32_t min_hit_groups;
  double confidence[SLK_MAX_THRESHOLDS];
} slk_classify_multi_opts;

SLK_API const char* slk_last_error(void);

/* ---- context ------------------------------------------------------------------------------------------- */
SLK_API int slk_ctx_create(int device, slk_ctx** out);
SLK_API void slk_ctx_destroy(slk_ctx* ctx);
SLK_API int slk_ctx_device(const slk_ctx* ctx);
/* pinned host memory for batch buffers (direct ByteBuffers on the JVM side wrap these) */
SLK_API int slk_host_alloc(size_t bytes, void** out);
SLK_API void slk_host_free(void* p);
SLK_API int slk_host_register(void* p, size_t bytes);
SLK_API int slk_host_unregister(void* p);

/* ---- parameters: SlackenMinimizerFormats.makeSplitter (slacken/SlackenMinimizerFormats.scala:31-42) ------ */
SLK_API int slk_params_init(int k, int m, int spaces, uint64_t toggle_mask, int canonical, slk_params* out);

/* ---- taxonomy: the parents array of slacken/Taxonomy.scala:159-160 (broadcast in KeyValueIndex.scala:44-47) - */
SLK_API int slk_taxonomy_create(slk_ctx* ctx, const int32_t* parents, int32_t n, slk_tax** out);
SLK_API void slk_taxonomy_destroy(slk_tax* tax);

/* ---- B3: library load, replaces KeyValueIndex.loadRecords + the join side (slacken/KeyValueIndex.scala:150-159,
 *          slacken/Classifier.scala:84). Columns of the Parquet table: id1:int64, taxon:int32. ------------------ */
/* id1 / taxon: host OR device pointers (the copies infer the direction; the distributed build hands over device memory) */
SLK_API int slk_index_from_records(slk_ctx* ctx, slk_tax* tax, const slk_params* params, const int64_t* id1,
                           const int32_t* taxon, uint64_t n, slk_index** out);
SLK_API void slk_index_destroy(slk_index* idx);
SLK_API uint64_t slk_index_size(const slk_index* idx);   /* number of records (distinct minimizers) */
/* copy the records back (for the Parquet writer, KeyValueIndex.writeRecords :125-139); order unspecified; the output
 * arrays may be host or device memory */
SLK_API int slk_index_records(slk_index* idx, int64_t* id1_out, int32_t* taxon_out, uint64_t cap, uint64_t* n_out);

/* ---- B2: library build, replaces SplitterMinimizers.find + groupBy(id1).agg(TaxonLCA)
 *          (slacken/Minimizers.scala:43-76, slacken/KeyValueIndex.scala:85-122). -------------------------------- */
SLK_API int slk_build_begin(slk_ctx* ctx, slk_tax* tax, const slk_params* params, uint64_t expected_bases,
                    slk_builder** out);
/* genome fragments with their taxon labels; fragments whose taxon is undefined in the taxonomy are skipped
 * (KeyValueIndex.scala:118-120). Ambiguous characters split a fragment (InputReader.removeInvalid). */
SLK_API int slk_build_add(slk_builder* b, const uint8_t* bases, const uint64_t* frag_off, const int32_t* frag_taxon,
                  uint32_t n_frag);
/* sort + LCA reduce + hash table construction; the builder is consumed (destroy it afterwards) */
SLK_API int slk_build_finish(slk_builder* b, slk_index** out);
SLK_API void slk_build_destroy(slk_builder* b);

/* ---- B1: classify, replaces KeyValueIndex.getSpans + Classifier.spansToGroupedHits + classifyHits
 *          (slacken/KeyValueIndex.scala:163-185, slacken/Classifier.scala:77-147,439-454). ---------------------- */
SLK_API int slk_classifier_create(slk_index* idx, slk_classifier** out);
SLK_API void slk_classifier_destroy(slk_classifier* c);
/* upper bound of merged hits for a batch: size hits_out with it to make SLK_E_NOSPACE impossible */
SLK_API uint64_t slk_classify_hits_bound(const slk_params* p, uint32_t n_reads, uint64_t total_bases, int paired);
/* Host-buffer entry point (what the JVM binding calls from mapPartitions). bases2/off2 = NULL for single-end.
 * taxon_out[n], flags_out[n] are mandatory; detail_out/hits_out may both be NULL (report-only, the
 * SQLClassifier path, slacken/Classifier.scala:259-410). hits_used receives the number of hit slots used. */
SLK_API int slk_classify_batch(slk_classifier* c, const slk_classify_opts* opts,
                       const uint8_t* bases1, const uint64_t* off1,
                       const uint8_t* bases2, const uint64_t* off2, uint32_t n_reads,
                       int32_t* taxon_out, uint8_t* flags_out, slk_read_detail* detail_out,
                       slk_hit* hits_out, uint64_t hits_cap, uint64_t* hits_used);
/* Same with every pointer in device memory (inputs already resident in HBM; outputs stay there).
 * hits_used_dev: one uint64 in device memory, zeroed by the call. Runs on the classifier's stream and
 * returns after the kernel has been enqueued; slk_classifier_sync waits for it. Device base buffers must be
 * readable up to the next 16-byte boundary past their end (any cudaMalloc allocation is). */
SLK_API int slk_classify_batch_dev(slk_classifier* c, const slk_classify_opts* opts,
                           const uint8_t* bases1, const uint64_t* off1,
                           const uint8_t* bases2, const uint64_t* off2, uint32_t n_reads,
                           int32_t* taxon_out, uint8_t* flags_out, slk_read_detail* detail_out,
                           slk_hit* hits_out, uint64_t hits_cap, uint64_t* hits_used_dev);
/* Packed input (the form BASELINE.json's north star names: the host hands 2-bit packed batches): every read starts a
 * new 32-base block; block b of a read holds bases 32b..32b+31 with base i in bits [2i,2i+1] of codes[b]
 * (A=0 C=1 G=2 T/U=3, BitRepresentation.scala:35-39) and bit i of mask[b] set for an ambiguous character;
 * boff[n+1] = block offsets (exclusive prefix of ceil(len/32)), len[n] = read lengths in bases. Same outputs and
 * semantics as the ASCII entry points; 72 instead of 158 bytes per 150 bp read cross PCIe. */
SLK_API int slk_classify_batch_packed(slk_classifier* c, const slk_classify_opts* opts,
                              const uint64_t* codes1, const uint32_t* mask1, const uint64_t* boff1, const uint32_t* len1,
                              const uint64_t* codes2, const uint32_t* mask2, const uint64_t* boff2, const uint32_t* len2,
                              uint32_t n_reads, int32_t* taxon_out, uint8_t* flags_out, slk_read_detail* detail_out,
                              slk_hit* hits_out, uint64_t hits_cap, uint64_t* hits_used);
/* The same for several confidence thresholds in one pass: taxon_out and flags_out are [n_thresholds][n_reads] (row t =
 * threshold t), detail_out and hits_out do not depend on the threshold. One scan and one table lookup per super-mer
 * whatever the number of thresholds (slacken/Classifier.scala:156-170). */
SLK_API int slk_classify_batch_packed_multi(slk_classifier* c, const slk_classify_multi_opts* opts,
                              const uint64_t* codes1, const uint32_t* mask1, const uint64_t* boff1, const uint32_t* len1,
                              const uint64_t* codes2, const uint32_t* mask2, const uint64_t* boff2, const uint32_t* len2,
                              uint32_t n_reads, int32_t* taxon_out, uint8_t* flags_out, slk_read_detail* detail_out,
                              slk_hit* hits_out, uint64_t hits_cap, uint64_t* hits_used);
/* Compact boundary: the fewest bytes across PCIe for the same results (the end-to-end rate of a multi-GPU box is bounded by
 * the host's memory traffic).
 *   in:  codes (as above; every read starts a new 32-base block) and len[n]; NO block offsets (a prefix sum the device does
 *        itself) and NO mask words: ambiguous characters travel as a list sorted by read, one entry per character,
 *        entry = read_index << 32 | mate << 31 | position (mate 0 / 1, position < len). Code bits at ambiguous positions
 *        are ignored. 44 bytes per 150-base read instead of 72.
 *   out: results_out[n] (16 bytes per read, threshold 0); taxon_more / flags_more [n_thresholds - 1][n_reads] for the further
 *        thresholds (NULL with one threshold); hits_out (may be NULL): the merged hits in READ ORDER, read i's hits right
 *        after read i - 1's, (hits_flags >> 2) of them each. */
typedef struct {
  int32_t taxon;        /* raw taxon id, 0 = unclassified */
  uint32_t len1;        /* lengthString parts as in slk_read_detail */
  uint32_t len2;        /* 0xFFFFFFFF for single-end */
  uint32_t hits_flags;  /* number of merged hits << 2 | SLK_READ_HAS_SPAN | SLK_READ_CLASSIFIED */
} slk_read_result;
SLK_API int slk_classify_batch_compact(slk_classifier* c, const slk_classify_multi_opts* opts,
                              const uint64_t* codes1, const uint32_t* len1, const uint64_t* codes2, const uint32_t* len2,
                              const uint64_t* ambiguous, uint64_t n_ambiguous, uint32_t n_reads, slk_read_result* results_out,
                              int32_t* taxon_more, uint8_t* flags_more, slk_hit* hits_out, uint64_t hits_cap, uint64_t* hits_used);
/* The same with 4-byte hits and 8-byte results (68 instead of 91 bytes per 150-base read across PCIe): a hit is
 * (label << 16 | k-mers), where label 0 = no record, 1..65534 = the taxon at position label - 1 of slk_index_taxa's list,
 * 0xFFFF = an ambiguous span; the word 0xFFFFFFFF is the mate-pair border (whose count is -(k - 1),
 * slacken/Classifier.scala:439-454). A merged hit of 65 535 or more k-mers does not fit: SLK_E_UNSUPPORTED, use
 * slk_classify_batch_compact for such reads. The results leave the two length fields out: they are what the hit list sums to
 * (len1 = k-mers of the hits before the border + k - 1, len2 likewise after it; slk_group.h / slacken/Classifier.scala:39-45). */
typedef struct slk_read_result_short {
  int32_t taxon;        /* raw taxon id, 0 = unclassified */
  uint32_t hits_flags;  /* number of merged hits << 2 | SLK_READ_HAS_SPAN | SLK_READ_CLASSIFIED */
} slk_read_result_short;
SLK_API int slk_classify_batch_compact_short(slk_classifier* c, const slk_classify_multi_opts* opts,
                              const uint64_t* codes1, const uint32_t* len1, const uint64_t* codes2, const uint32_t* len2,
                              const uint64_t* ambiguous, uint64_t n_ambiguous, uint32_t n_reads, slk_read_result_short* results_out,
                              int32_t* taxon_more, uint8_t* flags_more, uint32_t* hits_out, uint64_t hits_cap, uint64_t* hits_used);
SLK_API int slk_classify_packed_dev(slk_classifier* c, const slk_classify_opts* opts,
                            const uint64_t* codes1, const uint32_t* mask1, const uint64_t* boff1, const uint32_t* len1,
                            const uint64_t* codes2, const uint32_t* mask2, const uint64_t* boff2, const uint32_t* len2,
                            uint32_t n_reads, int32_t* taxon_out, uint8_t* flags_out, slk_read_detail* detail_out,
                            slk_hit* hits_out, uint64_t hits_cap, uint64_t* hits_used_dev);
/* Stage 1 on its own (2-bit encode + N mask): ASCII reads resident in HBM -> packed blocks. All pointers are device
 * pointers; boff_dev is an input (the caller knows the lengths). charToTwobit / isValid,
 * kmers/util/BitRepresentation.scala:127-143. */
SLK_API int slk_pack_reads_dev(slk_ctx* ctx, const uint8_t* bases_dev, const uint64_t* off_dev, uint32_t n_reads,
                       const uint64_t* boff_dev, uint64_t* codes_dev, uint32_t* mask_dev, uint32_t* len_dev);
SLK_API int slk_classifier_sync(slk_classifier* c);
/* the CUDA stream the classifier launches on (a cudaStream_t), for event timing by the caller */
SLK_API void* slk_classifier_stream(slk_classifier* c);
/* kernels launched by this classifier so far */
SLK_API uint64_t slk_classifier_launches(const slk_classifier* c);
/* totals over all launches so far: table probes issued (= super-mers looked up) and merged hits produced */
SLK_API int slk_classifier_stats(slk_classifier* c, uint64_t* probes, uint64_t* merged_hits);

/* CUDA events recorded on the classifier's launch stream, so a caller can time kernels on the device */
typedef struct slk_event slk_event;
SLK_API int slk_event_create(slk_ctx* ctx, slk_event** out);
SLK_API void slk_event_destroy(slk_event* e);
SLK_API int slk_event_record(slk_event* e, slk_classifier* c);
/* the same on the context's own stream (slk_pack_reads_dev, the build and the split-path kernels launch there) */
SLK_API int slk_event_record_ctx(slk_event* e, slk_ctx* ctx);
SLK_API int slk_event_elapsed_ms(slk_event* start, slk_event* end, float* ms);

/* ---- B4: report counts, replaces groupBy(sampleId, taxon).count (slacken/Classifier.scala:214-217). The host
 *          feeds the per-read results back (sample ids come from its --sample-regex), or the classifier
 *          accumulates them on the device when a counts object is attached. --------------------------------- */
SLK_API int slk_counts_create(slk_ctx* ctx, slk_tax* tax, int32_t n_samples, slk_counts** out);
SLK_API void slk_counts_destroy(slk_counts* cn);
/* attach: every following slk_classify_batch* call adds its reads (those with SLK_READ_HAS_SPAN) to sample
 * `sample` on the device, fused into the classify kernel. NULL detaches. */
SLK_API int slk_classifier_attach_counts(slk_classifier* c, slk_counts* cn, int32_t sample);
SLK_API int slk_counts_add(slk_counts* cn, const int32_t* taxon, const uint8_t* flags, const int32_t* sample_id,
                   uint32_t n);
/* dense per-taxon vector of one sample: per_taxon_out[t] for t in [0, n_taxa) (n_taxa = taxonomy size) */
SLK_API int slk_counts_fetch(slk_counts* cn, int32_t sample, int64_t* per_taxon_out, int32_t n_taxa);
/* device pointer to the [n_samples x n_taxa] int64 counter matrix (for an NCCL all-reduce across ranks) */
SLK_API void* slk_counts_device_ptr(slk_counts* cn);
SLK_API int slk_counts_reset(slk_counts* cn);

/* ---- synthetic workloads (bench / tests only; not part of the reference) -------------------------------- */
SLK_API int slk_synth_genome_dev(slk_ctx* ctx, uint64_t seed, uint64_t start, uint64_t n, uint8_t* out_dev);
SLK_API int slk_synth_reads_dev(slk_ctx* ctx, uint64_t gseed, uint64_t rseed, uint64_t n_genomes, uint64_t genome_len,
                        uint64_t first_read, uint64_t n_reads, uint32_t read_len, uint8_t* out_dev);
/* mate = 0: the same reads as slk_synth_reads_dev; mate = 1: their mates of a read pair (same genome, 250 bases further
 * along, opposite strand, errors of their own) -- the paired-end shape of BASELINE.json configs[3] */
SLK_API int slk_synth_mates_dev(slk_ctx* ctx, uint64_t gseed, uint64_t rseed, uint64_t n_genomes, uint64_t genome_len,
                        uint64_t first_read, uint64_t n_reads, uint32_t read_len, uint32_t mate, uint8_t* out_dev);
/* device-resident variant of slk_build_add (bases/frag_off/frag_taxon in device memory) */
SLK_API int slk_build_add_dev(slk_builder* b, const uint8_t* bases, const uint64_t* frag_off, const int32_t* frag_taxon,
                      uint32_t n_frag, uint64_t total_bases);
/* raw device memory helpers so a harness needs no CUDA binding of its own */
SLK_API int slk_dev_alloc(slk_ctx* ctx, size_t bytes, void** out);
SLK_API void slk_dev_free(slk_ctx* ctx, void* p);
SLK_API int slk_memcpy_h2d(slk_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
SLK_API int slk_memcpy_d2h(slk_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
SLK_API int slk_memcpy_d2d(slk_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes);
SLK_API int slk_ctx_sync(slk_ctx* ctx);
/* ---- Split path: a library sharded over several GPUs by minimizer hash range ------------------------------------------
 * Replaces the same join as B1 (slacken/Classifier.scala:84, spansToGroupedHits' left join of spans and records) when
 * the records do not fit one GPU (Spark does it with a shuffle). Every GPU holds the records
 * whose key it owns (slk_shard_of_records), classifies its own reads, and exchanges only span keys and taxa:
 *   slk_scan_spans_dev -> slk_route_spans_dev -> [all-to-all of keys] -> slk_probe_keys_dev on the owner
 *   -> [all-to-all of taxa] -> slk_resolve_spans_dev.
 * The two exchanges are the caller's (torch.distributed / NCCL in slacken_b200/sharded.py). All buffers are DEVICE
 * pointers except where a name ends in _host; every call returns when its work is done.
 *
 * A span word is (compressed minimizer << 16 | type << 14 | k-mer count), type 0 = sequence, 1 = ambiguous, 2 = mate
 * border (slacken/Supermers.scala:49-125). */
typedef struct slk_resolver slk_resolver;
/* owner (0 .. world-1) of every record of the Parquet table; host arrays. Owner d holds the d-th of `world` equal ranges
 * of the table-line hash of the minimizer. */
SLK_API int slk_shard_of_records(const slk_params* params, const int64_t* id1, uint64_t n, uint32_t world, uint8_t* shard_out);
/* the same for records that live in device memory (id1 and shard_out are device pointers) */
SLK_API int slk_shard_of_records_dev(slk_ctx* ctx, const slk_params* params, const int64_t* id1, uint64_t n, uint32_t world,
                                     uint8_t* shard_out);
/* the records of an index grouped by owner, in DEVICE memory: rows of owner d at [sum(counts_host[0..d)), +counts_host[d]) */
SLK_API int slk_index_records_by_owner_dev(slk_index* idx, uint32_t world, int64_t* id1_out, int32_t* taxon_out, uint64_t cap,
                                           uint64_t* counts_host);
/* One shard of a library range-partitioned over `world` GPUs from the records this rank owns (slk_shard_of_records): as
 * slk_index_from_records, but the table spreads the owner's RANGE of the line hash over all of its lines. An index of a
 * shard must be created with the `world` it was cut for; world = 1 is slk_index_from_records. */
SLK_API int slk_index_from_records_shard(slk_ctx* ctx, slk_tax* tax, const slk_params* params, const int64_t* id1,
                                         const int32_t* taxon, uint64_t n, uint32_t world, slk_index** out);
/* Distributed build (BASELINE configs[2]; the shuffle of groupBy(idColumns).agg(udafLca), slacken/KeyValueIndex.scala:85-93):
 *   every rank: slk_build_begin, slk_build_add* (its own genomes), slk_build_reduce (sort + LCA reduce; counts_out[d] =
 *   cells bound for owner d), slk_build_cells_dev (the cells, grouped by owner: the send buffer, owned by the builder) and
 *   slk_build_dense_taxa (the raw ids behind the 16-bit taxa inside the cells);
 *   [all-to-all of the cells, all-gather of the taxa lists: the caller's];
 *   every owner: slk_index_from_cell_runs on what it received: n_runs runs back to back in cells_dev (run r holds
 *   run_cells[r] cells and uses the run_dense[r] raw ids that follow those of run r-1 in dense_raw_host). The runs are
 *   ordered by table line, so the insert walks the table front to back; equal minimizers merge by LCA
 *   (slacken/LowestCommonAncestor.scala:152-170). */
SLK_API int slk_build_reduce(slk_builder* b, uint32_t world, uint64_t* counts_out);
SLK_API int slk_build_cells_dev(slk_builder* b, const uint64_t** cells_dev, uint64_t* n_out);
SLK_API int slk_build_dense_taxa(slk_builder* b, int32_t* raw_out, uint32_t cap, uint32_t* n_out);
SLK_API int slk_index_from_cell_runs(slk_ctx* ctx, slk_tax* tax, const slk_params* params, uint32_t world, uint32_t n_runs,
                                     const uint64_t* cells_dev, const uint64_t* run_cells, const int32_t* dense_raw_host,
                                     const uint32_t* run_dense, slk_index** out);
/* the taxa (raw ids, ancestors included) an index can answer with; out == NULL queries the count */
SLK_API int slk_index_taxa(slk_index* idx, int32_t* out, uint32_t cap, uint32_t* n_out);
/* the query side's view of the taxonomy: the union of slk_index_taxa over all shards (any order, duplicates allowed) */
SLK_API int slk_resolver_create(slk_ctx* ctx, slk_tax* tax, const slk_params* params, const int32_t* taxa, uint32_t n,
                                slk_resolver** out);
SLK_API void slk_resolver_destroy(slk_resolver* r);
/* KeyValueIndex.getSpans (slacken/KeyValueIndex.scala:163-173): ASCII fragments -> span words of fragment r at
 * spans[span_off[r] .. span_off[r+1]). spans == NULL only fills span_off and *n_spans_host (size query). */
SLK_API int slk_scan_spans_dev(slk_ctx* ctx, const slk_params* params, const uint8_t* bases1, const uint64_t* off1,
                               const uint8_t* bases2, const uint64_t* off2, uint32_t n_reads, uint64_t* span_off,
                               uint64_t* spans, uint64_t cap, uint64_t* n_spans_host);
/* the emit half alone, after a count-only call of slk_scan_spans_dev (span_off as that call left it). When the reads of
 * the batch are short enough the count-only call has already scanned them into a scratch owned by the context and this
 * call only compacts the rows (one scan instead of two); a (count, emit) pair on one slk_ctx must therefore not be
 * interleaved with another pair on the same slk_ctx from a second thread: give every scanning thread its own context. */
SLK_API int slk_emit_spans_dev(slk_ctx* ctx, const slk_params* params, const uint8_t* bases1, const uint64_t* off1,
                               const uint8_t* bases2, const uint64_t* off2, uint32_t n_reads, const uint64_t* span_off,
                               uint64_t* spans);
/* keys of the sequence spans grouped by owner (send_keys, counts_host[world]) and the span each came from (send_idx);
 * send_keys == NULL only fills counts_host */
SLK_API int slk_route_spans_dev(slk_ctx* ctx, const uint64_t* spans, uint64_t n_spans, uint32_t world, uint64_t* send_keys,
                                uint32_t* send_idx, uint64_t cap, uint64_t* counts_host);
/* the owner's half of the join: compressed keys -> raw taxon of the record, 0 = no record */
SLK_API int slk_probe_keys_dev(slk_index* idx, const uint64_t* keys, uint64_t n, int32_t* taxa);
/* spanToHit + classifyHits (slacken/KeyValueIndex.scala:176-185, slacken/Classifier.scala:124-147) from the span words
 * and the taxa that came back in send order. hits_out (optional) needs room for n_spans hits; the hits of fragment r
 * start at detail_out[r].hit_off = span_off[r]. */
SLK_API int slk_resolve_spans_dev(slk_resolver* r, const slk_classify_opts* opts, const uint64_t* spans, const uint64_t* span_off,
                                  uint64_t n_spans, uint32_t n_reads, int paired, const uint32_t* send_idx, const int32_t* taxa,
                                  uint64_t n_routed, int32_t* taxon_out, uint8_t* flags_out, slk_read_detail* detail_out,
                                  slk_hit* hits_out);

/* ---- NVLink mailbox: the two exchanges of the split path as stores into peer memory ----------------------------------------
 * The same join (slacken/Classifier.scala:84) with the all-to-alls fused into the kernels on either side of them: the
 * routing kernel stores every key straight into its owner's inbox over NVLink, the owner's lookup kernel stores every
 * taxon straight into the asker's reply area, and completion travels as one flag per (source, owner) pair that the
 * consuming kernels wait for on the device. No NCCL and no host synchronisation between the steps.
 *   every rank, per batch:  slk_scan_spans_dev -> slk_mailbox_route -> slk_mailbox_probe -> slk_mailbox_resolve
 * All three are collective (every rank calls them once per batch, with its own spans, possibly none); route and probe
 * only enqueue work. cap = room for the keys one rank sends to one owner in one batch (exceeding it fails loudly with
 * SLK_E_NOSPACE); a mailbox takes world * cap * 16 bytes of HBM. Index, resolver and mailbox must share one slk_ctx. */
typedef struct slk_mailbox slk_mailbox;
#define SLK_IPC_HANDLE_BYTES 64
/* handle_out (SLK_IPC_HANDLE_BYTES, may be NULL): what the other ranks' processes need to reach this mailbox */
SLK_API int slk_mailbox_create(slk_ctx* ctx, uint32_t rank, uint32_t world, uint64_t cap, slk_mailbox** out, uint8_t* handle_out);
/* one process per GPU: handles = the handle_out of every rank, in rank order (world * SLK_IPC_HANDLE_BYTES) */
SLK_API int slk_mailbox_connect(slk_mailbox* m, const uint8_t* handles);
/* one process driving every rank (ranks may even share a device): connects the mailboxes by pointer */
SLK_API int slk_mailbox_connect_local(slk_mailbox* const* boxes, uint32_t world);
SLK_API void slk_mailbox_destroy(slk_mailbox* m);
/* occupancy of the lookup kernels (256-thread blocks per SM, 1..8, default 8); a caller that overlaps the span scan of
 * the next batch (slk_scan_spans_dev has its own stream) sets 4 */
SLK_API int slk_mailbox_set_blocks_per_sm(slk_mailbox* m, uint32_t blocks_per_sm);
SLK_API int slk_mailbox_route(slk_mailbox* m, const uint64_t* spans, uint64_t n_spans);
SLK_API int slk_mailbox_probe(slk_mailbox* m, slk_index* idx);
SLK_API int slk_mailbox_resolve(slk_mailbox* m, slk_resolver* r, const slk_classify_opts* opts, const uint64_t* spans,
                                const uint64_t* span_off, uint64_t n_spans, uint32_t n_reads, int paired, int32_t* taxon_out,
                                uint8_t* flags_out, slk_read_detail* detail_out, slk_hit* hits_out);
/* slk_mailbox_resolve in two halves, for a caller that keeps the GPU busy across batches: _async launches the resolve of
 * batch e and returns; the caller may then issue slk_mailbox_route / slk_mailbox_probe of batch e+1 (they queue behind the
 * resolve kernels, which is all the protocol needs) and only then calls _wait, which returns when batch e's results are
 * complete and, when taxon_host / flags_host (pinned) are given, has copied n_reads of them there past the queued work. */
SLK_API int slk_mailbox_resolve_async(slk_mailbox* m, slk_resolver* r, const slk_classify_opts* opts, const uint64_t* spans,
                                      const uint64_t* span_off, uint64_t n_spans, uint32_t n_reads, int paired,
                                      int32_t* taxon_out, uint8_t* flags_out, slk_read_detail* detail_out, slk_hit* hits_out);
SLK_API int slk_mailbox_resolve_wait(slk_mailbox* m, slk_resolver* r, uint32_t n_reads, const int32_t* taxon_dev,
                                     const uint8_t* flags_dev, int32_t* taxon_host, uint8_t* flags_host);

/* ---- Bracken weights (slacken/BrackenWeights.scala:312-354) ------------------------------------------------------------------
 * All reads of length read_len of every genome fragment, self-classified against the library with the sliding window of
 * FragmentWindow (confidence 0, minHitGroups 2). The caller cuts the genomes like TaxonFragment.splitToMaxLength
 * (slacken/BrackenWeights.scala:152-164) and adds up the triples; host arrays in and out. */
typedef struct {
  int32_t dest;     /* taxon the reads classify to (0 = unclassified) */
  int32_t source;   /* taxon of the genome the reads come from */
  uint64_t reads;
} slk_bracken_triple;
SLK_API int slk_bracken_weights(slk_index* idx, const uint8_t* bases, const uint64_t* frag_off, const int32_t* frag_taxon,
                                uint32_t n_frag, uint32_t read_len, slk_bracken_triple* out, uint64_t cap, uint64_t* n_out);

/* test hook: the library's radix sort (K3a) on a host array, bits [begin_bit, end_bit), stable */
SLK_API int slk_debug_sort_u64(slk_ctx* ctx, uint64_t* keys, uint64_t n, int begin_bit, int end_bit);
/* test hook: out[i] = the scan's minimum of a[i] and b[i] (values below 2^62, compared on the FP64 pipe); host arrays */
SLK_API int slk_debug_min62(slk_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t n, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* SLACKEN_GPU_H */
