/* hlacount.h — C ABI of the B200-native scHLAcount quantification hot path.
 *
 * The reference (10XGenomics/scHLAcount, Rust) has no FFI; the seam this library drops into is the set of
 * functions src/main.rs imports from src/mapping.rs and src/em.rs (main.rs:40-47):
 *     mapping_wrapper        src/mapping.rs:310-405   (genotyping pseudoalignment -> EqClassDb)
 *     map_and_count_pseudo   src/mapping.rs:640-821   (4-way pseudoalignment + count())
 *     em_wrapper             src/em.rs:339-476        (eq_class_counts + squarem + caller)
 * and, below them, the un-vendored debruijn_mapping calls they make (Pseudoaligner::map_read, build_index,
 * utils::read_obj). Each entry point cites the reference interface it replaces. INTEGRATION.md shows the Rust
 * `extern "C"` block + build.rs a maintainer would add.
 *
 * Conventions: every function returns 0 on success and a negative hlac_status on failure; hlac_last_error()
 * returns a thread-local message. Handles are opaque. One session = one CUDA device + one stream; sessions are
 * not thread-safe. There is NO CPU fallback: every compute entry point fails with HLAC_ERR_CUDA when no device
 * is usable.
 */
#ifndef HLACOUNT_H_
#define HLACOUNT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  HLAC_OK = 0,
  HLAC_ERR_ARG = -1,
  HLAC_ERR_IO = -2,
  HLAC_ERR_FORMAT = -3,
  HLAC_ERR_CUDA = -4,
  HLAC_ERR_STATE = -5,
  HLAC_ERR_LIMIT = -6
} hlac_status;

#define HLAC_NOT_A_CELL 0xFFFFFFFFu /* batch.cell value for "CB not in the barcode list" (mapping.rs:760-763) */
#define HLAC_FLAG_PRIMARY 0x1u      /* batch.flags bit 0: !(secondary||supplementary) (mapping.rs:247-249)  */
#define HLAC_CODE_NONE 0xFFu        /* per-read debug code: read not scored                                 */
#define HLAC_K 24                   /* debruijn_mapping::config::KmerType = Kmer24 (pinned by the .idx)     */

typedef struct hlac_index hlac_index; /* a Pseudoaligner<KmerType> resident on one device (mapping.rs:318)   */
typedef struct hlac_count hlac_count; /* one map_and_count_pseudo run (mapping.rs:640-821)                   */
typedef struct hlac_geno hlac_geno;   /* one mapping_wrapper run + its EqClassDb (mapping.rs:310-405)        */
typedef struct hlac_comm hlac_comm;   /* one rank of a multi-GPU run: this rank's device + the NCCL communicator */

/* A batch of BamCrRead records (mapping.rs:64-75) in structure-of-arrays form. BAM decoding, region / CB / UB
 * handling stay on the host (mapping.rs:231-279); this is what crosses to the GPU.
 *   seq    n*words_per_read u64; read i = words [i*wpr, (i+1)*wpr). 2-bit codes A0 C1 G2 T3, base j of a read in
 *          word j/32 at bits 63-2*(j%32)..62-2*(j%32) (same layout as debruijn::DnaString storage); anything that
 *          is not ACGT is packed as A (DnaString::from_acgt_bytes). Bits beyond len must be zero.
 *   len    read length in bases (<= 32*words_per_read)
 *   cell   counting: matrix column (load_barcodes line index, main.rs:412-416) or HLAC_NOT_A_CELL;
 *          genotyping: any injective barcode id
 *   umi    any injective u32 key of the UB string (hlac_umi_key gives the packer the host driver uses)
 *   flags  HLAC_FLAG_PRIMARY
 * For hlac_*_push the pointers are HOST memory (pinned via hlac_host_alloc for full-speed copies); for
 * hlac_*_push_device they are DEVICE pointers on the session's device. */
typedef struct {
  const uint64_t* seq;
  const uint16_t* len;
  const uint32_t* cell;
  const uint32_t* umi;
  const uint8_t* flags;
  uint64_t n;
  uint32_t words_per_read;
  /* genotyping only, optional (NULL = position in the push stream): global ordinal of each read in BAM order, for a
   * session that sees only a shard of the stream (first-seen class ids are global, mapping.rs:147-154) */
  const uint64_t* ordinal;
} hlac_batch;

/* The same records for the counting path as ONE array of fixed-size records (words_per_read + 1 u64 per read): the
 * sequence words as in hlac_batch.seq, then a meta word
 *     umi << 32 | primary << 31 | len << 23 | cell        (len <= 255 bases; cell 23 bits, HLAC_PACKED_NOT_A_CELL)
 * 32 bytes per 91-bp read instead of the 35 of the five SoA arrays, and one host-to-device copy per push instead of
 * five - the counting path end to end is bound by exactly that copy. hlac_pack_records converts an hlac_batch. */
#define HLAC_PACKED_NOT_A_CELL 0x7FFFFFu
#define HLAC_PACKED_META(umi, primary, len, cell) \
  (((uint64_t)(uint32_t)(umi) << 32) | ((uint64_t)((primary) ? 1u : 0u) << 31) | ((uint64_t)((len) & 0xFFu) << 23) | \
   (uint64_t)((cell) & 0x7FFFFFu))
typedef struct {
  const uint64_t* rec; /* n * (words_per_read + 1) u64 */
  uint64_t n;
  uint32_t words_per_read;
} hlac_packed_batch;

/* The leanest host->device form for counting ("wire records"): end to end the path is bound by the H2D copy, so every byte
 * counts. A batch of equal-length reads as n records of words_per_record u32 words each, a big-endian bit stream per record
 * (bit 0 = MSB of word 0): 2 * read_len bits of bases (the device's own row layout: base j in word j / 16 at bits
 * 31-2(j%16)..30-2(j%16)), then primary (1 bit), cell (cell_bits; all ones = not in the whitelist), UMI key (umi_bits),
 * zero padding. A 91-bp read with a 10-bp UMI (21-bit key) and <= 131 070 columns is 7 words = 28 bytes (hlac_packed_batch:
 * 32, the five arrays of hlac_batch: 35). hlac_pack_wire converts an hlac_batch. */
typedef struct {
  const uint32_t* rec; /* n * words_per_record u32 */
  uint64_t n;
  uint32_t read_len, cell_bits, umi_bits;
  uint32_t words_per_record; /* hlac_wire_words(read_len, cell_bits, umi_bits) */
} hlac_wire_batch;

typedef struct { /* mapping.rs:108-116 `Metrics` */
  uint64_t num_reads, num_non_primary, num_not_cell_bc, num_cds_align, num_gen_align, num_not_aligned;
} hlac_metrics;

typedef struct { /* Vec<MatrixEntry> (mapping.rs:101-106), cells ascending then rows ascending (mapping.rs:897-906) */
  uint32_t* row;
  uint32_t* col;
  uint32_t* val;
  uint64_t n;
  uint32_t nrows;    /* matrix rows (mapping.rs:681-727) */
  uint32_t borrowed; /* 1: the arrays belong to the session (pinned), valid until its next _entries/_reset/_free */
} hlac_coo;

typedef struct {
  uint32_t n_nodes, n_tx, n_colours, lmer;
  uint64_t n_bases, n_kmers, dict_slots, bitmap_bytes;
} hlac_index_info;

/* EqClassCounts + reads_explained (em.rs:19-25, mapping.rs:169-228). Classes are CSR over tx ids.
 * `reads_*` = counts_reads (keys are the raw, possibly permuted read-level vectors), `umi_*` = counts_umi. */
typedef struct {
  uint32_t nitems;
  uint32_t nreads;
  uint64_t n_read_classes;
  uint64_t* reads_off;
  uint32_t* reads_tx;
  uint32_t* reads_cnt;
  uint64_t n_umi_classes;
  uint64_t* umi_off;
  uint32_t* umi_tx;
  uint32_t* umi_cnt;
  uint64_t* reads_explained; /* [nitems] */
} hlac_class_tables;

const char* hlac_last_error(void);
int hlac_device_count(int* n);
/* sizeof of hlac_batch, hlac_metrics, hlac_coo, hlac_index_info, hlac_class_tables, hlac_paths, hlac_calls (cap >= 7)
 * and hlac_eqclassdb, hlac_packed_batch, hlac_wire_batch (written if cap >= 8 / 9 / 10), so that a
 * binding (Rust repr(C), ctypes) can assert its own layouts against the library it loaded */
int hlac_abi_sizes(uint32_t* out, uint32_t cap);

/* Device memory of the library comes from a process-wide cache of freed blocks (no reference counterpart: the reference
 * has no device). hlac_device_cache_trim returns every cached (= currently unused) block of every device to the driver
 * and reports how many bytes that were; live sessions and indexes are not touched. HLAC_ALLOC_CACHE_MB bounds the cache
 * (default 8192, 0 = no caching). Either pointer may be null. */
int hlac_device_cache_trim(uint64_t* bytes_released, uint64_t* bytes_live);

/* Seed-scan stride of map_read's k-mer search ([EXT] debruijn_mapping find_kmer_match, `*kmer_pos += stride`). The
 * pinned revision is not vendored and the reference's golden matrices reproduce for strides 1..3 (SURVEY.md finding #4);
 * 1 (every position) is the default, 3 is what later upstream revisions use. Process-wide; sessions adopt the value
 * current when they are created. Also settable through the environment (HLAC_SEED_STRIDE=1|3). */
int hlac_set_seed_stride(int stride); /* 1 or 3 */
int hlac_get_seed_stride(void);

/* ---- index (replaces utils::read_obj(&hla_index), mapping.rs:318 / em.rs:346, and build_index, mapping.rs:652-662,
 * hla.rs:258) ------------------------------------------------------------------------------------------------ */
/* device = CUDA ordinal, or -1 for a host-only handle that supports inspection (hlac_index_get_info/_node/_tx_name)
 * but is rejected by every compute entry point. */
int hlac_index_from_idx_file(const char* path, int device, hlac_index** out); /* bincode Pseudoaligner<Kmer24> */
/* read_hla_cds (hla.rs:102-209) + build_index. allele_status == NULL: counting mode (use_filter = false,
 * mapping.rs:650,657); otherwise make_hla_index's filter (hla.rs:236-256). */
int hlac_index_from_fasta(const char* fasta_path, const char* allele_status, int device, hlac_index** out);
/* build_index::<Kmer24>(&seqs, &tx_names, _) on already-packed sequences (same 2-bit layout as hlac_batch.seq,
 * sequence i starts at word seq_word_start[i]); names are NUL-terminated. */
int hlac_index_from_sequences(const uint64_t* packed, const uint64_t* seq_word_start, const uint32_t* seq_len,
                              const char* const* names, uint32_t n_seq, int device, hlac_index** out);
int hlac_index_clone(const hlac_index* src, int device, hlac_index** out); /* the same index resident on another device */
int hlac_index_device(const hlac_index*);                                  /* CUDA ordinal, -1 = host-only handle */
int hlac_index_free(hlac_index*);
/* utils::write_obj(&index, hla_index) (hla.rs:260): the bincode Pseudoaligner<Kmer24> the reference reads with -i,
 * including its three boomphf tables (host work, no device needed) */
int hlac_index_write_idx(const hlac_index*, const char* path);
int hlac_index_get_info(const hlac_index*, hlac_index_info* out);
const char* hlac_index_tx_name(const hlac_index*, uint32_t i); /* Pseudoaligner.tx_names[i] */
/* node export (parity checks against the .idx fixture): ASCII sequence, ext byte (high nibble right), colour list */
int hlac_index_node(const hlac_index*, uint32_t node, char* seq, uint32_t seq_cap, uint32_t* len, uint32_t* exts,
                    uint32_t* colours, uint32_t col_cap, uint32_t* n_col);

/* ---- multi-GPU (the reference is single-threaded: mapping.rs:640-648 / :310-316 stay the seam) ----------
 * Reads are partitioned by barcode so that every (cell, UMI) molecule is complete on one rank; a session with a
 * communicator attached merges inside its *_finish. One rank = one device. The communicator is NCCL's, loaded at run time
 * (libnccl.so.2; a copy already in the process - torch's - is reused). Two ways to build one:
 *   one process per GPU : rank 0 calls hlac_comm_unique_id, ships the 128 bytes to the other ranks by any means
 *                         (torch.distributed / MPI / a file), every rank calls hlac_comm_init_rank;
 *   one process, N GPUs : hlac_comm_init_all (comms[i] = rank i on devices[i]); each rank's calls then come from its own
 *                         host thread (sc_hla_count --devices does this). */
#define HLAC_COMM_ID_BYTES 128
int hlac_comm_unique_id(uint8_t id[HLAC_COMM_ID_BYTES]);
int hlac_comm_init_rank(const uint8_t id[HLAC_COMM_ID_BYTES], int rank, int nranks, int device, hlac_comm** out);
int hlac_comm_init_all(const int* devices, int n, hlac_comm** comms);
int hlac_comm_free(hlac_comm*);
int hlac_comm_info(const hlac_comm*, int* rank, int* nranks, int* device, int* nccl_version);

/* ---- host-side helpers -------------------------------------------------------------------------------------- */
int hlac_host_alloc(size_t bytes, void** out); /* pinned host memory for batches */
int hlac_host_free(void* p);
/* DnaString::from_acgt_bytes (mapping.rs:252): ASCII -> 2-bit words (nwords >= ceil(len/32)), padding zeroed */
int hlac_pack_read(const char* ascii, uint32_t len, uint64_t* words, uint32_t nwords);
/* injective u32 key of a UB string: <=15 ACGT bases -> (1<<2n)|2-bit value; anything else needs interning by the
 * caller (returns HLAC_ERR_LIMIT) */
int hlac_umi_key(const char* umi, uint32_t len, uint32_t* key);

/* ---- counting: map_and_count_pseudo (mapping.rs:640-821) + count() (mapping.rs:823-909) ------------------------- */
int hlac_count_new(hlac_index* cds, hlac_index* genomic, uint32_t n_cells, int primary_only, hlac_count** out);
int hlac_count_free(hlac_count*);
uint32_t hlac_count_nrows(const hlac_count*);                  /* nrows (mapping.rs:681) */
const char* hlac_count_row_name(const hlac_count*, uint32_t i); /* rownames / labels.tsv (mapping.rs:699-725) */
/* run the session on a caller-owned CUDA stream (cudaStream_t / CUstream passed as void*; NULL = legacy default
 * stream) so that the caller can time it with its own events and order it against its own work */
int hlac_count_set_stream(hlac_count*, void* cuda_stream);
int hlac_count_reserve(hlac_count*, uint64_t total_reads);      /* optional: pre-size device buffers */
int hlac_count_push(hlac_count*, const hlac_batch* host_batch); /* the hot loop mapping.rs:740-807, async */
int hlac_count_push_device(hlac_count*, const hlac_batch* device_batch);
/* the same push for packed records (sessions with n_cells < HLAC_PACKED_NOT_A_CELL); rec is HOST memory for _push_packed,
 * DEVICE memory for _push_packed_device. Results are identical to pushing the equivalent hlac_batch. */
int hlac_count_push_packed(hlac_count*, const hlac_packed_batch* host_batch);
int hlac_count_push_packed_device(hlac_count*, const hlac_packed_batch* device_batch);
/* the same push for wire records; rec is HOST memory for _push_wire, DEVICE memory for _push_wire_device */
uint32_t hlac_wire_words(uint32_t read_len, uint32_t cell_bits, uint32_t umi_bits);
int hlac_count_push_wire(hlac_count*, const hlac_wire_batch* host_batch);
int hlac_count_push_wire_device(hlac_count*, const hlac_wire_batch* device_batch);
/* host helper: hlac_batch (host pointers; every read read_len bases, cells < 2^cell_bits - 1, UMI keys < 2^umi_bits) -> wire
 * records; rec must hold b->n * hlac_wire_words(..) u32 */
int hlac_pack_wire(const hlac_batch* b, uint32_t read_len, uint32_t cell_bits, uint32_t umi_bits, uint32_t* rec);
/* host helper: hlac_batch (host pointers, len <= 255, cell < HLAC_PACKED_NOT_A_CELL or HLAC_NOT_A_CELL) -> packed records;
 * rec must hold b->n * (b->words_per_read + 1) u64 */
int hlac_pack_records(const hlac_batch* b, uint64_t* rec);
/* reads the host driver dropped itself because their barcode is not in the whitelist (never scored, mapping.rs:760-763):
 * they only count in the metrics - n_reads into num_reads, n_non_primary into num_non_primary (mapping.rs:752-757),
 * n_not_cell into num_not_cell_bc (reads that reach mapping.rs:760: primary, or any when primary_only is off) */
int hlac_count_add_filtered(hlac_count*, uint64_t n_reads, uint64_t n_non_primary, uint64_t n_not_cell);
int hlac_count_sync(hlac_count*);                               /* host batch buffers reusable after this */
/* debug/parity: record per-read results (off by default), then fetch those of the LAST pushed batch:
 * code = gene*3+slot or HLAC_CODE_NONE */
int hlac_count_record_codes(hlac_count*, int on);
int hlac_count_last_codes(hlac_count*, uint8_t* codes, uint64_t n);
/* seed-filter counters since the last reset: l-mer bitmap tests and exact dictionary probes */
int hlac_count_probe_stats(hlac_count*, uint64_t* bitmap_tests, uint64_t* dict_probes);
/* count(): UMI consensus into the dense u32 [n_cells][nrows] matrix on the device; returns its device pointer so a
 * caller sharding reads by barcode can all-reduce it itself before extracting entries (hlac_count_attach_comm +
 * hlac_count_finish do that inside the library). */
int hlac_count_reduce(hlac_count*, uint32_t** dense_device, uint64_t* n_elems);
int hlac_count_entries(hlac_count*, hlac_coo* out, hlac_metrics* metrics); /* after reduce (and optional all-reduce) */
/* reduce + entries fused: grouping, consensus, COO extraction and the counters are enqueued back to back, the COO
 * triplets and counters land in host-mapped pinned memory and the host synchronises once. With a communicator attached
 * the dense matrices and the metrics of all ranks are summed in between (collective: every rank must call it). */
int hlac_count_finish(hlac_count*, hlac_coo* out, hlac_metrics* metrics);
/* multi-GPU counting: this session holds the reads of the barcodes assigned to its rank, with GLOBAL column indices
 * (n_cells = all columns). root < 0: all-reduce, every rank returns the whole matrix; root >= 0: reduce, only that rank
 * returns entries (the others get n = 0); metrics are global on every rank. comm == NULL detaches. */
int hlac_count_attach_comm(hlac_count*, hlac_comm* comm, int root);
/* scored reads grouped by the last finish / reduce, hash buckets in use, finishes that had to take the sort path */
int hlac_count_group_stats(hlac_count*, uint64_t* n_keys, uint32_t* n_buckets, uint64_t* fallbacks);
/* the launch shape the map kernel last ran with, as text (diagnostics; HLAC_COUNT_SHAPE="smem,threads,slots,idx" forces one,
 * HLAC_COUNT_TMA=1 enables the bulk-copy record prefetch) */
int hlac_count_shape(hlac_count*, char* out, size_t cap);
int hlac_count_reset(hlac_count*);                              /* drop accumulated reads, keep buffers */
int hlac_coo_free(hlac_coo*);
/* kernel timing of the session since the last reset, measured with CUDA events on the session stream:
 * ms[0] = map kernel(s), ms[1] = grouping + consensus, ms[2] = COO extraction, ms[3] = H2D copies. launches = kernels launched. */
int hlac_count_timing(hlac_count*, float ms[4], uint64_t* launches);

/* ---- genotyping: mapping_wrapper (mapping.rs:310-405) + em_wrapper's numeric half (em.rs:358-434) ---------------- */
int hlac_geno_new(hlac_index* idx, hlac_geno** out);
int hlac_geno_free(hlac_geno*);
/* forward_only = the --unmapped pass (mapping.rs:380-393). Batches must be pushed in BAM order: class ids are
 * first-seen ids (mapping.rs:147-154). */
int hlac_geno_push(hlac_geno*, const hlac_batch* host_batch, int forward_only);
int hlac_geno_push_device(hlac_geno*, const hlac_batch* device_batch, int forward_only);
/* debug/parity for the LAST pushed batch: coverage (0 = None) */
int hlac_geno_last_cov(hlac_geno*, uint32_t* cov, uint64_t n);
/* keep the path id of every read pushed (4 B/read) so that hlac_geno_read_class_ids works; call before the first push */
int hlac_geno_record_reads(hlac_geno*, int on);
/* counters: reads with Some(..) ("some_aln", mapping.rs:342), with cov > MIN_SCORE_CALL ("long_aln", :344), distinct
 * colour paths interned, seed-filter bitmap tests and dictionary probes */
int hlac_geno_stats(hlac_geno*, uint64_t* some_aln, uint64_t* long_aln, uint64_t* n_paths, uint64_t* bitmap_tests,
                    uint64_t* dict_probes);
/* EqClassDb::eq_class_counts (mapping.rs:169-228). Caller frees with hlac_class_tables_free. After this,
 * hlac_geno_read_class_ids returns, for every read pushed so far, its first-seen eq_class_id or 0xFFFFFFFF. */
int hlac_geno_finish(hlac_geno*, hlac_class_tables* out);
int hlac_geno_read_class_ids(hlac_geno*, uint32_t* ids, uint64_t n);
/* hlac_geno_finish in two halves, for multi-GPU runs: _begin resolves the classes and leaves the per-class UMI
 * histogram (u32[n_classes], classes in first-seen order) on the device so that ranks can all-reduce it; _end reads
 * it back and assembles the tables. _begin returns with the histogram complete; the caller's all-reduce must have
 * completed (stream-synchronised) before _end is called: the session reads it on its own stream. */
int hlac_geno_finish_begin(hlac_geno*, uint32_t** umi_hist_device, uint64_t* n_classes);
int hlac_geno_finish_end(hlac_geno*, hlac_class_tables* out);
/* the distinct colour paths a session has interned: hash pair, first-seen ordinal, read count, colour ids (CSR).
 * Multi-GPU (reads partitioned by barcode): every rank exports, all ranks gather the parts, build the
 * identical union with hlac_paths_merge and adopt it with hlac_geno_import_paths before hlac_geno_finish_begin. */
typedef struct {
  uint64_t n_paths;
  uint64_t* hash_a;
  uint64_t* hash_b;
  uint64_t* first_ord;
  uint32_t* count;
  uint64_t* off; /* n_paths + 1 */
  uint32_t* colours;
} hlac_paths;
/* Multi-GPU genotyping inside the library (collective: every rank calls hlac_geno_finish*): reads are partitioned by
 * barcode and pushed with their global ordinals (hlac_batch.ordinal); with a communicator attached hlac_geno_finish_begin
 * all-gathers the ranks' distinct colour paths, builds their union on the device (hash table keyed by the path hash;
 * first-seen ordinals by min, read counts by sum, colour paths compared on every duplicate), resolves only the union paths
 * this rank owns (hash mod n_ranks) and all-gathers the class fingerprints, and all-reduces the per-class UMI histogram, so
 * every rank returns the tables of the whole run. hlac_geno_export_paths / hlac_paths_merge / hlac_geno_import_paths below
 * are the same merge done by the caller through host memory (no communicator needed). */
int hlac_geno_attach_comm(hlac_geno*, hlac_comm* comm);
int hlac_geno_export_paths(hlac_geno*, hlac_paths* out);
int hlac_paths_merge(const hlac_paths* parts, uint32_t n_parts, hlac_paths* out);
int hlac_geno_import_paths(hlac_geno*, const hlac_paths* merged);
int hlac_paths_free(hlac_paths*);
int hlac_class_tables_free(hlac_class_tables*);
/* lean mode: hlac_geno_finish leaves counts_reads on the device (reads_off / reads_tx / reads_cnt are NULL,
 * n_read_classes is still reported); counts_umi, reads_explained and nreads are produced as usual. The reference's
 * EqClassCounts.counts_reads is only consumed by pair_reads_explained (em.rs:36-44), which hlac_geno_call evaluates
 * on the device-resident class vectors. */
int hlac_geno_set_lean(hlac_geno*, int on);
int hlac_geno_timing(hlac_geno*, float ms[4], uint64_t* launches);
/* the last path resolution (inside hlac_geno_finish*): kernel ms, distinct paths, no-truncate merges emulated, vector
 * elements walked, compulsory bytes (per merge one colour bitset, per path its colour ids in + its class vector out) */
int hlac_geno_resolve_stats(hlac_geno*, float* ms, uint64_t* paths, uint64_t* merges, uint64_t* element_steps, uint64_t* bytes);

/* squarem(&eq_class_counts) (em.rs:232-311) on the device, f64. */
int hlac_squarem(const hlac_class_tables* t, int device, double* theta_out, uint32_t* iters_out);
/* the caller (em.rs:361-434): called tx ids ascending; weights.tsv / pairs.tsv rows in file order. */
typedef struct {
  uint32_t n_called;
  uint32_t* called;
  uint32_t n_weights;
  uint32_t* w_tx;
  double* w_em;
  double* w_frac;
  uint32_t n_pairs;
  uint32_t* p_a;
  uint32_t* p_b;
  double* p_frac;
} hlac_calls;
int hlac_call_genotypes(const hlac_index* idx, const hlac_class_tables* t, const double* theta, hlac_calls* out);
int hlac_calls_free(hlac_calls*);
/* the same caller on a finished session (works with lean tables): pair overlaps from the class vectors on the device */
int hlac_geno_call(hlac_geno*, const hlac_class_tables* t, const double* theta, hlac_calls* out);

/* ---- the reference's counts file (replaces utils::write_obj(&eq_counts, ..) mapping.rs:403 and
 * utils::read_obj(&hla_counts) em.rs:347): bincode 1.x of EqClassDb (mapping.rs:84-90). Strings by id as CSR. ---- */
typedef struct {
  uint32_t n_barcodes;
  uint64_t* barcode_off; /* n_barcodes + 1 */
  uint8_t* barcode_bytes;
  uint32_t n_umis;
  uint64_t* umi_off;
  uint8_t* umi_bytes;
  uint32_t n_classes;  /* eq_classes: class vector by eq_class_id (first-seen order), as map_read returned it */
  uint64_t* class_off; /* n_classes + 1 */
  uint32_t* class_tx;
  uint64_t n_counts;   /* Vec<BusCount>: one triple per counted read, in read order */
  uint32_t* bc;
  uint32_t* umi;
  uint32_t* cls;
} hlac_eqclassdb;
int hlac_eqclassdb_read(const char* path, hlac_eqclassdb* out);  /* host only; caller frees with hlac_eqclassdb_free */
int hlac_eqclassdb_write(const char* path, const hlac_eqclassdb* db); /* host only; HashMaps emitted in id order */
int hlac_eqclassdb_free(hlac_eqclassdb*);
/* EqClassDb::eq_class_counts(nitems) (mapping.rs:169-228) on the device for a loaded triple list */
int hlac_eqclassdb_counts(const hlac_eqclassdb* db, uint32_t nitems, int device, hlac_class_tables* out);
/* the EqClassDb of a finished genotyping session (hlac_geno_record_reads must have been on, not lean): the caller
 * supplies what the session never sees - for every read pushed, in push order, an index into its own barcode and UMI
 * string tables (CSR) - and gets ids renumbered in first-counted order like EqClassDb::count (mapping.rs:128-163). */
int hlac_geno_eqclassdb(hlac_geno*, const hlac_class_tables* t, const uint32_t* read_bc, const uint32_t* read_umi,
                        uint64_t n_reads, const uint64_t* bc_off, const uint8_t* bc_bytes, uint32_t n_bc,
                        const uint64_t* umi_off, const uint8_t* umi_bytes, uint32_t n_umi, hlac_eqclassdb* out);

/* the same from explicit per-read class ids (first-seen eq_class ids as hlac_geno_read_class_ids returns them, 0xFFFFFFFF =
 * not counted), e.g. assembled by a driver that spread the reads over several sessions */
int hlac_eqclassdb_from_ids(const hlac_class_tables* t, const uint32_t* class_ids, const uint32_t* read_bc, const uint32_t* read_umi,
                            uint64_t n_reads, const uint64_t* bc_off, const uint8_t* bc_bytes, uint32_t n_bc,
                            const uint64_t* umi_off, const uint8_t* umi_bytes, uint32_t n_umi, hlac_eqclassdb* out);

/* ---- the three reference entry points, whole (host driver in C++; BAM/FASTA I/O on the host) ---------------------- */
/* mapping_wrapper(hla_index, outdir, bam, locus, use_unmapped) -> counts file path (mapping.rs:310-316).
 * region: samtools-style "chr:start-end" or NULL. The counts file is the reference's: bincode of EqClassDb
 * (mapping.rs:84-90,403), readable by the reference's em_wrapper; HLAC_COUNTS_FORMAT=tables writes this library's compact
 * table format (HLACCNT1) instead. hlac_em_wrapper reads both. */
int hlac_mapping_wrapper(const char* hla_index, const char* outdir, const char* bam, const char* region,
                         int use_unmapped, int device, char* counts_path_out, size_t cap);
/* em_wrapper(hla_index, hla_counts, cds_db, gen_db, outdir) -> (genomic, cds) paths (em.rs:339-345,475) */
int hlac_em_wrapper(const char* hla_index, const char* hla_counts, const char* cds_db, const char* gen_db,
                    const char* outdir, int device, char* gen_path_out, char* cds_path_out, size_t cap);
/* map_and_count_pseudo(bam, out_dir, barcodes, locus, cds, genomic, primary_only) (mapping.rs:640-648);
 * barcodes_path is the --cell-barcodes file (load_barcodes, main.rs:406-424). Writes labels.tsv into out_dir. */
int hlac_map_and_count_pseudo(const char* bam, const char* out_dir, const char* barcodes_path, const char* region,
                              const char* cds, const char* genomic, int primary_only, int device, hlac_coo* entries,
                              hlac_metrics* metrics, uint32_t* n_cells_out);
/* the same run spread over several GPUs of one box (sc_hla_count --devices 0,1,..): one counting session per device,
 * reads routed by barcode column (column % n_devices), the per-device matrices summed onto devices[0] through a
 * library-owned NCCL communicator inside hlac_count_finish. Results are identical to the single-device call. */
int hlac_map_and_count_pseudo_multi(const char* bam, const char* out_dir, const char* barcodes_path, const char* region,
                                    const char* cds, const char* genomic, int primary_only, const int* devices, int n_devices,
                                    hlac_coo* entries, hlac_metrics* metrics, uint32_t* n_cells_out);
/* mapping_wrapper on several GPUs of one box (sc_hla_count --devices): one genotyping session per device, reads routed by
 * barcode with their stream ordinals, hlac_geno_attach_comm merges inside the finish; the counts file is the same. */
int hlac_mapping_wrapper_multi(const char* hla_index, const char* outdir, const char* bam, const char* region, int use_unmapped,
                               const int* devices, int n_devices, char* counts_path_out, size_t cap);
/* host BAM reader self-check (BamSeqReader, mapping.rs:231-279): count of CB+UB records in the selection and an
 * FNV-1a checksum over (2-bit sequence, CB, UB, primary) in file order; region NULL = the unmapped tail ("*") */
int hlac_bam_scan(const char* bam, const char* region, uint64_t* n_records, uint64_t* checksum);
/* the batch path of the wrappers (records framed sequentially, parsed and packed on a pool of host threads, appended in
 * file order): record count and FNV-1a over (2-bit words carrying bases, len, first-seen barcode id, UMI key, flags) */
int hlac_bam_batch_checksum(const char* bam, const char* region, uint64_t* n_records, uint64_t* checksum);
/* sc_hla_count's _main (main.rs:151-298): argv as the CLI takes it. */
int hlac_main(int argc, const char* const* argv);

#ifdef __cplusplus
}
#endif
#endif /* HLACOUNT_H_ */
