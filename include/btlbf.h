/*
 * btlbf.h -- C ABI of libbtlbf_cuda.so: the B200 (sm_100a) k-mer insert/query hot path of
 * bcgsc/btl_bloomfilter (canonical ntHash / spaced-seed ntHash fused with BloomFilter and
 * CountingBloomFilter<uint8_t> insert/contains).
 *
 * Plain pointers and sizes only; every call returns an int status (BTLBF_OK == 0) and leaves a
 * message for btlbf_last_error() on failure.  Nothing here exits the process or throws.  There is
 * NO CPU fallback: every entry point fails with BTLBF_ERR_CUDA when no sm_100-class device is usable.
 *
 * Reference interfaces replaced (paths relative to the upstream tree):
 *   - the per-k-mer loop  ntHashIterator itr(seq,h,k); while(itr!=itr.end()){ bloom.insert(*itr); ++itr; }
 *     README.md:30-43, BloomFilterUtil.h:10-17 (insertSeq), Tests/AdHoc/ParallelFilter.cpp:78-122
 *       -> btlbf_insert_seqs / btlbf_insert_seqs_dev
 *   - the query twin (bloom.contains(*itr)), README.md:46-57, BloomFilter.hpp:252-262,
 *     CountingBloomFilter.hpp:190-196            -> btlbf_contains_seqs / _dev
 *   - CountingBloomFilter<uint8_t>::minCount, CountingBloomFilter.hpp:53-64 -> btlbf_mincount_seqs
 *   - CountingBloomFilter<uint8_t>::incrementAll, :164-183                   -> btlbf_increment_all_seqs
 *   - BloomFilter::insertAndCheck, BloomFilter.hpp:200-214, and
 *     CountingBloomFilter::insertAndCheck, CountingBloomFilter.hpp:206-214   -> btlbf_insert_and_check_seqs
 *   - ntHashIterator::operator* / stHashIterator::operator*, strandArray()
 *     (vendor/ntHashIterator.hpp:93-96, vendor/stHashIterator.hpp:94-104)    -> btlbf_hash_seqs
 *   - BloomFilter::getPop (BloomFilter.hpp:316-323), CountingBloomFilter::popCount /
 *     filtered_popcount (CountingBloomFilter.hpp:216-242)                    -> btlbf_filter_popcount / _count_ge
 *   - the raw filter array m_filter (BloomFilter.hpp:436, CountingBloomFilter.hpp:102) that
 *     storeFilter/loadFilter write/read after the TOML header               -> btlbf_filter_upload / _download
 *
 * Batch convention (shared with the test oracle):
 *   bases   : all sequences concatenated without separators (any byte values; validity follows the
 *             reference's seedTab: A C G T U a c g t u and raw bytes 1 3 4 5 7 hash, all else breaks k-mers)
 *   offsets : n_seqs+1 monotone offsets, offsets[0]==0, offsets[n_seqs]==n_bases
 *   window p: the k-mer whose first base is flat position p.  Per-window outputs are indexed by p:
 *             bit p of a little-endian bit array (byte p/8, mask 1<<(p%8)) of ceil(n_bases/32)*4 bytes,
 *             or element p of a per-window array.  Invalid windows (non-hashable byte inside, crossing
 *             a sequence end, sequence shorter than k) have valid=0, hit=0, count=0, hashes=0.
 *   The reference's iterator visits exactly the valid windows of each sequence in ascending p.
 */
#ifndef BTLBF_H
#define BTLBF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BTLBF_VERSION 100

enum {
	BTLBF_OK = 0,
	BTLBF_ERR_ARG = 1,     /* invalid argument (size not multiple of 8, k == 0, bad seeds, ...) */
	BTLBF_ERR_CUDA = 2,    /* CUDA runtime / device error, or no usable device */
	BTLBF_ERR_NOMEM = 3,   /* host or device allocation failed */
	BTLBF_ERR_STATE = 4    /* handle used in a way its kind does not support */
};

enum {
	BTLBF_BLOOM = 0,       /* BloomFilter: size = number of bits (multiple of 8), BloomFilter.hpp:389-399 */
	BTLBF_COUNTING8 = 1,   /* CountingBloomFilter<uint8_t>: size = number of 8-bit counters */
	BTLBF_BITVECTOR = 2    /* the level-1 bit vector of a multi-index Bloom filter (sdsl::bit_vector of `size`
	                        * bits, any size, 64-bit words, bit pos -> word pos>>6, mask 1<<(pos&63)):
	                        * insert_seqs = MIBFConstructSupport::insertBV (MIBFConstructSupport.hpp:76-87),
	                        * insert_and_check_seqs = insertBVColli (:55-74; found bit = all h bits were set,
	                        * the collision count is the number of found bits), contains_seqs, popcount.
	                        * Same bit layout as BTLBF_BLOOM (little-endian); no file format. */
};

typedef struct btlbf_ctx btlbf_ctx;       /* one per GPU: streams, staging buffers, constants */
typedef struct btlbf_filter btlbf_filter; /* one filter resident in that GPU's HBM */

const char *btlbf_last_error(void); /* thread-local, never NULL */
int btlbf_version(void);
int btlbf_device_count(int *count);

/* ---- context ---- */
int btlbf_ctx_create(int device, btlbf_ctx **ctx);
int btlbf_ctx_destroy(btlbf_ctx *ctx);
/* Run all kernels of this context on a caller-owned cudaStream_t (e.g. torch's current stream);
 * NULL restores the context's own stream (pass cudaStreamLegacy / cudaStreamPerThread for the default streams). */
int btlbf_ctx_set_stream(btlbf_ctx *ctx, void *cuda_stream);
int btlbf_ctx_sync(btlbf_ctx *ctx);
/* DEFERRED WORK.  Two things may still be pending when an insert call has returned AND its stream work has
 * finished: (1) the partitioned BloomFilter build (filters >= 96 MB) parks the hashed k-mers of successive
 * insert calls in partition buckets and ORs them into the filter (pass 2) only when the buckets are full or
 * when the filter is next touched through this library; (2) the legacy per-k-mer calls queue updates on the
 * host.  Every btlbf_* call that reads or writes a filter applies the deferred work first, so through this
 * ABI the deferral is not observable.  It IS observable to anything that reads the filter memory directly --
 * a pointer from btlbf_filter_device_ptr kept across later inserts, your own kernel, NCCL, a peer GPU: call
 * btlbf_ctx_flush (queues the deferred work on the active stream, does not block the host) or btlbf_ctx_sync
 * BEFORE every such access, whatever the "overlap" setting.  Filters over caller-owned memory
 * (btlbf_filter_wrap) are the exception: nothing stays parked for them when an insert call returns.
 * With option "overlap" = 1 pass 2 runs on a background stream (default 0: on the active stream);
 * btlbf_ctx_flush also makes the active stream wait for that.  A deferred pass that fails to launch is
 * reported by the call that triggered it. */
int btlbf_ctx_flush(btlbf_ctx *ctx);
int btlbf_ctx_aux_stream(btlbf_ctx *ctx, void **cuda_stream); /* the background cudaStream_t (for timing) */
/* number of kernels this context has launched so far (for accounting / tests) */
int btlbf_ctx_launch_count(btlbf_ctx *ctx, uint64_t *count);
/* diagnostics: "launches", "binned_launches" (passes 1 of the partitioned paths), "two_level_passes" */
int btlbf_ctx_counter(btlbf_ctx *ctx, const char *name, uint64_t *value);
/* Tuning / debugging knobs (none of them changes a result):
 *   hashing        "force_generic" (1: byte-class hashing path for every tile)
 *   host pipeline  "chunk_bases" (windows per pipeline stage), "query_chunk_factor" (partitioned queries run in
 *                  chunks of this many chunk_bases)
 *   partitioned    "bin_mode" / "bin_query_mode" (build / query: 0 auto, 1 always, -1 never), "bin_part_log2" (bits
 *   BloomFilter    per L2-resident partition), "bin_max_parts", "bin_slack_pct", "bin_accum_bytes" (sub-bucket storage
 *   paths          one accumulation of builds may use), "bin_kernel" (1: general-shape pass-1 kernels only),
 *                  "bin_two_level" / "bin_two_level_min" (optional two-level pass 2), "overlap" (pass 2 on the
 *                  background stream)
 *   query          "query_mode" (direct kernel: 0 all probes in flight, 1 early exit), "query_adaptive",
 *                  "query_adaptive_pct", "query_adaptive_min_tiles" (device-side choice between the partitioned and the
 *                  early-exit query from a sampled hit fraction)
 *   ordered        "cbf_batch" (windows per batch), "resv_log2", "list_log2", "drain_threshold", "ordered_coop"
 *   updates        (0: host-driven residual rounds), "ungrouped_commit"
 *   multi-GPU      "peer_unroll", "peer_grid", "peer_mode" (fused merge kernel shape; peer_mode 1 / 2 are
 *                  measurement-only half merges)
 *   device         "l2_fetch_granularity" */
int btlbf_ctx_set_option(btlbf_ctx *ctx, const char *key, int64_t value);

/* ---- filters ---- */
/* size: bits (BTLBF_BLOOM, multiple of 8) or counters (BTLBF_COUNTING8); zero-initialised like
 * BloomFilter::initSize / CountingBloomFilter's ctor.  threshold is CountingBloomFilter's
 * m_countThreshold (ignored for BTLBF_BLOOM). */
int btlbf_filter_create(btlbf_ctx *ctx, int kind, uint64_t size, unsigned hash_num,
                        unsigned kmer_size, unsigned threshold, btlbf_filter **filter);
/* Same, but over caller-owned device memory (e.g. a torch uint8 tensor) of >= round_up(bytes,16)
 * bytes, 16-byte aligned; the memory is used as is (not cleared). */
int btlbf_filter_wrap(btlbf_ctx *ctx, int kind, uint64_t size, unsigned hash_num,
                      unsigned kmer_size, unsigned threshold, void *device_ptr,
                      uint64_t capacity_bytes, btlbf_filter **filter);
int btlbf_filter_destroy(btlbf_filter *f);
int btlbf_filter_clear(btlbf_filter *f);
int btlbf_filter_info(btlbf_filter *f, int *kind, uint64_t *size, uint64_t *size_bytes,
                      unsigned *hash_num, unsigned *kmer_size, unsigned *threshold);
int btlbf_filter_set_threshold(btlbf_filter *f, unsigned threshold);
/* the raw array exactly as the reference keeps it in host memory / writes it to the file body */
int btlbf_filter_upload(btlbf_filter *f, const void *host, uint64_t nbytes);
int btlbf_filter_download(btlbf_filter *f, void *host, uint64_t nbytes);
int btlbf_filter_device_ptr(btlbf_filter *f, void **device_ptr, uint64_t *nbytes);
/* BLOOM: number of set bits (getPop); COUNTING8: number of non-zero counters (popCount) */
int btlbf_filter_popcount(btlbf_filter *f, uint64_t *count);
/* COUNTING8: number of counters >= threshold (filtered_popcount with an explicit threshold) */
int btlbf_filter_count_ge(btlbf_filter *f, unsigned threshold, uint64_t *count);
/* Spaced-seed mode (stHashIterator): n_seeds strings of exactly kmer_size chars, '1' = care,
 * anything else = don't care; h2 hashes per seed; requires hash_num == n_seeds*h2.
 * n_seeds == 0 returns the filter to contiguous ntHash mode. */
int btlbf_filter_set_seeds(btlbf_filter *f, const char *const *seeds, unsigned n_seeds, unsigned h2);
/* this |= other (BLOOM) or saturating this += other (COUNTING8); src is a device pointer to a raw
 * array of the same size (a peer GPU's mapped memory is fine): the local step of the multi-GPU merge */
int btlbf_filter_merge_from_device(btlbf_filter *f, const void *src_device, uint64_t nbytes);

/* the same on raw device arrays (slices of partial filters exchanged between GPUs): dst |= src (BLOOM)
 * or dst = saturating dst + src (COUNTING8), asynchronous on the context's stream */
int btlbf_merge_device_buffers(btlbf_ctx *ctx, int kind, void *dst_device, const void *src_device,
                               uint64_t nbytes);
/* ---- fused multi-GPU merge over peer memory (one process per GPU; replaces the reference-side pattern
 * "build one filter per thread/process, then OR them": BloomFilter.hpp has no merge, callers OR the
 * arrays byte by byte) ----
 * btlbf_merge_slice: the byte range [lo, hi) of an nbytes filter that rank reduces (16-byte granules).
 * btlbf_ipc_export / _open / _close: CUDA IPC plumbing so that every process can map its peers' filter
 * arrays (handle64 = 64 opaque bytes naming the allocation that contains device_ptr, offset = position
 * of device_ptr inside it; exchange both through any host channel, e.g. torch.distributed).
 * btlbf_merge_peers: ONE kernel on this GPU that reads range `rank` of all `world` partial filters
 * (bases[p] = rank p's array as mapped in this process, bases[rank] the local one) over NVLink, reduces
 * them (OR for BLOOM, saturating add for COUNTING8) and stores the result into all of them.  The caller
 * synchronises the ranks before (all partial builds done) and after (all ranges written). */
int btlbf_merge_slice(uint64_t nbytes, int world, int rank, uint64_t *lo, uint64_t *hi);
int btlbf_ipc_export(btlbf_ctx *ctx, const void *device_ptr, void *handle64, uint64_t *offset);
int btlbf_ipc_open(btlbf_ctx *ctx, const void *handle64, void **mapped_base);
int btlbf_ipc_close(btlbf_ctx *ctx, void *mapped_base);
int btlbf_merge_peers(btlbf_ctx *ctx, int kind, void *const *bases, int world, int rank, uint64_t nbytes);
/* The same merge for BLOOM filters with the reduction done INSIDE NVSwitch (NVLS).  mc_base is a multicast
 * address mapping the same byte range of every rank's partial filter (one replica per GPU bound to one multicast
 * object -- cuMulticastCreate / cuMulticastBindMem, or torch.distributed._symmetric_memory whose rendezvous
 * handle carries multicast_ptr; the filters are then btlbf_filter_wrap'ed around that symmetric allocation).
 * ONE kernel per GPU: multimem.ld_reduce.or of byte range `rank` over all replicas, multimem.st of the result to
 * all replicas; each NVLink direction of a GPU carries ~nbytes instead of 2 (world-1)/world nbytes + the loads.
 * Same bracketing by the caller as btlbf_merge_peers.  COUNTING8 is refused (saturating add has no multimem
 * reduction): use btlbf_merge_peers.  Context options "mm_unroll" (1, 2, 4, 8) and "mm_grid" tune the kernel;
 * "wrap_accumulate" = 1 lets wrapped filters defer pass 2 of the partitioned build like owned ones (the caller
 * then calls btlbf_ctx_flush before the merge, as PeerMerge / MultimemMerge in parallel.py do). */
int btlbf_merge_multimem(btlbf_ctx *ctx, int kind, void *mc_base, int world, int rank, uint64_t nbytes);
/* Both mechanisms in ONE kernel (BLOOM; world 2, 4 or 8): the first mm_pct per cent of this rank's byte range is
 * reduced inside the switch through mc_base, the rest with peer loads / stores through bases[] (as btlbf_merge_peers).
 * The two paths load different parts of the fabric; parallel.py times the split once per box (MultimemMerge.calibrate). */
int btlbf_merge_hybrid(btlbf_ctx *ctx, int kind, void *mc_base, void *const *bases, int world, int rank,
                       uint64_t nbytes, unsigned mm_pct);
/* Merge pipelined behind pass 2 of the build.  Pass 2 of the partitioned BloomFilter build is partition-major, so the
 * partitions of a partial filter are final one after the other.  btlbf_filter_flush_parts applies the parked k-mers of
 * chunk `chunk` of `n_chunks` (chunk = 0 .. n_chunks-1, in order) on the active stream and reports the byte range of
 * the array those partitions cover; btlbf_merge_peers_range merges exactly that range (this rank's 1/world share of it)
 * on a stream of the caller's choice and does NOT apply deferred work.  The caller orders the two with an event and a
 * cross-rank barrier per chunk (parallel.py: pipelined_flush_merge), so that the merge of chunk j runs over NVLink while
 * chunk j+1 is applied.  When the parked work cannot be split, chunk 0 applies all of it and reports the whole array. */
int btlbf_filter_flush_parts(btlbf_filter *f, unsigned chunk, unsigned n_chunks, uint64_t *byte_lo, uint64_t *byte_hi);
int btlbf_merge_peers_range(btlbf_ctx *ctx, int kind, void *const *bases, int world, int rank, uint64_t lo,
                            uint64_t hi, void *cuda_stream);
/* order-dependent updates (counting insert, insert_and_check): number of k-mers that had to wait for
 * the index-ordered residual rounds, and the number of such rounds, since the filter was created */
int btlbf_filter_ordered_stats(btlbf_filter *f, uint64_t *deferred, uint64_t *rounds);

/* ---- batched sequence operations, HOST buffers (copies are inside the call) ---- */
/* All outputs may be NULL when not wanted.  n_kmers = number of valid windows processed.
 * insert_seqs on a COUNTING8 filter and insert_and_check_seqs reproduce the reference's
 * single-threaded, read-order, position-order loop exactly (these updates are order-dependent). */
int btlbf_insert_seqs(btlbf_filter *f, const char *bases, const uint64_t *offsets, uint64_t n_seqs,
                      uint64_t *n_kmers);
int btlbf_contains_seqs(btlbf_filter *f, const char *bases, const uint64_t *offsets,
                        uint64_t n_seqs, uint8_t *hit_bits, uint8_t *valid_bits, uint64_t *n_kmers,
                        uint64_t *n_hits);
/* Asynchronous forms for streaming callers: return once the work is queued, so the H2D copy of the next
 * call overlaps the kernels of this one.  All host buffers (bases, offsets, outputs, counts_out) must stay
 * valid and untouched until btlbf_ctx_sync(); pinned memory makes the copies truly asynchronous.
 * counts_out: 2 host words that receive {n_kmers, n_hits} when the call completes (insert: may be NULL, in
 * which case the call is synchronous).  Up to 4 calls may be in flight; a fifth waits for the oldest. */
int btlbf_insert_seqs_async(btlbf_filter *f, const char *bases, const uint64_t *offsets, uint64_t n_seqs,
                            uint64_t *counts_out);
int btlbf_contains_seqs_async(btlbf_filter *f, const char *bases, const uint64_t *offsets, uint64_t n_seqs,
                              uint8_t *hit_bits, uint8_t *valid_bits, uint64_t *counts_out);
/* ---- 2-bit packed input.  For callers that keep reads as 2 bits per base: base i of the flat stream is
 * (codes[i >> 2] >> (2 * (i & 3))) & 3 with A = 0, C = 1, G = 2, T/U = 3 (the order of vendor/nthash.hpp:51), and
 * bit (i & 7) of invalid[i >> 3] is set when base i is not a base (N, IUPAC codes, anything seedTab maps to 0:
 * vendor/nthash.hpp:189-228); invalid may be NULL when every base is valid.  Same results as the ASCII calls on the
 * same sequences (the five raw bytes 1 3 4 5 7 that the reference also hashes have no packed form; btlbf_pack_seqs
 * refuses them).  offsets, outputs and the window convention are those of the ASCII calls and keep counting BASES.
 * Host buffers: ceil(n_bases / 4) and ceil(n_bases / 8) bytes.  Device buffers (_dev): 16-byte aligned and padded to
 * a multiple of 16 bytes.  A quarter to three eighths of the host->device bytes of the ASCII calls. ---- */
/* Context option "host_pack" = 1 makes the ASCII host-buffer calls (btlbf_insert_seqs / btlbf_contains_seqs and their _async
 * forms) do this packing themselves, chunk by chunk on "host_pack_threads" (default 8) host threads into pinned staging
 * buffers, while the GPU works on the previous chunk: same results, a quarter of the PCIe bytes (a chunk that holds one of the
 * raw bytes 1 3 4 5 7 travels as ASCII).  Off by default: at one GPU the ASCII path is not PCIe-bound. */
int btlbf_pack_seqs(const char *bases, uint64_t n_bases, uint8_t *codes, uint8_t *invalid, int threads,
                    uint64_t *n_invalid); /* host packer (threads = 0: all cores); invalid may be NULL if the input has none */
int btlbf_insert_seqs_packed(btlbf_filter *f, const uint8_t *codes, const uint8_t *invalid, const uint64_t *offsets,
                             uint64_t n_seqs, uint64_t *n_kmers);
int btlbf_contains_seqs_packed(btlbf_filter *f, const uint8_t *codes, const uint8_t *invalid, const uint64_t *offsets,
                               uint64_t n_seqs, uint8_t *hit_bits, uint8_t *valid_bits, uint64_t *n_kmers,
                               uint64_t *n_hits);
int btlbf_insert_seqs_packed_async(btlbf_filter *f, const uint8_t *codes, const uint8_t *invalid,
                                   const uint64_t *offsets, uint64_t n_seqs, uint64_t *counts_out);
int btlbf_contains_seqs_packed_async(btlbf_filter *f, const uint8_t *codes, const uint8_t *invalid,
                                     const uint64_t *offsets, uint64_t n_seqs, uint8_t *hit_bits,
                                     uint8_t *valid_bits, uint64_t *counts_out);
int btlbf_insert_seqs_packed_dev(btlbf_filter *f, const void *d_codes, const void *d_invalid, uint64_t n_bases,
                                 const uint64_t *d_offsets, uint64_t n_seqs, uint64_t *d_stats);
int btlbf_contains_seqs_packed_dev(btlbf_filter *f, const void *d_codes, const void *d_invalid, uint64_t n_bases,
                                   const uint64_t *d_offsets, uint64_t n_seqs, uint32_t *d_hit_bits,
                                   uint32_t *d_valid_bits, uint64_t *d_stats);
int btlbf_insert_and_check_seqs(btlbf_filter *f, const char *bases, const uint64_t *offsets,
                                uint64_t n_seqs, uint8_t *found_bits, uint8_t *valid_bits,
                                uint64_t *n_kmers);
int btlbf_mincount_seqs(btlbf_filter *f, const char *bases, const uint64_t *offsets,
                        uint64_t n_seqs, uint8_t *counts, uint8_t *valid_bits, uint64_t *n_kmers);
int btlbf_increment_all_seqs(btlbf_filter *f, const char *bases, const uint64_t *offsets,
                             uint64_t n_seqs, uint64_t *n_kmers);
/* ---- FASTA / FASTQ ingest (reference: the read-a-record / insertSeq loop of
 * swig/writeBloom_rolling.cpp:19-59 and BloomFilterUtil.h:10-17) ----
 * btlbf_insert_file / btlbf_query_file: parse the file with `threads` parser threads (0 = automatic; always
 * one for the order-dependent counting insert), stream the records through the batched calls above and
 * return the number of records, of k-mers inserted / queried and (query) of k-mers found.  Multi-line
 * FASTA, four-line FASTQ, CRLF line ends; sequences of any length (long ones are cut into pieces that
 * overlap by k-1 bases, so every k-mer is visited exactly once). */
int btlbf_insert_file(btlbf_filter *f, const char *path, int threads, uint64_t *n_seqs, uint64_t *n_kmers);
int btlbf_query_file(btlbf_filter *f, const char *path, int threads, uint64_t *n_seqs, uint64_t *n_kmers,
                     uint64_t *n_hits);
/* The file calls keep their pinned staging buffers (32 MiB each, parser threads + 6 of them) for the life of
 * the process; btlbf_ingest_release frees them. */
int btlbf_ingest_release(void);
/* The parser alone (no GPU involved): region `region` of `n_regions` equal cuts of the file, aligned to
 * line (FASTA) / record (FASTQ) starts.  btlbf_seqfile_next fills one flat batch: bases[0..*n_bases),
 * offsets[0..*n_seqs] (pieces; a piece that continues a cut sequence starts with its previous `overlap`
 * bases), *n_records = records that started in the batch, *done = 1 once the region is exhausted. */
typedef struct btlbf_seqfile btlbf_seqfile;
int btlbf_seqfile_open(const char *path, unsigned overlap, int n_regions, int region, btlbf_seqfile **reader);
int btlbf_seqfile_next(btlbf_seqfile *reader, char *bases, uint64_t cap_bases, uint64_t *offsets,
                       uint64_t cap_seqs, uint64_t *n_bases, uint64_t *n_seqs, uint64_t *n_records, int *done);
int btlbf_seqfile_close(btlbf_seqfile *reader);
int btlbf_filter_ctx(btlbf_filter *f, btlbf_ctx **ctx); /* the context a filter belongs to */

/* raw iterator output: hashes[p*H + i]; strands[p*H + i] (spaced seeds only, else zero).
 * seeds == NULL / n_seeds == 0: ntHashIterator with H = hash_num;
 * else stHashIterator with H = n_seeds*h2 (hash_num is ignored). */
int btlbf_hash_seqs(btlbf_ctx *ctx, unsigned hash_num, unsigned kmer_size, const char *const *seeds,
                    unsigned n_seeds, unsigned h2, const char *bases, const uint64_t *offsets,
                    uint64_t n_seqs, uint64_t *hashes, uint8_t *strands, uint8_t *valid_bits,
                    uint64_t *n_kmers);

/* ---- legacy per-k-mer interface: the caller supplies the hash_num precomputed hash values of each of
 * n_kmers k-mers (hashes[i*hash_num + j]), exactly what the reference's insert/contains(const uint64_t[])
 * take (BloomFilter.hpp:185-262, CountingBloomFilter.hpp:53-64,134-214).  Order-dependent updates
 * (counting insert; insert with `found` requested) are applied one k-mer after the other. ---- */
/* BLOOM: insert (found == NULL) or insertAndCheck (found[i] = 1 when every bit was already set);
 * COUNTING8: incrementMin; found[i] = minCount >= threshold before the update (insertAndCheck) */
int btlbf_insert_hashes(btlbf_filter *f, const uint64_t *hashes, uint64_t n_kmers, uint8_t *found);
int btlbf_contains_hashes(btlbf_filter *f, const uint64_t *hashes, uint64_t n_kmers, uint8_t *hit);
int btlbf_mincount_hashes(btlbf_filter *f, const uint64_t *hashes, uint64_t n_kmers, uint8_t *counts);
int btlbf_increment_all_hashes(btlbf_filter *f, const uint64_t *hashes, uint64_t n_kmers);

/* ---- file layout: "[BTLBloomFilter_v1]" / "[BTLCountingBloomFilter_v1]" TOML header, "[HeaderEnd]",
 * raw array -- byte-identical to storeFilter (BloomFilter.hpp:264-314, CountingBloomFilter.hpp:331-379)
 * and readable by loadFilter (BloomFilter.hpp:107-166, CountingBloomFilter.hpp:268-329).
 * dFPR / nEntry / tEntry are the BloomFilter members of those names (ignored for COUNTING8). ---- */
int btlbf_filter_store(btlbf_filter *f, const char *path, double dFPR, uint64_t nEntry, uint64_t tEntry);
int btlbf_filter_load(btlbf_ctx *ctx, const char *path, int kind, unsigned threshold, btlbf_filter **filter,
                      double *dFPR, uint64_t *nEntry, uint64_t *tEntry);
/* the header text alone (no device needed); *len = strlen, buf may be NULL to query the length */
int btlbf_format_header(int kind, uint64_t size, uint64_t size_bytes, unsigned hash_num, unsigned kmer_size,
                        double dFPR, uint64_t nEntry, uint64_t tEntry, char *buf, size_t cap, size_t *len);
/* the inverse (no device needed): loadHeader of BloomFilter.hpp:118-166 / CountingBloomFilter.hpp:283-329 on
 * header text held in memory -- the "[magic]" line through "[HeaderEnd]\n".  *header_len = bytes consumed (where
 * the raw array starts).  Fails like the reference does: wrong magic, missing "[HeaderEnd]", missing key.
 * Outputs may be NULL; dFPR / nEntry / tEntry are only meaningful for BTLBF_BLOOM. */
int btlbf_parse_header(int kind, const char *text, size_t len, uint64_t *size, uint64_t *size_bytes,
                       unsigned *hash_num, unsigned *kmer_size, double *dFPR, uint64_t *nEntry,
                       uint64_t *tEntry, size_t *header_len);

/* ---- the same operations on DEVICE-resident batches (asynchronous on the context's stream) ---- */
/* d_bases: n_bases bytes, 16-byte aligned; d_offsets: n_seqs+1 uint64; d_hit_bits / d_valid_bits:
 * ceil(n_bases/32) uint32 words; d_stats: 2 uint64 slots that are INCREMENTED by {n_kmers, n_hits}
 * (may be NULL).  Call btlbf_ctx_sync (or synchronise the stream) before reading results. */
int btlbf_insert_seqs_dev(btlbf_filter *f, const void *d_bases, uint64_t n_bases,
                          const uint64_t *d_offsets, uint64_t n_seqs, uint64_t *d_stats);
int btlbf_contains_seqs_dev(btlbf_filter *f, const void *d_bases, uint64_t n_bases,
                            const uint64_t *d_offsets, uint64_t n_seqs, uint32_t *d_hit_bits,
                            uint32_t *d_valid_bits, uint64_t *d_stats);
int btlbf_mincount_seqs_dev(btlbf_filter *f, const void *d_bases, uint64_t n_bases,
                            const uint64_t *d_offsets, uint64_t n_seqs, uint8_t *d_counts,
                            uint32_t *d_valid_bits, uint64_t *d_stats);

/* ---- synthetic inputs generated in HBM (bench / tests; replayable by the oracle) ---- */
/* base(i) = "ACGT"[(splitmix64(seed ^ (i>>5)) >> 2*(i&31)) & 3] for i in [start, start+n) */
int btlbf_synth_genome_dev(btlbf_ctx *ctx, void *d_out, uint64_t start, uint64_t n, uint64_t seed);
/* read r (first_read <= r < first_read+n_reads) = genome[s .. s+read_len) with
 * s = g_start + splitmix64(read_seed + r) % (g_len - read_len): reads sampled from the region
 * [g_start, g_start+g_len) of the synthetic genome */
int btlbf_synth_reads_dev(btlbf_ctx *ctx, void *d_out, uint64_t first_read, uint64_t n_reads,
                          unsigned read_len, uint64_t g_start, uint64_t g_len, uint64_t genome_seed,
                          uint64_t read_seed);
/* random-sector microbenchmark that establishes the measured roofline denominator:
 * mode 0: one 4-byte load per access, mode 1: one atomicOr per access, over a device array of
 * `bytes` bytes at pseudo-random 4-byte-aligned addresses; elapsed_ms from CUDA events. */
int btlbf_random_access_probe(btlbf_ctx *ctx, void *d_array, uint64_t bytes, uint64_t n_access,
                              int mode, float *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* BTLBF_H */
