/* autocycler_gpu.h — C ABI of libautocycler_gpu.so, the B200 implementation of Autocycler's
 * `compress` hot path.
 *
 * The reference (rrwick/Autocycler v0.6.1, Rust) has no FFI for this path; the path sits behind
 * in-crate calls.  Each entry point below names the reference interface it replaces (file:line
 * under the reference's src/), so that a Rust `extern "C"` shim can bind them in
 * place of those calls:
 *
 *   KmerGraph::new + add_sequences          kmer_graph.rs:79-90      -> ac_create, ac_add_sequence, ac_upload
 *   UnitigGraph::from_kmer_graph            unitig_graph.rs:36-48    -> ac_build
 *   (the UnitigGraph / Unitig fields)       unitig_graph.rs:28-33, unitig.rs:30-45 -> ac_counts_get, ac_unitigs_copy
 *   simplify_structure                      graph_simplification.rs:26-40 -> ac_simplify
 *   merge_linear_paths (downstream, 8f)     graph_simplification.rs:315-371 -> ac_merge_linear_paths
 *   UnitigGraph::save_gfa                   unitig_graph.rs:317-331  -> ac_gfa_size, ac_gfa_copy
 *   compress (the whole subcommand)         compress.rs:32-50        -> ac_compress_dir
 *
 * Conventions: every function returns 0 on success and a negative AC_E* code on failure; the message
 * is available from ac_last_error(handle) (or ac_last_error(NULL) when no handle exists).  No C++
 * exception crosses the boundary.  The caller owns every buffer it passes; inputs are copied during
 * the call; outputs are written into caller-allocated buffers sized from ac_counts_get / ac_gfa_size.
 * A handle must be used from one host thread at a time (the reference path is single-threaded and
 * !Send).  There is no CPU fallback: without a CUDA device ac_create fails with AC_ENODEVICE.
 */
#ifndef AUTOCYCLER_GPU_H
#define AUTOCYCLER_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AC_OK 0
#define AC_EINVAL (-1)      /* bad argument (even k, k out of range, non-ACGT. byte, call order) */
#define AC_ENODEVICE (-2)   /* no usable CUDA device */
#define AC_ECUDA (-3)       /* CUDA runtime error */
#define AC_ERANGE (-4)      /* buffer too small / input too large */
#define AC_EIO (-5)         /* file system error (ac_compress_dir) */
#define AC_EINPUT (-6)      /* the reference's own input errors (misc.rs:130-136 quit_with_error) */

typedef struct ac_handle ac_handle;

typedef struct {
    uint32_t k;             /* odd; 3..511 (compress.rs:56-58 restricts the CLI to 11..501) */
    int32_t device;         /* CUDA device ordinal */
    void* stream;           /* cudaStream_t to run on, or NULL for a private (non-blocking) stream; the default stream is named by its handle (cudaStreamLegacy / cudaStreamPerThread) */
    uint32_t keep_positions;/* non-zero: ac_unitigs_copy can return full forward/reverse position lists */
    int32_t n_devices;      /* > 1: this ONE process drives several GPUs: the assemblies are sharded by file over */
    const int32_t* devices; /* devices[0..n_devices) (devices[0] finishes the graph; `device` and `stream` are ignored); 0/1: one GPU, `device` */
} ac_config;

typedef struct {
    uint64_t n_kmers;       /* both strands == KmerGraph.kmers.len() printed at compress.rs:152 */
    uint64_t n_unitigs;
    uint64_t n_links;       /* UnitigGraph::link_count().1 (unitig_graph.rs:478-507) */
    uint64_t total_length;  /* UnitigGraph::total_length (unitig_graph.rs:474-476) */
    uint64_t seq_bytes;     /* sum of unitig sequence lengths for ac_unitigs_copy (trimmed) */
    uint64_t n_fwd_pos, n_rev_pos;   /* total entries of forward_positions / reverse_positions */
    uint64_t n_next;        /* total entries of forward_next + reverse_next */
    uint64_t n_sequences;
    uint64_t n_path_steps;  /* sum of path lengths over all sequences */
    uint64_t length_before_simplify;   /* total_length() of the graph as built, before simplify_structure moved bases (compress.rs:164) */
} ac_counts;

/* Unitigs in the graph's current order (after ac_build: the order renumber_unitigs gives, numbers 1..U).
 * All arrays are caller-allocated from ac_counts; offsets arrays have n_unitigs+1 entries. */
typedef struct {
    uint32_t* number;       /* [U] Unitig.number */
    uint64_t* seq_off;      /* [U+1] */
    uint8_t* seq;           /* [seq_bytes] forward_seq, ASCII */
    double* depth;          /* [U] Unitig.depth */
    uint64_t* fpos_off;     /* [U+1]; positions need ac_config.keep_positions, else pass NULL */
    uint32_t* fpos;         /* [n_fwd_pos] Position.pos (position.rs:20) */
    uint16_t* fpos_id_strand; /* [n_fwd_pos] Position.seq_id_and_strand (position.rs:21; bit 15 = forward strand) */
    uint64_t* rpos_off; uint32_t* rpos; uint16_t* rpos_id_strand;
    uint64_t* next_off;     /* [2U+1] forward_next of unitig i at 2i, reverse_next at 2i+1 */
    int32_t* next;          /* [n_next] signed unitig numbers, negative = reverse strand (UnitigStrand::signed_number) */
} ac_unitigs;

typedef struct {            /* milliseconds */
    float h2d, pack, insert, adjacency, boundaries, runs, unitigs, links, seed_sort, emit, d2h, device_total;
    float host_graph, host_simplify, host_gfa;
    float sample, device_simplify, device_gfa;   /* sizing pass of the k-mer table; expand_repeats passes + renumbering and the GFA text when they run on the device */
    float insert_kernel, reserved0;              /* the hash-insert kernel alone (CUDA events right around its launch; `insert` also holds the table initialisation and the counter read-back) */
    uint64_t insert_occurrences;   /* k-mer occurrences hashed by the insert kernel (forward windows; each feeds both strands) */
    uint64_t table_capacity, table_used;
    uint64_t kernel_launches;      /* cumulative launches of this library's kernels in the process */
    uint64_t h2d_bytes, d2h_bytes; /* bytes moved by ac_upload / ac_build */
} ac_timings;

const char* ac_last_error(const ac_handle* h);
const char* ac_version(void);

int ac_create(ac_handle** out, const ac_config* cfg);
void ac_destroy(ac_handle* h);

/* One padded, end-repaired forward strand (Sequence.forward_seq after compress.rs:125, bytes in "ACGT.",
 * k/2 dots or repaired bases at both ends) with the fields save_gfa prints (unitig_graph.rs:352-360). */
int ac_add_sequence(ac_handle* h, uint16_t seq_id, const uint8_t* fwd_padded, uint64_t padded_len,
                    const char* filename, const char* contig_header);
int ac_clear_sequences(ac_handle* h);
int ac_upload(ac_handle* h);            /* host -> HBM copy of the added sequences */
/* Multi-GPU: copy only the strands of sequences [seq_lo, seq_hi) over this process's PCIe link; the other ranks' blocks are brought
 * over NVLink by the caller's collective, each into the device range ac_strand_block names for it (blocks tile the strand buffer). */
int ac_upload_shard(ac_handle* h, uint32_t seq_lo, uint32_t seq_hi);
int ac_strand_block(ac_handle* h, uint32_t seq_lo, uint32_t seq_hi, void** dev_ptr, uint64_t* n_bytes);
int ac_build(ac_handle* h);             /* k-mer table, unitigs, links, renumber: the graph after from_kmer_graph */
int ac_simplify(ac_handle* h);          /* simplify_structure */
/* compress.rs:42-47 in one call — build_kmer_graph, build_unitig_graph, simplify_unitig_graph and the bytes save_gfa writes — as ONE
 * device pipeline: repeat expansion, the closing renumbering and the H/S/L/P text are produced by kernels and only the text and the
 * counts compress prints travel back.  Afterwards ac_gfa_* return the file, ac_counts_get the counts (length_before_simplify = the
 * graph as built); the graph arrays stay in HBM and are fetched the first time a call needs them (ac_unitigs_copy, ac_path_copy,
 * ac_merge_linear_paths, ...).  `autocycler compress`, ac_compress_dir and bench.py use this call. */
int ac_compress(ac_handle* h);
/* merge_linear_paths (graph_simplification.rs:315-371), what cluster/trim/resolve/clean first do to a loaded compress
 * graph (cluster.rs:804): chains of exclusively linked unitigs become one unitig numbered max+1, max+2, ...; merged
 * unitigs follow the surviving ones in the S lines.  use_paths != 0 keeps sequence-path ends fixed (the reference's
 * `seqs` argument); 0 is its `&vec![]` form: everything mergeable is merged and the P lines lose their paths. */
int ac_merge_linear_paths(ac_handle* h, int use_paths);
/* renumber_unitigs (unitig_graph.rs:295-315) on the current graph, e.g. after ac_merge_linear_paths as trim.rs:266-268 does. */
int ac_renumber_unitigs(ac_handle* h);

/* Multi-GPU form of ac_build: one process per GPU, every process adds and uploads ALL sequences, owns the
 * contiguous block [seq_lo, seq_hi) of them (index = order of ac_add_sequence), and the caller (e.g. torch.distributed over
 * NCCL) moves the exported records between the processes.  Buffers are device memory on the handle's device.
 *   ac_build_local -> ac_entries_count/export -> [all-gather] -> ac_entries_merge (every other rank's records)
 *   -> ac_runs_local -> ac_runs_export -> [gather to rank 0] -> rank 0: ac_runs_import (all ranks, rank order) -> ac_build_finish
 * Records are opaque: 16 bytes per k-mer entry, 16 bytes per unitig occurrence.  ac_runs_import_padded reads the buffer a padded
 * gather leaves behind (rank r's counts[r] records start at record r * stride_records); ac_compress_finish is ac_compress for the
 * importing rank (simplify_structure and the GFA text on the device as well). */
int ac_build_local(ac_handle* h, uint32_t seq_lo, uint32_t seq_hi, uint32_t multi);
int ac_entries_count(ac_handle* h, uint64_t* n);
int ac_entries_export(ac_handle* h, void* dst, uint64_t cap_records);
int ac_entries_merge(ac_handle* h, const void* src, uint64_t n);
int ac_runs_local(ac_handle* h, uint64_t* n_runs);
int ac_runs_export(ac_handle* h, void* dst, uint64_t cap_records);
int ac_runs_import(ac_handle* h, const void* src, uint64_t n);
int ac_runs_import_padded(ac_handle* h, const void* src, uint64_t stride_records, const uint64_t* counts, uint32_t n_ranks);
int ac_build_finish(ac_handle* h);      /* on the rank that imported the runs: the graph after from_kmer_graph */
int ac_compress_finish(ac_handle* h);
/* The same with the P lines printed where the sequences live (the GFA of many assemblies is mostly P lines; one rank printing and
 * copying out all of them is what bounds the scaling):
 *   rank 0 (it must own the first block of sequences): ac_compress_finish_split — ac_gfa_data then ends after the L lines —
 *   -> ac_path_tokens_export: one uint32 per occurrence, "(final unitig number - 1) << 1 | strand", rank r's counts[r] tokens at
 *   dst + r * stride_tokens (the counts given to ac_runs_import_padded) -> [scatter] -> every rank: ac_path_lines_render with the tokens
 *   of its own occurrences -> ac_path_lines_data: the P lines of its own sequences (unitig_graph.rs:352-360), pinned host memory.
 * input_assemblies.gfa = rank 0's text followed by the ranks' path lines in rank order. */
int ac_compress_finish_split(ac_handle* h);
int ac_path_tokens_export(ac_handle* h, void* dst, uint64_t stride_tokens, const uint64_t* counts, uint32_t n_ranks);
int ac_path_lines_render(ac_handle* h, const void* tokens, uint64_t n_tokens);
int ac_path_lines_data(ac_handle* h, const char** data, uint64_t* n_bytes);   /* borrowed: valid until the next call on h */
int ac_counts_get(const ac_handle* h, ac_counts* out);
int ac_unitigs_copy(const ac_handle* h, ac_unitigs* out);
int ac_path_copy(const ac_handle* h, uint64_t seq_index, int32_t* out, uint64_t cap, uint64_t* n);  /* get_unitig_path_for_sequence_i32 */
int ac_gfa_size(ac_handle* h, uint64_t* n_bytes);
int ac_gfa_copy(ac_handle* h, char* buf, uint64_t cap);
int ac_gfa_data(ac_handle* h, const char** data, uint64_t* n_bytes);   /* the same bytes, borrowed: valid until the next call on h */
int ac_timings_get(const ac_handle* h, ac_timings* out);

/* `autocycler compress -i assemblies_dir -a autocycler_dir --kmer k --max_contigs m -t threads`
 * (main.rs:126-147, compress.rs:32-50): writes input_assemblies.gfa and input_assemblies.yaml.
 * ac_compress_dir_devices is the same on several GPUs of one box (`autocycler compress --devices 0,1,...`): the assemblies are
 * sharded by file, one contiguous block of files per device, the k-mer buckets are read across NVLink by the peers' merge kernels. */
int ac_compress_dir(const char* assemblies_dir, const char* autocycler_dir, uint32_t k, uint32_t max_contigs,
                    uint32_t threads, int32_t device, int32_t verbose);
int ac_compress_dir_devices(const char* assemblies_dir, const char* autocycler_dir, uint32_t k, uint32_t max_contigs,
                            uint32_t threads, const int32_t* devices, int32_t n_devices, int32_t verbose);

/* Host-side stage A of compress (compress.rs:98-133): directory scan, FASTA load, padding, end repair.
 * Fills a handle created with the same k, ready for ac_upload.  Returns the number of assemblies. */
int ac_load_sequences(ac_handle* h, const char* assemblies_dir, uint32_t max_contigs, uint32_t threads,
                      uint64_t* assembly_count);
/* Read back one loaded sequence (for tests of stage A): padded forward strand and header fields. */
int ac_sequence_get(const ac_handle* h, uint64_t index, uint16_t* seq_id, uint64_t* length,
                    char* fwd_padded, uint64_t cap_fwd, char* filename, uint64_t cap_fn, char* header, uint64_t cap_hd);

/* UnitigGraph::from_gfa_lines (unitig_graph.rs:55-174): replaces whatever the handle holds by the graph and the sequences of an
 * Autocycler GFA (H/S/L/P lines; `length` bytes of text), so that the calls above that work on a built graph — ac_gfa_*,
 * ac_sequence_reconstruct (decompress.rs:83-105), ac_merge_linear_paths, ac_simplify, ac_renumber_unitigs, ac_counts_get,
 * ac_unitigs_copy, ac_path_copy — apply to a file written earlier.  Host only.  Any DP:f: value the reference parses (f64) and the
 * CL:Z: segment colours (UnitigType, unitig.rs:72-86) are carried and written back; lines may end in "\r\n" (misc.rs:51-61).
 * Errors: the reference's own (missing tags, unknown unitigs, non-zero overlaps ...). */
int ac_load_gfa(ac_handle* h, const char* gfa_text, uint64_t length);

/* Deployment helper, not part of the reference: restricts the calling thread (and the threads created after it) to the CPUs of the
 * NUMA node the CUDA device is attached to, so that pinned host buffers are local to both.  Returns the node (>= 0) or a negative
 * AC_E* with the affinity unchanged.  Call it before ac_create; `autocycler compress` and bench.py do. */
int ac_bind_host_to_device(int32_t device);

/* pairwise_contig_distances (cluster.rs:132-151), the all-against-all step of `autocycler cluster`: out[a * S + b] for the S sequences of
 * the handle (built or loaded graph); the unitig-set intersections are computed on the GPU.  ac_distance_matrix_text renders
 * save_distance_matrix's file (cluster.rs:160-176); `out` may be NULL to query the length. */
int ac_pairwise_distances(ac_handle* h, double* out, uint64_t cap);
int ac_distance_matrix_text(ac_handle* h, char* out, uint64_t cap, uint64_t* length);

/* `autocycler decompress` (decompress.rs:27-137): every contig of the GFA's paths written back per original file under out_dir
 * (gzip when the name ends in .gz) and/or as one FASTA (out_file); either may be NULL, not both. */
int ac_decompress_gfa(const char* in_gfa, const char* out_dir, const char* out_file, int32_t device, int32_t verbose);

/* reconstruct_original_sequence (unitig_graph.rs:383-400; decompress.rs:83-105 writes these out): the sequence spelled by
 * the path of loaded/added sequence `index` through the current graph.  `out` may be null to query `length` only. */
int ac_sequence_reconstruct(const ac_handle* h, uint64_t index, char* out, uint64_t cap, uint64_t* length);

#ifdef __cplusplus
}
#endif
#endif
