/*
 * autometa_b200.h — C ABI of libautometa_b200.so
 *
 * B200 (sm_100a) kernels for the one data-parallel hot path of KwanLab/Autometa:
 *   autometa.common.kmers.count / normalize / embed (PCA stage) and the DBSCAN
 *   eps-neighbourhood sweep inside autometa.binning.recursive_dbscan.
 *
 * The reference is pure Python and has NO FFI for this path: the
 * boundary a maintainer binds is this header, called from the Python functions
 * that keep the reference's signatures (autometa_b200/common/kmers.py and
 * autometa_b200/binning/recursive_dbscan.py).  Each entry point below cites the
 * reference lines it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.
 *   - every function returns 0 on success, a negative AM_E* code on failure;
 *     am_last_error() returns a thread-local message for the last failure.
 *   - "d_" pointers are device pointers on the CUDA device that is current for
 *     the calling thread; "h_" pointers are host pointers.  All buffers are
 *     allocated and freed by the caller (as torch tensors in the Python host).
 *   - device functions are asynchronous on `stream` (a cudaStream_t passed as
 *     void*); they never synchronise and never allocate.
 *   - there is no CPU fallback: device entry points fail with AM_ECUDA when no
 *     sm_100 device/context is usable.
 */
#ifndef AUTOMETA_B200_H
#define AUTOMETA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AM_OK 0
#define AM_EINVAL (-1)   /* bad argument                                   */
#define AM_ECUDA (-2)    /* CUDA runtime / launch error                    */
#define AM_ECAPACITY (-3) /* caller-provided buffer too small (see docs)   */
#define AM_EFORMAT (-4)  /* malformed input (e.g. FASTA)                   */

#define AM_NORM_AM_CLR 0
#define AM_NORM_CLR 1
#define AM_NORM_ILR 2

/* bases per packed group: one 16-byte vector load = one lane's segment */
#define AM_GROUP_BASES 64
/* windows per work-list tile (32 lanes x 64 bases) */
#define AM_TILE_BASES 2048

int am_version(void);
const char *am_last_error(void);
/* number of CUDA kernels launched by this library since load (bench.py's gpu_launches) */
int64_t am_launch_count(void);
/* SMs that kernels with a static work split (Gram, projection) leave free, so that a latency-bound kernel of
 * another stream (the eigensolve of the previous pass, pipeline.FeatureStream) runs beside them; 0 = none. */
int am_set_reserved_sms(int n);
int am_get_reserved_sms(void);

/* ---------------------------------------------------------------------------
 * k-mer column table — replaces init_kmers / _revcomp
 * (autometa/common/kmers.py:38-53, 56-89).  Encoding A=0,T=1,C=2,G=3 (the
 * reference's enumeration order, kmers.py:75); column = rank of min(code, rc).
 * ------------------------------------------------------------------------- */
int am_kmer_num_columns(int k);                 /* 512 for k=5; <0 on bad k (1..7) */
int am_kmer_column_lut(int k, int32_t *h_lut);  /* h_lut[4^k]: MSB-first code -> column */
/* h_names: ncols * (k+1) bytes, NUL-terminated strings in column order */
int am_kmer_column_names(int k, char *h_names);

/* ---------------------------------------------------------------------------
 * FASTA ingest (host) — replaces Bio.SeqIO.parse as used at kmers.py:143-148,
 * 234-235: id = first whitespace-delimited token of the header line; sequence
 * = the record's lines concatenated with whitespace removed, case preserved.
 *
 * am_fasta_index: one pass over an in-memory FASTA buffer.
 *   h_seq_out   [>= len bytes]  sequence bytes of all records, concatenated
 *   h_offsets   [cap+1] int64   record i occupies h_seq_out[off[i], off[i+1])
 *   h_id_begin/h_id_end [cap]   byte span of record i's id inside h_buf
 *   *n_records  number of records found (if > cap returns AM_ECAPACITY with
 *               *n_records set to the required capacity)
 * ------------------------------------------------------------------------- */
int am_fasta_index(const uint8_t *h_buf, int64_t len, uint8_t *h_seq_out,
                   int64_t *h_offsets, int64_t *h_id_begin, int64_t *h_id_end,
                   int64_t cap, int64_t *n_records);
/* the same with an explicit thread count (<= 0: all host threads; 1: one pass on the calling thread): the buffer
 * is cut at line starts, headers and sequence bytes are counted per range, then every range writes from its
 * prefix-summed record number and output position. am_fasta_index = am_fasta_index_mt(..., 0).  With all four
 * output pointers null only the records are counted (*n_records), so the caller can size its arrays. */
int am_fasta_index_mt(const uint8_t *h_buf, int64_t len, uint8_t *h_seq_out,
                   int64_t *h_offsets, int64_t *h_id_begin, int64_t *h_id_end,
                   int64_t cap, int64_t *n_records, int n_threads);

/* ---------------------------------------------------------------------------
 * 2-bit packing.  Layout in HBM ("packed assembly"):
 *   codes      uint32 words, 16 bases per word, base b at bits [2(b%16), +2),
 *              A=0 T=1 C=2 G=3; any other byte (N, IUPAC, lower case — the
 *              reference counts upper-case ACGT only, kmers.py:197-204) is
 *              stored as 0 and recorded as an exception.
 *   starts[i]  first packed base of contig i; a multiple of 64 so that every
 *              contig begins on a 16-byte boundary.  am_pack_plan fills it and
 *              returns the total number of packed bases (multiple of 64, incl.
 *              one trailing pad group that kernels may over-read).
 *   exc_bitmap one bit per 64-base group: group holds >= 1 invalid base
 *   exc_rank   per 32 groups: number of flagged groups before this bitmap word
 *   exc_mask   one uint64 per flagged group (bit j = base j invalid), in group order
 * ------------------------------------------------------------------------- */
int am_pack_plan(const int64_t *h_offsets, int64_t n_contigs, int64_t *h_starts,
                 int32_t *h_lengths, int64_t *total_packed_bases);
/* sizes (in elements) of the arrays for a packed assembly of `total_packed_bases` */
int64_t am_pack_code_words(int64_t total_packed_bases);   /* uint32 count */
int64_t am_pack_bitmap_words(int64_t total_packed_bases); /* uint32 count (bitmap and rank) */
/* host packer; exc_capacity = room in h_exc_mask; *n_exc = flagged groups found
 * (AM_ECAPACITY if it does not fit: call again with a larger buffer). */
int am_pack_host(const uint8_t *h_seq, const int64_t *h_offsets, int64_t n_contigs,
                 const int64_t *h_starts, int64_t total_packed_bases,
                 uint32_t *h_codes, uint32_t *h_exc_bitmap, uint32_t *h_exc_rank,
                 uint64_t *h_exc_mask, int64_t exc_capacity, int64_t *n_exc,
                 int n_threads);
/* device packer: ASCII bytes already in HBM -> the same layout.  d_seq must be
 * 4-byte aligned and readable up to 4 bytes past the last base; d_exc_mask holds
 * exc_capacity entries; d_n_exc receives the number of flagged groups (if it
 * exceeds exc_capacity the surplus masks were dropped: call again with a larger
 * buffer).  d_work: am_pack_work_words(total) uint32 of scratch. */
int64_t am_pack_work_words(int64_t total_packed_bases);
int am_pack_device(const uint8_t *d_seq, const int64_t *d_offsets,
                   const int64_t *d_starts, int64_t n_contigs,
                   int64_t total_packed_bases, uint32_t *d_codes,
                   uint32_t *d_exc_bitmap, uint32_t *d_exc_rank,
                   uint64_t *d_exc_mask, int64_t exc_capacity, int64_t *d_n_exc,
                   void *d_work, void *stream);

/* FASTA-ingest by-products.  am_pack_device_gc is am_pack_device
 * that also counts, per contig and case-insensitively, the bases
 * Bio.SeqUtils.gc_fraction(seq, "remove") counts: d_gc_at[2i] = #{G,C,S},
 * d_gc_at[2i+1] = #{A,T,W,U} (uint32 [n_contigs x 2], zeroed by the call; NULL =
 * plain am_pack_device).  am_gc_finish turns the counts into
 * Metagenome.gc_content's column, autometa/common/metagenome.py:303-320:
 * gc_content[i] = gc / (gc + at) * 100 in IEEE double (0 when gc + at == 0).
 * Contig lengths are the offsets' differences (am_pack_plan). */
int am_pack_device_gc(const uint8_t *d_seq, const int64_t *d_offsets,
                      const int64_t *d_starts, int64_t n_contigs,
                      int64_t total_packed_bases, uint32_t *d_codes,
                      uint32_t *d_exc_bitmap, uint32_t *d_exc_rank,
                      uint64_t *d_exc_mask, int64_t exc_capacity,
                      int64_t *d_n_exc, uint32_t *d_gc_at, void *d_work,
                      void *stream);
int am_gc_finish(const uint32_t *d_gc_at, int64_t n_contigs,
                 double *d_gc_content, void *stream);

/* ---------------------------------------------------------------------------
 * Work list for the count kernel: contigs are cut into items of at most
 * `max_item_bases` windows (a multiple of AM_TILE_BASES) so that a multi-Mbp
 * contig is spread over many warps.  Item = {contig, first window, n windows,
 * flags}; flags bit0 = contig has several items (accumulate with atomics into a
 * pre-zeroed row).  h_multi_rows lists those contigs.
 * ------------------------------------------------------------------------- */
int am_count_plan(const int32_t *h_lengths, int64_t n_contigs, int k,
                  int32_t max_item_bases, int32_t *h_items /*4 x cap*/,
                  int64_t item_cap, int64_t *n_items, int32_t *h_multi_rows,
                  int64_t multi_cap, int64_t *n_multi);

/* ---------------------------------------------------------------------------
 * Kernel 1 — canonical k-mer count.  Replaces record_counter / seq_counter /
 * mp_counter (kmers.py:122-265): for contig i of length L the first L-k windows
 * are counted (kmers.py:186,193 — the last k-mer is dropped), a window counts
 * iff all k bases are upper-case ACGT, canonical column per am_kmer_column_lut.
 *   d_counts      [n_contigs x ncols] int32, row-major, fully overwritten
 *   d_col_present [ncols] int32, set to 1 for columns with any count (must be
 *                 zeroed by the caller; feeds am_clr's column drop, kmers.py:369)
 *   d_row_sum     [n_contigs] int32 total counted windows per contig (may be NULL)
 *   d_work        >= 16 bytes of device scratch (work-queue counter)
 * ------------------------------------------------------------------------- */
int am_count_kmers(const uint32_t *d_codes, const uint32_t *d_exc_bitmap,
                   const uint32_t *d_exc_rank, const uint64_t *d_exc_mask,
                   const int64_t *d_starts, const int32_t *d_items, int64_t n_items,
                   const int32_t *d_multi_rows, int64_t n_multi, int64_t n_contigs,
                   int k, int32_t *d_counts, int32_t *d_col_present,
                   int32_t *d_row_sum, void *d_work, void *stream);

/* ---------------------------------------------------------------------------
 * Kernel 2 — row-wise log-ratio normalisation.  Replaces autometa_clr
 * (kmers.py:333-373) and the skbio multiplicative_replacement + clr / ilr calls
 * (kmers.py:418-421).
 *   d_counts   [n_rows_in x D] int32 (NA == 0)
 *   d_row_index[n_rows_out] int64 source row of each output row, or NULL = identity
 *   d_col_index[D_keep] int32 source column of each retained column, or NULL = all
 *   d_out      [n_rows_out x D_out] float64; D_out = D_keep (am_clr, clr) or D_keep-1 (ilr)
 *   d_colsum   [D_out] float64 or NULL: column sums of the output are ADDED here
 *              (caller zeroes) — the PCA mean comes for free.  Deterministic:
 *              per-CTA partials in d_work (am_normalize_work_bytes(D) bytes) are
 *              reduced in a fixed order.  Not available for ilr (use am_colstats).
 *   d_colmaxabs[D_out] float64 or NULL (needs d_colsum): column max|out| is max-ed in
 *              (caller zeroes) — sizes the fixed-point scale of am_gram_i8.
 * am_clr: out = log1p(x) - mean_j log1p(x)   (== log(((x+1)/S)/gmean((x+1)/S)))
 * clr/ilr: p = x/S; zeros -> 1/D^2, others * (1 - z/D^2); log; centre; (ilr: Helmert basis)
 * ------------------------------------------------------------------------- */
int64_t am_normalize_work_bytes(int D);
int am_normalize(const int32_t *d_counts, int64_t n_rows_out, int D,
                 const int64_t *d_row_index, const int32_t *d_col_index, int D_keep,
                 int method, double *d_out, double *d_colsum, double *d_colmaxabs, void *d_work,
                 void *stream);

/* Fused normalise + quantise for the k = 5 table (D = 512, S = 6; am_clr, clr or ilr): one pass
 * writes the fp64 rows, adds the column sums and writes the int8 digit planes of
 * round((out - centre_j) sc_j) (layout of am_quantize_planes) that am_gram_planes and
 * am_project_planes consume.  d_center[512] is a provisional centre (e.g. the column means of a
 * row sample normalised first); d_sc from am_plane_scales.  ilr has 511 output columns: d_out,
 * d_colsum and the planes keep a pitch of 512 with an all-zero last column (entry 511 of
 * d_center / d_sc is ignored). */
int am_normalize_planes(const int32_t *d_counts, int64_t n_rows_out, int D,
                        const int64_t *d_row_index, int method, double *d_out, double *d_colsum,
                        const double *d_center, const double *d_sc, int S, int8_t *d_planes,
                        void *d_work, void *stream);

/* ---------------------------------------------------------------------------
 * Kernel 3 — PCA (sklearn 1.9 PCA, svd_solver="covariance_eigh", the branch
 * kmers.py:605 takes for N >= 10*D; sklearn/decomposition/_pca.py:560,604-646).
 *   am_colsum      colsum[D] += sum_i X[i,:]                       (_pca.py:560)
 *   am_gram        G[D x D] += (X - mu)^T (X - mu) over this rank's rows
 *                  (centred before the product instead of _pca.py:604-610's
 *                  X^T X - n mu mu^T: same matrix, no cancellation)
 *   am_eigh        all eigenpairs of the symmetric D x D matrix, descending
 *                  (_pca.py:611-624); d_evecs row j = eigenvector j with sklearn's
 *                  svd_flip(u_based_decision=False) sign (largest-|.| entry > 0)
 *   am_project     scores[N x P] = (X - mu) @ comps[P x D]^T        (_base.py:151-159)
 * d_work for am_gram / am_eigh: see am_gram_work_bytes / am_eigh_work_bytes.
 * ------------------------------------------------------------------------- */
int64_t am_colsum_work_bytes(int D);
int am_colsum(const double *d_X, int64_t n_rows, int D, double *d_colsum, void *d_work,
              void *stream);
/* column sums and column max|x| in one pass: colsum[D] += , colmaxabs[D] = max(colmaxabs, .)
 * (caller zeroes both); the max-abs sizes the fixed-point scale of am_gram_i8 */
int64_t am_colstats_work_bytes(int D);
int am_colstats(const double *d_X, int64_t n_rows, int D, double *d_colsum, double *d_colmaxabs,
                void *d_work, void *stream);
/* Covariance from the Gram in one launch (_pca.py:604-610): mu[D] = colsum / n and, in place,
 * G <- (G - n (mu - centre)(mu - centre)^T) / (n - 1); d_center may be NULL when G is already about
 * the mean; d_n points to the (possibly all-reduced) row count in device memory. */
int am_cov_finish(double *d_G, int D, const double *d_colsum, const double *d_center,
                  const double *d_n, double *d_mu, void *stream);
/* flags[0] = min(col_present), flags[1] = min(row_sum): am_clr's all-NA column / row tests
 * (kmers.py:369) in one launch. */
int am_count_flags(const int32_t *d_col_present, int ncols, const int32_t *d_row_sum,
                   int64_t n_contigs, int32_t *d_flags, void *stream);
/* SUM-reducible form of the same two tests, for the single all-reduce of the multi-GPU path:
 * d_out[0..ncols) = 1.0 where the column occurs in this shard, d_out[ncols] = number of contigs
 * without any counted window (kmers.py:369). */
int am_stats_pack(const int32_t *d_col_present, int ncols, const int32_t *d_row_sum,
                  int64_t n_contigs, double *d_out, void *stream);
/* Move a Gram taken about d_c_old to one about d_c_new (NULL = the origin for either):
 * G += n [(mu - c_new)(mu - c_new)^T - (mu - c_old)(mu - c_old)^T], mu = colsum / n.  Lets every
 * rank quantise about its own provisional centre and still all-reduce one buffer [G | colsum | n]
 * (sklearn's own form is the Gram about the origin, _pca.py:604-610). */
int am_gram_recenter(double *d_G, int D, const double *d_colsum, const double *d_c_old,
                     const double *d_c_new, const double *d_n, void *stream);
/* Column presence (0/1) and row sums of the rows d_row_index[0..n_rows) (NULL = all rows) of an int32
 * count table [. x D]: the two statistics am_count_kmers reports for an assembly, for a partition of a
 * table that arrived as TSV (large_data_mode.py:454-470 subsets counts per taxon before normalising). */
int am_table_stats(const int32_t *d_counts, const int64_t *d_row_index, int64_t n_rows, int D,
                   int32_t *d_col_present, int32_t *d_row_sum, void *stream);
/* Helpers of the launch-lean pass (pipeline.FeatureStream): x = x * a + b on an fp64 vector; an async memset;
 * the tail of the PCA fit — eigenvalues clamped at zero in place (_pca.py:630) and
 * ratio[i] = eigenvalue[i] / total[0] (_pca.py:645-646). */
int am_axpb_f64(double *d_x, int64_t n, double a, double b, void *stream);
int am_zero_async(void *d_ptr, int64_t nbytes, void *stream);
int am_pca_finish(double *d_evals, const double *d_total, int P, double *d_ratio, void *stream);
int64_t am_gram_work_bytes(int D);
int am_gram(const double *d_X, int64_t n_rows, int D, const double *d_mu,
            double *d_G, void *d_work, void *stream);
/* Tensor-core Gram (tcgen05.mma kind::i8, TMEM accumulators, TMA-staged operands): the same
 * G[D x D] += (X - mu)^T (X - mu) as am_gram, computed exactly on S balanced base-256 digit
 * planes of the fixed-point values round((x - mu_j) 2^(8S-2-e_j)), 2^e_j > colmaxabs_j + |mu_j|.
 * S in 2..7 (6: quantisation step 2^-46 of the column range).  d_colmaxabs[D] >= max_i |X[i,j]|
 * (from am_colstats or the normalise kernel).  d_work: am_gram_i8_work_bytes(n_rows, D, S)
 * bytes (digit planes S*n_rows*D int8 + partial tiles). */
int64_t am_gram_i8_work_bytes(int64_t n_rows, int D, int S);
/* The same pipeline in pieces, so the digit planes can be produced elsewhere (fused into the
 * normalise pass: am_normalize_planes) and shared with am_project_planes:
 *   am_plane_scales   e[j] = exponent with 2^e[j] > max|x_j| + |centre_j| (max|x_j| = d_colmaxabs[j],
 *                     or the scalar `bound` for every column when d_colmaxabs is NULL);
 *                     sc[j] = 2^(8S-2-e[j])
 *   am_quantize_planes planes[S][n_rows][pitch] int8, pitch = D rounded up to 16: balanced
 *                     base-256 digits (most significant plane first) of round((x - centre_j) sc[j])
 *   am_gram_planes    G[D x D] += sum_r (x_r - centre)(x_r - centre)^T from the planes.  The centre
 *                     need not be the mean: the caller subtracts n (mu - centre)(mu - centre)^T. */
int64_t am_planes_bytes(int64_t n_rows, int D, int S);
int am_plane_scales(const double *d_center, const double *d_colmaxabs, double bound, int D, int S,
                    int32_t *d_e, double *d_sc, void *stream);
int am_quantize_planes(const double *d_X, int64_t n_rows, int D, const double *d_center,
                       const double *d_sc, int S, int8_t *d_planes, void *stream);
int64_t am_gram_planes_work_bytes(int64_t n_rows, int D, int S);
int am_gram_planes(const int8_t *d_planes, int64_t n_rows, int D, int S, const int32_t *d_e,
                   double *d_G, void *d_work, void *stream);
int am_gram_i8(const double *d_X, int64_t n_rows, int D, const double *d_mu,
               const double *d_colmaxabs, int S, double *d_G, void *d_work, void *stream);
int64_t am_eigh_work_bytes(int D);
int am_eigh(const double *d_A, int D, double *d_evals, double *d_evecs,
            void *d_work, int max_sweeps, double tol, void *stream);
/* Leading P eigenpairs only (what kmers.embed keeps: pca_dimensions = 50, kmers.py:601-605)
 * by Householder tridiagonalisation inside one 16-CTA cluster (matrix resident in distributed
 * shared memory), Sturm-count multisection, inverse iteration and back-transformation — the
 * dsyevx route.  3 <= D <= 576, P <= min(D, 128).  d_evals[P] descending; d_evecs[P x D] rows
 * with the svd_flip sign; d_trace[1] (may be NULL) = sum of ALL eigenvalues (= trace), the
 * denominator of explained_variance_ratio_ (_pca.py:645-646). */
int64_t am_eigh_topk_work_bytes(int D, int P);
/* am_eigh_topk that also tells when its first kernel — the 16-CTA tridiagonalisation cluster — is RESIDENT:
 * CTA 0 stores started_value to *d_started (device memory, 4-byte aligned) as its first action.  Paired with
 * am_stream_wait_geq on another stream, it lets the caller start an SM-filling kernel only once the cluster has
 * its SMs, so that the two run side by side instead of the cluster queueing behind the persistent kernel. */
int am_eigh_topk_signal(const double *d_A, int D, int P, double *d_evals, double *d_evecs,
                        double *d_trace, void *d_work, uint32_t *d_started, uint32_t started_value,
                        void *stream);
/* Hold `stream` until *d_flag >= value (cuStreamWaitValue32: a stream-level wait, no kernel spins). */
int am_stream_wait_geq(void *stream, const uint32_t *d_flag, uint32_t value);
int am_eigh_topk(const double *d_A, int D, int P, double *d_evals, double *d_evecs,
                 double *d_trace, void *d_work, void *stream);
/* Tensor-core projection from the digit planes (tcgen05 kind::i8, K-major operands):
 * scores[n_rows x P] = (X - mu) @ comps^T, where the planes hold round((x - centre_j) 2^(46-e_j))
 * (am_normalize_planes / am_quantize_planes with S = 6) and mu is the exact mean.  P <= 64. */
int64_t am_project_planes_work_bytes(int D, int S);
int am_project_planes(const int8_t *d_planes, int64_t n_rows, int D, int S, const int32_t *d_e,
                      const double *d_center, const double *d_mu, const double *d_comps, int P,
                      double *d_scores, void *d_work, void *stream);
int am_project(const double *d_X, int64_t n_rows, int D, const double *d_mu,
               const double *d_comps, int P, double *d_scores, void *stream);

/* ---------------------------------------------------------------------------
 * Kernel 4 — DBSCAN(min_samples=1) eps step.  Replaces
 * DBSCAN(eps, min_samples=1).fit(X).labels_ at recursive_dbscan.py:109: every
 * point is a core point, so labels = connected components of the graph
 * {(i,j): sum_d (X[i,d]-X[j,d])^2 <= eps*eps} (fp64, column order, inclusive —
 * sklearn/neighbors/_binary_tree.pxi.tp:1952-1957), numbered by ascending
 * minimum member row (sklearn/cluster/_dbscan_inner.pyx visiting order).
 *   am_eps_union          unions all pairs within eps into the forest d_parent
 *                         (int32[N], caller initialises to identity once and
 *                         keeps it across increasing eps: the sweep is
 *                         incremental — components only merge)
 *   am_labels_from_parent compress + relabel; d_n_clusters int32[1]
 * d_work: am_dbscan_work_bytes(N) bytes of device scratch.
 * ------------------------------------------------------------------------- */
int64_t am_dbscan_work_bytes(int64_t n_points);
int am_eps_union(const double *d_X, int64_t n_points, int F, double eps,
                 const double *h_min /*F*/, const double *h_max /*F*/,
                 int32_t *d_parent, void *d_work, void *stream);
int am_labels_from_parent(int32_t *d_parent, int64_t n_points, int64_t *d_labels,
                          int32_t *d_n_clusters, void *d_work, void *stream);
int am_iota_i32(int32_t *d_v, int64_t n, void *stream);

/* ---------------------------------------------------------------------------
 * Per-eps cluster metrics, cutoff filter and median completeness of the passing clusters —
 * replaces add_metrics / apply_binning_metrics_filter (autometa/binning/utilities.py:125-262)
 * as used inside the sweep (recursive_dbscan.py:180-192).
 *   d_labels[n_points] int64 (0-based, from am_labels_from_parent), d_n_clusters int32[1];
 *   contig x marker counts in COO form (rows index points), n_markers = reference marker count M;
 *   outputs per cluster c < n_clusters (entries beyond are left untouched): completeness, purity (NaN when nothing present),
 *   coverage / GC sigma (ddof 1, NaN for singletons), present / single-copy marker counts;
 *   d_summary[3] = {n_clusters, median completeness of passing clusters (-inf if none), n passing}.
 * d_work: am_metrics_work_bytes(n_points, M) bytes, zero-filled once by the caller; the call leaves it
 * zero-filled again, so one buffer (of the largest size) can serve sweeps of any (n_points, M). */
int64_t am_metrics_work_bytes(int64_t n_points, int n_markers);
int am_cluster_metrics(const int64_t *d_labels, int64_t n_points, const int32_t *d_n_clusters,
                       const int32_t *d_marker_rows, const int32_t *d_marker_cols,
                       const int32_t *d_marker_vals, int64_t nnz, int n_markers,
                       const double *d_coverage, const double *d_gc, double completeness_cutoff,
                       double purity_cutoff, double coverage_stddev_cutoff,
                       double gc_content_stddev_cutoff, double *d_completeness, double *d_purity,
                       double *d_cov_std, double *d_gc_std, int32_t *d_present, int32_t *d_single,
                       double *d_summary, void *d_work, void *stream);

/* ---------------------------------------------------------------------------
 * The eps step with its grid read from device memory, and the grid of an eps on the host.
 *   am_eps_grid       out2 = {1 / cell edge, eps * eps} for the bounding box [h_min, h_max] — the two
 *                     eps-dependent values of a step (what am_eps_union derives from its `eps` argument)
 *   am_eps_union_dyn  am_eps_union with those two doubles read from d_dyn[2] when the kernels run, so the
 *                     same launch sequence (or its captured graph) serves every eps of a sweep */
int am_eps_grid(double eps, const double *h_min, const double *h_max, int F, double *out2);
int am_eps_union_dyn(const double *d_X, int64_t n_points, int F, const double *h_min /*F*/,
                     const double *d_dyn /*2*/, int32_t *d_parent, void *d_work, void *stream);

/* ---------------------------------------------------------------------------
 * The WHOLE eps sweep of recursive_dbscan (autometa/binning/recursive_dbscan.py:172-215) resident on the
 * device: besides the step itself (am_eps_union_dyn -> am_labels_from_parent -> am_cluster_metrics) a
 * one-thread controller kernel advances the loop state exactly as the reference's Python does —
 *     take = median >= best_median (:194-196);  clustering_rounds[n_clusters] += 1 (:199-202);
 *     > 10 -> step *= 10 (:204-205);  median == 0 -> clustering_rounds[0] += 1 (:206-207);
 *     clustering_rounds[0] >= 10 -> stop (:208-209);  eps += step (:215);  n_clusters <= 1 -> stop (:179)
 * (IEEE double add / multiply without contraction = Python float arithmetic) — writes the next step's grid
 * to where am_eps_union_dyn reads it, and a copy kernel keeps the labels and per-cluster metrics of the best
 * eps so far.  The host therefore neither reads anything back per eps nor decides anything: it queues steps
 * (one captured CUDA graph per step, replayed) and looks at the `done` word once per batch.  Steps queued
 * after the sweep has ended change nothing (the state is frozen and the step at an unchanged eps is
 * idempotent), so batches may overshoot.
 *
 * All buffers belong to the caller (device unless named h_*):
 *   d_ctl            am_sweep_ctl_bytes(n_points, trace_cap) bytes: AM_SWEEP_CTL_WORDS doubles of loop state
 *                    (indices AM_SWEEP_*), then int32 clustering_rounds[n_points + 1], then the trace,
 *                    trace_cap x 4 doubles {eps, n_clusters, median, best_median} per executed step
 *   d_metrics_work   am_metrics_work_bytes, zero-filled once; d_dbscan_work am_dbscan_work_bytes
 *   d_* / d_best_*   the per-cluster outputs of am_cluster_metrics for the current / the best eps
 * am_sweep_init resets the state (eps 0.3, step 0.1, best -inf, forest = identity).  am_sweep_steps queues
 * `repeats` steps as plain launches; am_sweep_graph_create captures one step on `stream` (which must not be
 * the legacy default stream) into an executable graph, am_sweep_graph_launch replays it `repeats` times.
 * Both then copy the AM_SWEEP_CTL_WORDS state doubles to h_ctl (pinned) on the same stream when h_ctl is
 * not null. */
#define AM_SWEEP_CTL_WORDS 32
#define AM_SWEEP_EPS 0        /* eps of the next step to run                       */
#define AM_SWEEP_STEP 1       /* current increment                                 */
#define AM_SWEEP_BEST 2       /* best median completeness so far (-inf: none yet)  */
#define AM_SWEEP_DONE 3       /* 1 once the loop has ended                         */
#define AM_SWEEP_NSTEPS 4     /* steps executed                                    */
#define AM_SWEEP_BEST_STEP 5  /* 0-based step the kept labels come from (-1: none) */
#define AM_SWEEP_BEST_NC 6    /* number of clusters at that step                   */
#define AM_SWEEP_LAST_NC 7
#define AM_SWEEP_LAST_MED 8
#define AM_SWEEP_TAKE 9       /* 1 while the current step is the new best          */
#define AM_SWEEP_INV_CELL 10  /* d_dyn of am_eps_union_dyn: {1 / cell, eps^2}      */
#define AM_SWEEP_EPS2 11
#define AM_SWEEP_SPAN 12      /* 4 doubles: h_max - h_min                          */
#define AM_SWEEP_BEST_EPS 16  /* eps of the kept step                              */
typedef struct am_sweep_args {
  const double *d_X;            /* [n_points x F] row-major */
  int64_t n_points;
  int32_t F;
  int32_t n_markers;            /* reference marker count M */
  double h_min[4], h_max[4];    /* bounding box of X */
  int32_t *d_parent;            /* [n_points] forest */
  void *d_dbscan_work;
  int64_t *d_labels;            /* [n_points] current step */
  int32_t *d_n_clusters;        /* [1] */
  const int32_t *d_marker_rows, *d_marker_cols, *d_marker_vals;
  int64_t nnz;
  const double *d_coverage, *d_gc;
  double cutoffs[4];            /* completeness, purity, coverage sigma, GC sigma */
  double *d_completeness, *d_purity, *d_cov_std, *d_gc_std;
  int32_t *d_present, *d_single;
  double *d_summary;            /* [4] */
  void *d_metrics_work;
  void *d_ctl;
  int64_t *d_best_labels;
  double *d_best_completeness, *d_best_purity, *d_best_cov_std, *d_best_gc_std;
  int32_t *d_best_present, *d_best_single;
  int32_t trace_cap;
  int32_t reserved;
} am_sweep_args;
int64_t am_sweep_args_size(void); /* sizeof(am_sweep_args): lets a binding check its struct layout */
int64_t am_sweep_ctl_bytes(int64_t n_points, int trace_cap);
int am_sweep_init(const am_sweep_args *a, void *stream);
int am_sweep_steps(const am_sweep_args *a, int repeats, double *h_ctl, void *stream);
int am_sweep_graph_create(const am_sweep_args *a, void *stream, void **graph_out);
int am_sweep_graph_launch(void *graph, const am_sweep_args *a, int repeats, double *h_ctl, void *stream);
int am_sweep_graph_destroy(void *graph);
/* The controller's state machine on the host, for tests: advances ctl / rounds by one step with the given
 * outcome (n_clusters, median) exactly as the device kernel does; returns 1 when the step is the new best. */
int am_sweep_advance_host(double *ctl, int32_t *rounds, int F, int n_clusters, double median);

/* ---------------------------------------------------------------------------
 * Host-side conversion between the row-major count matrix and the per-column (values, mask) arrays of the
 * frame `count` returns (autometa/common/kmers.py:205,325-326: zero counts are pd.NA, contigs with L <= k are
 * all-NA rows, :187-192).  Blocked and threaded (n_threads <= 0: all host threads).
 *   am_counts_to_columns  values_t[c][r] = counts[r][c]; mask_t[c][r] = (counts[r][c] == 0 || !countable[r])
 *                         (countable may be null: every row countable)
 *   am_columns_to_counts  the inverse for `normalize`: column c has n_rows values of elem_size 4 (int32) or 8
 *                         (int64) bytes and a byte mask; counts[r][c] = masked ? 0 : value, isna[r][c] = masked;
 *                         *h_bad = number of unmasked values outside 0 .. 2^31-1 (the caller raises) */
int am_counts_to_columns(const int32_t *h_counts, int64_t n_rows, int64_t n_cols, const uint8_t *h_countable,
                         int32_t *h_values_t, uint8_t *h_mask_t, int n_threads);
int am_columns_to_counts(const void *const *h_col_values, const uint8_t *const *h_col_masks, int elem_size,
                         int64_t n_rows, int64_t n_cols, int32_t *h_counts, uint8_t *h_isna, int64_t *h_bad,
                         int n_threads);

/* ---------------------------------------------------------------------------
 * Host-side TSV I/O of the k-mer tables, all host threads.
 * Replaces `df.to_csv(out, sep="\t", index=True, header=True)` and
 * `pd.read_csv(fpath, sep="\t", index_col="contig")` at
 * autometa/common/kmers.py:116,328,431,704.  `header` = the complete header line
 * including its '\n'; `ids`/`id_offsets[n_rows+1]` = the (already quoted where the
 * CSV dialect requires it) index labels.  Floats are written as Python repr,
 * NaN / count 0 as empty cells: the files are byte-identical to pandas'.  The
 * parser reproduces pandas' default C-engine float conversion bit for bit and
 * returns AM_EINVAL for tables outside its fast path (quoted labels, ragged rows,
 * non-numeric cells). */
int am_tsv_write_f64(const char *path, const uint8_t *header, int64_t header_len,
                     const uint8_t *ids, const int64_t *id_offsets, int64_t n_rows,
                     int64_t n_cols, const double *values, int n_threads);
int am_tsv_write_i32(const char *path, const uint8_t *header, int64_t header_len,
                     const uint8_t *ids, const int64_t *id_offsets, int64_t n_rows,
                     int64_t n_cols, const int32_t *counts, int n_threads);
int am_tsv_shape(const uint8_t *buf, int64_t len, int64_t *n_rows, int64_t *n_fields);
int am_tsv_parse_f64(const uint8_t *buf, int64_t len, int64_t n_rows, int64_t n_fields,
                     int64_t *id_begin, int64_t *id_end, double *values,
                     uint8_t *col_all_int, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* AUTOMETA_B200_H */
