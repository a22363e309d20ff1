/*
 * b2ingest.h — C ABI of libb2ingest.so: the B200 (sm_100a) ingest + label-aggregation
 * hot path of Elmer-Carvalho/Image-Classification-System, re-built from scratch.
 *
 * Conventions
 *   - Every entry point returns int: 0 = B2_OK, < 0 = b2_status error.  The message of the
 *     last error on the calling thread is b2_last_error().  No C++ exception crosses the ABI.
 *   - Pointers named d_* are DEVICE pointers on the current CUDA device of the calling
 *     thread (in Python: tensor.data_ptr()); h_* are host pointers.  The library never
 *     frees or retains caller memory; outputs are caller-allocated.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  All
 *     work is enqueued on it and the call returns without synchronising unless stated.
 *   - Re-entrant: the device-pointer entry points keep no global mutable state except immutable
 *     per-shape coefficient tables owned by b2_resize_plan objects; the host-pointer entry points
 *     (b2_*_host, b2_ingest_ring_*) share a mutex-guarded pool of page-locked staging buffers and a
 *     cache of resize plans.  Several host threads may call concurrently (the reference reaches this
 *     path from up to five service threads); a b2_ingest_ring or a b2_comm belongs
 *     to one thread at a time.
 *   - There is no CPU fallback.  Without a Blackwell GPU every compute entry point fails.
 *
 * Each entry point cites the reference interface (file:line under the reference repo) it
 * replaces or — where the reference has no implementation — the BASELINE.json config that
 * requires it.
 */
#ifndef B2INGEST_H
#define B2INGEST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum b2_status {
    B2_OK = 0,
    B2_ERR_BAD_ARG = -1,      /* null pointer, zero/oversized dimension, misaligned buffer   */
    B2_ERR_CUDA = -2,         /* a CUDA runtime call failed; see b2_last_error()              */
    B2_ERR_NOT_SORTED = -3,   /* b2_label_tally_status: rows out of image order in sorted mode */
    B2_ERR_NO_DEVICE = -4,    /* no CUDA device / not compute capability 10.x                */
    B2_ERR_WORKSPACE = -5,    /* workspace smaller than the matching *_workspace_bytes()      */
    B2_ERR_NCCL = -6          /* NCCL missing in the process or a collective failed            */
} b2_status;

/* ---- library ------------------------------------------------------------------------- */
int b2_version(void);                       /* ABI version, currently 2                       */
const char *b2_last_error(void);            /* thread-local, never NULL                        */
int b2_init(int device);                    /* cudaSetDevice + capability check (sm_100)       */
int b2_device_sm_count(int device, int *sm_count);
/* Release what the host-pointer entry points keep between calls (pool of page-locked staging buffers, cached resize
 * plans).  Objects the caller created (plans, rings, communicators) are destroyed by the caller first; no other call
 * may be in flight.  The library stays usable: everything is re-created on demand. */
int b2_shutdown(void);

/* ---- a1: content hash -------------------------------------------------------------------
 * Replaces hashlib.sha256(data).hexdigest() at app/services/webdav_sync.py:59 (called :445),
 * app/services/activity_api_sync.py:798 and app/api/routes/images.py:62, for a batch.
 * Message i is the byte range d_data[d_offsets[i] .. d_offsets[i] + d_lengths[i]).
 * One lane hashes one message; a warp owns 32 consecutive slots of `d_order` (a
 * permutation of 0..n-1, normally messages sorted by length so a warp's lanes finish
 * together; NULL = identity).  Batches of up to 2 x SM-count message warps run as warp
 * PAIRS (one warp expands the message schedule into shared memory, the other runs the
 * rounds: 1.4x the per-message speed); larger ones with one warp per 32 messages.
 * Message starts that are 16-byte aligned take the 128-bit-load path.  d_digests receives
 * n x 32 raw digest bytes (FIPS 180-4 byte order).
 */
int b2_sha256_batch(const uint8_t *d_data, const uint64_t *d_offsets, const uint64_t *d_lengths,
                    const uint32_t *d_order, uint32_t n, uint8_t *d_digests, void *stream);

/* 32-byte digests -> 64 lowercase ASCII hex chars each (the String(64) primary key,
 * app/db/models.py:212).  d_hex receives n x 64 chars, no terminator. */
int b2_digest_hex(const uint8_t *d_digests, uint32_t n, char *d_hex, void *stream);

/* ---- a4: dedupe decision ----------------------------------------------------------------
 * Replaces the sequential lookup-then-insert loop of WebDAVSync._process_image_batch,
 * app/services/webdav_sync.py:311-400 (lookup :324, insert+flush :329-354, update
 * :371-398, counters :354/:398/:400), for a batch of n digests:
 *   d_valid[i] == 0 (or NULL = all valid) marks an image skipped before the lookup
 *       (:314 invalid, :320 failed download): not counted, is_new = 0, first_index = -1.
 *   d_existing: m digests already in table `imagens`, sorted ascending in memcmp order
 *       (NULL/0 = empty table).
 *   d_seq (NULL = i): arrival order; "first occurrence" = smallest seq (multi-GPU callers
 *       pass the global image index after the all-gather).
 * Outputs: d_is_new[i] = 1 iff i is the first valid occurrence of its digest and the digest
 * is not in d_existing (the insert branch); d_first_index[i] / d_last_index[i] (each may be NULL)
 * = batch index of the first / last valid occurrence of the same digest (identity columns
 * come from the first, nome_img/caminho_img from the last); d_counts[3] =
 * {processed, created, updated}.  d_workspace: b2_dedupe_workspace_bytes(n) bytes.
 */
uint64_t b2_dedupe_workspace_bytes(uint32_t n);
int b2_dedupe(const uint8_t *d_digests, const uint8_t *d_valid, const uint32_t *d_seq, uint32_t n,
              const uint8_t *d_existing, uint64_t m,
              uint8_t *d_is_new, int32_t *d_first_index, int32_t *d_last_index,
              uint32_t *d_counts, void *d_workspace, uint64_t workspace_bytes, void *stream);

/* Lookup only (app/api/routes/images.py:65: PK lookup per uploaded file):
 * d_found_index[i] = position of digest i in the sorted d_existing, or -1. */
int b2_lookup_sorted(const uint8_t *d_digests, uint32_t n, const uint8_t *d_existing, uint64_t m,
                     int64_t *d_found_index, void *stream);

/* ---- a12: thumbnail / preview tensor ------------------------------------------------------
 * Absent in the reference; required by BASELINE.json configs 1,2,3,5 ("hash + 256x256
 * thumbnail").  Semantics = Pillow (the reference's pinned image library,
 * requirements.txt:9): Image.resize((out_w,out_h), Image.BILINEAR) on decoded RGB HWC
 * uint8, bit-exact (two-pass, uint8 intermediate, 22-bit fixed-point taps), plus an
 * optional float32 CHW preview ((u8/255 - mean[c]) * inv_std[c]).
 * A plan holds the per-axis tap tables for one (in_h, in_w) -> (out_h, out_w) shape.
 */
typedef struct b2_resize_plan b2_resize_plan;
int b2_resize_plan_create(int in_h, int in_w, int out_h, int out_w, b2_resize_plan **plan);
int b2_resize_plan_destroy(b2_resize_plan *plan);
/* Host copies of the tap tables (for tests): bounds = out x {first, count}; coeffs = out x ksize. */
int b2_resize_plan_taps(const b2_resize_plan *plan, int axis /*0=horizontal,1=vertical*/,
                        int *ksize, int32_t *h_bounds, int32_t *h_coeffs, uint64_t coeffs_capacity);
/* Image i of the call starts at d_rgb + d_offsets[i] (HWC, row pitch in_w*3, no padding) and
 * its outputs go to slot s = d_out_slot ? d_out_slot[i] : i :
 *   d_thumb   + s * out_h*out_w*3   (uint8 HWC)
 *   d_preview + s * 3*out_h*out_w   (float32 CHW; NULL = skip). */
int b2_resize_normalize_batch(const b2_resize_plan *plan, const uint8_t *d_rgb,
                              const uint64_t *d_offsets, const uint32_t *d_out_slot, uint32_t n,
                              uint8_t *d_thumb, float *d_preview,
                              const float mean[3], const float inv_std[3], void *stream);
/* ---- a13: per-image label tally + Fleiss partials ---------------------------------------
 * Absent in the reference (it only groups one user's rows, app/crud/classificacao_crud.py:
 * 318-322); required by BASELINE.json configs 1,4,5.  Rows are SoA (image_idx int32,
 * class_idx uint8, active uint8 — the dictionary-encoded `classificacoes` columns id_img,
 * id_opc, ativo, app/db/models.py:224-241).  Only rows with active != 0 count
 * (classificacao_crud.py:314).  This shard owns images [image_base, image_base+n_images).
 *   flags & B2_TALLY_SORTED: rows are ordered by image_idx (an index scan on id_img); each
 *       CTA builds tiles of images in shared memory and writes d_counts with plain coalesced
 *       stores (d_counts need not be zeroed).
 *   otherwise: any order; d_counts is zeroed then built with global atomics.
 * d_counts: int32[n_images * k].  d_partials: int64[k + B2_PARTIALS_EXTRA] =
 *   { T_0..T_{k-1}, S2 = sum n_ij^2, R = sum n_i, images with n_i >= 1, images with
 *     n_i >= 2, sum n_i (n_i - 1), rows tallied (active or not, image and class in range),
 *     adjacent row pairs out of image order (sorted mode only) },
 *   ZEROED by the call and accumulated with integer atomics — exact, so kappa derived from
 *   them is identical on 1/2/4/8 GPUs after an integer all-reduce.
 * The call never synchronises.  Once the partials are on the host, b2_label_tally_status()
 * turns the last two entries into B2_OK / B2_ERR_NOT_SORTED / B2_ERR_BAD_ARG (a row with
 * image_idx or class_idx out of range); on error d_counts is unspecified.
 * d_agree_hist (may be NULL): int64[B2_AGREE_BINS], zeroed by the call — the general-n Fleiss kappa without a
 *   second pass over the count matrix and still exact for any sharding: bin n (2 <= n < B2_AGREE_BINS) =
 *   sum over the images with n_i = n of (sum_j n_ij^2 - n_i), so that sum_i P_i = sum_n bin[n] / (n (n - 1)) is
 *   computed on the host from integers (after the same integer all-reduce as the partials); bin 0 = images with
 *   n_i >= B2_AGREE_BINS (general kappa then needs b2_fleiss_partials' d_sum_pi), bin 1 = 0.
 * Row arrays must be 16-byte aligned.
 */
#define B2_TALLY_SORTED 1u
#define B2_PARTIALS_EXTRA 7
#define B2_AGREE_BINS 1024
int b2_label_tally(const int32_t *d_image_idx, const uint8_t *d_class_idx, const uint8_t *d_active,
                   uint64_t rows, uint32_t image_base, uint32_t n_images, uint32_t k, uint32_t flags,
                   int32_t *d_counts, int64_t *d_partials, int64_t *d_agree_hist, void *stream);
int b2_label_tally_status(const int64_t *h_partials, uint32_t k, uint64_t rows);   /* host only */
/* Partials from an existing count matrix (e.g. after a row-sharded all-reduce of counts); the
 * last two entries stay 0.  d_sum_pi (may be NULL): sum over images with n_i >= 2 of
 * (sum_j n_ij^2 - n_i)/(n_i(n_i-1)), float64, reduced in a fixed order (reproducible for a given
 * n_images, k); needs b2_fleiss_workspace_bytes() of 8-byte aligned workspace.  d_agree_hist (may be NULL): as
 * in b2_label_tally. */
uint64_t b2_fleiss_workspace_bytes(uint32_t n_images);
int b2_fleiss_partials(const int32_t *d_counts, uint32_t n_images, uint32_t k,
                       int64_t *d_partials, double *d_sum_pi, int64_t *d_agree_hist,
                       void *d_workspace, uint64_t workspace_bytes, void *stream);

/* ---- next row (f3): per-user aggregation ---------------------------------------------------
 * Bulk form of COUNT(DISTINCT id_img) WHERE id_con = ? AND ativo
 * (app/api/routes/classificacoes.py:224-230) for every annotator at once: rows sorted by
 * (annotator_idx, image_idx); d_distinct[a] = number of distinct images with an active row. */
int b2_distinct_images_per_annotator(const int32_t *d_annotator_idx, const int32_t *d_image_idx,
                                     const uint8_t *d_active, uint64_t rows, uint32_t n_annotators,
                                     uint32_t *d_distinct, void *stream);

/* ---- next row (f2 i): dictionary-encode rows of `classificacoes` for the tally ---------------------
 * Row r: id_img = 64 lowercase hex characters at d_img_hex + 64 r (String(64) foreign key,
 * app/db/models.py:229), id_opc = 16-byte UUID at d_opc_uuid + 16 r (:231), ativo = d_ativo[r] (:233).
 * Dictionaries: d_image_keys = the n_images stored digests sorted in memcmp order (the table b2_dedupe
 * takes as `existing`; image index = position), d_option_keys = the k <= 255 option UUIDs of the environment
 * sorted in memcmp order (class index = position).  Outputs the SoA arrays b2_label_tally reads:
 * d_image_idx[r] (-1 = unknown image or a key that is not 64 characters of [0-9a-f]), d_class_idx[r]
 * (255 = unknown option), d_active[r] = ativo != 0; d_unknown[2] = rows with an unknown image / option.
 * Keys and dictionaries must be 16-byte aligned. */
int b2_encode_label_rows(const char *d_img_hex, const uint8_t *d_opc_uuid, const uint8_t *d_ativo,
                         uint64_t rows, const uint8_t *d_image_keys, uint64_t n_images,
                         const uint8_t *d_option_keys, uint32_t k, int32_t *d_image_idx,
                         uint8_t *d_class_idx, uint8_t *d_active, uint64_t *d_unknown, void *stream);

/* ---- host-buffer entry points: the end-to-end form of the path --------------------------------
 * What the reference holds when the path starts is bytes in HOST memory: downloaded files
 * (app/services/webdav_sync.py:441, the 50-image batch loop :273-283) and rows fetched from table
 * `classificacoes`.  These entry points take host pointers only and own the device side (staging
 * buffers, streams, events), so a caller needs nothing but an FFI.  Page-locked buffers
 * (b2_host_alloc, or any pinned allocation) make the copies asynchronous and full speed; pageable
 * memory works, slower. */
int b2_host_alloc(void **p, uint64_t bytes);
int b2_host_free(void *p);

/* b2_ingest_ring: the streaming form of the whole ingest step for LISTINGS of any mix of image sizes — what the
 * reference's sync loop holds after its downloads (webdav_sync.py:273-283, :441; BASELINE configs 2, 3, 5).
 * Depth is bounded in BYTES: one device staging ring of `ring_bytes` is carved first-in first-out into chunks of
 * consecutive listing entries (about `chunk_bytes` each, 0 = 1 GiB; smaller for small listings) holding
 * [metadata | images | thumbnails | previews]; a chunk is released when its hash kernel, its resize kernels and
 * the read-back of its outputs are done, and submit() blocks only while the ring is full.  SHA-256 is serial per
 * message (~60 MB/s per message whatever the batch), so PCIe is only kept busy with >= ~1 000 messages hashing at
 * once: size the ring at a second's worth of input (tens of GB) and keep several listings in flight.
 *
 * submit(): entry i of the listing is
 *     h_pixels[i] -> h_hw[2i] x h_hw[2i+1] x 3 bytes of decoded RGB (HWC); NULL pointer or a zero dimension = no
 *                    thumbnail for this entry (its output slots are zeroed if the chunk has other thumbnails,
 *                    else untouched).  h_pixels may be NULL altogether (hash + dedupe only).
 *     h_files[i]  -> h_file_lens[i] bytes = the message that is hashed (the downloaded file).  h_files == NULL:
 *                    the pixel buffer is the message (BASELINE's synthetic images).
 *     h_valid[i]  == 0 (h_valid may be NULL = all valid): skipped before the lookup (webdav_sync.py:314, :320) —
 *                    not copied, not counted, is_new = 0, first/last index = -1.
 * Outputs, all in host memory and in LISTING ORDER, complete when wait(ticket) returns: h_digests (n x 32),
 * h_is_new (n), h_first_index / h_last_index (n each, may be NULL), h_counts[3] = {processed, created, updated}
 * of b2_dedupe over the listing against h_existing_sorted (m digests in memcmp order, may be NULL / 0),
 * h_thumbs (n x out_h*out_w*3, uint8 HWC; NULL only when h_pixels is NULL), h_previews (n x 3*out_h*out_w float32
 * CHW, may be NULL).  Input and output buffers must stay valid until wait() returns; page-locked buffers make every
 * copy asynchronous.  Up to max_listings listings in flight; tickets may be waited for in any order.
 * poll(): *done = 1 when wait() would not block; *images_flushed (may be NULL) = leading entries of the listing
 * whose thumbnails / previews are already in the host buffers (a consumer can start on them).
 * A ring belongs to one thread at a time.  On any error inside submit() everything the ring had in flight is
 * drained before the call returns, so no copy still references the caller's buffers. */
typedef struct b2_ingest_ring b2_ingest_ring;
int b2_ingest_ring_create(int device, uint64_t ring_bytes, uint64_t chunk_bytes, uint32_t max_listings,
                          int out_h, int out_w, int want_preview, b2_ingest_ring **ring_out);
int b2_ingest_ring_destroy(b2_ingest_ring *r);
int b2_ingest_ring_submit(b2_ingest_ring *r, const uint8_t *const *h_pixels, const uint32_t *h_hw,
                          const uint8_t *const *h_files, const uint64_t *h_file_lens, const uint8_t *h_valid,
                          uint32_t n, const uint8_t *h_existing_sorted, uint64_t m,
                          uint8_t *h_digests, uint8_t *h_is_new, int32_t *h_first_index, int32_t *h_last_index,
                          uint32_t *h_counts /*3*/, uint8_t *h_thumbs, float *h_previews, uint64_t *ticket);
int b2_ingest_ring_wait(b2_ingest_ring *r, uint64_t ticket, uint64_t *h2d_bytes, uint64_t *d2h_bytes,
                        uint32_t *kernel_launches);
int b2_ingest_ring_poll(b2_ingest_ring *r, uint64_t ticket, int *done, uint32_t *images_flushed);
int b2_ingest_ring_stats(const b2_ingest_ring *r, uint64_t *ring_bytes, uint64_t *bytes_in_flight,
                         uint32_t *chunks_in_flight, uint64_t *stalls);
/* SHA-256 of n byte strings anywhere in host memory (h_msgs[i] -> h_lens[i] bytes: e.g. the buffers of the
 * downloaded files, webdav_sync.py:441-445): packed into a recycled page-locked buffer, one copy, one kernel,
 * digests (n*32) and optionally their lowercase hex form (n*64, no terminator) back.  Blocking. */
int b2_sha256_host(int device, const uint8_t *const *h_msgs, const uint64_t *h_lens, uint32_t n,
                   uint8_t *h_digests, char *h_hex);
/* b2_dedupe for digests, validity flags (NULL = all valid) and the sorted table in host memory.  Blocking. */
int b2_dedupe_host(int device, const uint8_t *h_digests, const uint8_t *h_valid, uint32_t n,
                   const uint8_t *h_existing_sorted, uint64_t m, uint8_t *h_is_new,
                   int32_t *h_first_index, int32_t *h_last_index, uint32_t *h_counts /*3*/);
/* Decoded RGB images of any mix of shapes in host memory (h_rgb[i] -> h_hw[2i] x h_hw[2i+1] x 3 bytes, HWC) ->
 * uint8 HWC thumbnails (n * out_h*out_w*3) and, if not NULL, float32 CHW previews in host memory.  Images are
 * grouped by shape, one launch per group with a tap plan cached inside the library.  Blocking. */
int b2_thumbnails_host(int device, const uint8_t *const *h_rgb, const uint32_t *h_hw, uint32_t n,
                       uint32_t out_h, uint32_t out_w, uint8_t *h_thumbs, float *h_previews,
                       const float mean[3], const float inv_std[3]);
/* Label rows in host memory -> h_partials (int64[k + B2_PARTIALS_EXTRA]) and, if not NULL, the count
 * matrix h_counts (int32[n_images * k]).  Blocking; returns the verdict of b2_label_tally_status. */
int b2_label_tally_host(int device, const int32_t *h_image_idx, const uint8_t *h_class_idx,
                        const uint8_t *h_active, uint64_t rows, uint32_t image_base, uint32_t n_images,
                        uint32_t k, uint32_t flags, int32_t *h_counts, int64_t *h_partials,
                        int64_t *h_agree_hist /* NULL or int64[B2_AGREE_BINS] */);
/* b2_distinct_images_per_annotator for rows (sorted by (annotator, image)) in host memory.  Blocking. */
int b2_distinct_images_host(int device, const int32_t *h_annotator_idx, const int32_t *h_image_idx,
                            const uint8_t *h_active, uint64_t rows, uint32_t n_annotators, uint32_t *h_distinct);

/* ---- (e) multi-GPU: one process per GPU, NCCL over NVLink / NVSwitch -----------------------------------------
 * SURVEY.md section 8(e).  Images and label rows shard with no data-path collective; the two exchanges are an
 * all-gather of digests for the cross-rank dedupe decision and an integer all-reduce of the label partials.  NCCL is
 * bound at run time (the copy already loaded in the process, else libnccl.so.2, else $B2_NCCL_LIB); without it these
 * return B2_ERR_NCCL.  b2_comm_unique_id() is called on ONE rank, its 128 bytes travel to the others by whatever the
 * host has (file, socket, torch.distributed store), then every rank calls b2_comm_init (collective, blocking).
 * Collectives are enqueued on the caller's stream and return without synchronising. */
#define B2_COMM_ID_BYTES 128
typedef struct b2_comm b2_comm;
int b2_comm_unique_id(uint8_t *id_out /*B2_COMM_ID_BYTES*/);
int b2_comm_init(int device, int rank, int world, const uint8_t *id /*B2_COMM_ID_BYTES*/, b2_comm **comm_out);
int b2_comm_destroy(b2_comm *c);
int b2_comm_info(const b2_comm *c, int *rank, int *world, int *nccl_version);
/* d_all (world x n_per_rank x 32) = every rank's d_local (n_per_rank x 32) in rank order (ncclAllGather). */
int b2_allgather_digests(b2_comm *c, const uint8_t *d_local, uint32_t n_per_rank, uint8_t *d_all, void *stream);
/* In-place sum over ranks of count int64 values: class totals + partials (+ agreement histogram) of b2_label_tally
 * (ncclAllReduce, ncclInt64, ncclSum).  kappa computed from the result is bit-identical for any GPU count. */
int b2_allreduce_i64(b2_comm *c, int64_t *d_values, uint64_t count, void *stream);
/* The dedupe decision of WebDAVSync._process_image_batch (webdav_sync.py:311-400) for a listing SHARDED over the
 * ranks in any way (dist.shard_by_bytes for mixed sizes): rank r holds n_local digests, their global listing
 * positions d_seq (unique over all ranks) and optional validity flags; n_max = the largest n_local of any rank.
 * Shards are padded to n_max with invalid entries, digests / positions / flags are all-gathered in one NCCL group
 * and every rank runs the same b2_dedupe keyed on the listing position, so "first seen wins" means first in the
 * LISTING whatever rank holds it.  Outputs for this rank's entries: d_is_new, d_first_seq / d_last_seq (may be NULL)
 * = listing position of the first / last occurrence of the same content (-1 = invalid entry); d_counts[3] =
 * {processed, created, updated} of the WHOLE listing (identical on every rank).  d_workspace: 256-byte aligned,
 * b2_dedupe_global_workspace_bytes(world, n_max) bytes. */
/* The label all-reduce fused into the tally kernel ("fuse the partial reduction into the tally
 * kernel epilogue").  The partial vector is 8.6 KB: an NCCL all-reduce of that size is pure latency (~30 us at eight
 * ranks, as long as config 4's 12.5 M-row tally).  b2_comm_enable_peer_reduce (collective, blocking, once) gives
 * every rank a mailbox in its own HBM that its peers map through CUDA IPC; b2_label_tally_reduce then runs the SAME
 * slab kernel as b2_label_tally whose last CTA writes the rank's vector into every peer's mailbox over NVLink,
 * publishes a flag, waits for the peers' flags and sums the slots in rank order — one kernel, no NCCL call, the same
 * integers in the same order on every rank.  d_partials_hist: int64[k + B2_PARTIALS_EXTRA + B2_AGREE_BINS] (partials,
 * then the agreement histogram), replaced by the sum over ranks; d_counts stays this rank's slab.  Needs
 * B2_TALLY_SORTED; any-order rows (or a communicator without mailboxes) fall back to b2_label_tally +
 * b2_allreduce_i64.  b2_peer_allreduce_i64: the same exchange for any int64 vector (one small kernel).  A peer that
 * never arrives is given up after ~3 s (b2_comm_peer_status reports it); the kernels never hang. */
int b2_comm_enable_peer_reduce(b2_comm *c, uint32_t max_values);
int b2_comm_peer_status(b2_comm *c, int *timed_out);
int b2_peer_allreduce_i64(b2_comm *c, int64_t *d_values, uint32_t count, void *stream);
int b2_label_tally_reduce(b2_comm *c, const int32_t *d_image_idx, const uint8_t *d_class_idx, const uint8_t *d_active,
                          uint64_t rows, uint32_t image_base, uint32_t n_images, uint32_t k, uint32_t flags,
                          int32_t *d_counts, int64_t *d_partials_hist, void *stream);
uint64_t b2_dedupe_global_workspace_bytes(uint32_t world, uint32_t n_max);
int b2_dedupe_global(b2_comm *c, const uint8_t *d_digests, const uint8_t *d_valid, const uint32_t *d_seq,
                     uint32_t n_local, uint32_t n_max, const uint8_t *d_existing, uint64_t m,
                     uint8_t *d_is_new, int64_t *d_first_seq, int64_t *d_last_seq, uint32_t *d_counts /*3*/,
                     void *d_workspace, uint64_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B2INGEST_H */
