/*
 * jne_dat.h -- batched, byte-compatible EIGENVALS_V6 .dat writer / reader.
 *
 * Wire format fixed by the reference's DATA_FORMAT.md:15-72 and written today record by record by
 * src/data_storage/writer.rs:206-259 (AppendOnlyWriter::append_eigenvalues) behind a single writer thread
 * (src/data_storage/thread_manager.rs:59-61).  At ~10^7 records/s that per-record path is the end-to-end
 * bottleneck; these functions encode whole batches into one buffer and scan files without materialising
 * a Vec per record.  Host-only code; no GPU is touched.
 *
 *   header  : "EIGENVALS_V6" | model u8 | dim u8 | steps u32 LE                       (18 bytes)
 *   record  : ULEB128(seed u32) | count u8 | count x f64 LE
 *   trailer : "EOF_MARK" | total records u64 LE | eigenvalues per record u8           (17 bytes)
 */
#ifndef JNE_DAT_H
#define JNE_DAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jne_dat_writer jne_dat_writer;

/* ULEB128 codec for u32 (src/data_storage/uleb128.rs:68-140).  encode returns the byte count (1..5).
 * decode returns the bytes consumed, or -1 incomplete encoding, -2 encoding too long (> 5 bytes),
 * -3 value too large for u32 -- the three error kinds of Uleb128Error. */
int jne_uleb128_encode(uint32_t value, uint8_t out[5]);
int jne_uleb128_decode(const uint8_t* bytes, size_t len, uint32_t* value);
int jne_uleb128_encoded_size(uint32_t value);

/* Expected file size for seeds 1..=num_runs (src/data_storage/file_format.rs:14-27). */
uint64_t jne_dat_expected_file_size(uint64_t num_runs, uint32_t eigenvalues_per_run);

/* Open for appending, like AppendOnlyWriter::with_expected_size (src/data_storage/writer.rs:28-178):
 * a new file gets the header; an existing file is validated (model / dim / steps mismatch -> JNE_ERR_IO with
 * the reference's message "... mismatch: file has ..., expected ..."), its 17-byte trailer is truncated
 * (writer.rs:181-203) and a torn last record is dropped; a file with a foreign magic header is recreated
 * (writer.rs:116-150).  *existing_records receives the number of complete records already present. */
int jne_dat_open(const char* path, uint8_t model, uint8_t dim, uint32_t steps,
                 uint64_t* existing_records, jne_dat_writer** out);

/* Append n records: seeds[i] with eigs[i*p .. i*p+p).  One encode pass, one write.  p must be 1..255 and
 * constant within a file ("Eigenvalue count mismatch", writer.rs:224-239). */
int jne_dat_append_batch(jne_dat_writer* w, const uint32_t* seeds, const double* eigs, uint64_t n, uint32_t p);
/* The same with record i's values at eigs[i*stride .. i*stride+p): appends one model's block of the rows of a fused
 * multi-model batch (jne_eigs_batch_multi) without a de-interleaving copy.  stride >= p. */
int jne_dat_append_batch_strided(jne_dat_writer* w, const uint32_t* seeds, const double* eigs, uint64_t n, uint32_t p,
                                 uint64_t stride);
/* The same with the records encoded by `threads` host threads straight into the file's pages (a random-access batch,
 * below, filled in contiguous ranges).  Same bytes as the serial call. */
int jne_dat_append_batch_strided_mt(jne_dat_writer* w, const uint32_t* seeds, const double* eigs, uint64_t n, uint32_t p,
                                    uint64_t stride, int threads);

/* Random-access batch: reserves the bytes of n records (seeds[i], p values each -- the sizes follow from the seeds) at
 * the end of the file, maps them, and lets several threads fill disjoint record ranges in any order; the file ends up
 * byte-identical to appending the same records serially.  This is how run_models_simulation lets every device's host
 * thread encode its rows straight from the pinned staging buffer into the file (no intermediate array, no writer
 * phase).  `seeds` must stay valid until jne_dat_batch_end.  Crash consistency: until the batch is committed its first
 * five bytes are 0xFF (an invalid ULEB128 at which the reference's scan reader, reader.rs:163-166, stops), so an
 * interrupted job resumes at the start of the batch.
 *   jne_dat_batch_fill   records first .. first+count-1 from rows (record i at rows + (i - first) * stride); thread-safe
 *                        for disjoint ranges; threads > 1 splits a large range over short-lived helper threads
 *   jne_dat_batch_end    commit != 0: every record must have been filled; the batch becomes part of the file.
 *                        commit == 0: the file is truncated back to where the batch began.  Frees the batch. */
typedef struct jne_dat_batch jne_dat_batch;
int jne_dat_batch_begin(jne_dat_writer* w, const uint32_t* seeds, uint64_t n, uint32_t p, jne_dat_batch** out);
int jne_dat_batch_fill(jne_dat_batch* b, uint64_t first, uint64_t count, const double* rows, uint64_t stride, int threads);
int jne_dat_batch_end(jne_dat_batch* b, int commit);

/* Flush buffered records to the OS (the reference flushes every 10 000 records, config.rs:5). */
int jne_dat_flush(jne_dat_writer* w);

/* Write the trailer and close (AppendOnlyWriter::finish, writer.rs:262-302).  Frees w. */
int jne_dat_finish(jne_dat_writer* w);

/* Close WITHOUT a trailer (simulates an interrupted run; the file stays readable by the scan path). Frees w. */
void jne_dat_abandon(jne_dat_writer* w);

const char* jne_dat_last_error(void);   /* thread-local message of the last failing jne_dat_* call */

/* Header + record count of a file (read_append_file, src/data_storage/reader.rs:23-73): fast path through the
 * trailer, else a scan that stops at a torn record (reader.rs:157-216).  has_trailer: 1 if a valid trailer exists. */
int jne_dat_info(const char* path, uint8_t* model, uint8_t* dim, uint32_t* steps,
                 uint64_t* n_records, uint32_t* eigenvalues_per_run, int* has_trailer);

/* Read up to `capacity` records into seeds[capacity] and eigs[capacity * p]; *n_read receives the count. */
int jne_dat_read(const char* path, uint32_t* seeds, double* eigs, uint64_t capacity, uint32_t p, uint64_t* n_read);

/* Resume scan (row f4; replaces check_append_progress + get_remaining_seeds, src/data_storage/progress.rs:11-61,
 * which build a HashSet of every completed seed): walks the record headers only, sets bit (seed-1) of `bitmap`
 * (ceil(num_runs/8) bytes, zeroed by the callee) for every completed seed in 1..=num_runs, and returns the number
 * of completed records in *completed.  A missing file is 0 completed; a parameter mismatch is JNE_ERR_IO. */
int jne_dat_completed_bitmap(const char* path, uint8_t model, uint8_t dim, uint32_t steps, uint64_t num_runs,
                             uint8_t* bitmap, uint64_t* completed);

/* Seeds of 1..=num_runs whose bit is clear, ascending (get_remaining_seeds, progress.rs:56-61).
 * Returns the count; writes at most `capacity` seeds to out (out may be NULL to just count). */
uint64_t jne_dat_remaining_seeds(const uint8_t* bitmap, uint64_t num_runs, uint32_t* out, uint64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* JNE_DAT_H */
