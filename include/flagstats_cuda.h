/*
 * flagstats_cuda.h -- C ABI of libflagstats_cuda.so, the B200 (sm_100a)
 * implementation of libflagstats' flagstat hot path.
 *
 * Plain C, plain pointers and sizes; no CUDA or torch types in any signature
 * (streams travel as void*).  Every entry point names the reference interface
 * it replaces or extends; file:line are in mklarqvist/libflagstats @93f68238.
 *
 * Counter layout (unchanged from the reference, libflagstats.h:69-112,118-142):
 *   flags[ 0..15]  QC-pass records, indexed by SAM bit number
 *   flags[16..31]  QC-fail records
 *   written slots: {2,6,7,8,10,11,12,13,14} in both halves, 25 (= n_fail) and
 *   9 (= n_pass, the convention of the kernels FLAGSTATS_u16 dispatches to,
 *   libflagstats.h:429,1212,1843).  All other slots are never touched.
 * All FLAGSTAT_* entry points ACCUMULATE into flags (caller zeroes), like the
 * reference kernels; POSPOPCNT_cuda_u16* ZERO out[] first, like
 * STORM_pospopcnt_u16 (libalgebra.h:3497-3498).
 *
 * Return value: 0 on success (as every reference kernel), otherwise a negative
 * FLAGSTAT_CUDA_E* code or a positive cudaError_t; on failure flags[] is left
 * untouched.  Nothing here aborts or throws.  There is no CPU fallback: with no
 * usable device every compute entry returns FLAGSTAT_CUDA_ENODEV.
 */
#ifndef FLAGSTATS_CUDA_H_
#define FLAGSTATS_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLAGSTAT_CUDA_ENODEV (-1) /* no CUDA device / driver */
#define FLAGSTAT_CUDA_EINVAL (-2) /* bad argument */
#define FLAGSTAT_CUDA_ENOMEM (-3) /* host allocation failed */
#define FLAGSTAT_CUDA_ESTATE (-4) /* handle used out of order */
#define FLAGSTAT_CUDA_ETIMEOUT (-5) /* a peer GPU never delivered its counters */
#define FLAGSTAT_CUDA_EFORMAT (-6) /* malformed block container / LZ4 block */
#define FLAGSTAT_CUDA_EIO (-7) /* cannot open or read the file */

/* ---- the reference's plugin signature ---------------------------------- */

/* Same type as FLAGSTATS_func, libflagstats.h:2970. */
typedef int (*FLAGSTATS_cuda_func)(const uint16_t*, uint32_t, uint32_t*);

/*
 * Drop-in sibling of FLAGSTAT_scalar / _sse4 / _avx2 / _avx512
 * (libflagstats.h:170,184,967,1646).  `array` may be a host pointer (pageable
 * or pinned: staged over PCIe in overlapped chunks) or a device / managed
 * pointer (kernel runs in place).  Any 2-byte alignment.  Synchronous.
 * Counters wrap mod 2^32 exactly like the reference's `++`.
 */
int FLAGSTAT_cuda(const uint16_t* array, uint32_t len, uint32_t* flags);

/* Same, with 64-bit length and counters (the reference cannot express
 * len >= 2^32 or counts >= 2^32; libflagstats.h:170 uses uint32_t). */
int FLAGSTAT_cuda_u64(const uint16_t* array, uint64_t len, uint64_t* flags);

/*
 * Device-resident, asynchronous form: d_array and d_flags (uint64_t[32]) are
 * device pointers on the current device, `stream` is a cudaStream_t (NULL =
 * legacy default stream).  Enqueues the kernel and returns; accumulates.
 * This is what a caller holding the FLAG column in HBM uses, and what the
 * multi-GPU path runs per shard before the 32-counter all-reduce.
 */
int FLAGSTAT_cuda_device(const uint16_t* d_array, uint64_t len, uint64_t* d_flags, void* stream);

/* The same, launched so that it may START before the previous kernel of `stream` has
 * finished (programmatic dependent launch): its CTAs stream d_array while that kernel
 * drains; the adds into d_flags still happen after it has completed, in stream order.
 * For back-to-back calls over resident columns (+5 % on 1.65 GB columns: no launch gap, no
 * kernel tail).  Contract: d_array must not be written by the kernel enqueued immediately
 * before this call on that stream.  Falls back to a plain launch on streams that cannot
 * take the attribute. */
int FLAGSTAT_cuda_device_overlapped(const uint16_t* d_array, uint64_t len, uint64_t* d_flags,
                                    void* stream);

/* Run-time half of the dispatch test the reference does with cpuid
 * (libflagstats.h:3000-3019): number of usable CUDA devices (0 = none). */
int FLAGSTAT_cuda_available(void);

/* Length at or above which FLAGSTATS_get_function should pick FLAGSTAT_cuda for
 * HOST data (the analogue of the 1024/512/256 thresholds, :3000,3006,3016).
 * Default 1,048,576 records: the measured crossover of one synchronous call on PAGEABLE host
 * memory against FLAGSTAT_avx512 on one core of a B200 host (227 vs 253 us; at the reference's
 * 512,000-record block it is a tie, 129 vs 126 us, because staging pageable memory costs one
 * core about what the AVX-512 kernel costs -- so the reference's own block loop stays on the
 * CPU by default).  Callers with PINNED buffers win from 131,072 records (24.7 vs 31.7 us) and
 * lower it through env FLAGSTAT_CUDA_MIN_LEN or the setter.  profiles/r4c_dropin_time.jsonl. */
uint32_t FLAGSTAT_cuda_min_len(void);
void FLAGSTAT_cuda_set_min_len(uint32_t n);

/* ---- samtools mode: the reference benchmark's flagstat_loop caller -----------
 *
 * benchmark/flagstats.cpp:43-71 keeps a bam_flagstat_t next to the 32 counters and
 * prints samtools' report from it (:577-588).  Ten of its eleven FLAG-derived
 * fields follow from the 32 counters; the eleventh, n_pair_all ("paired in
 * sequencing", :58-59), does not -- FLAGSTAT_scalar_update has it commented out
 * (libflagstats.h:132) and the Python wrapper approximates it as READ1 + READ2
 * (python/libflagstats.pyx:35).  The *_samtools entries count it exactly, in the
 * same single pass: flags[0] / flags[16] += n_pair_all of QC-pass / QC-fail
 * records (slots the FLAGSTAT_* contract leaves free), every other slot as
 * FLAGSTAT_cuda_u64.  Same pointer kinds, accumulate and error contract. */
int FLAGSTAT_cuda_samtools_u64(const uint16_t* array, uint64_t len, uint64_t* flags /*[32]*/);
/* async, device pointers (cf. FLAGSTAT_cuda_device) */
int FLAGSTAT_cuda_samtools_device(const uint16_t* d_array, uint64_t len, uint64_t* d_flags,
                                  void* stream);

/* bam_flagstat_t, benchmark/flagstats.cpp:43-49 (same fields, same order; each
 * [QC-pass, QC-fail]).  n_diffchr / n_diffhigh need RNAME / MAPQ, not the FLAG
 * word, and are never written. */
typedef struct {
    long long n_reads[2], n_mapped[2], n_pair_all[2], n_pair_map[2], n_pair_good[2];
    long long n_sgltn[2], n_read1[2], n_read2[2];
    long long n_dup[2];
    long long n_diffchr[2], n_diffhigh[2];
    long long n_secondary[2], n_supp[2];
} FLAGSTAT_cuda_bam_flagstat;

/* flagstat_loop (benchmark/flagstats.cpp:51-71) over a whole column: ADDS to *s. */
int FLAGSTAT_cuda_samtools(const uint16_t* array, uint64_t len, FLAGSTAT_cuda_bam_flagstat* s);
/* The same struct from 32 counters produced by a *_samtools entry (host memory); ADDS. */
int FLAGSTAT_cuda_samtools_from_counters(const uint64_t* flags /*[32]*/,
                                         FLAGSTAT_cuda_bam_flagstat* s);
/* The report of benchmark/flagstats.cpp:577-588, byte for byte (percent() of :73-78).
 * Writes at most capacity - 1 characters + NUL; returns the length of the full
 * report (snprintf convention) or a negative FLAGSTAT_CUDA_E* code.  Host only. */
int FLAGSTAT_cuda_samtools_report(const FLAGSTAT_cuda_bam_flagstat* s, char* buf, size_t capacity);

/* ---- raw positional popcount (STORM_pospopcnt_u16, libalgebra.h:3496) ---- */

int POSPOPCNT_cuda_u16(const uint16_t* data, size_t len, uint32_t* out /*[16]*/);
int POSPOPCNT_cuda_u16_u64(const uint16_t* data, uint64_t len, uint64_t* out /*[16]*/);
/* async, device pointers, ACCUMULATES into d_out[16] (caller zeroes) */
int POSPOPCNT_cuda_device(const uint16_t* d_data, uint64_t len, uint64_t* d_out, void* stream);

/* ---- block streaming (caller pattern of benchmark/flagstats.cpp:288-358:
 *      one shared counters[32] accumulated over 1,024,000-byte blocks) ------- */

typedef struct FLAGSTAT_cuda_stream FLAGSTAT_cuda_stream;

/* device: CUDA ordinal; block_records: capacity of one slot (512000 for the
 * reference's block); n_slots: pinned ring depth (>= 2; 0 = default 4). */
int FLAGSTAT_cuda_stream_open(FLAGSTAT_cuda_stream** s, int device, uint32_t block_records,
                              int n_slots);
/* Same with the transport spelled out.  mode: how a block crosses PCIe --
 * _DMA: cudaMemcpyAsync into a device twin of the ring, then the kernel;
 * _ZEROCOPY: no device staging at all, the kernel's own asynchronous loads read
 * the pinned block over PCIe (transfer and counting are one launch).
 * coalesce: consecutive full blocks shipped per DMA + launch (1 = every block on
 * its own; 0 = default 8: 97 % of the pinned-memcpy rate on a Gen5 x16 B200, profiles/r1l_stream.jsonl).  The ring then holds n_slots groups of `coalesce`
 * blocks; a short block closes its group early.  Counters are only complete
 * after _finish either way. */
#define FLAGSTAT_CUDA_STREAM_DMA 0
#define FLAGSTAT_CUDA_STREAM_ZEROCOPY 1
int FLAGSTAT_cuda_stream_open_ex(FLAGSTAT_cuda_stream** s, int device, uint32_t block_records,
                                 int n_slots, int mode, int coalesce);
/* Zero-copy producer API: get the next pinned slot (blocks until the slot's
 * previous transfer has finished), fill it, then submit n records. */
uint16_t* FLAGSTAT_cuda_stream_acquire(FLAGSTAT_cuda_stream* s);
int FLAGSTAT_cuda_stream_submit(FLAGSTAT_cuda_stream* s, uint32_t n_records);
/* Convenience: acquire + memcpy from any host pointer + submit. */
int FLAGSTAT_cuda_stream_push(FLAGSTAT_cuda_stream* s, const uint16_t* block, uint32_t n_records);
/* Wait for everything submitted so far, ADD the totals into flags[32] and
 * reset the device accumulator.  The handle stays usable. */
int FLAGSTAT_cuda_stream_finish(FLAGSTAT_cuda_stream* s, uint64_t* flags);
int FLAGSTAT_cuda_stream_close(FLAGSTAT_cuda_stream* s);
/* Measurement aid: re-submits the ring's current contents as n_blocks full blocks
 * (acquire + submit, no host-side fill -- the producer-already-wrote-the-slot
 * case), finishes into flags[32] and returns the wall-clock seconds of the
 * whole loop including the final synchronisation. */
int FLAGSTAT_cuda_stream_selftime(FLAGSTAT_cuda_stream* s, uint32_t n_blocks, uint64_t* flags,
                                  double* seconds);

/* ---- the reference's FLAG files (benchmark/flagstats.cpp) ------------------
 *
 * _RAW: a plain uint16 stream (".bin"), consumed in the reference's 1,024,000-byte
 *       blocks through the pinned ring (flagstat_raw, flagstats.cpp:415-468).
 * _LZ4: the container lz4f()/lz4hc() write (:110-186) and lz4_decompress() reads
 *       (:288-358): repeated [int32 raw_size][int32 comp_size][LZ4 block].  The
 *       blocks cross PCIe COMPRESSED and are decoded on the GPU (one CTA per
 *       block; FLAGSTAT_cuda_set_lz4_variant), then counted from HBM.  N = raw_size >> 1
 *       records per block like :323.
 * _ZSTD: the container zstd() writes (:192-215) and zstd_decompress() reads (:636-676): the
 *       same records around one Zstandard FRAME each (ZSTD_compress / ZSTD_decompress,
 *       :90-98).  Shipped compressed, decoded on the GPU (RFC 8878: Huffman / FSE / repeat
 *       offsets; no dictionaries, checksum not verified; an entropy stage with one lane
 *       per frame writes sequence descriptors, a CTA per frame then executes them with the
 *       copy phase of the LZ4 decoder; environment FLAGSTAT_CUDA_ZSTD_VARIANT=0 selects the
 *       first version, one thread per frame start to end, for A/B), counted from HBM.
 * flags[32] is accumulated into; *n_records (may be NULL) receives the number of
 * records counted.  Runs on the current device; synchronous. */
#define FLAGSTAT_CUDA_FILE_RAW 0
#define FLAGSTAT_CUDA_FILE_LZ4 1
#define FLAGSTAT_CUDA_FILE_ZSTD 2
/* OR into `format`: count like FLAGSTAT_cuda_samtools_u64 (flags[0] / flags[16] =
 * n_pair_all) -- the reference's "samtools" readers of the same files
 * (flagstat_loop per block, flagstats.cpp:496-519 raw, :547-590 LZ4). */
#define FLAGSTAT_CUDA_FILE_SAMTOOLS 0x100
int FLAGSTAT_cuda_file_u64(const char* path, int format, uint64_t* flags, uint64_t* n_records);
/* the same for a container already in host memory */
int FLAGSTAT_cuda_container_u64(const void* bytes, uint64_t n_bytes, int format, uint64_t* flags,
                                uint64_t* n_records);
/* Decode-only aid (tests / tools): n_blocks LZ4 blocks at comp_off[]/comp_size[] of
 * `comp` are decoded to raw_off[] of `raw` (host memory, raw_total bytes);
 * status[b] = bytes produced (must equal raw_size[b]) or < 0 for a malformed block. */
int FLAGSTAT_cuda_lz4_decode(const void* comp, uint64_t comp_bytes, const uint64_t* comp_off,
                             const uint32_t* comp_size, const uint64_t* raw_off,
                             const uint32_t* raw_size, uint32_t n_blocks, void* raw,
                             uint64_t raw_total, int* status);

/* The same for Zstandard frames (one frame per block, as in the _ZSTD container). */
int FLAGSTAT_cuda_zstd_decode(const void* comp, uint64_t comp_bytes, const uint64_t* comp_off,
                              const uint32_t* comp_size, const uint64_t* raw_off,
                              const uint32_t* raw_size, uint32_t n_blocks, void* raw,
                              uint64_t raw_total, int* status);

/* ---- FLAG ingest (benchmark/utility.cpp:29-32) -------------------------------
 * `samtools view FILE | cut -f 2` text -- one decimal FLAG per line -- to the
 * uint16 column, on the GPU: one record per line (std::getline: a last line
 * without '\n' counts), value = (uint16_t)atoi(line).  `out` (host, may be NULL)
 * receives the column if out_capacity records fit (else EINVAL with *n_records =
 * the room needed); `flags` (may be NULL) is accumulated with the counters of
 * the column straight from device memory. */
int FLAGSTAT_cuda_ingest_text(const char* text, uint64_t n_bytes, uint16_t* out, uint64_t out_capacity,
                              uint64_t* n_records, uint64_t* flags);

/* ---- several GPUs from one process (range shards, host-side sum) --------- */

/* Splits [0,len) into n_devices contiguous ranges (boundaries on 8-record
 * multiples), runs one shard per device concurrently and sums the 32 counters
 * on the host.  array must be a HOST pointer.  The one-process-per-GPU form
 * with an NCCL all-reduce lives in libflagstats_b200/sharded.py on top of
 * FLAGSTAT_cuda_device. */
int FLAGSTAT_cuda_multi_u64(const uint16_t* array, uint64_t len, uint64_t* flags, int n_devices);

/* ---- one process (or thread) per GPU: fused count + counter exchange -------
 *
 * The reference has no multi-device code.  The FLAG column is
 * range-sharded; each rank counts its shard and the 32 counters are summed
 * across ranks.  FLAGSTAT_cuda_device_allreduce does both in ONE kernel launch
 * per rank: the last CTA of each rank's kernel stores the rank's 32 totals
 * straight into every peer's exchange buffer over NVLink (peer-mapped memory; each
 * counter as two self-validating 8-byte words {32 bits | epoch tag}: no flag, no fence),
 * reads the peers' totals out of its own buffer as they land and writes the GLOBAL
 * counters to d_flags.  No NCCL launch, no host round trip.
 *
 * Setup: every rank calls _create (current device = its GPU), the 64-byte
 * handles are exchanged by whatever the host program uses to talk
 * (torch.distributed in libflagstats_b200/sharded.py), then _connect maps the
 * peers.  Ranks living in ONE process use _connect_local instead.
 * Every rank must make the same sequence of *_allreduce calls (it is a
 * collective).  world <= 16. */
typedef struct FLAGSTAT_cuda_xchg FLAGSTAT_cuda_xchg;
#define FLAGSTAT_CUDA_XCHG_HANDLE_BYTES 64
int FLAGSTAT_cuda_xchg_create(FLAGSTAT_cuda_xchg** x, int rank, int world,
                              void* handle_out /* 64 bytes, may be NULL when world == 1 */);
int FLAGSTAT_cuda_xchg_connect(FLAGSTAT_cuda_xchg* x, const void* all_handles /* world x 64 B */);
int FLAGSTAT_cuda_xchg_connect_local(FLAGSTAT_cuda_xchg** xs /* [world], by rank */, int world);
/* d_flags: uint64_t[32] on this rank's device, receives the GLOBAL counters:
 * accumulate != 0 adds them (the FLAGSTAT_* contract), 0 overwrites (saves the
 * caller a memset).  Asynchronous on `stream`; current device must be the
 * handle's. */
int FLAGSTAT_cuda_device_allreduce(FLAGSTAT_cuda_xchg* x, const uint16_t* d_array, uint64_t len,
                                   uint64_t* d_flags, int accumulate, void* stream);
int FLAGSTAT_cuda_samtools_device_allreduce(FLAGSTAT_cuda_xchg* x, const uint16_t* d_array,
                                            uint64_t len, uint64_t* d_flags, int accumulate,
                                            void* stream);
int POSPOPCNT_cuda_device_allreduce(FLAGSTAT_cuda_xchg* x, const uint16_t* d_data, uint64_t len,
                                    uint64_t* d_out /*[16]*/, int accumulate, void* stream);
/* Deferred collection: the same count + push, but the call does NOT wait for the peers' totals.
 * The global counters of this call are written to d_flags by the NEXT *_allreduce* call on the
 * handle (which first collects what its predecessor left pending, then pushes its own totals)
 * or by FLAGSTAT_cuda_xchg_collect -- so d_flags must stay valid until then.  For back-to-back
 * steps: a rank waits only for its peers' PREVIOUS step and may run one step ahead of the
 * slowest rank.  Measured on 2, 4 and 8 B200s with overlapped steps it ties with waiting in the
 * same launch: use it where the caller's own schedule benefits from
 * counters that arrive one call late, not for speed.  Every rank must
 * issue the same sequence of deferred / immediate / collect calls.  With one rank d_flags is
 * written at once. */
int FLAGSTAT_cuda_device_allreduce_deferred(FLAGSTAT_cuda_xchg* x, const uint16_t* d_array,
                                            uint64_t len, uint64_t* d_flags, int accumulate,
                                            void* stream);
/* Writes the pending counters of the last deferred call (one warp; asynchronous on `stream`,
 * which must be ordered after that call).  No-op when nothing is pending. */
int FLAGSTAT_cuda_xchg_collect(FLAGSTAT_cuda_xchg* x, void* stream);
/* Overlapped steps (default off; returns the previous setting).  When on, a *_allreduce call
 * on this handle may START before the previous kernel of the same stream has finished
 * (programmatic dependent launch): its CTAs stream their input while the previous call's
 * last CTA is still exchanging counters, which hides the exchange latency and the kernel
 * tail of back-to-back calls.  Everything a call writes (d_flags, the exchange buffers) still
 * happens after the previous kernel has completed, in stream order.  The caller's side of
 * the contract: d_array must not be written by the kernel enqueued immediately before the
 * call on that stream (data produced earlier, or on the other side of an event / a
 * synchronisation, is fine).  Needs a stream that accepts the attribute (not the NULL
 * stream on every driver); otherwise the call falls back to a plain launch. */
int FLAGSTAT_cuda_xchg_set_overlap(FLAGSTAT_cuda_xchg* x, int on);
/* Peers that never show up: the kernel gives up after the timeout (default 20 s),
 * leaves d_flags untouched, marks the exchange failed on every rank it can reach, and
 * _status returns FLAGSTAT_CUDA_ETIMEOUT.  The failure is sticky: later *_allreduce* calls on
 * the handle return FLAGSTAT_CUDA_ESTATE, and ALL ranks must destroy and recreate their
 * handles (the epoch-parity protocol assumes every rank completed every epoch).
 * _status synchronises with the device. */
int FLAGSTAT_cuda_xchg_set_timeout_ms(FLAGSTAT_cuda_xchg* x, uint32_t ms);
int FLAGSTAT_cuda_xchg_status(FLAGSTAT_cuda_xchg* x);
int FLAGSTAT_cuda_xchg_destroy(FLAGSTAT_cuda_xchg* x);

/* ---- diagnostics / test & bench support ---------------------------------- */

const char* FLAGSTAT_cuda_strerror(int code);
const char* FLAGSTAT_cuda_version(void);
/* Kernels this library has launched in this process (all threads). */
uint64_t FLAGSTAT_cuda_launch_count(void);
/* Select a kernel variant for A/B tests (0 = default).  Returns the previous, or
 * FLAGSTAT_CUDA_EINVAL for a variant this build does not contain: the product library carries
 * 0 (default), 1 (integer-only mask select), 3 (TMA-staged) and 8 (compare-mask forms); the
 * other measured-and-superseded variants need a -DFSB_ALL_VARIANTS build. */
int FLAGSTAT_cuda_set_variant(int variant);
/* LZ4 block decoder for A/B tests: 2 = one CTA per block, parse and copy phases (default),
 * 1 = one warp per block, 32 sequences per step, 0 = one warp per block, one sequence per
 * step.  Returns the previous.  Env FLAGSTAT_CUDA_LZ4_VARIANT sets the initial value. */
int FLAGSTAT_cuda_set_lz4_variant(int variant);
/* Name of the kernel instantiation the selected variant launches (as ncu prints it);
 * mode 0 = flagstat, 1 = pospopcnt, 2 = samtools. */
const char* FLAGSTAT_cuda_kernel_name(int mode);
/* Persistent-grid size override: CTAs per SM (0 = default). */
int FLAGSTAT_cuda_set_ctas_per_sm(int n);
/* A/B builds (-DFSB_ALL_VARIANTS) only: hand the work of the default kernel out dynamically (every
 * warp claims 8 KiB chunks from per-launch counters) instead of splitting it statically.
 * min_chunks: -1 = static (the only value the product library accepts), 0 = dynamic from 6
 * chunks per resident warp on, > 0 = dynamic from that many chunks on.  groups_per_chunk: 1 or
 * 2 (8 / 16 KiB per claim), 0 = leave unchanged.  Measured: balances the CTAs, does not shorten
 * the launch (profiles/r4h_*). */
int FLAGSTAT_cuda_set_dynamic(long long min_chunks, int groups_per_chunk);

/* Deterministic synthetic FLAG columns, pure functions of the GLOBAL record
 * index.  Device twins of oracle_synth_uniform/_hiseqx; d_out is
 * a device pointer; async on `stream`. */
int FLAGSTAT_cuda_synth_uniform(uint16_t* d_out, uint64_t start, uint64_t n, uint64_t seed,
                                uint16_t mask, void* stream);
int FLAGSTAT_cuda_synth_hiseqx(uint16_t* d_out, uint64_t start, uint64_t n, uint64_t seed,
                               uint32_t qcfail_ppm, void* stream);

/* Minimal device-memory helpers so C callers need no CUDA headers. */
void* FLAGSTAT_cuda_malloc(size_t bytes);
void* FLAGSTAT_cuda_malloc_host(size_t bytes); /* pinned */
int FLAGSTAT_cuda_free(void* p);
int FLAGSTAT_cuda_free_host(void* p);
int FLAGSTAT_cuda_memcpy_h2d(void* d, const void* h, size_t bytes);
int FLAGSTAT_cuda_memcpy_d2h(void* h, const void* d, size_t bytes);
int FLAGSTAT_cuda_memset(void* d, int value, size_t bytes);
int FLAGSTAT_cuda_sync(void);
/* Times `iters` back-to-back device-resident launches with CUDA events on an
 * internal stream; returns 0 and the mean milliseconds per launch. pospopcnt_mode: 0 = flagstat,
 * 1 = pospopcnt, 2 = samtools mode. */
int FLAGSTAT_cuda_time_device(const uint16_t* d_array, uint64_t len, uint64_t* d_flags, int iters,
                              int pospopcnt_mode, float* ms_per_launch);
/* The same over n_rot copies of the column laid out stride_records apart (launch i reads copy
 * i % n_rot): columns smaller than the 126 MB L2 are then timed from HBM rather than from the
 * cache a back-to-back loop over ONE copy would hit.  mode as pospopcnt_mode above; + 4: the
 * launches are overlapped ones (FLAGSTAT_cuda_device_overlapped). */
int FLAGSTAT_cuda_time_device_rot(const uint16_t* d_base, uint64_t len, uint64_t stride_records,
                                  uint32_t n_rot, uint64_t* d_flags, int iters, int mode,
                                  float* ms_per_launch);
/* Read-only HBM probe over device memory (16-byte aligned): LDG.128 + one XOR per 16
 * bytes, nothing else; mean milliseconds per pass.  The roofline a stream that only
 * reads can reach on this device (a copy pays bus turn-arounds a read does not). */
int FLAGSTAT_cuda_read_probe(const void* d_bytes, uint64_t n_bytes, int iters, float* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* FLAGSTATS_CUDA_H_ */
