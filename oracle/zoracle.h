/*
 * zoracle.h -- CPU restatement ("oracle") of the zlib-rs hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by
 * or executed from the shipped library (zlib_rs_b200/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it, and only as the checker / CPU baseline.
 *
 * The reference (trifectatechfoundation/zlib-rs @ aeded1c, v0.6.7) is Rust and
 * cannot be compiled in this image (no rustc/cargo).  This is a plain-C
 * restatement of its algorithms for: deflate (all levels / strategies,
 * zlib-rs/src/deflate.rs + deflate/algorithm/{*}.rs + deflate/longest_match.rs +
 * deflate/hash_calc.rs), inflate (zlib-rs/src/inflate.rs + inflate/{*}.rs),
 * adler32 (zlib-rs/src/adler32.rs) and crc32 (zlib-rs/src/crc32.rs,
 * crc32/combine.rs).  It is pinned by the reference's own literal golden
 * vectors (tests/golden/, extracted by tests/golden/extract_vectors.py) and by
 * round trips through the system's stock zlib 1.3.
 *
 * Parity status: pinned on every literal vector the reference's tests hold for this path (tests/golden/kat.json: 15 deflate
 * vectors over all levels / strategies / flush modes, the inflate and checksum vectors, the hash and slide_hash tables) AND at
 * corpus scale for level 9: the oracle's level-9 stream of silesia-small.tar equals the reference repository's own
 * silesia-small.tar.gz byte for byte (tests/test_oracle.py).  The other levels have no corpus-scale golden in the reference
 * (it asserts equality with zlib-ng, which is not vendored: SURVEY.md section 8c); they share the block / tree / bit-writer code
 * that the level-9 pin exercises.  deflateSetDictionary is checked against stock zlib only (no literal vector exists).
 */
#ifndef ZORACLE_H
#define ZORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* return codes: zlib-rs/src/lib.rs:224-236 */
#define ZO_OK 0
#define ZO_STREAM_END 1
#define ZO_NEED_DICT 2
#define ZO_ERRNO (-1)
#define ZO_STREAM_ERROR (-2)
#define ZO_DATA_ERROR (-3)
#define ZO_MEM_ERROR (-4)
#define ZO_BUF_ERROR (-5)
#define ZO_VERSION_ERROR (-6)

/* flush values: zlib-rs/src/lib.rs DeflateFlush / InflateFlush */
#define ZO_NO_FLUSH 0
#define ZO_PARTIAL_FLUSH 1
#define ZO_SYNC_FLUSH 2
#define ZO_FULL_FLUSH 3
#define ZO_FINISH 4
#define ZO_BLOCK 5
#define ZO_TREES 6

/* strategies: zlib-rs/src/deflate.rs:1377-1385 */
#define ZO_DEFAULT_STRATEGY 0
#define ZO_FILTERED 1
#define ZO_HUFFMAN_ONLY 2
#define ZO_RLE 3
#define ZO_FIXED 4

/* gzip header (zlib-rs/src/c_api.rs:170-199), write side only */
typedef struct zo_gz_header {
    int text;
    unsigned long time;
    int xflags;
    int os;
    uint8_t *extra;
    unsigned extra_len;
    unsigned extra_max;
    uint8_t *name;
    unsigned name_max;
    uint8_t *comment;
    unsigned comm_max;
    int hcrc;
    int done;
} zo_gz_header;

/* Same public fields as z_stream (zlib-rs/src/c_api.rs:56-71) minus the
 * allocator hooks: the oracle always uses malloc/free. */
typedef struct zo_stream {
    const uint8_t *next_in;
    uint32_t avail_in;
    uint64_t total_in;
    uint8_t *next_out;
    uint32_t avail_out;
    uint64_t total_out;
    const char *msg;
    void *state;
    int data_type;
    uint64_t adler;
} zo_stream;

/* ---- checksums ---- */
uint32_t zo_adler32(uint32_t adler, const uint8_t *buf, size_t len);
uint32_t zo_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2);
uint32_t zo_crc32(uint32_t crc, const uint8_t *buf, size_t len);
uint32_t zo_crc32_combine(uint32_t c1, uint32_t c2, uint64_t len2);
uint32_t zo_crc32_combine_gen(uint64_t len2);
uint32_t zo_crc32_combine_op(uint32_t c1, uint32_t c2, uint32_t op);

/* ---- deflate ---- */
int zo_deflate_init(zo_stream *strm, int level, int window_bits, int mem_level, int strategy);
int zo_deflate(zo_stream *strm, int flush);
int zo_deflate_end(zo_stream *strm);
int zo_deflate_reset(zo_stream *strm);
int zo_deflate_set_header(zo_stream *strm, zo_gz_header *head);
int zo_deflate_set_dictionary(zo_stream *strm, const uint8_t *dict, size_t len);
int zo_deflate_params(zo_stream *strm, int level, int strategy);
size_t zo_deflate_bound(zo_stream *strm, size_t source_len);
size_t zo_compress_bound(size_t source_len);
/* one-shot, mirrors deflate::compress_with_flush (zlib-rs/src/deflate.rs:2880) */
int zo_compress_ex(uint8_t *dest, size_t *dest_len, const uint8_t *src, size_t src_len,
                   int level, int window_bits, int mem_level, int strategy, int final_flush);
int zo_compress2(uint8_t *dest, size_t *dest_len, const uint8_t *src, size_t src_len, int level);

/* hash primitives exposed for the KATs (zlib-rs/src/deflate/hash_calc.rs:143-166) */
uint32_t zo_hash_standard(uint32_t le32);
uint32_t zo_hash_roll(uint32_t h, uint32_t byte);
void zo_slide_hash_chain(uint16_t *table, size_t n, uint16_t wsize);

/* Parse trace for tests of the GPU engine: receives every symbol the level-6
 * (deflate_medium) parser tallies, in order.  dist==0 => literal lc. */
typedef void (*zo_sym_trace_fn)(void *ctx, uint64_t abs_pos, unsigned dist, unsigned lc_or_len);
void zo_deflate_set_trace(zo_stream *strm, zo_sym_trace_fn fn, void *ctx);

/* ---- inflate ---- */
int zo_inflate_init(zo_stream *strm, int window_bits);
int zo_inflate(zo_stream *strm, int flush);
int zo_inflate_end(zo_stream *strm);
int zo_inflate_reset(zo_stream *strm);
int zo_inflate_set_dictionary(zo_stream *strm, const uint8_t *dict, size_t len);
/* one-shot, mirrors inflate::uncompress2 (zlib-rs/src/inflate.rs:195-277) */
int zo_uncompress2(uint8_t *dest, size_t *dest_len, const uint8_t *src, size_t *src_len);
int zo_uncompress(uint8_t *dest, size_t *dest_len, const uint8_t *src, size_t src_len);

#ifdef __cplusplus
}
#endif
#endif /* ZORACLE_H */
