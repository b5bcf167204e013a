/* A plain C client of the zlib ABI of libz_b200.so, written the way zlib's own example clients are (cf. the reference's
 * libz-rs-sys-cdylib/zpipe.c:63-145 call sequence): deflateInit / deflate(Z_NO_FLUSH ... Z_FINISH) / deflateEnd over 16 KiB
 * chunks, then inflateInit / inflate / inflateEnd, compares the round trip and prints the compressed size and adler.
 * TEST INFRASTRUCTURE: proves that a C program including the header and linking -lz_b200 needs nothing else.
 * usage: pipe_client <file> [level]     exit 0 ok, 77 no CUDA device (Z_MEM_ERROR from deflateInit), 1 failure */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "zlib_b200.h"

#define CHUNK 16384

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s file [level]\n", argv[0]); return 1; }
    int level = argc > 2 ? atoi(argv[2]) : Z_DEFAULT_COMPRESSION;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 1; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char *src = malloc(n ? n : 1), *comp = malloc(compressBound(n) + 64), *back = malloc(n ? n : 1);
    if (fread(src, 1, n, f) != (size_t)n) { fprintf(stderr, "short read\n"); return 1; }
    fclose(f);

    z_stream d;
    memset(&d, 0, sizeof d);
    int rc = deflateInit(&d, level);
    if (rc == Z_MEM_ERROR) { fprintf(stderr, "deflateInit: Z_MEM_ERROR (%s)\n", d.msg ? d.msg : ""); return 77; }
    if (rc != Z_OK) { fprintf(stderr, "deflateInit rc=%d\n", rc); return 1; }
    size_t clen = 0;
    long pos = 0;
    unsigned char out[CHUNK];
    do {
        long take = n - pos < CHUNK ? n - pos : CHUNK;
        d.next_in = src + pos;
        d.avail_in = (uInt)take;
        pos += take;
        int flush = pos >= n ? Z_FINISH : Z_NO_FLUSH;
        do {
            d.next_out = out;
            d.avail_out = CHUNK;
            rc = deflate(&d, flush);
            if (rc == Z_STREAM_ERROR) { fprintf(stderr, "deflate rc=%d\n", rc); return 1; }
            size_t have = CHUNK - d.avail_out;
            memcpy(comp + clen, out, have);
            clen += have;
        } while (d.avail_out == 0);
    } while (pos < n || rc != Z_STREAM_END);
    uLong adler = d.adler;
    if (deflateEnd(&d) != Z_OK) { fprintf(stderr, "deflateEnd\n"); return 1; }

    z_stream i;
    memset(&i, 0, sizeof i);
    if (inflateInit(&i) != Z_OK) { fprintf(stderr, "inflateInit\n"); return 1; }
    size_t blen = 0, cpos = 0;
    do { /* zpipe's inf(): feed a chunk, drain the output, until Z_STREAM_END */
        size_t take = clen - cpos < CHUNK ? clen - cpos : CHUNK;
        if (take == 0) break;
        i.next_in = comp + cpos;
        i.avail_in = (uInt)take;
        cpos += take;
        do {
            i.next_out = out;
            i.avail_out = CHUNK;
            rc = inflate(&i, Z_NO_FLUSH);
            if (rc == Z_NEED_DICT || rc == Z_DATA_ERROR || rc == Z_MEM_ERROR || rc == Z_STREAM_ERROR) {
                fprintf(stderr, "inflate rc=%d %s\n", rc, i.msg ? i.msg : "");
                return 1;
            }
            size_t have = CHUNK - i.avail_out;
            if (blen + have > (size_t)n) { fprintf(stderr, "inflate produced too much\n"); return 1; }
            memcpy(back + blen, out, have);
            blen += have;
        } while (i.avail_out == 0);
    } while (rc != Z_STREAM_END);
    if (rc != Z_STREAM_END) { fprintf(stderr, "inflate ended with rc=%d\n", rc); return 1; }
    inflateEnd(&i);
    if (blen != (size_t)n || memcmp(back, src, n) != 0) { fprintf(stderr, "round trip mismatch (%zu vs %ld)\n", blen, n); return 1; }
    /* one-shot entry points too */
    uLongf c2 = compressBound(n) + 64;
    if (compress2(comp, &c2, src, n, level) != Z_OK || c2 != clen) { fprintf(stderr, "compress2 differs from the streamed result (%lu vs %zu)\n", (unsigned long)c2, clen); return 1; }
    uLongf u2 = n;
    if (uncompress(back, &u2, comp, c2) != Z_OK || u2 != (uLongf)n) { fprintf(stderr, "uncompress\n"); return 1; }
    printf("ok in=%ld out=%zu adler=%08lx crc=%08lx version=%s\n", n, clen, adler, crc32(0, src, (uInt)n), zlibVersion());
    return 0;
}
