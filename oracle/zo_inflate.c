/*
 * zo_inflate.c -- oracle (TEST INFRASTRUCTURE ONLY, see zoracle.h).
 *
 * Plain-C restatement of the reference's inflate: the Mode state machine
 * (zlib-rs/src/inflate.rs:896-1839 dispatch, :567-895 len_and_friends), the
 * call wrapper with its window / checksum epilogue (:2376-2457), table
 * construction (inflate/inftrees.rs:42-245), the 32 KiB ring window
 * (inflate/window.rs:95-168) and uncompress2 (:195-277).  Only the slow
 * (bit-at-a-time) decode path is restated; inflate_fast_help (:1880-2158) is a
 * result-identical accelerator of the same loop.
 */
#include "zoracle.h"
#include <stdlib.h>
#include <string.h>

#define MAX_WBITS 15
#define MIN_WBITS 8
#define ENOUGH_LENS 1332
#define ENOUGH_DISTS 592
#define WSIZE 32768u

typedef struct { uint8_t op, bits; uint16_t val; } code;

enum imode { HEAD, FLAGS, TIME, OS, EXLEN, EXTRA, NAME, COMMENT, HCRC, SYNC, MEM, LENGTH, TYPE, TYPEDO, STORED,
             COPYBLOCK, CHECK, LEN_, LEN, LIT, LENEXT, DIST, DISTEXT, MATCH, TABLE, LENLENS, CODELENS, DICTID, DICT,
             DONE, BAD };
enum { CT_CODES, CT_LENS, CT_DISTS };

typedef struct istate {
    int mode, flush;
    unsigned wrap, wbits;
    int last, have_dict, sane;
    int gzip_flags;
    size_t dmax;
    uint32_t checksum, crc_value;
    uint64_t total;
    /* bit reader (inflate/bitreader.rs) */
    uint64_t hold;
    unsigned bits;
    const uint8_t *in, *in_end;
    /* writer (inflate/writer.rs): output of this call */
    uint8_t *out;
    size_t out_len, out_cap, out_available;
    /* window (inflate/window.rs) */
    uint8_t *window;
    size_t whave, wnext;
    size_t length, offset, extra, back, was;
    size_t nlen, ndist, ncode, have, next;
    const code *lencode, *distcode;
    unsigned lenbits, distbits;
    uint16_t lens[320], work[288];
    code len_codes[ENOUGH_LENS], dist_codes[ENOUGH_DISTS], codes_codes[128];
    const char *error_message;
} istate;

static code lenfix[512], distfix[32];
static int fixed_ready;

static const uint16_t lbase[31] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258, 0, 0};
static const uint8_t lext[31] = {16, 16, 16, 16, 16, 16, 16, 16, 17, 17, 17, 17, 18, 18, 18, 18, 19, 19, 19, 19, 20, 20, 20, 20, 21, 21, 21, 21, 16, 77, 202};
static const uint16_t dbase[32] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577, 0, 0};
static const uint8_t dext[32] = {16, 16, 16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 21, 21, 22, 22, 23, 23, 24, 24, 25, 25, 26, 26, 27, 27, 28, 28, 29, 29, 64, 64};

static uint32_t rev32(uint32_t v)
{
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
}

/* inflate/inftrees.rs:42-245.  Returns 0 on success. */
static int inflate_table(int type, const uint16_t *lens, size_t codes, code *table, unsigned bits, uint16_t *work,
                         unsigned *root_out, size_t *used_out)
{
    uint16_t count[16] = {0}, offs[16] = {0};
    unsigned min = 15, max = 0, len, root, curr, drop;
    for (size_t i = 0; i < codes; i++)
        if (lens[i]) { count[lens[i]]++; if (lens[i] > max) max = lens[i]; if (lens[i] < min) min = lens[i]; }
    if (max == 0) {
        code c = {64, 1, 0};
        table[0] = table[1] = c;
        *root_out = 1; *used_out = 2;
        return 0;
    }
    root = bits < min ? min : bits > max ? max : bits;
    int64_t left = 1;
    for (len = 1; len <= 15; len++) { left = (left << 1) - count[len]; if (left < 0) return -1; }
    if (left > 0 && (type == CT_CODES || max != 1)) return -1;
    for (len = 1; len < 15; len++) offs[len + 1] = (uint16_t)(offs[len] + count[len]);
    for (size_t sym = 0; sym < codes; sym++) if (lens[sym]) work[offs[lens[sym]]++] = (uint16_t)sym;
    const uint16_t *base; const uint8_t *extra; unsigned match;
    if (type == CT_CODES) { base = NULL; extra = NULL; match = 20; }
    else if (type == CT_LENS) { base = lbase; extra = lext; match = 257; }
    else { base = dbase; extra = dext; match = 0; }
    size_t used = (size_t)1 << root;
    if ((type == CT_LENS && used > ENOUGH_LENS) || (type == CT_DISTS && used > ENOUGH_DISTS)) return 1;
    size_t huff = 0, next = 0, low = (size_t)-1, mask = used - 1, sym = 0;
    uint32_t rhuff = 0;
    len = min; curr = root; drop = 0;
    for (;;) {
        code here;
        here.bits = (uint8_t)(len - drop);
        if (work[sym] >= match) { here.op = extra[work[sym] - match]; here.val = base[work[sym] - match]; }
        else if ((unsigned)work[sym] + 1 < match) { here.op = 0; here.val = work[sym]; }
        else { here.op = 96; here.val = 0; }
        size_t incr = (size_t)1 << (len - drop), fill = (size_t)1 << curr, mn = fill;
        do { fill -= incr; table[next + (huff >> drop) + fill] = here; } while (fill != 0);
        rhuff += 0x80000000u >> (len - 1);
        huff = rev32(rhuff);
        sym++;
        if (--count[len] == 0) {
            if (len == max) break;
            len = lens[work[sym]];
        }
        if (len > root && (huff & mask) != low) {
            if (drop == 0) drop = root;
            next += mn;
            curr = len - drop;
            int l2 = 1 << curr;
            while (curr + drop < max) {
                l2 -= count[curr + drop];
                if (l2 <= 0) break;
                curr++;
                l2 <<= 1;
            }
            used += (size_t)1 << curr;
            if ((type == CT_LENS && used > ENOUGH_LENS) || (type == CT_DISTS && used > ENOUGH_DISTS)) return 1;
            low = huff & mask;
            table[low].op = (uint8_t)curr;
            table[low].bits = (uint8_t)root;
            table[low].val = (uint16_t)next;
        }
    }
    if (huff != 0) { code h = {64, (uint8_t)(len - drop), 0}; table[next + huff] = h; }
    *root_out = root; *used_out = used;
    return 0;
}

static void fixed_tables(void)
{
    if (fixed_ready) return;
    uint16_t lens[288], work[288];
    unsigned root; size_t used;
    size_t sym = 0;
    while (sym < 144) lens[sym++] = 8;
    while (sym < 256) lens[sym++] = 9;
    while (sym < 280) lens[sym++] = 7;
    while (sym < 288) lens[sym++] = 8;
    inflate_table(CT_LENS, lens, 288, lenfix, 9, work, &root, &used);
    for (sym = 0; sym < 32; sym++) lens[sym] = 5;
    inflate_table(CT_DISTS, lens, 32, distfix, 5, work, &root, &used);
    fixed_ready = 1;
}

/* window.rs:95-168 (always a 32 KiB ring, inflate.rs:2266-2269) */
static void window_extend(istate *s, const uint8_t *p, size_t len, int update_checksum)
{
    if (update_checksum) {
        if (s->gzip_flags != 0) s->crc_value = zo_crc32(s->crc_value, p, len);
        else s->checksum = zo_adler32(s->checksum, p, len);
    }
    if (len >= WSIZE) {
        memcpy(s->window, p + (len - WSIZE), WSIZE);
        s->wnext = 0;
        s->whave = WSIZE;
        return;
    }
    size_t dist = WSIZE - s->wnext < len ? WSIZE - s->wnext : len;
    memcpy(s->window + s->wnext, p, dist);
    if (len > dist) {
        memcpy(s->window, p + dist, len - dist);
        s->wnext = len - dist;
        s->whave = WSIZE;
    } else {
        s->wnext += dist;
        if (s->wnext == WSIZE) s->wnext = 0;
        if (s->whave < WSIZE) s->whave += dist;
    }
}

static int reset_keep(zo_stream *strm)
{
    istate *s = (istate *)strm->state;
    strm->total_in = strm->total_out = 0;
    s->total = 0;
    strm->msg = NULL;
    if (s->wrap) strm->adler = s->wrap & 1;
    s->mode = HEAD;
    s->checksum = 1;
    s->last = 0; s->have_dict = 0; s->sane = 1;
    s->gzip_flags = -1;
    s->dmax = 32768;
    s->hold = 0; s->bits = 0;
    s->next = 0;
    s->lencode = lenfix; s->distcode = distfix; s->lenbits = 0; s->distbits = 0;
    s->back = (size_t)-1;
    return ZO_OK;
}

int zo_inflate_reset(zo_stream *strm)
{
    if (!strm || !strm->state) return ZO_STREAM_ERROR;
    istate *s = (istate *)strm->state;
    s->whave = s->wnext = 0;
    s->error_message = NULL;
    return reset_keep(strm);
}

int zo_inflate_init(zo_stream *strm, int window_bits) /* inflate.rs:2233-2327 */
{
    if (!strm) return ZO_STREAM_ERROR;
    fixed_tables();
    strm->msg = NULL;
    int wrap;
    if (window_bits < 0) {
        wrap = 0;
        if (window_bits < -MAX_WBITS) return ZO_STREAM_ERROR;
        window_bits = -window_bits;
    } else {
        wrap = (window_bits >> 4) + 5;
        if (window_bits < 48) window_bits &= MAX_WBITS;
    }
    if (window_bits != 0 && (window_bits < MIN_WBITS || window_bits > MAX_WBITS)) return ZO_STREAM_ERROR;
    istate *s = (istate *)calloc(1, sizeof(istate));
    if (!s) return ZO_MEM_ERROR;
    s->window = (uint8_t *)calloc(WSIZE + 64, 1);
    if (!s->window) { free(s); return ZO_MEM_ERROR; }
    s->wrap = (unsigned)wrap & 0xff;
    s->wbits = (unsigned)window_bits;
    strm->state = s;
    return zo_inflate_reset(strm);
}

int zo_inflate_end(zo_stream *strm)
{
    if (!strm || !strm->state) return ZO_STREAM_ERROR;
    istate *s = (istate *)strm->state;
    free(s->window);
    free(s);
    strm->state = NULL;
    return ZO_OK;
}

int zo_inflate_set_dictionary(zo_stream *strm, const uint8_t *dict, size_t len) /* :2611-2640 */
{
    if (!strm || !strm->state) return ZO_STREAM_ERROR;
    istate *s = (istate *)strm->state;
    if (s->wrap != 0 && s->mode != DICT) return ZO_STREAM_ERROR;
    if (s->mode == DICT && zo_adler32(1, dict, len) != s->checksum) return ZO_DATA_ERROR;
    window_extend(s, dict, len, 0);
    s->have_dict = 1;
    return ZO_OK;
}

#define BITS(n) (s->hold & (((uint64_t)1 << (n)) - 1))
#define DROPBITS(n) do { s->hold >>= (n); s->bits -= (unsigned)(n); } while (0)
#define INITBITS() do { s->hold = 0; s->bits = 0; } while (0)
#define PULLBYTE() do { if (s->in == s->in_end) return ZO_OK; s->hold |= (uint64_t)*s->in++ << s->bits; s->bits += 8; } while (0)
#define NEEDBITS(n) do { while (s->bits < (unsigned)(n)) PULLBYTE(); } while (0)
#define BADMSG(m) do { s->mode = BAD; s->error_message = (m); return ZO_DATA_ERROR; } while (0)
#define CRC2(v) do { uint8_t b_[2] = {(uint8_t)(v), (uint8_t)((v) >> 8)}; s->checksum = zo_crc32(s->checksum, b_, 2); } while (0)
#define HCRC_ON() ((s->gzip_flags & 0x0200) && (s->wrap & 4))

static uint32_t zswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00) | ((v << 8) & 0xff0000) | (v << 24); }

static int dispatch(istate *s)
{
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    for (;;) {
        switch (s->mode) {
        case HEAD:
            if (s->wrap == 0) { s->mode = TYPEDO; break; }
            NEEDBITS(16);
            if ((s->wrap & 2) && s->hold == 0x8b1f) {
                if (s->wbits == 0) s->wbits = 15;
                s->checksum = 0;
                CRC2(s->hold);
                INITBITS();
                s->mode = FLAGS;
                break;
            }
            if (!(s->wrap & 1) || ((BITS(8) << 8) + (s->hold >> 8)) % 31) BADMSG("incorrect header check");
            if (BITS(4) != 8) BADMSG("unknown compression method");
            DROPBITS(4);
            {
                unsigned len = (unsigned)BITS(4) + 8;
                if (s->wbits == 0) s->wbits = len;
                if (len > 15 || len > s->wbits) BADMSG("invalid window size");
                s->dmax = (size_t)1 << len;
            }
            s->gzip_flags = 0;
            s->checksum = 1;
            s->mode = (s->hold & 0x200) ? DICTID : TYPE;
            INITBITS();
            break;
        case FLAGS:
            NEEDBITS(16);
            s->gzip_flags = (int)s->hold;
            if ((s->gzip_flags & 0xff) != 8) BADMSG("unknown compression method");
            if (s->gzip_flags & 0xe000) BADMSG("unknown header flags set");
            if (HCRC_ON()) CRC2(s->hold);
            INITBITS();
            s->mode = TIME;
            break;
        case TIME:
            NEEDBITS(32);
            if (HCRC_ON()) { CRC2(s->hold); CRC2(s->hold >> 16); }
            INITBITS();
            s->mode = OS;
            break;
        case OS:
            NEEDBITS(16);
            if (HCRC_ON()) CRC2(s->hold);
            INITBITS();
            s->mode = EXLEN;
            break;
        case EXLEN:
            if (s->gzip_flags & 0x0400) {
                NEEDBITS(16);
                s->length = (size_t)s->hold;
                if (HCRC_ON()) CRC2(s->hold);
                INITBITS();
            }
            s->mode = EXTRA;
            break;
        case EXTRA:
            if (s->gzip_flags & 0x0400) {
                size_t avail = (size_t)(s->in_end - s->in);
                size_t copy = s->length < avail ? s->length : avail;
                if (copy) {
                    if (HCRC_ON()) s->checksum = zo_crc32(s->checksum, s->in, copy);
                    s->in += copy;
                    s->length -= copy;
                }
                if (s->length) return ZO_OK;
            }
            s->length = 0;
            s->mode = NAME;
            break;
        case NAME:
        case COMMENT: {
            int flag = s->mode == NAME ? 0x0800 : 0x1000;
            if (s->gzip_flags & flag) {
                if (s->in == s->in_end) return ZO_OK;
                const uint8_t *z = (const uint8_t *)memchr(s->in, 0, (size_t)(s->in_end - s->in));
                size_t n = z ? (size_t)(z - s->in) + 1 : (size_t)(s->in_end - s->in);
                if (HCRC_ON()) s->checksum = zo_crc32(s->checksum, s->in, n);
                s->in += n;
                if (!z && s->in == s->in_end) return ZO_OK;
            }
            s->length = 0;
            s->mode = s->mode == NAME ? COMMENT : HCRC;
            break;
        }
        case HCRC:
            if (s->gzip_flags & 0x0200) {
                NEEDBITS(16);
                if ((s->wrap & 4) && (uint32_t)s->hold != (s->checksum & 0xffff)) BADMSG("header crc mismatch");
                INITBITS();
            }
            if ((s->wrap & 4) && s->gzip_flags != 0) { s->crc_value = 0; s->checksum = 0; }
            s->mode = TYPE;
            break;
        case DICTID:
            NEEDBITS(32);
            s->checksum = zswap32((uint32_t)s->hold);
            INITBITS();
            s->mode = DICT;
            break;
        case DICT:
            if (!s->have_dict) return ZO_NEED_DICT;
            s->checksum = 1;
            s->mode = TYPE;
            break;
        case TYPE:
            if (s->flush == ZO_BLOCK || s->flush == ZO_TREES) return ZO_OK;
            s->mode = TYPEDO;
            break;
        case TYPEDO:
            if (s->last) {
                DROPBITS(s->bits & 7);
                s->mode = CHECK;
                break;
            }
            NEEDBITS(3);
            s->last = (int)BITS(1);
            DROPBITS(1);
            switch (BITS(2)) {
            case 0: DROPBITS(2); s->mode = STORED; break;
            case 1:
                s->lencode = lenfix; s->lenbits = 9; s->distcode = distfix; s->distbits = 5;
                s->mode = LEN_;
                DROPBITS(2);
                if (s->flush == ZO_TREES) return ZO_OK;
                break;
            case 2: DROPBITS(2); s->mode = TABLE; break;
            default: DROPBITS(2); BADMSG("invalid block type");
            }
            break;
        case STORED:
            DROPBITS(s->bits & 7);
            NEEDBITS(32);
            {
                uint32_t h = (uint32_t)BITS(32);
                if ((uint16_t)h != (uint16_t)~(h >> 16)) BADMSG("invalid stored block lengths");
                s->length = h & 0xffff;
            }
            INITBITS();
            s->mode = COPYBLOCK;
            if (s->flush == ZO_TREES) return ZO_OK;
            break;
        case COPYBLOCK:
            while (s->length) {
                size_t copy = s->length;
                if (copy > s->out_cap - s->out_len) copy = s->out_cap - s->out_len;
                if (copy > (size_t)(s->in_end - s->in)) copy = (size_t)(s->in_end - s->in);
                if (copy == 0) return ZO_OK;
                memcpy(s->out + s->out_len, s->in, copy);
                s->out_len += copy;
                s->in += copy;
                s->length -= copy;
            }
            s->mode = TYPE;
            break;
        case TABLE:
            NEEDBITS(14);
            s->nlen = (size_t)BITS(5) + 257; DROPBITS(5);
            s->ndist = (size_t)BITS(5) + 1; DROPBITS(5);
            s->ncode = (size_t)BITS(4) + 4; DROPBITS(4);
            if (s->nlen > 286 || s->ndist > 30) BADMSG("too many length or distance symbols");
            s->have = 0;
            s->mode = LENLENS;
            break;
        case LENLENS: {
            while (s->have < s->ncode) { NEEDBITS(3); s->lens[order[s->have++]] = (uint16_t)BITS(3); DROPBITS(3); }
            while (s->have < 19) s->lens[order[s->have++]] = 0;
            unsigned root; size_t used;
            if (inflate_table(CT_CODES, s->lens, 19, s->codes_codes, 7, s->work, &root, &used)) BADMSG("invalid code lengths set");
            s->next = used;
            s->lencode = s->codes_codes; s->lenbits = root;
            s->have = 0;
            s->mode = CODELENS;
            break;
        }
        case CODELENS: {
            while (s->have < s->nlen + s->ndist) {
                code here;
                for (;;) { here = s->lencode[BITS(s->lenbits)]; if (here.bits <= s->bits) break; PULLBYTE(); }
                if (here.val < 16) { DROPBITS(here.bits); s->lens[s->have++] = here.val; continue; }
                unsigned len = 0; size_t copy;
                if (here.val == 16) {
                    NEEDBITS(here.bits + 2); DROPBITS(here.bits);
                    if (s->have == 0) BADMSG("invalid bit length repeat");
                    len = s->lens[s->have - 1];
                    copy = 3 + (size_t)BITS(2); DROPBITS(2);
                } else if (here.val == 17) {
                    NEEDBITS(here.bits + 3); DROPBITS(here.bits);
                    copy = 3 + (size_t)BITS(3); DROPBITS(3);
                } else {
                    NEEDBITS(here.bits + 7); DROPBITS(here.bits);
                    copy = 11 + (size_t)BITS(7); DROPBITS(7);
                }
                if (s->have + copy > s->nlen + s->ndist) BADMSG("invalid bit length repeat");
                while (copy--) s->lens[s->have++] = (uint16_t)len;
            }
            if (s->lens[256] == 0) BADMSG("invalid code -- missing end-of-block");
            unsigned root; size_t used;
            if (inflate_table(CT_LENS, s->lens, s->nlen, s->len_codes, 10, s->work, &root, &used)) BADMSG("invalid literal/lengths set");
            s->lencode = s->len_codes; s->lenbits = root; s->next = used;
            if (inflate_table(CT_DISTS, s->lens + s->nlen, s->ndist, s->dist_codes, 9, s->work, &root, &used)) BADMSG("invalid distances set");
            s->distcode = s->dist_codes; s->distbits = root; s->next += used;
            s->mode = LEN_;
            if (s->flush == ZO_TREES) return ZO_OK;
            break;
        }
        case LEN_:
            s->mode = LEN;
            break;
        case LEN: {
            s->back = 0;
            code here, last;
            for (;;) { here = s->lencode[BITS(s->lenbits)]; if (here.bits <= s->bits) break; PULLBYTE(); }
            if (here.op && (here.op & 0xf0) == 0) {
                last = here;
                for (;;) {
                    here = s->lencode[last.val + (BITS(last.bits + last.op) >> last.bits)];
                    if ((unsigned)last.bits + here.bits <= s->bits) break;
                    PULLBYTE();
                }
                DROPBITS(last.bits);
                s->back += last.bits;
            }
            DROPBITS(here.bits);
            s->back += here.bits;
            s->length = here.val;
            if (here.op == 0) { s->mode = LIT; break; }
            if (here.op & 32) { s->back = (size_t)-1; s->mode = TYPE; break; }
            if (here.op & 64) BADMSG("invalid literal/length code");
            s->extra = here.op & 15;
            s->mode = LENEXT;
            break;
        }
        case LIT:
            if (s->out_len == s->out_cap) return ZO_OK;
            s->out[s->out_len++] = (uint8_t)s->length;
            s->mode = LEN;
            break;
        case LENEXT:
            if (s->extra) { NEEDBITS(s->extra); s->length += (size_t)BITS(s->extra); DROPBITS(s->extra); s->back += s->extra; }
            s->was = s->length;
            s->mode = DIST;
            break;
        case DIST: {
            code here, last;
            for (;;) { here = s->distcode[BITS(s->distbits)]; if (here.bits <= s->bits) break; PULLBYTE(); }
            if ((here.op & 0xf0) == 0) {
                last = here;
                for (;;) {
                    here = s->distcode[last.val + (BITS(last.bits + last.op) >> last.bits)];
                    if ((unsigned)last.bits + here.bits <= s->bits) break;
                    PULLBYTE();
                }
                DROPBITS(last.bits);
                s->back += last.bits;
            }
            DROPBITS(here.bits);
            if (here.op & 64) BADMSG("invalid distance code");
            s->offset = here.val;
            s->extra = here.op & 15;
            s->mode = DISTEXT;
            break;
        }
        case DISTEXT:
            if (s->extra) { NEEDBITS(s->extra); s->offset += (size_t)BITS(s->extra); DROPBITS(s->extra); s->back += s->extra; }
            s->mode = MATCH;
            break;
        case MATCH:
            while (s->length) {
                if (s->out_len == s->out_cap) return ZO_OK;
                size_t left = s->out_cap - s->out_len, copy = s->out_len;
                if (s->offset > copy) {
                    copy = s->offset - copy;
                    if (copy > s->whave) {
                        if (s->sane) BADMSG("invalid distance too far back");
                        return ZO_DATA_ERROR;
                    }
                    size_t from;
                    if (copy > s->wnext) { copy -= s->wnext; from = WSIZE - copy; }
                    else from = s->wnext - copy;
                    if (copy > s->length) copy = s->length;
                    if (copy > left) copy = left;
                    memcpy(s->out + s->out_len, s->window + from, copy);
                    s->out_len += copy;
                } else {
                    copy = s->length < left ? s->length : left;
                    const uint8_t *src = s->out + s->out_len - s->offset;
                    uint8_t *dst = s->out + s->out_len;
                    for (size_t i = 0; i < copy; i++) dst[i] = src[i];
                    s->out_len += copy;
                }
                s->length -= copy;
            }
            s->mode = LEN;
            break;
        case CHECK:
            if (s->wrap) {
                NEEDBITS(32);
                s->total += s->out_len;
                if (s->wrap & 4) {
                    if (s->gzip_flags != 0) { s->crc_value = zo_crc32(s->crc_value, s->out, s->out_len); s->checksum = s->crc_value; }
                    else s->checksum = zo_adler32(s->checksum, s->out, s->out_len);
                }
                uint32_t given = s->gzip_flags != 0 ? (uint32_t)s->hold : zswap32((uint32_t)s->hold);
                s->out_available = s->out_cap - s->out_len;
                if ((s->wrap & 4) && given != s->checksum) BADMSG("incorrect data check");
                INITBITS();
            }
            s->mode = LENGTH;
            break;
        case LENGTH:
            if (s->wrap && s->gzip_flags != 0) {
                NEEDBITS(32);
                if ((s->wrap & 4) && (uint32_t)s->hold != (uint32_t)s->total) BADMSG("incorrect length check");
                INITBITS();
            }
            s->mode = DONE;
            return ZO_STREAM_END;
        case DONE: return ZO_STREAM_END;
        case BAD: s->error_message = "repeated call with bad state"; return ZO_DATA_ERROR;
        case MEM: return ZO_MEM_ERROR;
        default: return ZO_STREAM_ERROR;
        }
    }
}

int zo_inflate(zo_stream *strm, int flush) /* inflate.rs:2376-2457 */
{
    if (!strm || !strm->state) return ZO_STREAM_ERROR;
    istate *s = (istate *)strm->state;
    if (strm->next_out == NULL || (strm->next_in == NULL && strm->avail_in != 0)) return ZO_STREAM_ERROR;
    if (s->mode == TYPE) s->mode = TYPEDO;
    s->flush = flush;
    s->in = strm->next_in;
    s->in_end = strm->next_in + strm->avail_in;
    s->out = strm->next_out;
    s->out_len = 0;
    s->out_cap = strm->avail_out;
    s->out_available = strm->avail_out;
    int err = dispatch(s);
    size_t in_read = (size_t)(s->in - strm->next_in);
    size_t out_written = s->out_available - (s->out_cap - s->out_len);
    strm->total_in += in_read;
    s->total += out_written;
    strm->total_out = s->total;
    strm->avail_in = (uint32_t)(s->in_end - s->in);
    strm->next_in = s->in;
    strm->avail_out = (uint32_t)(s->out_cap - s->out_len);
    strm->next_out = s->out + s->out_len;
    strm->adler = s->checksum;
    window_extend(s, s->out, out_written, (s->wrap & 4) != 0);
    if (s->error_message) strm->msg = s->error_message;
    strm->data_type = (int)s->bits | (s->last ? 64 : 0) | (s->mode == TYPE ? 128 : (s->mode == LEN_ || s->mode == COPYBLOCK) ? 256 : 0);
    if (((in_read == 0 && out_written == 0) || flush == ZO_FINISH) && err == ZO_OK) return ZO_BUF_ERROR;
    return err;
}

int zo_uncompress2(uint8_t *dest, size_t *dest_len, const uint8_t *src, size_t *src_len) /* :195-277 */
{
    uint8_t buf[1];
    uint64_t left, len = *src_len;
    size_t out_cap = *dest_len, result_len = out_cap;
    uint8_t *d;
    if (out_cap == 0) { left = 1; d = buf; } else { left = out_cap; result_len = 0; d = dest; }
    zo_stream strm;
    memset(&strm, 0, sizeof strm);
    strm.next_in = src;
    int err = zo_inflate_init(&strm, 15);
    if (err != ZO_OK) return err;
    strm.next_out = d;
    for (;;) {
        if (strm.avail_out == 0) { strm.avail_out = (uint32_t)(left < 0xffffffffu ? left : 0xffffffffu); left -= strm.avail_out; }
        if (strm.avail_in == 0) { strm.avail_in = (uint32_t)(len < 0xffffffffu ? len : 0xffffffffu); len -= strm.avail_in; }
        err = zo_inflate(&strm, ZO_NO_FLUSH);
        if (err != ZO_OK) break;
    }
    *src_len -= (size_t)(len + strm.avail_in);
    if (out_cap != 0) result_len = (size_t)strm.total_out;
    else if (strm.total_out != 0 && err == ZO_BUF_ERROR) left = 1;
    uint32_t avail_out = strm.avail_out;
    zo_inflate_end(&strm);
    *dest_len = result_len;
    if (err == ZO_STREAM_END) return ZO_OK;
    if (err == ZO_NEED_DICT) return ZO_DATA_ERROR;
    if (err == ZO_BUF_ERROR && (left + avail_out) != 0) return ZO_DATA_ERROR;
    return err;
}

int zo_uncompress(uint8_t *dest, size_t *dest_len, const uint8_t *src, size_t src_len)
{
    return zo_uncompress2(dest, dest_len, src, &src_len);
}
