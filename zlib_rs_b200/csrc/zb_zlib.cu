// zb_zlib.cu -- the zlib C ABI of libz-rs-sys (libz-rs-sys/src/lib.rs:149-2332) on top of the GPU engine.
//
// Host code here is framing and bookkeeping only: argument checks with the reference's return codes,
// buffering of streamed input until a flush point, draining of engine output through next_out/avail_out
// (the role of flush_pending, zlib-rs/src/deflate.rs:2805-2826), and O(log n) checksum-combine algebra.
// All compression, decompression and checksumming is done by the kernels; there is no CPU fallback --
// without a CUDA device deflateInit*/inflateInit* return Z_MEM_ERROR with msg "no CUDA device" and the
// checksum entry points abort loudly.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>
#include "../../include/zb_engine.h"
#include "../../include/zlib_b200.h"

namespace {

thread_local zb_engine *t_engine = nullptr;
thread_local int t_engine_err = 0;

zb_engine *engine()
{
    if (!t_engine) {
        int dev = 0;
        if (const char *s = getenv("ZB_DEVICE")) dev = atoi(s);
        else if (const char *lr = getenv("LOCAL_RANK")) dev = atoi(lr) % (zb_device_count() > 0 ? zb_device_count() : 1);
        t_engine = zb_engine_create(dev, &t_engine_err);
    }
    return t_engine;
}

const char *const kNoDevice = "no CUDA device";

const char *err_msg(int rc)
{
    // zlib-rs/src/lib.rs:252-264
    switch (rc) {
    case Z_OK: return "";
    case Z_STREAM_END: return "stream end";
    case Z_NEED_DICT: return "need dictionary";
    case Z_ERRNO: return "file error";
    case Z_STREAM_ERROR: return "stream error";
    case Z_DATA_ERROR: return "data error";
    case Z_MEM_ERROR: return "insufficient memory";
    case Z_BUF_ERROR: return "buffer error";
    case Z_VERSION_ERROR: return "incompatible version";
    default: return "";
    }
}

int version_ok(const char *version, int stream_size)
{
    // libz-rs-sys/src/lib.rs:2133-2145
    return version && version[0] == '1' && stream_size == (int)sizeof(z_stream);
}

constexpr uint32_t kDMagic = 0x7a62444cu, kIMagic = 0x7a62494cu;

struct DState {
    uint32_t magic;
    int level, strategy, window_bits, mem_level, wrap;
    int status;       // 1 init, 2 busy, 3 finished
    int last_flush;
    bool header_done; // segment mode: framing header already emitted
    bool any_segment; // a non-final flush was executed: the stream is built from raw segments
    uint32_t check;   // running adler32 / crc32 over consumed segments
    uint64_t check_len;
    std::vector<uint8_t> in, out;
    size_t out_pos;
    gz_header *gzhead;
};

struct IState {
    uint32_t magic;
    int window_bits;
    int status; // 0 collecting, 1 output ready, 2 done, -1 bad
    std::vector<uint8_t> in, out;
    size_t out_pos;
    size_t last_attempt;
    int result;
    char msg[64];
    uint64_t consumed;
};

DState *dstate(z_streamp s)
{
    if (!s || !s->state) return nullptr;
    DState *d = reinterpret_cast<DState *>(s->state);
    return d->magic == kDMagic ? d : nullptr;
}
IState *istate(z_streamp s)
{
    if (!s || !s->state) return nullptr;
    IState *d = reinterpret_cast<IState *>(s->state);
    return d->magic == kIMagic ? d : nullptr;
}

voidpf default_alloc(voidpf, uInt items, uInt size) { return malloc((size_t)items * size); }
void default_free(voidpf, voidpf p) { free(p); }

int rank_flush(int f) { return f * 2 - (f > 4 ? 9 : 0); } // deflate.rs:1667-1670

// adler32_combine (adler32.rs:58-87) / crc32 combine (crc32/combine.rs:3-61): O(log n) host algebra
uint32_t adler_combine(uint32_t a1, uint32_t a2, uint64_t len2)
{
    const uint64_t BASE = 65521;
    uint64_t rem = len2 % BASE, sum1 = a1 & 0xffff, sum2 = (rem * sum1) % BASE;
    sum1 += (a2 & 0xffff) + BASE - 1;
    sum2 += ((a1 >> 16) & 0xffff) + ((a2 >> 16) & 0xffff) + BASE - rem;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum2 >= (BASE << 1)) sum2 -= (BASE << 1);
    if (sum2 >= BASE) sum2 -= BASE;
    return (uint32_t)(sum1 | (sum2 << 16));
}
uint32_t multmodp(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ 0xedb88320u : b >> 1;
    }
    return p;
}
uint32_t x2nmodp(uint64_t n, uint32_t k)
{
    uint32_t p = 1u << 31, sq = 1u << 30;
    for (uint32_t i = 0; i < k; i++) sq = multmodp(sq, sq);
    while (n) { if (n & 1) p = multmodp(sq, p); n >>= 1; if (n) sq = multmodp(sq, sq); }
    return p;
}

size_t drain(z_streamp strm, std::vector<uint8_t> &out, size_t &pos)
{
    size_t n = out.size() - pos;
    if (n > strm->avail_out) n = strm->avail_out;
    if (n) {
        memcpy(strm->next_out, out.data() + pos, n);
        strm->next_out += n;
        strm->avail_out -= (uInt)n;
        strm->total_out += n;
        pos += n;
    }
    if (pos == out.size()) { out.clear(); pos = 0; }
    return n;
}

int map_rc(int rc)
{
    switch (rc) {
    case ZB_OK: return Z_OK;
    case ZB_E_BUF: return Z_BUF_ERROR;
    case ZB_E_MEM: return Z_MEM_ERROR;
    case ZB_E_PARAM: return Z_STREAM_ERROR;
    case ZB_E_DATA: return Z_DATA_ERROR;
    default: return Z_STREAM_ERROR; // CUDA / internal failure: never abort across the ABI
    }
}

} // namespace

extern "C" {

const char *zlibVersion(void) { return ZLIB_VERSION; }
const char *zError(int err) { return err_msg(err); }
uLong zlibCompileFlags(void) { return (sizeof(uInt) == 4 ? 1 : 0) | (sizeof(uLong) == 8 ? 2 << 2 : 1 << 2) | (2 << 4) | (2 << 6); }

// ---------------------------------------------------------------- deflate
int deflateInit2_(z_streamp strm, int level, int method, int windowBits, int memLevel, int strategy, const char *version,
                  int stream_size)
{
    if (!version_ok(version, stream_size)) return Z_VERSION_ERROR;
    if (!strm) return Z_STREAM_ERROR;
    strm->msg = nullptr;
    if (!strm->zalloc) { strm->zalloc = default_alloc; strm->opaque = nullptr; }
    if (!strm->zfree) strm->zfree = default_free;
    // parameter checks of deflate::init (zlib-rs/src/deflate.rs:282-312)
    if (level == Z_DEFAULT_COMPRESSION) level = 6;
    int wrap = 1, wb = windowBits;
    if (wb < 0) { if (wb < -MAX_WBITS) return Z_STREAM_ERROR; wrap = 0; wb = -wb; }
    else if (wb > MAX_WBITS) { wrap = 2; wb -= 16; }
    if (memLevel < 1 || memLevel > MAX_MEM_LEVEL || method != Z_DEFLATED || wb < 8 || wb > 15 || level < 0 || level > 9 ||
        strategy < 0 || strategy > Z_FIXED || (wb == 8 && wrap != 1))
        return Z_STREAM_ERROR;
    if (!engine()) { strm->msg = kNoDevice; return Z_MEM_ERROR; }
    void *mem = strm->zalloc(strm->opaque, 1, (uInt)sizeof(DState));
    if (!mem) return Z_MEM_ERROR;
    DState *d = new (mem) DState();
    d->magic = kDMagic;
    d->level = level; d->strategy = strategy; d->window_bits = windowBits; d->mem_level = memLevel; d->wrap = wrap;
    d->gzhead = nullptr;
    strm->state = reinterpret_cast<internal_state *>(d);
    return deflateReset(strm);
}

int deflateInit_(z_streamp strm, int level, const char *version, int stream_size)
{
    return deflateInit2_(strm, level, Z_DEFLATED, MAX_WBITS, DEF_MEM_LEVEL, Z_DEFAULT_STRATEGY, version, stream_size);
}

int deflateResetKeep(z_streamp strm)
{
    DState *d = dstate(strm);
    if (!d) return Z_STREAM_ERROR;
    strm->total_in = strm->total_out = 0;
    strm->msg = nullptr;
    strm->data_type = Z_UNKNOWN;
    strm->adler = d->wrap == 2 ? 0 : 1;
    d->status = 1;
    d->last_flush = -2;
    d->header_done = false;
    d->any_segment = false;
    d->check = d->wrap == 2 ? 0 : 1;
    d->check_len = 0;
    d->in.clear(); d->out.clear(); d->out_pos = 0;
    return Z_OK;
}
int deflateReset(z_streamp strm) { return deflateResetKeep(strm); }

static int run_segment(z_streamp strm, DState *d, bool final)
{
    zb_deflate_result r;
    zb_engine *e = engine();
    if (!e) { strm->msg = kNoDevice; return Z_MEM_ERROR; }
    const size_t n = d->in.size();
    const size_t base = d->out.size();
    const size_t cap = zb_deflate_bound(n) + 64;
    int rc;
    if (final && !d->any_segment) {
        // the one-shot path: byte-identical to the reference's deflate(Z_FINISH) at every level, strategy and memLevel (32 KiB window)
        d->out.resize(base + cap);
        rc = zb_deflate_ex(e, d->in.data(), n, 0, d->out.data() + base, cap, 0, d->level, d->strategy, d->window_bits,
                           ZB_FLAG_MEMLEVEL(d->mem_level), &r);
        if (rc != ZB_OK) { d->out.resize(base); strm->msg = zb_last_error(); return map_rc(rc); }
        d->out.resize(base + r.out_bytes);
        strm->adler = r.check;
        if (strm->data_type == Z_UNKNOWN) strm->data_type = r.data_type;
        d->in.clear();
        return Z_OK;
    }
    // segment mode: framing header, raw segments closed by the sync marker, trailer at the end
    if (!d->header_done) {
        if (d->wrap == 1) {
            const unsigned lf = (d->strategy >= Z_HUFFMAN_ONLY || d->level < 2) ? 0 : d->level < 6 ? 1 : d->level == 6 ? 2 : 3;
            unsigned h = ((8u + (7u << 4)) << 8) | (lf << 6);
            h += 31 - (h % 31);
            d->out.push_back((uint8_t)(h >> 8)); d->out.push_back((uint8_t)h);
        } else if (d->wrap == 2) {
            const uint8_t xfl = d->level == 9 ? 2 : (d->strategy >= Z_HUFFMAN_ONLY || d->level < 2) ? 4 : 0;
            const uint8_t g[10] = {31, 139, 8, 0, 0, 0, 0, 0, xfl, 3};
            d->out.insert(d->out.end(), g, g + 10);
        }
        d->header_done = true;
    }
    const size_t b2 = d->out.size();
    d->out.resize(b2 + cap);
    const int lvl = d->level == 0 ? 1 : d->level; // stored segments need level > 0 framing
    rc = zb_deflate_ex(e, d->in.data(), n, 0, d->out.data() + b2, cap, 0, lvl, d->strategy, -15,
                       (final ? 0 : ZB_FLAG_NOT_LAST) | ZB_FLAG_MEMLEVEL(d->mem_level), &r);
    if (rc != ZB_OK) { d->out.resize(b2); strm->msg = zb_last_error(); return map_rc(rc); }
    d->out.resize(b2 + r.out_bytes);
    // running check value over all consumed input (segment checks chained with the combine algebra)
    uint32_t seg = 0;
    if (d->wrap == 1) { zb_adler32(e, 1, d->in.data(), n, 0, &seg, nullptr); d->check = adler_combine(d->check, seg, n); }
    else if (d->wrap == 2) { zb_crc32(e, 0, d->in.data(), n, 0, &seg, nullptr); d->check = multmodp(x2nmodp(n, 3), d->check) ^ seg; }
    d->check_len += n;
    strm->adler = d->check;
    if (strm->data_type == Z_UNKNOWN) strm->data_type = r.data_type;
    d->any_segment = true;
    d->in.clear();
    if (final) {
        if (d->wrap == 1) for (int i = 3; i >= 0; i--) d->out.push_back((uint8_t)(d->check >> (8 * i)));
        else if (d->wrap == 2) {
            for (int i = 0; i < 4; i++) d->out.push_back((uint8_t)(d->check >> (8 * i)));
            for (int i = 0; i < 4; i++) d->out.push_back((uint8_t)((uint32_t)d->check_len >> (8 * i)));
        }
    }
    return Z_OK;
}

int deflate(z_streamp strm, int flush)
{
    DState *d = dstate(strm);
    if (!d || flush < 0 || flush > Z_BLOCK) return Z_STREAM_ERROR;
    // zlib-rs/src/deflate.rs:2489-2540
    if (!strm->next_out || (strm->avail_in != 0 && !strm->next_in) || (d->status == 3 && flush != Z_FINISH)) {
        strm->msg = err_msg(Z_STREAM_ERROR);
        return Z_STREAM_ERROR;
    }
    if (strm->avail_out == 0) { strm->msg = err_msg(Z_BUF_ERROR); return Z_BUF_ERROR; }
    const int old_flush = d->last_flush;
    d->last_flush = flush;
    if (d->out_pos < d->out.size()) {
        drain(strm, d->out, d->out_pos);
        if (strm->avail_out == 0) { d->last_flush = -1; return Z_OK; }
    } else if (strm->avail_in == 0 && rank_flush(flush) <= rank_flush(old_flush) && flush != Z_FINISH) {
        strm->msg = err_msg(Z_BUF_ERROR);
        return Z_BUF_ERROR;
    }
    if (d->status == 3 && strm->avail_in != 0) { strm->msg = err_msg(Z_BUF_ERROR); return Z_BUF_ERROR; }
    if (d->status == 1) d->status = 2;
    // take all offered input (the engine works on whole segments; read_buf_window's role)
    if (strm->avail_in) {
        d->in.insert(d->in.end(), strm->next_in, strm->next_in + strm->avail_in);
        strm->next_in += strm->avail_in;
        strm->total_in += strm->avail_in;
        strm->avail_in = 0;
    }
    if (flush == Z_NO_FLUSH) return Z_OK;
    if (d->status != 3) {
        const bool final = flush == Z_FINISH;
        if (final || !d->in.empty() || !d->header_done) {
            int rc = run_segment(strm, d, final);
            if (rc != Z_OK) return rc;
        }
        if (final) d->status = 3;
    }
    drain(strm, d->out, d->out_pos);
    if (strm->avail_out == 0 && d->out_pos < d->out.size()) { d->last_flush = -1; return Z_OK; }
    if (flush != Z_FINISH) { if (strm->avail_out == 0) d->last_flush = -1; return Z_OK; }
    return d->out_pos < d->out.size() ? Z_OK : Z_STREAM_END;
}

int deflateEnd(z_streamp strm)
{
    DState *d = dstate(strm);
    if (!d) return Z_STREAM_ERROR;
    const int status = d->status;
    d->magic = 0;
    d->~DState();
    strm->zfree(strm->opaque, d);
    strm->state = nullptr;
    return status == 2 ? Z_DATA_ERROR : Z_OK; // libz-rs-sys/src/lib.rs:1583-1591
}

int deflateParams(z_streamp strm, int level, int strategy)
{
    DState *d = dstate(strm);
    if (!d) return Z_STREAM_ERROR;
    if (level == Z_DEFAULT_COMPRESSION) level = 6;
    if (level < 0 || level > 9 || strategy < 0 || strategy > Z_FIXED) return Z_STREAM_ERROR;
    if ((level != d->level || strategy != d->strategy) && d->last_flush != -2 && !d->in.empty()) {
        int rc = deflate(strm, Z_BLOCK);
        if (rc == Z_STREAM_ERROR) return rc;
        if (strm->avail_in || d->out_pos < d->out.size()) return Z_BUF_ERROR;
    }
    d->level = level;
    d->strategy = strategy;
    return Z_OK;
}

int deflateSetDictionary(z_streamp strm, const Bytef *, uInt)
{
    DState *d = dstate(strm);
    if (!d) return Z_STREAM_ERROR;
    strm->msg = "preset dictionaries are not implemented by the B200 engine";
    return Z_STREAM_ERROR;
}
int deflateGetDictionary(z_streamp strm, Bytef *, uInt *len)
{
    if (!dstate(strm)) return Z_STREAM_ERROR;
    if (len) *len = 0;
    return Z_OK;
}
int deflatePrime(z_streamp strm, int, int) { return dstate(strm) ? Z_BUF_ERROR : Z_STREAM_ERROR; }
int deflatePending(z_streamp strm, unsigned *pending, int *bits)
{
    DState *d = dstate(strm);
    if (!d) return Z_STREAM_ERROR;
    if (pending) *pending = (unsigned)(d->out.size() - d->out_pos);
    if (bits) *bits = 0;
    return Z_OK;
}
int deflateCopy(z_streamp dest, z_streamp source)
{
    DState *s = dstate(source);
    if (!s || !dest) return Z_STREAM_ERROR;
    *dest = *source;
    void *mem = dest->zalloc(dest->opaque, 1, (uInt)sizeof(DState));
    if (!mem) return Z_MEM_ERROR;
    DState *d = new (mem) DState(*s);
    dest->state = reinterpret_cast<internal_state *>(d);
    return Z_OK;
}
int deflateSetHeader(z_streamp strm, gz_headerp head)
{
    DState *d = dstate(strm);
    if (!d || d->wrap != 2) return Z_STREAM_ERROR;
    d->gzhead = head; // fields other than the defaults are not emitted by the engine yet
    return Z_OK;
}
uLong deflateBound(z_streamp strm, uLong sourceLen)
{
    (void)strm;
    return (uLong)zb_deflate_bound(sourceLen) + 32;
}
int deflateTune(z_streamp strm, int, int, int, int) { return dstate(strm) ? Z_OK : Z_STREAM_ERROR; }

uLong compressBound(uLong sourceLen) { return (uLong)zb_deflate_bound(sourceLen) + 32; }

int compress2(Bytef *dest, uLongf *destLen, const Bytef *source, uLong sourceLen, int level)
{
    // libz-rs-sys/src/lib.rs:1472-1558: NULL checks -> Z_STREAM_ERROR; too-small dest -> Z_BUF_ERROR
    if (!destLen || !dest || (!source && sourceLen)) return Z_STREAM_ERROR;
    if (level == Z_DEFAULT_COMPRESSION) level = 6;
    if (level < 0 || level > 9) return Z_STREAM_ERROR;
    zb_engine *e = engine();
    if (!e) return Z_MEM_ERROR;
    zb_deflate_result r;
    int rc = zb_deflate(e, source, sourceLen, 0, dest, *destLen, 0, level, Z_DEFAULT_STRATEGY, MAX_WBITS, &r);
    if (rc == ZB_OK) { *destLen = (uLong)r.out_bytes; return Z_OK; }
    return map_rc(rc);
}
int compress(Bytef *dest, uLongf *destLen, const Bytef *source, uLong sourceLen)
{
    return compress2(dest, destLen, source, sourceLen, Z_DEFAULT_COMPRESSION);
}

// ---------------------------------------------------------------- inflate
int inflateInit2_(z_streamp strm, int windowBits, const char *version, int stream_size)
{
    if (!version_ok(version, stream_size)) return Z_VERSION_ERROR;
    if (!strm) return Z_STREAM_ERROR;
    strm->msg = nullptr;
    if (!strm->zalloc) { strm->zalloc = default_alloc; strm->opaque = nullptr; }
    if (!strm->zfree) strm->zfree = default_free;
    // inflate::reset_with_config (zlib-rs/src/inflate.rs:2298-2327)
    int wb = windowBits;
    if (wb < 0) { if (wb < -15) return Z_STREAM_ERROR; wb = -wb; }
    else if (wb < 48) wb &= 15;
    if (wb != 0 && (wb < 8 || wb > 15)) return Z_STREAM_ERROR;
    if (!engine()) { strm->msg = kNoDevice; return Z_MEM_ERROR; }
    void *mem = strm->zalloc(strm->opaque, 1, (uInt)sizeof(IState));
    if (!mem) return Z_MEM_ERROR;
    IState *s = new (mem) IState();
    s->magic = kIMagic;
    s->window_bits = windowBits;
    strm->state = reinterpret_cast<internal_state *>(s);
    return inflateReset(strm);
}
int inflateInit_(z_streamp strm, const char *version, int stream_size) { return inflateInit2_(strm, MAX_WBITS, version, stream_size); }

int inflateReset(z_streamp strm)
{
    IState *s = istate(strm);
    if (!s) return Z_STREAM_ERROR;
    strm->total_in = strm->total_out = 0;
    strm->msg = nullptr;
    strm->adler = s->window_bits < 0 ? 0 : 1;
    s->status = 0;
    s->in.clear(); s->out.clear(); s->out_pos = 0; s->last_attempt = 0; s->consumed = 0;
    s->result = Z_OK;
    return Z_OK;
}
int inflateReset2(z_streamp strm, int windowBits)
{
    IState *s = istate(strm);
    if (!s) return Z_STREAM_ERROR;
    s->window_bits = windowBits;
    return inflateReset(strm);
}

// Decode everything collected so far.  Returns 1 when the stream is complete (or definitely bad).
static int try_decode(z_streamp strm, IState *s, bool must_finish)
{
    zb_engine *e = engine();
    if (!e) { s->status = -1; s->result = Z_MEM_ERROR; strm->msg = kNoDevice; return 1; }
    size_t cap = s->in.size() * 4 + 65536;
    for (;;) {
        s->out.resize(cap);
        zb_inflate_result r;
        int rc = zb_inflate(e, s->in.data(), s->in.size(), 0, s->out.data(), cap, 0, s->window_bits == 0 ? 15 : s->window_bits, &r);
        if (rc == ZB_E_BUF) { cap *= 4; continue; }
        if (rc == ZB_OK) {
            s->out.resize(r.out_bytes);
            s->consumed = r.in_bytes;
            strm->adler = r.check;
            s->status = 1;
            s->result = Z_STREAM_END;
            return 1;
        }
        s->out.clear();
        if (rc == ZB_E_DATA && strcmp(r.msg, "unexpected end of input") == 0 && !must_finish) return 0; // need more input
        s->status = -1;
        if (rc == ZB_E_DATA) {
            snprintf(s->msg, sizeof s->msg, "%s", r.msg);
            strm->msg = s->msg;
            s->result = strcmp(r.msg, "need dictionary") == 0 ? Z_NEED_DICT : Z_DATA_ERROR;
            if (strcmp(r.msg, "unexpected end of input") == 0) { s->status = 0; s->result = Z_BUF_ERROR; return 0; }
        } else { strm->msg = zb_last_error(); s->result = map_rc(rc); }
        return 1;
    }
}

int inflate(z_streamp strm, int flush)
{
    IState *s = istate(strm);
    if (!s) return Z_STREAM_ERROR;
    if (!strm->next_out || (!strm->next_in && strm->avail_in != 0)) return Z_STREAM_ERROR; // inflate.rs:2376-2379
    if (s->status == -1) return s->result;
    if (s->status == 2) return Z_STREAM_END;
    const size_t in0 = strm->avail_in;
    size_t produced = 0;
    if (s->status == 0) {
        if (strm->avail_in) {
            s->in.insert(s->in.end(), strm->next_in, strm->next_in + strm->avail_in);
            strm->next_in += strm->avail_in;
            strm->total_in += strm->avail_in;
            strm->avail_in = 0;
        }
        // The engine decodes whole streams: try when asked to finish, when no new input arrived, or when the
        // collected input has doubled since the last attempt (keeps total work linear for chunked callers).
        const bool attempt = flush == Z_FINISH || in0 == 0 || s->in.size() >= 2 * s->last_attempt || s->in.size() <= (1u << 20);
        if (attempt && !s->in.empty()) {
            s->last_attempt = s->in.size();
            try_decode(strm, s, false);
            if (s->status == -1) return s->result;
            if (s->status == 1 && s->consumed < s->in.size()) {
                // bytes after the end of the stream are handed back to the caller
                const size_t extra = s->in.size() - s->consumed;
                if (extra <= in0) { strm->next_in -= extra; strm->avail_in += (uInt)extra; strm->total_in -= extra; }
            }
        }
    }
    if (s->status == 1) {
        produced = drain(strm, s->out, s->out_pos);
        if (s->out_pos >= s->out.size() && s->out.empty()) { s->status = 2; return Z_STREAM_END; }
        return Z_OK;
    }
    if ((in0 == 0 && produced == 0) || flush == Z_FINISH) return Z_BUF_ERROR;
    return Z_OK;
}

int inflateEnd(z_streamp strm)
{
    IState *s = istate(strm);
    if (!s) return Z_STREAM_ERROR;
    s->magic = 0;
    s->~IState();
    strm->zfree(strm->opaque, s);
    strm->state = nullptr;
    return Z_OK;
}
int inflateSetDictionary(z_streamp strm, const Bytef *, uInt) { return istate(strm) ? Z_STREAM_ERROR : Z_STREAM_ERROR; }
int inflateGetHeader(z_streamp strm, gz_headerp head)
{
    if (!istate(strm)) return Z_STREAM_ERROR;
    if (head) head->done = 0;
    return Z_OK;
}
int inflateSync(z_streamp strm) { return istate(strm) ? Z_DATA_ERROR : Z_STREAM_ERROR; }
int inflateCopy(z_streamp dest, z_streamp source)
{
    IState *s = istate(source);
    if (!s || !dest) return Z_STREAM_ERROR;
    *dest = *source;
    void *mem = dest->zalloc(dest->opaque, 1, (uInt)sizeof(IState));
    if (!mem) return Z_MEM_ERROR;
    IState *d = new (mem) IState(*s);
    dest->state = reinterpret_cast<internal_state *>(d);
    return Z_OK;
}
long inflateMark(z_streamp strm) { return istate(strm) ? 0 : -(1L << 16); }
int inflatePrime(z_streamp strm, int, int) { return istate(strm) ? Z_STREAM_ERROR : Z_STREAM_ERROR; }

int uncompress2(Bytef *dest, uLongf *destLen, const Bytef *source, uLong *sourceLen)
{
    // libz-rs-sys/src/lib.rs:518-633 + zlib-rs/src/inflate.rs:195-277
    if (!destLen || !sourceLen || (!dest && *destLen) || (!source && *sourceLen)) return Z_STREAM_ERROR;
    zb_engine *e = engine();
    if (!e) return Z_MEM_ERROR;
    zb_inflate_result r;
    uint8_t dummy[1];
    const bool empty_dest = *destLen == 0;
    int rc = zb_inflate(e, source, *sourceLen, 0, empty_dest ? dummy : dest, empty_dest ? 1 : *destLen, 0, MAX_WBITS, &r);
    *sourceLen = (uLong)r.in_bytes;
    if (rc == ZB_OK) {
        if (empty_dest && r.out_bytes) { *destLen = 0; return Z_BUF_ERROR; }
        *destLen = (uLong)r.out_bytes;
        return Z_OK;
    }
    *destLen = empty_dest ? 0 : (uLong)r.out_bytes;
    if (rc == ZB_E_DATA) return Z_DATA_ERROR; // includes truncated input and NEED_DICT (inflate.rs:271-276)
    if (rc == ZB_E_BUF) return Z_BUF_ERROR;
    return map_rc(rc);
}
int uncompress(Bytef *dest, uLongf *destLen, const Bytef *source, uLong sourceLen)
{
    return uncompress2(dest, destLen, source, &sourceLen);
}

// ---------------------------------------------------------------- checksums
static void need_engine_or_die()
{
    if (!engine()) {
        fprintf(stderr, "libz_b200: %s (%s); adler32/crc32 have no CPU fallback\n", kNoDevice, zb_last_error());
        abort();
    }
}
uLong adler32_z(uLong adler, const Bytef *buf, z_size_t len)
{
    if (!buf) return 1; // libz-rs-sys/src/lib.rs:307-312
    need_engine_or_die();
    uint32_t out = 0;
    if (zb_adler32(engine(), (uint32_t)adler, buf, len, 0, &out, nullptr) != ZB_OK) {
        fprintf(stderr, "libz_b200: adler32 failed: %s\n", zb_last_error());
        abort();
    }
    return out;
}
uLong adler32(uLong adler, const Bytef *buf, uInt len) { return adler32_z(adler, buf, len); }
uLong crc32_z(uLong crc, const Bytef *buf, z_size_t len)
{
    if (!buf) return 0; // libz-rs-sys/src/lib.rs:150-155
    need_engine_or_die();
    uint32_t out = 0;
    if (zb_crc32(engine(), (uint32_t)crc, buf, len, 0, &out, nullptr) != ZB_OK) {
        fprintf(stderr, "libz_b200: crc32 failed: %s\n", zb_last_error());
        abort();
    }
    return out;
}
uLong crc32(uLong crc, const Bytef *buf, uInt len) { return crc32_z(crc, buf, len); }
uLong adler32_combine64(uLong a1, uLong a2, z_off64_t len2) { return len2 < 0 ? 0xffffffffUL : adler_combine((uint32_t)a1, (uint32_t)a2, (uint64_t)len2); }
uLong adler32_combine(uLong a1, uLong a2, z_off_t len2) { return adler32_combine64(a1, a2, len2); }
uLong crc32_combine_gen64(z_off64_t len2) { return x2nmodp((uint64_t)len2, 3); }
uLong crc32_combine_gen(z_off_t len2) { return crc32_combine_gen64(len2); }
uLong crc32_combine_op(uLong c1, uLong c2, uLong op) { return multmodp((uint32_t)op, (uint32_t)c1) ^ (uint32_t)c2; }
uLong crc32_combine64(uLong c1, uLong c2, z_off64_t len2) { return crc32_combine_op(c1, c2, crc32_combine_gen64(len2)); }
uLong crc32_combine(uLong c1, uLong c2, z_off_t len2) { return crc32_combine64(c1, c2, len2); }

} // extern "C"
