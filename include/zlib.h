/* zlib.h -- shim: a C client written against <zlib.h> (e.g. the reference's libz-rs-sys-cdylib/zpipe.c) compiles against
 * libz_b200.so's header unchanged with -I<repo>/include. */
#ifndef ZLIB_H
#define ZLIB_H
#include "zlib_b200.h"
#endif
