/* mini-hts: bamutil.h:8 includes <htslib/kstring.h> but uses nothing from it. TEST INFRASTRUCTURE ONLY. */
#ifndef MINIHTS_KSTRING_H
#define MINIHTS_KSTRING_H
#include <stddef.h>
typedef struct kstring_t { size_t l, m; char *s; } kstring_t;
#endif
