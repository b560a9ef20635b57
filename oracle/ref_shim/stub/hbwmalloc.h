/* Stand-in for memkind's <hbwmalloc.h>: the KNL "high-bandwidth" allocator mapped onto
 * plain 64-byte-aligned host allocations, so the reference's DDR/MCDRAM split degenerates
 * to one memory tier.  Test infrastructure only -- see oracle/README.md. */
#pragma once
#include <stdlib.h>
static inline void *hbw_malloc(size_t bytes)
{
	void *p = NULL;
	return posix_memalign(&p, 64, bytes ? bytes : 64) ? NULL : p;
}
static inline void hbw_free(void *p) { free(p); }
static inline void *hbw_realloc(void *p, size_t bytes) { return realloc(p, bytes); }
static inline int hbw_posix_memalign(void **p, size_t align, size_t bytes)
{
	return posix_memalign(p, align, bytes);
}
