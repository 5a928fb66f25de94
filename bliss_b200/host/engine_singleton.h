/* Process-wide engine used by the bliss.h wrappers (bl_analyze, bl_*_sort, ...). */
#ifndef BLX_ENGINE_SINGLETON_H
#define BLX_ENGINE_SINGLETON_H
#include "../../include/blx.h"

/* Locks the process-wide engine (creating it on first use on device $BLISS_DEVICE or 0) and
 * returns it; NULL (after printing the reason to stderr) when no B200 is usable.
 * Pair every non-NULL return with bl_engine_release(). */
blx_engine *bl_engine_acquire(void);
void bl_engine_release(void);
#endif
