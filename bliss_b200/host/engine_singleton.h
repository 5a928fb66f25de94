/* Process-wide engine pool used by the bliss.h wrappers (bl_analyze, bl_*_sort, ...). */
#ifndef BLX_ENGINE_SINGLETON_H
#define BLX_ENGINE_SINGLETON_H
#include "../../include/blx.h"

/* Takes an engine out of the process-wide pool for the calling thread (creating one on device $BLISS_DEVICE or 0
 * when all are in use, up to 16) and returns it; NULL (after printing the reason to stderr) when no B200 is usable.
 * Pair every non-NULL return with bl_engine_release() on the same thread. */
blx_engine *bl_engine_acquire(void);
void bl_engine_release(void);
#endif
