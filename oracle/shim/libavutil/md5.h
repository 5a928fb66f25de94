/* TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for <libavutil/md5.h>
 * (include/bliss.h:6 of the reference includes it; the hot path never calls it). */
#ifndef ORACLE_SHIM_MD5_H
#define ORACLE_SHIM_MD5_H
#endif
